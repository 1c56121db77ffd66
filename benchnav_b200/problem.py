"""Minimal problem-description objects with the attribute surface the engine introspects.

The engine reads the reference's own ``GridMap`` / ``UnicycleModel`` / ``Objectives`` instances
(src/environments/grid_map.py:12-61, src/simulator/problem_formulation/robot_model.py:15-57,
objectives.py:11-27) through the attributes listed in SURVEY 8b.  Where the reference package is not
installed (the GPU box, the tests, bench.py) these data carriers expose the same attributes; they hold
data only -- every computation happens in the CUDA library.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch


@dataclass
class ModelConfig:
    mode: str = "inference"
    inference_metric: Optional[str] = "expected_value"
    confidence_value: Optional[float] = None


class GridSpec:
    """Geometry of a square 2.5D grid map; limits follow grid_map.py:42-50."""

    def __init__(self, grid_size: int, resolution: float, device: str = "cpu", distributions=None) -> None:
        # distributions: {"predictions": d, "latent_models": d} with d.mean / d.stddev [G,G] (grid_map.py:24-33)
        self.distributions = distributions if distributions is not None else {}
        self.grid_size = int(grid_size)
        self.resolution = resolution
        self.center_x = self.center_y = grid_size * resolution / 2
        self.x_limits = (self.center_x - grid_size / 2 * resolution, self.center_x + grid_size / 2 * resolution)
        self.y_limits = (self.center_y - grid_size / 2 * resolution, self.center_y + grid_size / 2 * resolution)
        self.device = device


@dataclass
class SlipDistribution:
    """Per-cell Normal slip model: the two tensors a ``torch.distributions.Normal`` exposes as mean / stddev."""

    mean: torch.Tensor
    stddev: torch.Tensor


class _RiskHolder:
    def __init__(self, risks: torch.Tensor) -> None:
        self._risks = risks


class UnicycleProblem:
    """Dynamics-side carrier: grid geometry, inference-mode risk map, action bounds (robot_model.py:54-57)."""

    def __init__(self, grid: GridSpec, risks: torch.Tensor, delta_t: float = 0.1,
                 min_action: Sequence[float] = (0.0, -1.0), max_action: Sequence[float] = (1.0, 1.0)) -> None:
        if tuple(risks.shape) != (grid.grid_size, grid.grid_size):
            raise ValueError("risk map must be [grid_size, grid_size]")
        self._grid_map = grid
        self._model_config = ModelConfig()
        self._traversability_model = _RiskHolder(risks.to(torch.float32))
        self.min_action = torch.tensor(list(min_action), dtype=torch.float32)
        self.max_action = torch.tensor(list(max_action), dtype=torch.float32)
        self._delta_t = float(delta_t)

    def transit(self, state, action, delta_t: float = 0.1):  # signature carrier only (default dt is introspected)
        raise NotImplementedError("UnicycleProblem carries data; rollouts run inside the CUDA engine")


class GoalObjectives:
    """Objectives-side carrier (objectives.py:25-27)."""

    def __init__(self, dynamics: UnicycleProblem, goal_pos: torch.Tensor, stuck_threshold: float) -> None:
        self._dynamics = dynamics
        self._goal_pos = goal_pos
        self._stuck_threshold = stuck_threshold
