"""Device-side mirrors of the environment calls either side of the planner:

* ``collision_check``        ``PlanetaryEnv.collision_check`` (src/simulator/planetary_env.py:221-232)
* ``traversability``         ``UnicycleModel.get_traversability`` (robot_model.py:102-112) in either mode
* ``BatchedPlanetaryEnv``    ``PlanetaryEnv.reset/step/collision_check`` (planetary_env.py:141-232) for E
                             environments stepped in one launch, states resident in HBM (closes the loop of
                             BASELINE config 3 without host round trips)

All arithmetic is in libbnvmppi.so (``bnv_trav_lookup`` / ``bnv_env_step``); rendering, gym plumbing and random
start/goal placement are out of scope.  Observation-mode draws come from the engine's Philox stream keyed by
``seed`` and a call counter, or are injected (``xi``) for parity tests.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _cabi
from .mppi import _slip_distribution


def _grid_of(grid_map, pitch: int) -> _cabi.Grid:
    return _cabi.make_grid(grid_map.grid_size, pitch, grid_map.resolution, grid_map.x_limits, grid_map.y_limits)


def _cuda(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device, torch.float32).contiguous()


def traversability(grid_map, states: torch.Tensor, *, mean: torch.Tensor, std: Optional[torch.Tensor] = None,
                   xi: Optional[torch.Tensor] = None, seed: int = 0, counter: int = 0,
                   stuck_threshold: Optional[float] = None):
    """Traversability at ``states`` [..., >=2] on the map ``mean`` ([G,G] risk map, inference mode) or the slip
    distribution (``mean``, ``std``) (observation mode).  Returns trav [...] and, with ``stuck_threshold``, also the
    uint8 collision mask.  Tensors must live on a CUDA device (the map's device decides)."""
    dev = mean.device
    if dev.type != "cuda":
        raise RuntimeError("benchnav_b200.env runs on CUDA (sm_100a) only; there is no CPU fallback")
    lib = _cabi.load()
    mean = _cuda(mean, dev)
    std_ptr = None
    if std is not None:
        std = _cuda(std, dev)
        if std.stride(0) != mean.stride(0):
            raise ValueError("mean and std maps must share their layout")
        std_ptr = std.data_ptr()
    pos = _cuda(states, dev)
    lead, stride = pos.shape[:-1], pos.shape[-1]
    n = int(pos.numel() // stride)
    xi_ptr = None
    if xi is not None:
        xi = _cuda(xi, dev)
        if xi.numel() != n:
            raise ValueError("xi must hold one normal per position")
        xi_ptr = xi.data_ptr()
    trav = torch.empty(lead, device=dev, dtype=torch.float32)
    stuck = torch.empty(lead, device=dev, dtype=torch.uint8) if stuck_threshold is not None else None
    grid = _grid_of(grid_map, mean.stride(0))
    with torch.cuda.device(dev):
        _cabi.check(lib.bnv_trav_lookup(C.byref(grid), mean.data_ptr(), std_ptr, 0, 0, pos.data_ptr(), n, stride, xi_ptr,
                                        int(seed), int(counter), None, float(stuck_threshold or 0.0), trav.data_ptr(),
                                        stuck.data_ptr() if stuck is not None else None,
                                        torch.cuda.current_stream(dev).cuda_stream))
    return (trav, stuck) if stuck is not None else trav


def collision_check(grid_map, states: torch.Tensor, stuck_threshold: float, *, mean: torch.Tensor,
                    std: Optional[torch.Tensor] = None, xi: Optional[torch.Tensor] = None, seed: int = 0,
                    counter: int = 0) -> torch.Tensor:
    """planetary_env.py:221-232: ``get_traversability(states) <= stuck_threshold`` (bool, shape of states[..., 0])."""
    _, stuck = traversability(grid_map, states, mean=mean, std=std, xi=xi, seed=seed, counter=counter,
                              stuck_threshold=stuck_threshold)
    return stuck.bool()


class BatchedPlanetaryEnv:
    """E planetary environments advanced together on the device (planetary_env.py:141-232).

    ``grid_maps``: E reference-shaped ``GridMap`` objects of common geometry whose ``distributions["latent_models"]``
    hold the true slip models; ``start_pos`` / ``goal_pos`` [E,2].  ``step(actions [E,2])`` returns
    ``(robot_states [E,3], rewards [E], is_terminated [E] bool, is_truncated bool)`` like the reference's ``step`` per
    environment."""

    def __init__(self, grid_maps, start_pos: torch.Tensor, goal_pos: torch.Tensor, delta_t: float = 0.1,
                 time_limit: float = 100, stuck_threshold: float = 0.1, goal_threshold: float = 1.0, seed: int = 0,
                 device=torch.device("cuda"), min_action=(0.0, -1.0), max_action=(1.0, 1.0),
                 graph_capturable: bool = False) -> None:
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("benchnav_b200.BatchedPlanetaryEnv runs on CUDA (sm_100a) only; there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device, self._lib = dev, _cabi.load()
        self._grid_maps = list(grid_maps)
        self._num_envs = E = len(self._grid_maps)
        g0 = self._grid_maps[0]
        for gm in self._grid_maps[1:]:
            if (gm.grid_size, gm.resolution, tuple(gm.x_limits), tuple(gm.y_limits)) != \
                    (g0.grid_size, g0.resolution, tuple(g0.x_limits), tuple(g0.y_limits)):
                raise ValueError("all environments of a batch must share the grid geometry")
        ms = [_slip_distribution(gm, "latent_models") for gm in self._grid_maps]
        self._mean = torch.stack([_cuda(m, "cpu") for m, _ in ms]).to(dev).contiguous()
        self._std = torch.stack([_cuda(s, "cpu") for _, s in ms]).to(dev).contiguous()
        self._grid = _grid_of(g0, self._mean.stride(1))
        self._delta_t, self._time_limit = float(delta_t), time_limit
        self.stuck_threshold, self._goal_threshold = stuck_threshold, float(goal_threshold)
        self._seed, self._counter = int(seed), 0
        # graph_capturable: the draw counter lives in device memory and is advanced by a (capturable) in-place add, so a
        # captured step()/collision_check() draws fresh normals on every replay of the graph
        self._counter_dev = torch.zeros(1, dtype=torch.int64, device=dev) if graph_capturable else None
        self._u_min = (C.c_float * 2)(*min_action)
        self._u_max = (C.c_float * 2)(*max_action)
        assert tuple(start_pos.shape) == (E, 2) and tuple(goal_pos.shape) == (E, 2)
        self._start_pos, self._goal_pos = _cuda(start_pos, dev), _cuda(goal_pos, dev)
        self._reward = torch.full((E,), float("nan"), device=dev)
        self._terminated = torch.zeros(E, dtype=torch.uint8, device=dev)
        self._elapsed_time = 0
        self._robot_state = self._initial_state()

    def _initial_state(self) -> torch.Tensor:
        d = self._goal_pos - self._start_pos  # planetary_env.py:134-138: heading towards the goal
        return torch.cat([self._start_pos, torch.atan2(d[:, 1], d[:, 0]).unsqueeze(1)], dim=1).contiguous()

    def reset(self, seed: Optional[int] = None) -> torch.Tensor:
        if seed is not None:
            self._seed = int(seed)
        self._counter, self._elapsed_time = 0, 0
        if self._counter_dev is not None:
            self._counter_dev.zero_()
        self._robot_state = self._initial_state()
        self._reward.fill_(float("nan"))
        return self._robot_state

    def step(self, actions: torch.Tensor, xi: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, bool]:
        E = self._num_envs
        assert tuple(actions.shape) == (E, 2)
        actions = _cuda(actions, self._device)
        xi_ptr = None
        if xi is not None:
            xi = _cuda(xi, self._device)
            assert tuple(xi.shape) == (E,)
            xi_ptr = xi.data_ptr()
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_env_step(
                C.byref(self._grid), self._mean.data_ptr(), self._std.data_ptr(), self._mean.stride(0), E,
                self._robot_state.data_ptr(), actions.data_ptr(), self._goal_pos.data_ptr(), xi_ptr, self._seed,
                0 if self._counter_dev is not None else self._counter,
                self._counter_dev.data_ptr() if self._counter_dev is not None else None, self._u_min, self._u_max,
                self._delta_t, self._goal_threshold, self._reward.data_ptr(), self._terminated.data_ptr(),
                torch.cuda.current_stream(self._device).cuda_stream))
        self._advance_counter()
        self._elapsed_time += self._delta_t  # planetary_env.py:210
        self._keepalive = (actions, xi)
        return self._robot_state, self._reward, self._terminated.bool(), self._elapsed_time > self._time_limit

    def collision_check(self, states: torch.Tensor, xi: Optional[torch.Tensor] = None) -> torch.Tensor:
        """states [E,P,3] (environment e's positions against environment e's map) -> bool [E,P]."""
        E = self._num_envs
        assert states.dim() == 3 and states.shape[0] == E
        pos = _cuda(states, self._device)
        n, rows = int(pos.shape[0] * pos.shape[1]), int(pos.shape[1])
        xi_ptr = None
        if xi is not None:
            xi = _cuda(xi, self._device)
            xi_ptr = xi.data_ptr()
        stuck = torch.empty(E, rows, device=self._device, dtype=torch.uint8)
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_trav_lookup(
                C.byref(self._grid), self._mean.data_ptr(), self._std.data_ptr(), self._mean.stride(0), rows,
                pos.data_ptr(), n, pos.shape[-1], xi_ptr, self._seed,
                (1 << 40) + (0 if self._counter_dev is not None else self._counter),
                self._counter_dev.data_ptr() if self._counter_dev is not None else None,
                float(self.stuck_threshold), None, stuck.data_ptr(), torch.cuda.current_stream(self._device).cuda_stream))
        self._advance_counter()
        return stuck.bool()

    def closed_loop_step(self, actions: torch.Tensor, planned: torch.Tensor, done: torch.Tensor,
                         steps_to_goal: torch.Tensor, step_no: torch.Tensor, collisions: torch.Tensor,
                         planner=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Everything Tutorial 3.3's loop does between two planner calls, in ONE launch (``bnv_closed_loop_step``):
        ``step`` with the first planned action (robots with ``done[e]`` set stop), ``collision_check`` of the planned
        trajectory into ``collisions`` [E,T+1] (uint8), and the loop's books -- ``done`` (uint8 [E], in place),
        ``steps_to_goal`` (int64 [E]), ``step_no`` (int64 scalar, +1) and this environment's draw counter (+2).

        ``actions`` [E,T,2] and ``planned`` [E,1,T+1,3] (or [E,T+1,3]) are the planner's outputs, untouched.  With
        ``planner`` (a ``BatchedMPPI`` in ``graph_capturable(True, external_advance=True)`` mode) its iteration counter
        is advanced too, so a captured control step is two kernels.  Needs ``graph_capturable=True`` at construction
        (device-resident draw counter).  Bit-identical to ``step`` + ``collision_check`` + the torch bookkeeping."""
        if self._counter_dev is None:
            raise RuntimeError("closed_loop_step needs BatchedPlanetaryEnv(..., graph_capturable=True)")
        E = self._num_envs
        T = int(actions.shape[1])
        assert tuple(actions.shape) == (E, T, 2) and planned.numel() == E * (T + 1) * 3
        for t, dt_ in ((actions, torch.float32), (planned, torch.float32), (done, torch.uint8), (collisions, torch.uint8),
                       (steps_to_goal, torch.int64), (step_no, torch.int64)):
            if not (t.is_cuda and t.dtype == dt_ and t.is_contiguous()):
                raise ValueError("closed_loop_step takes contiguous CUDA tensors of the documented dtypes")
        if getattr(self, "_loop_ticket", None) is None:
            self._loop_ticket = torch.zeros(1, dtype=torch.int32, device=self._device)
        it_ptr = planner.iteration_counter_ptr if planner is not None else 0
        with torch.cuda.device(self._device):
            _cabi.check(self._lib.bnv_closed_loop_step(
                C.byref(self._grid), self._mean.data_ptr(), self._std.data_ptr(), self._mean.stride(0), E, T,
                self._robot_state.data_ptr(), actions.data_ptr(), planned.data_ptr(), self._goal_pos.data_ptr(),
                self._seed, self._counter_dev.data_ptr(), self._u_min, self._u_max, self._delta_t, self._goal_threshold,
                float(self.stuck_threshold), self._reward.data_ptr(), self._terminated.data_ptr(), collisions.data_ptr(),
                done.data_ptr(), steps_to_goal.data_ptr(), step_no.data_ptr(), it_ptr or None,
                self._loop_ticket.data_ptr(), torch.cuda.current_stream(self._device).cuda_stream))
        self._counter += 2
        self._elapsed_time += self._delta_t
        return self._robot_state, self._reward, self._terminated

    def _advance_counter(self) -> None:
        self._counter += 1
        if self._counter_dev is not None:
            self._counter_dev += 1
