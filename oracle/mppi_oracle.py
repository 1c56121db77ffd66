"""CPU oracle for the MPPI hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  ``benchnav_b200`` never does:
the product path is the sm_100a CUDA library and fails loudly without it.

What this is
------------
A functional restatement, in plain PyTorch-on-CPU tensor arithmetic, of one
control iteration of the reference planner (all citations are relative to the
reference repository root):

* ``MPPI.forward``                         src/planners/local_planners/mppi.py:130-219
* ``MPPI.get_top_samples``                 src/planners/local_planners/mppi.py:221-240
* ``UnicycleModel.transit``                src/simulator/problem_formulation/robot_model.py:59-100
* ``Objectives.stage_cost/terminal_cost``  src/simulator/problem_formulation/objectives.py:29-65
* ``TraversabilityModel.get_traversability`` (inference mode)
                                           src/simulator/problem_formulation/traversability_model.py:70-72
* ``GridMap.get_values_at_positions`` / ``get_grid_indices_from_positions``
                                           src/environments/grid_map.py:145-181, 183-210

The arithmetic itself lives in PyTorch's ATen CPU kernels (third-party; the
reference pins ``torch==2.00`` in pyproject.toml:21, this image has 2.11.0).
The oracle therefore uses the *same ATen ops in the same order* (fp32 ``sub``,
true ``div``, ``floor``, ``clamp``, ``cos``/``sin``, ``remainder``, ``norm``,
``sum``, ``softmax``) so that in fp32 it reproduces the reference bit for bit
on the recorded states and weights.  Run with ``dtype=torch.float64`` it is the
"truth" used to size the fp32 tolerance.

Parity pinning
--------------
The reference's own tests contain no assertions or fixtures for this path
(test/test_mppi.py is a visual demo needing absent datasets), so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build
container: ``tests/golden/make_golden.py`` imports the unmodified reference,
runs ``MPPI.forward`` on synthetic maps, and commits inputs/outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this module
against them (bit-exact for recorded states, weights and controls).

Vocabulary (matches DESIGN.md)
------------------------------
``noise``   [K,T,2]  the sigma-scaled control noise, exactly ``MPPI._action_noises``
``u_prev``  [T,2]    previous optimal sequence (``_previous_action_seq``), no time shift
``risk``    [G,G]    ``TraversabilityModel._risks`` (row = y cell, column = x cell)
``rec``     [K,T+1,3] what the reference leaves in ``_state_seq_batch``: slot t<T
                      holds the *pre-clamp, pre-wrap* successor of step t (in-place
                      aliasing in ``transit``), slot T the clamped/wrapped final state
"""

from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Dict, Optional, Sequence, Tuple

import torch

STUCK_PENALTY = 1e4  # objectives.py:53


@dataclass
class Problem:
    """Everything ``forward`` reads through ``dynamics`` / ``objectives``."""

    risk: torch.Tensor  # [G,G] float32
    resolution: float
    x_min: float
    y_min: float
    x_max: float
    y_max: float
    goal: Tuple[float, float]
    stuck_threshold: float
    u_min: Tuple[float, float] = (0.0, -1.0)  # robot_model.py:54-57
    u_max: Tuple[float, float] = (1.0, 1.0)
    dt: float = 0.1  # robot_model.py:60 (MPPI never overrides it)
    # stochastic-slip mode (BASELINE config 4): `risk` holds the slip MEAN and `slip_std` its standard deviation;
    # lookups then sample the cell's Normal (observation-mode lookup, traversability_model.py:65-69)
    slip_std: Optional[torch.Tensor] = None

    @property
    def grid_size(self) -> int:
        return int(self.risk.shape[0])


def make_problem(risk: torch.Tensor, resolution: float, goal: Sequence[float],
                 stuck_threshold: float) -> Problem:
    """Limits exactly as GridMap computes them (grid_map.py:42-50)."""
    g = int(risk.shape[0])
    center = g * resolution / 2
    lo = center - g / 2 * resolution
    hi = center + g / 2 * resolution
    return Problem(risk=risk.to(torch.float32), resolution=float(resolution), x_min=lo, y_min=lo,
                   x_max=hi, y_max=hi, goal=(float(goal[0]), float(goal[1])),
                   stuck_threshold=float(stuck_threshold))


# --------------------------------------------------------------------------- lookup
def cell_indices(p: Problem, xy: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """grid_map.py:195-210 -- (pos - min) / r -> floor -> int -> clamp.

    The subtrahend is a default-dtype (fp32) tensor in the reference and the
    divisor a Python float, i.e. a true division in the tensor's dtype.
    """
    origin = torch.tensor([p.x_min, p.y_min], dtype=torch.float32).to(xy.dtype)
    idx = ((xy - origin) / p.resolution).floor().int().clamp(0, p.grid_size - 1)
    return idx[..., 0].long(), idx[..., 1].long()


def traversability(p: Problem, xy: torch.Tensor, xi: Optional[torch.Tensor] = None) -> torch.Tensor:
    """traversability_model.py:70-72 + grid_map.py:167: 1 - clamp(risk[iy, ix], 0, 1).

    With ``xi`` (standard normals, one per position) the observation-mode lookup instead
    (traversability_model.py:65-69, grid_map.py:169-178): ``Normal(mean[cell], std[cell]).sample()`` is ATen's
    ``normal_(0, 1).mul_(std).add_(mean)``, i.e. ``xi * std + mean`` with two roundings.
    """
    ix, iy = cell_indices(p, xy)
    if xi is None:
        return 1 - torch.clamp(p.risk.to(xy.dtype)[iy, ix], 0, 1)
    assert p.slip_std is not None, "stochastic lookup needs Problem.slip_std"
    sample = xi.to(xy.dtype).mul(p.slip_std.to(xy.dtype)[iy, ix]).add(p.risk.to(xy.dtype)[iy, ix])
    return 1 - torch.clamp(sample, 0, 1)


# --------------------------------------------------------------------------- dynamics
def unicycle_step(p: Problem, s: torch.Tensor, u: torch.Tensor, xi: Optional[torch.Tensor] = None,
                  dt: Optional[float] = None, return_trav: bool = False):
    """One ``transit`` (robot_model.py:75-95) written without aliasing.

    Returns ``(raw, nxt)``: ``raw`` is the un-clamped / un-wrapped successor (what
    the reference's in-place ``+=`` leaves in the *input* slot) and ``nxt`` the
    clamped/wrapped state it returns.  ``xi``: lookup normals (stochastic / observation mode).
    """
    tau = traversability(p, s[:, :2], xi)
    if dt is not None:
        p = replace(p, dt=dt)
    v = torch.clamp(u[:, 0], p.u_min[0], p.u_max[0])
    w = torch.clamp(u[:, 1], p.u_min[1], p.u_max[1])
    th = s[:, 2]
    x = s[:, 0] + tau * v * torch.cos(th) * p.dt  # ((tau*v)*cos)*dt, then add
    y = s[:, 1] + tau * v * torch.sin(th) * p.dt
    th_raw = th + tau * w * p.dt
    th_wrapped = (th_raw + torch.pi) % (2 * torch.pi) - torch.pi
    raw = torch.stack([x, y, th_raw], dim=1)
    nxt = torch.stack([torch.clamp(x, p.x_min, p.x_max), torch.clamp(y, p.y_min, p.y_max), th_wrapped], dim=1)
    if return_trav:
        return raw, nxt, tau
    return raw, nxt


def rollout(p: Problem, state: torch.Tensor, controls: torch.Tensor, xi: Optional[torch.Tensor] = None
            ) -> torch.Tensor:
    """mppi.py:160-165 (and :209-214 for the batch-1 optimal rollout).

    ``controls`` [B,T,2] -> ``rec`` [B,T+1,3] with the aliasing quirk reproduced.
    ``xi`` [B,T]: lookup normals of the T transits (stochastic mode).
    """
    b, t_h = controls.shape[0], controls.shape[1]
    rec = torch.zeros(b, t_h + 1, 3, dtype=controls.dtype)
    cur = state.to(controls.dtype).repeat(b, 1)
    for t in range(t_h):
        raw, cur = unicycle_step(p, cur, controls[:, t, :], None if xi is None else xi[:, t])
        rec[:, t, :] = raw
    rec[:, t_h, :] = cur
    return rec


# --------------------------------------------------------------------------- costs
def stage_cost(p: Problem, s: torch.Tensor, xi: Optional[torch.Tensor] = None,
               goal: Optional[Sequence[float]] = None) -> torch.Tensor:
    """objectives.py:46-53 on one [B,3] slice of ``rec`` (``goal`` overrides the problem's: DWA's sub-goal)."""
    goal = torch.tensor(p.goal if goal is None else goal, dtype=torch.float32).to(s.dtype)
    dist = torch.norm(s[:, :2] - goal, dim=1)
    stuck = traversability(p, s[:, :2], xi) <= p.stuck_threshold
    return dist + STUCK_PENALTY * stuck


def sample_costs(p: Problem, rec: torch.Tensor, controls: torch.Tensor, u_prev: torch.Tensor,
                 sigmas: torch.Tensor, lam: float, xi_stage: Optional[torch.Tensor] = None,
                 xi_term: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mppi.py:168-190: sum_t stage + terminal + sum_t lambda * u_prev[t]^T Sigma^-1 v[k,t].
    ``xi_stage`` [K,T] / ``xi_term`` [K]: lookup normals of the cost evaluations (stochastic mode)."""
    k, t_h = controls.shape[0], controls.shape[1]
    dt = controls.dtype
    inv_cov = torch.inverse(torch.diag(sigmas.to(torch.float32) ** 2)).to(dt)  # mppi.py:94-97
    stage = torch.zeros(k, t_h, dtype=dt)
    act = torch.zeros(k, t_h, dtype=dt)
    for t in range(t_h):
        stage[:, t] = stage_cost(p, rec[:, t, :], None if xi_stage is None else xi_stage[:, t])
        act[:, t] = u_prev[t] @ inv_cov @ controls[:, t].T
    terminal = stage_cost(p, rec[:, -1, :], xi_term)
    return torch.sum(stage, dim=1) + terminal + torch.sum(lam * act, dim=1)


# --------------------------------------------------------------------------- one iteration
def mppi_iteration(p: Problem, state: torch.Tensor, u_prev: torch.Tensor, noise: torch.Tensor,
                   sigmas: torch.Tensor, lam: float, dtype: torch.dtype = torch.float32,
                   xi: Optional[torch.Tensor] = None, xi_opt: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """One ``MPPI.forward`` (mppi.py:146-217) with the noise injected.

    Returns u_opt [T,2], opt_rec [1,T+1,3], rec [K,T+1,3], weights [K], costs [K], controls [K,T,2].

    Stochastic-slip mode (BASELINE config 4; no native reference -- SURVEY 8c): the same function with every
    ``get_traversability`` call sampling the cell's slip prediction.  ``xi`` [K,2T+1] holds, per sample, the normals of
    (transit 0, stage 0, transit 1, stage 1, ..., terminal), ``xi_opt`` [T] those of the optimal rollout's transits.
    """
    xi_tr = xi_st = xi_term = None
    if xi is not None:
        t_h = noise.shape[1]
        xi = xi.to(dtype)
        xi_tr, xi_st, xi_term = xi[:, 0:2 * t_h:2], xi[:, 1:2 * t_h:2], xi[:, 2 * t_h]
    state = state.to(dtype)
    u_prev = u_prev.to(dtype)
    noise = noise.to(dtype)
    lo = torch.tensor(p.u_min, dtype=torch.float32).to(dtype)
    hi = torch.tensor(p.u_max, dtype=torch.float32).to(dtype)
    controls = torch.clamp(u_prev + noise, lo, hi)  # mppi.py:152-157
    rec = rollout(p, state, controls, xi_tr)
    costs = sample_costs(p, rec, controls, u_prev, sigmas, lam, xi_st, xi_term)
    weights = torch.softmax(-costs / lam, dim=0)  # mppi.py:193
    u_opt = torch.sum(weights.view(-1, 1, 1) * controls, dim=0)  # mppi.py:196-199
    opt_rec = rollout(p, state, u_opt.repeat(1, 1, 1),
                      None if xi_opt is None else xi_opt.to(dtype).view(1, -1))  # mppi.py:202-214
    return {"u_opt": u_opt, "opt_rec": opt_rec, "rec": rec, "weights": weights, "costs": costs,
            "controls": controls}


def top_samples(rec: torch.Tensor, weights: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """mppi.py:229-240: top-n by weight, returned in descending weight order."""
    assert n <= weights.shape[0]
    idx = torch.topk(weights, n).indices
    w = weights[idx]
    order = torch.argsort(w, descending=True)
    return rec[idx][order], w[order]


# --------------------------------------------------------------------------- sharded softmax
def shard_partial(costs: torch.Tensor, controls: torch.Tensor, lam: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-shard (m, s, U): m = min cost, s = sum exp(-(c-m)/lam), U[t,j] = sum_k e_k v[k,t,j]."""
    m = costs.min()
    e = torch.exp(-(costs - m) / lam)
    return m, e.sum(), torch.sum(e.view(-1, 1, 1) * controls, dim=0)


def merge_partials(parts: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], lam: float
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Log-sum-exp merge: equals softmax(-c/lam) over the union of the shards (mppi.py:193-199)."""
    m_all = torch.stack([p_[0] for p_ in parts]).min()
    s_tot = torch.zeros((), dtype=parts[0][1].dtype)
    u_tot = torch.zeros_like(parts[0][2])
    for m, s, u in parts:
        a = torch.exp(-(m - m_all) / lam)
        s_tot = s_tot + a * s
        u_tot = u_tot + a * u
    return m_all, s_tot, u_tot / s_tot


# --------------------------------------------------------------------------- multi-iteration driver (bench/reference arm)
class OracleSolver:
    """Stateful wrapper with the reference's call pattern (``forward`` keeps ``u_prev``, no time shift).

    Used as the CPU baseline ("port") in bench.py: it performs the same sequence of
    ATen calls per iteration as the reference loop (T transits, T stage costs with their
    own lookups, 2T small matmuls, the batch-1 optimal rollout), so its wall time on the
    host cores stands in for the reference's.
    """

    def __init__(self, p: Problem, horizon: int, num_samples: int, sigmas: Sequence[float], lam: float,
                 seed: int = 42, dtype: torch.dtype = torch.float32) -> None:
        self.p, self.T, self.K, self.lam, self.dtype = p, horizon, num_samples, float(lam), dtype
        self.sigmas = torch.tensor(list(sigmas), dtype=torch.float32)
        self.gen = torch.Generator().manual_seed(seed)
        self.u_prev = torch.zeros(horizon, 2, dtype=dtype)
        self.last: Optional[Dict[str, torch.Tensor]] = None

    def draw_noise(self) -> torch.Tensor:
        """mppi.py:149-151: eps ~ N(0,I) [K,T,2], scaled by L = diag(sigma)."""
        eps = torch.randn(self.K, self.T, 2, generator=self.gen, dtype=torch.float32)
        return eps * self.sigmas

    def forward(self, state: torch.Tensor, noise: Optional[torch.Tensor] = None):
        if noise is None:
            noise = self.draw_noise()
        out = mppi_iteration(self.p, state, self.u_prev, noise, self.sigmas, self.lam, self.dtype)
        self.u_prev = out["u_opt"]  # mppi.py:217 (no shift)
        out["noise"] = noise
        self.last = out
        return out["u_opt"], out["opt_rec"]

    def get_top_samples(self, n: int):
        assert self.last is not None
        return top_samples(self.last["rec"], self.last["weights"], n)


def wrap_angle(theta: float) -> float:
    """Scalar helper: python-style modulo wrap into [-pi, pi)."""
    return (theta + math.pi) % (2 * math.pi) - math.pi
