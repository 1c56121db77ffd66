"""CPU oracle for the rows either side of the MPPI iteration -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

Restated here in plain PyTorch-on-CPU ops, in the reference's evaluation order (citations relative to the reference
repository root):

* ``PlanetaryEnv.collision_check``             src/simulator/planetary_env.py:221-232
* ``PlanetaryEnv.step``                        src/simulator/planetary_env.py:189-219
* ``TraversabilityModel._infer_risk_map``      src/simulator/problem_formulation/traversability_model.py:28-51
* ``DWA._generate_actions / _simulate_state_sequences / _compute_costs / _select_sub_goal / forward /
  get_top_samples``                            src/planners/local_planners/dwa.py:116-299

Random draws are always INJECTED (standard normals ``xi`` or ready-made samples), never drawn here: the reference
uses ``Normal.sample()`` = ATen ``normal_(0,1).mul_(std).add_(mean)``, reproduced by ``mppi_oracle.traversability``.

Parity pinning: ``tests/golden/make_golden_ext.py`` runs the unmodified reference classes (with an instrumented
sampler that records the normals) and commits inputs/outputs as ``tests/golden/{env,risk,dwa,stoch}_*.npz``;
``tests/test_oracle_ext_golden.py`` checks this module against them bit for bit.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import mppi_oracle as orc


# --------------------------------------------------------------------------------------------- N1 / N3: environment
def collision_check(p: orc.Problem, states: torch.Tensor, threshold: float, xi: Optional[torch.Tensor] = None
                    ) -> torch.Tensor:
    """planetary_env.py:221-232: ``get_traversability(states) <= stuck_threshold`` for states [B,P,3].
    ``xi`` [B,P]: the lookup normals (observation mode, the environment's own dynamics); None = inference mode."""
    trav = orc.traversability(p, states[..., :2], xi)
    return trav <= threshold


def env_step(p: orc.Problem, state: torch.Tensor, action: torch.Tensor, goal: torch.Tensor, delta_t: float,
             goal_threshold: float, xi: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """planetary_env.py:203-217 for E environments at once: states [E,3], actions [E,2], goals [E,2], xi [E].
    Returns (next_state [E,3], reward [E] = the traversability drawn for the step, terminated [E] bool)."""
    _, nxt, trav = orc.unicycle_step(p, state, action, xi=xi, dt=delta_t, return_trav=True)
    terminated = torch.stack([torch.norm(nxt[e, :2] - goal[e]) < goal_threshold for e in range(nxt.shape[0])])
    return nxt, trav, terminated


# --------------------------------------------------------------------------------------------- N2: risk map
def risk_map(mean: torch.Tensor, std: torch.Tensor, metric: str, confidence: Optional[float] = None,
             samples: Optional[torch.Tensor] = None) -> torch.Tensor:
    """traversability_model.py:36-51 with the Monte-Carlo samples [S,G,G] injected (= ``distributions.sample((S,))``)."""
    if metric == "expected_value":
        return mean
    assert samples is not None and confidence is not None
    var = torch.quantile(samples, confidence, dim=0)
    if metric == "var":
        return var
    mask = samples > var.unsqueeze(0)
    tail = torch.where(mask, samples, torch.tensor(torch.nan))
    return torch.nanmean(tail, dim=0)


def risk_map_closed_form(mean: torch.Tensor, std: torch.Tensor, metric: str, confidence: Optional[float] = None
                         ) -> torch.Tensor:
    """Limit of the estimator above for a Normal slip model (float64): mean + std * z_q (VaR),
    mean + std * phi(z_q) / (1 - q) (CVaR)."""
    mean, std = mean.double(), std.double()
    if metric == "expected_value":
        return mean
    n = torch.distributions.Normal(torch.tensor(0.0, dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64))
    z = n.icdf(torch.tensor(confidence, dtype=torch.float64))
    if metric == "var":
        return mean + std * z
    return mean + std * torch.exp(n.log_prob(z)) / (1.0 - confidence)


# --------------------------------------------------------------------------------------------- N4: DWA
def dwa_generate_actions(prev_action: torch.Tensor, u_min: Sequence[float], u_max: Sequence[float],
                         a_lim: torch.Tensor, delta_t: float, num_lin_vel: int, num_ang_vel: int) -> torch.Tensor:
    """dwa.py:160-184."""
    lo = torch.tensor(list(u_min), dtype=torch.float32)
    hi = torch.tensor(list(u_max), dtype=torch.float32)
    v_min = torch.max(lo[0], prev_action[0] - a_lim[0] * delta_t)
    v_max = torch.min(hi[0], prev_action[0] + a_lim[0] * delta_t)
    w_min = torch.max(lo[1], prev_action[1] - a_lim[1] * delta_t)
    w_max = torch.min(hi[1], prev_action[1] + a_lim[1] * delta_t)
    vs = torch.linspace(v_min, v_max, num_lin_vel, dtype=torch.float32)
    ws = torch.linspace(w_min, w_max, num_ang_vel, dtype=torch.float32)
    return torch.cartesian_prod(vs, ws)


def dwa_select_sub_goal(path: torch.Tensor, state: torch.Tensor, lookahead: float) -> torch.Tensor:
    """dwa.py:270-285."""
    deltas = path - state[:2]
    distances = torch.norm(deltas, dim=1)
    angles = torch.atan2(deltas[:, 1], deltas[:, 0]) - state[2]
    valid = (angles.abs() < torch.pi / 2) & (distances > lookahead)
    if valid.any():
        d_min = distances[valid].min()
        return path[torch.where(distances == d_min)[0][0]]
    return path[-1]


def dwa_forward(p: orc.Problem, state: torch.Tensor, actions: torch.Tensor, horizon: int,
                path: Optional[torch.Tensor] = None, lookahead: float = 1.0) -> Dict[str, torch.Tensor]:
    """dwa.py:136-149 (+ :186-258): constant-action rollouts (same in-place quirk of ``transit`` as MPPI: slot t < T of
    ``rec`` holds the raw successor of step t), sequential cost accumulation, argmin, softmax(-cost) weights.

    With a reference ``path`` the stage cost follows a sub-goal, selected -- as the reference does, dwa.py:225-228 --
    from ``state_seq_batch[0, 0, :]`` AFTER the simulation: not the robot state but the raw (unclamped, unwrapped)
    successor of the first action's first step that the in-place update left there."""
    k = actions.shape[0]
    controls = actions.unsqueeze(1).repeat(1, horizon, 1)
    rec = orc.rollout(p, state, controls)
    cost = torch.zeros(k, dtype=torch.float32)
    sub_goal = None if path is None else dwa_select_sub_goal(path, rec[0, 0, :], lookahead)
    goal = None if sub_goal is None else sub_goal.tolist()
    for t in range(horizon):
        cost += orc.stage_cost(p, rec[:, t, :], goal=goal)
    cost += orc.stage_cost(p, rec[:, -1, :])  # terminal cost always uses the final goal (dwa.py:233, objectives.py:65)
    idx = torch.argmin(cost, dim=0)
    return {"opt_action": actions[idx].unsqueeze(0), "opt_states": rec[idx].unsqueeze(0), "rec": rec,
            "weights": torch.softmax(-cost, dim=0), "costs": cost, "index": idx, "sub_goal": sub_goal}


def dwa_top_samples(rec: torch.Tensor, weights: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """dwa.py:287-299."""
    order = torch.argsort(weights, descending=True)
    return rec[order], weights[order]
