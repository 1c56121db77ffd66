"""The oracles of the rows either side of the MPPI iteration (oracle/env_oracle.py) and of the stochastic-slip
iteration (oracle/mppi_oracle.py with xi) pinned against outputs of the reference classes themselves.

Fixtures: tests/golden/{env,risk,dwa,stoch}_*.npz, produced by tests/golden/make_golden_ext.py from the unmodified
reference on CPU with an instrumented Normal sampler (records the standard normals it draws).
"""

import numpy as np
import pytest
import torch

from oracle import env_oracle as eo
from oracle import mppi_oracle as orc
from tests.helpers import ext_problem, t2n


def test_env_step_episode_is_bit_exact(golden_cases):
    c = golden_cases["env_g48"]
    p = ext_problem(c, stochastic=True)
    state = torch.from_numpy(c["start_state"]).clone()
    goal = torch.from_numpy(c["goal"]).view(1, 2)
    for t in range(c["actions"].shape[0]):
        nxt, rew, term = eo.env_step(p, state.view(1, 3), torch.from_numpy(c["actions"][t]).view(1, 2), goal,
                                     float(c["delta_t"]), float(c["goal_threshold"]),
                                     torch.from_numpy(c["xi_steps"][t:t + 1]))
        np.testing.assert_array_equal(t2n(nxt[0]), c["states"][t])
        np.testing.assert_array_equal(t2n(rew[0]), c["rewards"][t])
        assert bool(term[0]) == bool(c["terminated"][t])
        state = nxt[0]
    nxt, rew, term = eo.env_step(p, torch.from_numpy(c["near_goal_state"]).view(1, 3),
                                 torch.from_numpy(c["near_goal_action"]).view(1, 2), goal, float(c["delta_t"]),
                                 float(c["goal_threshold"]), torch.from_numpy(c["near_goal_xi"]).view(1))
    np.testing.assert_array_equal(t2n(nxt[0]), c["near_goal_next"])
    assert bool(term[0]) and bool(c["near_goal_term"])


def test_collision_check_is_bit_exact(golden_cases):
    c = golden_cases["env_g48"]
    p = ext_problem(c, stochastic=True)
    got = eo.collision_check(p, torch.from_numpy(c["coll_points"]), float(c["coll_threshold"]),
                             torch.from_numpy(c["coll_xi"]))
    np.testing.assert_array_equal(t2n(got), c["coll_result"])
    assert 0 < int(got.sum()) < got.numel()


@pytest.mark.parametrize("name", ["risk_g12_s1000_q90", "risk_g20_s37_q75"])
def test_risk_map_is_bit_exact_and_close_to_closed_form(golden_cases, name):
    c = golden_cases[name]
    mean, std, q = torch.from_numpy(c["mean"]), torch.from_numpy(c["std"]), float(c["confidence"])
    for metric in ("var", "cvar"):
        got = eo.risk_map(mean, std, metric, q, torch.from_numpy(c[f"samples_{metric}"]))
        np.testing.assert_array_equal(t2n(got), c[f"risk_{metric}"])
        # Monte-Carlo error of the reference's estimator against the closed form it converges to
        closed = eo.risk_map_closed_form(mean, std, metric, q)
        s_n = int(c["num_samples"])
        assert float((got.double() - closed).abs().max()) < 6.0 * float(std.max()) / np.sqrt(s_n) / (1 - q) ** 0.5
    np.testing.assert_array_equal(t2n(eo.risk_map(mean, std, "expected_value")), c["risk_expected_value"])


def test_dwa_calls_are_bit_exact(golden_cases):
    c = golden_cases["dwa_g64"]
    p = ext_problem(c, stochastic=False)
    a_lim = torch.from_numpy(c["a_lim"])
    nv, nw, T, dt = int(c["num_lin_vel"]), int(c["num_ang_vel"]), int(c["horizon"]), float(c["delta_t"])
    path = torch.from_numpy(c["path"])
    seen_path = False
    for i in range(int(c["n_calls"])):
        state = torch.from_numpy(c[f"state_{i}"])
        actions = eo.dwa_generate_actions(torch.from_numpy(c[f"prev_action_{i}"]), p.u_min, p.u_max, a_lim, dt, nv, nw)
        np.testing.assert_array_equal(t2n(actions), c[f"actions_{i}"])
        has_path = bool(c[f"has_path_{i}"])
        out = eo.dwa_forward(p, state, actions, T, path if has_path else None, float(c["lookahead"]))
        if has_path:
            seen_path = True
            np.testing.assert_array_equal(t2n(out["sub_goal"]), c[f"sub_goal_{i}"])
        np.testing.assert_array_equal(t2n(out["rec"]), c[f"rec_{i}"])
        np.testing.assert_array_equal(t2n(out["weights"]), c[f"weights_{i}"])
        np.testing.assert_array_equal(t2n(out["opt_action"]), c[f"opt_action_{i}"])
        np.testing.assert_array_equal(t2n(out["opt_states"]), c[f"opt_states_{i}"])
        ts, tw = eo.dwa_top_samples(out["rec"], out["weights"])
        np.testing.assert_array_equal(t2n(tw), c[f"top_weights_{i}"])
        np.testing.assert_array_equal(t2n(ts), c[f"top_states_{i}"])
    assert seen_path


@pytest.mark.parametrize("name", ["stoch_g64_k384_t20", "stoch_g50_k131_t7"])
def test_stochastic_iteration_is_bit_exact(golden_cases, name):
    c = golden_cases[name]
    p = ext_problem(c, stochastic=True)
    for i in range(int(c["n_calls"])):
        out = orc.mppi_iteration(p, torch.from_numpy(c[f"state_{i}"]), torch.from_numpy(c[f"u_prev_{i}"]),
                                 torch.from_numpy(c[f"noise_{i}"]), torch.from_numpy(c["sigmas"]), float(c["lam"]),
                                 xi=torch.from_numpy(c[f"xi_{i}"]), xi_opt=torch.from_numpy(c[f"xi_opt_{i}"]))
        np.testing.assert_array_equal(t2n(out["rec"]), c[f"rec_{i}"])
        np.testing.assert_array_equal(t2n(out["weights"]), c[f"weights_{i}"])
        np.testing.assert_array_equal(t2n(out["u_opt"]), c[f"u_opt_{i}"])
        np.testing.assert_array_equal(t2n(out["opt_rec"]), c[f"opt_rec_{i}"])


def test_stochastic_lookup_shares_the_cell_between_stage_cost_and_next_transit(golden_cases):
    """cell(raw recorded state t) == cell(clamped state fed to transit t+1): the engine fetches (mean, std) once per
    step and applies two different draws to it."""
    c = golden_cases["stoch_g50_k131_t7"]
    p = ext_problem(c, stochastic=True)
    rec = torch.from_numpy(c["rec_1"])
    T = int(c["T"])
    for t in range(T):
        raw = rec[:, t, :2]
        clamped = torch.stack([raw[:, 0].clamp(p.x_min, p.x_max), raw[:, 1].clamp(p.y_min, p.y_max)], dim=1)
        a, b = orc.cell_indices(p, raw), orc.cell_indices(p, clamped)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
