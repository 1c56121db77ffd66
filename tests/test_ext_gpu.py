"""GPU parity tests of the widened path (through the C ABI): stochastic-slip MPPI (BASELINE config 4), batched
environments (config 3), and the rows either side of the iteration (collision check / environment
step, risk-map inference, DWA) -- against golden outputs of the reference classes and the CPU oracles.
All tests here need a B200 (`-m gpu`)."""

import numpy as np
import pytest
import torch

from oracle import env_oracle as eo
from oracle import mppi_oracle as orc
from tests.gpu_common import TOL_REC, assert_iteration_close, engine_outputs, make_solver, oracle_outputs
from tests.helpers import ext_problem

pytestmark = pytest.mark.gpu


def _stoch_solver(mean, std, res, goal, thr, K, T, sigmas, lam, seed=42):
    from benchnav_b200 import MPPI
    from benchnav_b200.problem import GoalObjectives, GridSpec, SlipDistribution, UnicycleProblem

    d = SlipDistribution(mean, std)
    grid = GridSpec(int(mean.shape[0]), res, distributions={"predictions": d, "latent_models": d})
    dyn = UnicycleProblem(grid, mean)
    obj = GoalObjectives(dyn, torch.as_tensor(goal), thr)
    return MPPI(T, K, 3, 2, dyn, obj, torch.as_tensor(sigmas, dtype=torch.float32), lam, device=torch.device("cuda"),
                seed=seed, stochastic_slip=True)


# ------------------------------------------------------------------------------------------------ config 4
@pytest.mark.parametrize("name", ["stoch_g64_k384_t20", "stoch_g50_k131_t7"])
def test_stochastic_golden_calls(golden_cases, name):
    """Reference MPPI.forward with observation-mode lookups on the predicted slip: its own noise and lookup normals
    injected call by call."""
    c = golden_cases[name]
    solver = _stoch_solver(torch.from_numpy(c["mean"]), torch.from_numpy(c["std"]), float(c["resolution"]),
                           c["goal"].tolist(), float(c["thr"]), int(c["K"]), int(c["T"]), c["sigmas"], float(c["lam"]))
    p = ext_problem(c, stochastic=True)
    for i in range(int(c["n_calls"])):
        solver._previous_action_seq.copy_(torch.from_numpy(c[f"u_prev_{i}"]))
        u, opt = solver.forward(torch.from_numpy(c[f"state_{i}"]), noise=torch.from_numpy(c[f"noise_{i}"]),
                                xi=torch.from_numpy(c[f"xi_{i}"]), xi_opt=torch.from_numpy(c[f"xi_opt_{i}"]))
        eng = engine_outputs(solver, u, opt)
        ref = {"u_opt": c[f"u_opt_{i}"], "opt_rec": c[f"opt_rec_{i}"], "rec": c[f"rec_{i}"], "weights": c[f"weights_{i}"]}
        ref["costs"] = orc.mppi_iteration(p, torch.from_numpy(c[f"state_{i}"]), torch.from_numpy(c[f"u_prev_{i}"]),
                                          torch.from_numpy(c[f"noise_{i}"]), torch.from_numpy(c["sigmas"]), float(c["lam"]),
                                          xi=torch.from_numpy(c[f"xi_{i}"]), xi_opt=torch.from_numpy(c[f"xi_opt_{i}"]))["costs"].numpy()
        assert_iteration_close(eng, ref, f"{name}[{i}]")


def _synthetic_slip(g, seed=0):
    from benchnav_b200.synthetic import make_terrain

    terr = make_terrain(g, 0.5, seed)
    mean = terr["slip_mean"].clone()
    c = int(8.0 / 0.5)
    mean[c - 1:c + 2, c - 1:c + 2] = torch.clamp(mean[c - 1:c + 2, c - 1:c + 2], max=0.2)
    return mean, terr["slip_std"].clone()


@pytest.mark.parametrize("g,K,T", [(128, 2048, 30), (256, 32768, 50), (64, 999, 9)])
def test_stochastic_in_engine_draws_replayed_through_oracle(g, K, T):
    """Full contract with in-kernel Philox noise AND lookup normals (config 4 at full size: 256x256, K=32768, T=50):
    the drawn control noise is read back from the engine, the lookup normals from the stand-alone draw kernel (same
    Philox calls), and the iteration is replayed through the oracle."""
    mean, std = _synthetic_slip(g)
    lim = g * 0.5
    goal, thr, sig, lam = [0.375 * lim, 0.375 * lim], 0.3, [0.5, 0.5], 0.5
    solver = _stoch_solver(mean, std, 0.5, goal, thr, K, T, sig, lam, seed=7)
    p = orc.make_problem(mean, 0.5, goal, thr)
    p.slip_std = std
    state = torch.tensor([min(8.0, 0.25 * lim), min(8.0, 0.25 * lim), 0.785398], dtype=torch.float32)
    u_prev = torch.zeros(T, 2)
    for it in range(2):
        u, opt = solver.forward(state)
        eng = engine_outputs(solver, u, opt)
        noise = solver._action_noises.cpu()
        xi, xi_opt = (t.cpu() for t in solver.draw_lookup_normals(it))
        assert abs(float(xi.mean())) < 0.02 and abs(float(xi.std()) - 1.0) < 0.02
        ref = orc.mppi_iteration(p, state, u_prev, noise, torch.tensor(sig), lam, xi=xi, xi_opt=xi_opt)
        assert_iteration_close(eng, {k: v.numpy() for k, v in ref.items()}, f"stoch G{g} K{K} T{T} it{it}")
        u_prev = u.cpu()
    # a second solver with the same seed reproduces the stream bit for bit
    again = _stoch_solver(mean, std, 0.5, goal, thr, K, T, sig, lam, seed=7)
    u2, _ = again.forward(state)
    again.forward(state)
    torch.cuda.synchronize()
    assert torch.equal(again._state_seq_batch, solver._state_seq_batch)


def test_stochastic_differs_from_deterministic_and_rejects_partial_injection():
    mean, std = _synthetic_slip(64)
    solver = _stoch_solver(mean, std, 0.5, [24.0, 24.0], 0.3, 512, 20, [0.5, 0.5], 0.5)
    det = make_solver(mean, 0.5, [24.0, 24.0], 0.3, 512, 20, [0.5, 0.5], 0.5)
    state = torch.tensor([8.0, 8.0, 0.7])
    noise = torch.randn(512, 20, 2, generator=torch.Generator().manual_seed(0)) * 0.5
    xi = torch.randn(512, 41, generator=torch.Generator().manual_seed(1))
    solver.forward(state, noise=noise, xi=xi, xi_opt=torch.zeros(20))
    det.forward(state, noise=noise)
    torch.cuda.synchronize()
    assert not torch.equal(solver._state_seq_batch, det._state_seq_batch)
    with pytest.raises(ValueError):
        solver.forward(state, noise=noise)
    # zero std: the stochastic lookup degenerates to the deterministic one on the mean map
    flat = _stoch_solver(mean, torch.zeros_like(std), 0.5, [24.0, 24.0], 0.3, 512, 20, [0.5, 0.5], 0.5)
    flat.forward(state, noise=noise, xi=xi, xi_opt=torch.ones(20))
    torch.cuda.synchronize()
    assert torch.equal(flat._state_seq_batch, det._state_seq_batch)
    np.testing.assert_allclose(flat._weights.cpu().numpy(), det._weights.cpu().numpy(), rtol=0, atol=1e-7)


# ------------------------------------------------------------------------------------------------ config 3
def _batch_problems(E, g, seed0=0):
    from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
    from benchnav_b200.synthetic import benchmark_problem

    dyns, objs, risks, goals, states = [], [], [], [], []
    for e in range(E):
        risk, start, goal, thr = benchmark_problem(g, 0.5, seed=seed0 + e)
        goal = goal + torch.tensor([0.7 * e, -0.4 * e])
        start = start + torch.tensor([0.3 * e, 0.2 * e, 0.1 * e])
        dyn = UnicycleProblem(GridSpec(g, 0.5), risk)
        dyns.append(dyn)
        objs.append(GoalObjectives(dyn, goal, thr))
        risks.append(risk)
        goals.append(goal)
        states.append(start)
    return dyns, objs, risks, goals, torch.stack(states), thr


@pytest.mark.parametrize("E,K,T", [(5, 777, 30), (8, 4096, 30), (3, 33, 7)])
def test_batched_matches_oracle_and_single_solvers(E, K, T):
    """Config 3's meaning: E independent reference solvers.  Injected noise; every environment is compared
    with the oracle and must be BIT-EQUAL to the single-solver engine run on the same inputs."""
    from benchnav_b200 import BatchedMPPI

    g, sig, lam = 64, [0.5, 0.5], 0.5
    dyns, objs, risks, goals, states, thr = _batch_problems(E, g)
    solver = BatchedMPPI(T, K, dyns, objs, torch.tensor(sig), lam, seed=3)
    gen = torch.Generator().manual_seed(11)
    u_prev = torch.zeros(E, T, 2)
    for it in range(2):
        noise = torch.randn(E, K, T, 2, generator=gen) * torch.tensor(sig)
        u, opt = solver.forward(states, noise=noise)
        torch.cuda.synchronize()
        for e in range(E):
            p = orc.make_problem(risks[e], 0.5, goals[e].tolist(), thr)
            ref = oracle_outputs(p, states[e], u_prev[e], noise[e], sig, lam)
            eng = {"u_opt": u[e].cpu().numpy(), "opt_rec": opt[e].cpu().numpy(), "weights": solver._weights[e].cpu().numpy(),
                   "costs": solver._costs[e].cpu().numpy(), "rec": solver._state_seq_batch[e].cpu().numpy()}
            assert_iteration_close(eng, ref, f"batch env {e} it {it}")
            if it == 0 and e in (0, E - 1):
                single = make_solver(risks[e], 0.5, goals[e].tolist(), thr, K, T, sig, lam)
                us, _ = single.forward(states[e], noise=noise[e])
                torch.cuda.synchronize()
                assert torch.equal(single._state_seq_batch, solver._state_seq_batch[e])
                assert torch.equal(single.costs, solver._costs[e])
                np.testing.assert_allclose(us.cpu().numpy(), eng["u_opt"], rtol=0, atol=2e-6)
        u_prev = u.cpu()
        np.testing.assert_array_equal(solver._previous_action_seq.cpu().numpy(), u_prev.numpy())


def test_batched_config3_full_size_in_engine_noise():
    """64 environments x K=4096 x T=30 in one launch (more CTAs than the device holds: the non-cooperative schedule),
    in-engine Philox noise; a sample of environments replayed through the oracle; per-environment streams differ."""
    from benchnav_b200 import BatchedMPPI

    E, K, T, g, sig, lam = 64, 4096, 30, 64, [0.5, 0.5], 0.5
    dyns, objs, risks, goals, states, thr = _batch_problems(E, g)
    solver = BatchedMPPI(T, K, dyns, objs, torch.tensor(sig), lam, seed=5)
    u, opt = solver.forward(states)
    torch.cuda.synchronize()
    w = solver._weights.double().sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(w, 1.0, atol=1e-5)
    assert not torch.equal(solver._action_noises[0], solver._action_noises[1])
    for e in (0, 17, 63):
        p = orc.make_problem(risks[e], 0.5, goals[e].tolist(), thr)
        ref = oracle_outputs(p, states[e], torch.zeros(T, 2), solver._action_noises[e].cpu(), sig, lam)
        eng = {"u_opt": u[e].cpu().numpy(), "opt_rec": opt[e].cpu().numpy(), "weights": solver._weights[e].cpu().numpy(),
               "costs": solver._costs[e].cpu().numpy(), "rec": solver._state_seq_batch[e].cpu().numpy()}
        assert_iteration_close(eng, ref, f"config3 env {e}")
    # top-n per environment (mppi.py:221-240)
    ts, tw = solver.get_top_samples(50)
    torch.cuda.synchronize()
    for e in (0, 40):
        want_s, want_w = orc.top_samples(solver._state_seq_batch[e].cpu(), solver._weights[e].cpu(), 50)
        np.testing.assert_array_equal(tw[e].cpu().numpy(), want_w.numpy())
        if len(np.unique(want_w.numpy())) == 50:  # no ties: the order is unambiguous
            np.testing.assert_array_equal(ts[e].cpu().numpy(), want_s.numpy())


def test_batched_forward_host_equals_forward():
    """BatchedMPPI.forward_host (host states in, host results out, one staged copy each way inside the library) against
    forward() of a twin solver: bit-equal over chained iterations, caller-owned output buffers honoured."""
    from benchnav_b200 import BatchedMPPI

    E, K, T, g, sig, lam = 6, 1000, 20, 64, [0.5, 0.5], 0.5
    dyns, objs, risks, goals, states, thr = _batch_problems(E, g)
    a = BatchedMPPI(T, K, dyns, objs, torch.tensor(sig), lam, seed=7)
    b = BatchedMPPI(T, K, dyns, objs, torch.tensor(sig), lam, seed=7)
    out = (torch.empty(E, T, 2), torch.empty(E, 1, T + 1, 3))
    for it in range(3):
        u1, o1 = a.forward(states)
        u2, o2 = b.forward_host(states, out=out if it else None)
        assert not u2.is_cuda and o2.shape == (E, 1, T + 1, 3)
        assert torch.equal(u1.cpu(), u2) and torch.equal(o1.cpu(), o2), f"iteration {it}"
        states = o2[:, 0, 1, :].clone()  # every environment moves on
        states[:, 2] = ((states[:, 2] + np.pi) % (2 * np.pi)) - np.pi
    with pytest.raises(ValueError):
        b.forward_host(states, out=(torch.empty(E, T, 3), out[1]))


# ------------------------------------------------------------------------------------------------ N1 / N3
class _GM:
    def __init__(self, c, mean, std):
        from benchnav_b200.problem import SlipDistribution

        g = mean.shape[0]
        self.grid_size, self.resolution = g, float(c["resolution"])
        lim = c["limits"].tolist()
        self.x_limits, self.y_limits = (lim[0], lim[1]), (lim[2], lim[3])
        d = SlipDistribution(mean, std)
        self.distributions = {"latent_models": d, "predictions": d}


def test_env_step_golden_episode(golden_cases):
    """PlanetaryEnv.step (planetary_env.py:189-219): the reference's 40-step episode with its own draws injected, run
    as environment 0 and 2 of a batch of three (environment 1 gets other actions)."""
    from benchnav_b200 import BatchedPlanetaryEnv

    c = golden_cases["env_g48"]
    mean, std = torch.from_numpy(c["mean"]), torch.from_numpy(c["std"])
    gms = [_GM(c, mean, std), _GM(c, mean.flip(0).contiguous(), std), _GM(c, mean, std)]
    start = torch.from_numpy(c["start_state"][:2]).repeat(3, 1)
    goal = torch.from_numpy(c["goal"]).repeat(3, 1)
    env = BatchedPlanetaryEnv(gms, start, goal, delta_t=float(c["delta_t"]), goal_threshold=float(c["goal_threshold"]))
    s0 = env.reset()
    np.testing.assert_allclose(s0[0].cpu().numpy(), c["start_state"], atol=1e-6)
    for t in range(c["actions"].shape[0]):
        a = torch.from_numpy(c["actions"][t])
        acts = torch.stack([a, torch.tensor([0.9, 0.4]), a])
        xi = torch.tensor([c["xi_steps"][t], 0.3, c["xi_steps"][t]])
        st, rew, term, trunc = env.step(acts, xi=xi)
        st, rew, term = st.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy()
        for e in (0, 2):
            np.testing.assert_allclose(st[e], c["states"][t], rtol=0, atol=2e-5)
            np.testing.assert_allclose(rew[e], c["rewards"][t], rtol=0, atol=1e-6)
            assert bool(term[e]) == bool(c["terminated"][t])
        assert not trunc
    env._robot_state.copy_(torch.from_numpy(c["near_goal_state"]).repeat(3, 1))
    st, rew, term, _ = env.step(torch.from_numpy(c["near_goal_action"]).repeat(3, 1),
                                xi=torch.from_numpy(c["near_goal_xi"]).repeat(3))
    assert bool(term[0]) and bool(c["near_goal_term"])
    np.testing.assert_allclose(st[0].cpu().numpy(), c["near_goal_next"], rtol=0, atol=2e-5)


def test_collision_check_golden_and_modes(golden_cases):
    """PlanetaryEnv.collision_check (planetary_env.py:221-232), observation mode with the reference's draws; the
    inference-mode lookup on a risk map; the engine's own draws are standard normal."""
    from benchnav_b200 import env as benv

    c = golden_cases["env_g48"]
    dev = torch.device("cuda")
    mean, std = torch.from_numpy(c["mean"]).to(dev), torch.from_numpy(c["std"]).to(dev)
    gm = _GM(c, mean, std)
    pts = torch.from_numpy(c["coll_points"]).to(dev)
    got = benv.collision_check(gm, pts, float(c["coll_threshold"]), mean=mean, std=std, xi=torch.from_numpy(c["coll_xi"]).to(dev))
    np.testing.assert_array_equal(got.cpu().numpy(), c["coll_result"])
    p = ext_problem(c, stochastic=True)
    trav = benv.traversability(gm, pts, mean=mean)
    np.testing.assert_array_equal(trav.cpu().numpy(), orc.traversability(p, pts.cpu()[..., :2]).numpy())
    # Philox draws: recover xi = (1 - trav - mean) / std where the clamp is inactive; must look standard normal
    big = torch.rand(200000, 3, device=dev) * 20.0 + 2.0
    flat_mean, flat_std = torch.full_like(mean, 0.5), torch.full_like(std, 0.05)
    t1 = benv.traversability(gm, big, mean=flat_mean, std=flat_std, seed=3, counter=1)
    t2 = benv.traversability(gm, big, mean=flat_mean, std=flat_std, seed=3, counter=2)
    z = ((1.0 - t1) - 0.5) / 0.05
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01 and not torch.equal(t1, t2)
    assert torch.equal(t1, benv.traversability(gm, big, mean=flat_mean, std=flat_std, seed=3, counter=1))


# ------------------------------------------------------------------------------------------------ N2
@pytest.mark.parametrize("name", ["risk_g12_s1000_q90", "risk_g20_s37_q75"])
def test_risk_map_golden_and_closed_form(golden_cases, name):
    from benchnav_b200 import infer_risk_map

    c = golden_cases[name]
    dev = torch.device("cuda")
    mean, std, q = torch.from_numpy(c["mean"]).to(dev), torch.from_numpy(c["std"]).to(dev), float(c["confidence"])
    for metric in ("var", "cvar"):
        got = infer_risk_map(mean, std, metric, q, method="monte_carlo", samples=torch.from_numpy(c[f"samples_{metric}"]).to(dev))
        np.testing.assert_allclose(got.cpu().numpy(), c[f"risk_{metric}"], rtol=0, atol=2e-6)
        closed = infer_risk_map(mean, std, metric, q)
        want = eo.risk_map_closed_form(mean.cpu(), std.cpu(), metric, q).numpy()
        np.testing.assert_allclose(closed.cpu().numpy(), want, rtol=0, atol=1e-6)
    np.testing.assert_array_equal(infer_risk_map(mean, std, "expected_value").cpu().numpy(), c["risk_expected_value"])


def test_risk_map_in_engine_draws_at_benchmark_size():
    """256x256 map, 1000 draws per cell (the reference's default): the engine's own samples replayed through the
    reference estimator; the estimate sits within Monte-Carlo error of the closed form."""
    from benchnav_b200 import infer_risk_map

    mean, std = (t.cuda() for t in _synthetic_slip(256))
    for metric, q in (("var", 0.9), ("cvar", 0.95)):
        got, samples = infer_risk_map(mean, std, metric, q, method="monte_carlo", num_samples=1000, seed=5, return_samples=True)
        z = (samples - mean) / std.clamp_min(1e-6)
        assert abs(float(z.mean())) < 2e-3 and abs(float(z.std()) - 1.0) < 2e-3
        want = eo.risk_map(mean.cpu(), std.cpu(), metric, q, samples.cpu())
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=0, atol=2e-6)
        closed = infer_risk_map(mean, std, metric, q)
        assert float((got - closed).abs().max()) < 8.0 * float(std.max()) / np.sqrt(1000 * (1 - q))
    with pytest.raises(AssertionError):
        infer_risk_map(mean, std, "var", None)


# ------------------------------------------------------------------------------------------------ N4
def test_dwa_golden_calls(golden_cases):
    """DWA.forward / get_top_samples (dwa.py:116-299) call by call, with and without a reference path."""
    from benchnav_b200 import DWA
    from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

    c = golden_cases["dwa_g64"]
    risk = torch.from_numpy(c["risk"])
    dyn = UnicycleProblem(GridSpec(risk.shape[0], float(c["resolution"])), risk)
    obj = GoalObjectives(dyn, torch.from_numpy(c["goal"]), float(c["thr"]))
    T = int(c["horizon"])
    solver = DWA(T, 3, 2, dyn, obj, torch.from_numpy(c["a_lim"]), float(c["delta_t"]), lookahead_distance=float(c["lookahead"]),
                 num_lin_vel=int(c["num_lin_vel"]), num_ang_vel=int(c["num_ang_vel"]), device=torch.device("cuda"))
    p = ext_problem(c, stochastic=False)
    for i in range(int(c["n_calls"])):
        if bool(c[f"has_path_{i}"]) and solver.reference_path is None:
            solver.update_reference_path(torch.from_numpy(c["path"]))
        if i > 0:  # same chain as the reference: the previous optimum seeds the window
            np.testing.assert_allclose(solver._previous_action_seq[0].cpu().numpy(), c[f"prev_action_{i}"], rtol=0, atol=1e-6)
        a, s = solver.forward(torch.from_numpy(c[f"state_{i}"]))
        torch.cuda.synchronize()
        np.testing.assert_allclose(solver._actions.cpu().numpy(), c[f"actions_{i}"], rtol=0, atol=1e-6)
        if bool(c[f"has_path_{i}"]):
            np.testing.assert_array_equal(solver._sub_goal.cpu().numpy(), c[f"sub_goal_{i}"])
        rec = solver._state_seq_batch.cpu().numpy()
        bad = np.abs(rec - c[f"rec_{i}"]).reshape(rec.shape[0], -1).max(axis=1) > TOL_REC
        # K = 100 constant-action rollouts: at most ONE may differ -- a rollout whose position lands within an ulp of
        # a cell border can take the neighbouring cell's traversability (the engine's sin/cos is <= 2 ulp, not
        # bit-equal to ATen's Sleef kernel), after which its remaining states follow the other cell's value
        assert bad.sum() <= 1, f"dwa[{i}]: {bad.sum()} of {bad.size} rollouts off by > {TOL_REC}"
        np.testing.assert_allclose(solver._weights.cpu().numpy(), c[f"weights_{i}"], rtol=0, atol=2e-4)
        assert a.shape == (1, 2) and s.shape == (1, T + 1, 3)
        np.testing.assert_allclose(a.cpu().numpy(), c[f"opt_action_{i}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(s.cpu().numpy(), c[f"opt_states_{i}"], rtol=0, atol=1e-4)
        ts, tw = solver.get_top_samples()
        np.testing.assert_allclose(tw.cpu().numpy(), c[f"top_weights_{i}"], rtol=0, atol=2e-4)
        assert ts.shape == (100, T + 1, 3) and bool((tw[:-1] >= tw[1:]).all())


# ------------------------------------------------------------------------------------------------ closed loop
def test_closed_loop_planner_and_environment_follow_the_oracle():
    """Tutorial 3.3's loop (solver.forward -> env.step -> collision_check -> get_top_samples) on the device for three
    environments at once, against the same loop run with the CPU oracles on identical injected draws."""
    from benchnav_b200 import BatchedMPPI, BatchedPlanetaryEnv

    E, K, T, g, sig, lam, steps = 3, 512, 20, 64, [0.5, 0.5], 0.5, 12
    dyns, objs, risks, goals, states0, thr = _batch_problems(E, g)
    stds = [torch.full((g, g), 0.03) for _ in range(E)]

    class GM:
        def __init__(self, mean, std):
            from benchnav_b200.problem import SlipDistribution

            self.grid_size, self.resolution = g, 0.5
            self.x_limits = self.y_limits = (0.0, g * 0.5)
            self.distributions = {"latent_models": SlipDistribution(mean, std)}

    planner = BatchedMPPI(T, K, dyns, objs, torch.tensor(sig), lam, seed=9)
    env = BatchedPlanetaryEnv([GM(risks[e], stds[e]) for e in range(E)], states0[:, :2], torch.stack(goals),
                              stuck_threshold=0.1)
    env._robot_state.copy_(states0)  # same start heading as the oracle loop
    gen = torch.Generator().manual_seed(21)
    ref_state = states0.clone()
    ref_uprev = torch.zeros(E, T, 2)
    for step in range(steps):
        noise = torch.randn(E, K, T, 2, generator=gen) * torch.tensor(sig)
        xi = torch.randn(E, generator=gen)
        xi_c = torch.randn(E, T + 1, generator=gen)
        u, seq = planner.forward(env._robot_state, noise=noise)
        coll = env.collision_check(seq[:, 0], xi=xi_c)
        top_s, top_w = planner.get_top_samples(16)
        st, rew, term, trunc = env.step(u[:, 0, :], xi=xi)
        torch.cuda.synchronize()
        for e in range(E):
            p = orc.make_problem(risks[e], 0.5, goals[e].tolist(), thr)
            ref = orc.mppi_iteration(p, ref_state[e], ref_uprev[e], noise[e], torch.tensor(sig), lam)
            ps = orc.make_problem(risks[e], 0.5, goals[e].tolist(), thr)
            ps.slip_std = stds[e]
            want_coll = eo.collision_check(ps, ref["opt_rec"], 0.1, xi_c[e].view(1, -1))
            nxt, r_rew, r_term = eo.env_step(ps, ref_state[e].view(1, 3), ref["u_opt"][0].view(1, 2), goals[e].view(1, 2),
                                             0.1, 1.0, xi[e].view(1))
            # two free-running closed loops (engine state / oracle state): step s starts from states and mean sequences
            # that already differ by the accumulated deviation of steps 0..s-1, so the bounds grow linearly -- 2e-3 per
            # step on u* (the single-call tolerance) and 2e-3 * v_max * dt * 5 = 1e-3 per step on the state (the applied
            # control moves the robot by at most dt per step; the factor 5 covers the heading's lever arm)
            np.testing.assert_allclose(u[e].cpu().numpy(), ref["u_opt"].numpy(), rtol=0, atol=2e-3 * (step + 1))
            np.testing.assert_allclose(st[e].cpu().numpy(), nxt[0].numpy(), rtol=0, atol=1e-3 * (step + 1))
            np.testing.assert_allclose(rew[e].item(), r_rew[0].item(), rtol=0, atol=1e-6)
            assert bool(term[e]) == bool(r_term[0])
            assert int((coll[e].cpu() != want_coll[0]).sum()) <= 1  # a position within rounding of a cell border
            assert top_s.shape == (E, 16, T + 1, 3) and bool((top_w[e][:-1] >= top_w[e][1:]).all())
            ref_state[e] = nxt[0]
            ref_uprev[e] = ref["u_opt"]
    moved = (env._robot_state[:, :2].cpu() - states0[:, :2]).norm(dim=1)
    assert float(moved.min()) > 0.2  # the robots actually drive


def test_closed_loop_example_reaches_the_goals():
    """examples/closed_loop.py: eight environments driven to their goals by the batched planner (the functional
    outcome of Tutorial 3.3's loop)."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "closed_loop.py")
    spec = importlib.util.spec_from_file_location("closed_loop_example", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    runs = {}
    for use_graph in (True, False):  # one CUDA-graph launch per control step, or host-issued launches
        for fused in (True, False):  # environment step + collision check + the loop's books as ONE kernel, or separately
            steps, dist, _ = mod.run(envs=8, samples=2048, horizon=30, max_steps=900, verbose=False, use_graph=use_graph,
                                     fused=fused)
            assert int((steps > 0).sum()) >= 7, (use_graph, fused, steps.tolist(), dist.tolist())  # one straggler tolerated
            assert float(dist.min()) < 1.0
            runs[(use_graph, fused)] = (steps, dist)
    # the fused kernel makes the same draws as the separate calls: identical closed loops, captured or not
    ref_steps, ref_dist = runs[(False, False)]
    for key, (steps, dist) in runs.items():
        assert torch.equal(steps, ref_steps), (key, steps.tolist(), ref_steps.tolist())
        assert torch.equal(dist, ref_dist), key


# ------------------------------------------------------------------------------------------------ CUDA graphs
def test_forward_captured_in_a_cuda_graph_replays_the_same_noise_stream():
    """With the iteration counter in device memory, forward() is captured once and replayed: every replay advances
    the Philox stream exactly like an uncaptured call (bit-equal recorded states and controls)."""
    from tests.gpu_common import make_solver
    from benchnav_b200.synthetic import benchmark_problem

    risk, start, goal, thr = benchmark_problem(64, 0.5, seed=0)
    K, T = 1024, 20
    plain = make_solver(risk, 0.5, goal.tolist(), thr, K, T, [0.5, 0.5], 0.5, seed=11)
    graphed = make_solver(risk, 0.5, goal.tolist(), thr, K, T, [0.5, 0.5], 0.5, seed=11)
    state = start.cuda()
    ref = []
    for _ in range(4):
        u, opt = plain.forward(state)
        torch.cuda.synchronize()
        ref.append((u.clone(), opt.clone(), plain._state_seq_batch.clone(), plain._weights.clone()))
    graphed.graph_capturable(True)
    u0, opt0 = graphed.forward(state)  # iteration 0, uncaptured but on the device counter
    torch.cuda.synchronize()
    assert torch.equal(u0, ref[0][0]) and torch.equal(graphed._state_seq_batch, ref[0][2])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            ug, optg = graphed.forward(state)
    torch.cuda.current_stream().wait_stream(side)
    for i in range(1, 4):  # the capture itself launched nothing; replays are iterations 1, 2, 3
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(ug, ref[i][0]), f"replay {i}: controls differ"
        assert torch.equal(optg, ref[i][1])
        assert torch.equal(graphed._state_seq_batch, ref[i][2])
        assert torch.equal(graphed._weights, ref[i][3])
    graphed.graph_capturable(False)  # back to by-value counters, the count carries over
    u4, _ = graphed.forward(state)
    u4p, _ = plain.forward(state)
    torch.cuda.synchronize()
    assert torch.equal(u4, u4p)


# ------------------------------------------------------------------------------------------------ more shapes
def test_batched_and_stochastic_without_a_staged_window():
    """Fine resolution + long reach: the traversability window exceeds shared memory, lookups go to the global map
    (batched: per-environment base pointer; stochastic: interleaved (mean, std) pairs); non-power-of-two resolution."""
    from benchnav_b200 import BatchedMPPI
    from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

    g, res, K, T, sig, lam, thr = 200, 0.06, 300, 60, [0.5, 0.5], 0.5, 0.3
    gen = torch.Generator().manual_seed(5)
    E = 3
    risks = [torch.rand(g, g, generator=gen) * 0.8 for _ in range(E)]
    goals = [torch.tensor([9.0 - e, 8.0 + 0.5 * e]) for e in range(E)]
    states = torch.tensor([[3.0, 3.0, 0.3], [6.0, 5.0, -2.0], [11.9, 0.05, 1.0]])
    dyns = [UnicycleProblem(GridSpec(g, res), r) for r in risks]
    objs = [GoalObjectives(d, gl, thr) for d, gl in zip(dyns, goals)]
    solver = BatchedMPPI(T, K, dyns, objs, torch.tensor(sig), lam, seed=2)
    noise = torch.randn(E, K, T, 2, generator=gen) * torch.tensor(sig)
    u, opt = solver.forward(states, noise=noise)
    torch.cuda.synchronize()
    for e in range(E):
        p = orc.make_problem(risks[e], res, goals[e].tolist(), thr)
        ref = oracle_outputs(p, states[e], torch.zeros(T, 2), noise[e], sig, lam)
        eng = {"u_opt": u[e].cpu().numpy(), "opt_rec": opt[e].cpu().numpy(), "weights": solver._weights[e].cpu().numpy(),
               "costs": solver._costs[e].cpu().numpy(), "rec": solver._state_seq_batch[e].cpu().numpy()}
        assert_iteration_close(eng, ref, f"no-window batch env {e}")
    # stochastic, same geometry
    std = torch.rand(g, g, generator=gen) * 0.2
    st = _stoch_solver(risks[0], std, res, goals[0].tolist(), thr, K, T, sig, lam, seed=4)
    xi = torch.randn(K, 2 * T + 1, generator=gen)
    xi_opt = torch.randn(T, generator=gen)
    u1, opt1 = st.forward(states[0], noise=noise[0], xi=xi, xi_opt=xi_opt)
    p = orc.make_problem(risks[0], res, goals[0].tolist(), thr)
    p.slip_std = std
    ref = orc.mppi_iteration(p, states[0], torch.zeros(T, 2), noise[0], torch.tensor(sig), lam, xi=xi, xi_opt=xi_opt)
    assert_iteration_close(engine_outputs(st, u1, opt1), {k: v.numpy() for k, v in ref.items()}, "no-window stochastic")


def test_batch_of_one_equals_the_single_solver():
    from benchnav_b200 import BatchedMPPI

    dyns, objs, risks, goals, states, thr = _batch_problems(1, 64)
    K, T, sig, lam = 640, 25, [0.5, 0.5], 0.5
    b = BatchedMPPI(T, K, dyns, objs, torch.tensor(sig), lam, seed=8)
    s = make_solver(risks[0], 0.5, goals[0].tolist(), thr, K, T, sig, lam, seed=8)
    for _ in range(2):  # in-engine Philox noise: same seed, same stream (environment 0 adds nothing to the counter)
        ub, ob = b.forward(states)
        us, os_ = s.forward(states[0])
        torch.cuda.synchronize()
        assert torch.equal(b._action_noises[0], s._action_noises)
        assert torch.equal(b._state_seq_batch[0], s._state_seq_batch)
        np.testing.assert_allclose(ub[0].cpu().numpy(), us.cpu().numpy(), rtol=0, atol=2e-6)
        np.testing.assert_allclose(ob[0].cpu().numpy(), os_.cpu().numpy(), rtol=0, atol=1e-5)


def test_error_behaviour_of_the_widened_path():
    from benchnav_b200 import DWA, MPPI, BatchedMPPI, _cabi
    from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

    sig = torch.tensor([0.5, 0.5])
    dyn = UnicycleProblem(GridSpec(16, 0.5), torch.zeros(16, 16))
    obj = GoalObjectives(dyn, torch.tensor([4.0, 4.0]), 0.3)
    with pytest.raises(_cabi.BnvError) as ei:  # a horizon whose slabs cannot fit shared memory even with one warp
        MPPI(3000, 64, 3, 2, dyn, obj, sig, 0.5)
    assert ei.value.code == -3
    other = UnicycleProblem(GridSpec(32, 0.5), torch.zeros(32, 32))
    with pytest.raises(ValueError):  # environments of a batch share the grid geometry
        BatchedMPPI(10, 64, [dyn, other], [obj, GoalObjectives(other, torch.tensor([4.0, 4.0]), 0.3)], sig, 0.5)
    with pytest.raises(ValueError):
        BatchedMPPI(10, 64, [dyn], [obj, obj], sig, 0.5)
    b = BatchedMPPI(10, 64, [dyn, dyn], [obj, obj], sig, 0.5)
    with pytest.raises(AssertionError):
        b.forward(torch.zeros(3, 3))
    with pytest.raises(ValueError):
        b.forward(torch.zeros(2, 3), noise=torch.zeros(2, 64, 9, 2))
    with pytest.raises(_cabi.BnvError):  # top samples before any forward
        b.get_top_samples(4)
    with pytest.raises(AssertionError):  # dwa.py:70-72
        DWA(10, 3, 2, dyn, obj, torch.tensor([0.5]), 0.1)
    with pytest.raises(TypeError):  # stochastic slip needs the slip distribution on the grid map
        MPPI(10, 64, 3, 2, dyn, obj, sig, 0.5, stochastic_slip=True)
    d = DWA(10, 3, 2, dyn, obj, torch.tensor([0.5, 1.5]), 0.1)
    with pytest.raises(AssertionError):  # dwa.py:129-131
        d.forward(torch.zeros(2))
    with pytest.raises(AssertionError):  # dwa.py:154-156
        d.update_reference_path(torch.zeros(5, 3))
