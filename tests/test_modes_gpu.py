"""GPU parity tests of the solver modes that round 1 compiled but never ran:
``record_states=False`` (every kRecord=false instantiation), ``noise_source="torch"`` and the general-angle
instantiations (kFastAngles=false, taken when ``dt * max|omega| >= 3`` rad per step).  All through the C ABI."""

import ctypes as C
from dataclasses import replace

import numpy as np
import pytest
import torch

from benchnav_b200 import MPPI, _cabi
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem
from oracle import mppi_oracle as orc
from tests.gpu_common import assert_iteration_close, engine_outputs, make_solver, oracle_outputs, solver_from_golden
from tests.helpers import problem_from_golden

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------ record_states=False
@pytest.mark.parametrize("name", ["kat_g64_k1000_t25", "ragged_g50_k777_t7", "tiny_g8_k33_t1"])
def test_lean_solver_golden_calls(golden_cases, name):
    """Without recorded states the iteration's results are unchanged (u*, weights, optimal state sequence, costs):
    the reference's golden outputs call by call; `_state_seq_batch` is absent and get_top_samples refuses."""
    case = golden_cases[name]
    lean = solver_from_golden(case, record_states=False)
    full = solver_from_golden(case)
    p = problem_from_golden(case)
    assert lean._state_seq_batch is None
    for i in range(int(case["n_calls"])):
        for s in (lean, full):
            s._previous_action_seq.copy_(torch.from_numpy(case[f"u_prev_{i}"]))
        u, opt = lean.forward(torch.from_numpy(case[f"state_{i}"]), noise=torch.from_numpy(case[f"noise_{i}"]))
        uf, optf = full.forward(torch.from_numpy(case[f"state_{i}"]), noise=torch.from_numpy(case[f"noise_{i}"]))
        eng = engine_outputs(lean, u, opt)
        ref = {"u_opt": case[f"u_opt_{i}"], "opt_rec": case[f"opt_rec_{i}"], "weights": case[f"weights_{i}"],
               "costs": oracle_outputs(p, case[f"state_{i}"], case[f"u_prev_{i}"], case[f"noise_{i}"], case["sigmas"],
                                       float(case["lam"]))["costs"]}
        assert_iteration_close(eng, ref, f"lean {name}[{i}]")
        # and bit-identical to the recording solver: dropping the slab changes no arithmetic
        assert torch.equal(u, uf) and torch.equal(opt, optf)
        assert torch.equal(lean._weights, full._weights) and torch.equal(lean.costs, full.costs)
        # get_top_samples of the lean solver re-rolls the selected samples: bit-identical to the recorded rows
        n = min(int(case["K"]), 16)
        ts_l, tw_l = lean.get_top_samples(n)
        ts_f, tw_f = full.get_top_samples(n)
        torch.cuda.synchronize()
        assert torch.equal(tw_l, tw_f)
        w_np = tw_f.cpu().numpy()
        uniq = np.concatenate([[True], np.diff(w_np) != 0])
        uniq[:-1] &= uniq[1:]  # rows whose weight is unique (ties may come back in either order)
        assert torch.equal(ts_l[torch.from_numpy(uniq)], ts_f[torch.from_numpy(uniq)])


@pytest.mark.parametrize("K,T,G", [(16384, 50, 256), (40000, 30, 64)])
def test_lean_solver_full_size_philox(K, T, G):
    """In-engine noise, full size (one co-resident wave and a multi-wave grid): oracle on the drawn noise."""
    risk, start, goal, thr = benchmark_problem(G, 0.5, seed=0)
    sig, lam = [0.5, 0.5], 0.5
    s = make_solver(risk, 0.5, goal, thr, K, T, sig, lam, record_states=False)
    p = orc.make_problem(risk, 0.5, goal.tolist(), thr)
    u_prev = torch.zeros(T, 2)
    for it in range(2):
        u, opt = s.forward(start)
        eng = engine_outputs(s, u, opt)
        assert eng["rec"] is None
        ref = oracle_outputs(p, start, u_prev, s._action_noises.cpu(), sig, lam)
        rec_ref = ref.pop("rec")
        assert_iteration_close(eng, ref, f"lean philox K{K} it{it}")
        # re-rolled top samples against the oracle's recorded states of the same sample indices
        ts, tw = s.get_top_samples(64)
        torch.cuda.synchronize()
        w_all = s._weights.cpu()
        order = torch.argsort(w_all, descending=True, stable=True)[:64]
        np.testing.assert_array_equal(tw.cpu().numpy(), w_all[order].numpy())
        distinct = torch.ones(64, dtype=torch.bool)
        distinct[1:] &= tw.cpu()[1:] != tw.cpu()[:-1]
        distinct[:-1] &= tw.cpu()[:-1] != tw.cpu()[1:]
        bad = (ts.cpu()[distinct] - torch.from_numpy(rec_ref)[order][distinct]).abs().reshape(int(distinct.sum()), -1).max(dim=1).values > 1e-4
        assert int(bad.sum()) <= 1, f"{int(bad.sum())} re-rolled top samples off"
        u_prev = torch.from_numpy(eng["u_opt"])


# ------------------------------------------------------------------------------------------ noise_source="torch"
def test_torch_noise_source_parity_and_draw_order():
    """noise_source='torch': (1) the iteration on the torch-drawn noise equals the oracle on the same tensor; (2) the
    draws are the ones the reference makes on a CUDA device -- torch.manual_seed(seed), one throw-away
    MultivariateNormal.rsample in the constructor (mppi.py:99-107), one per forward (mppi.py:149-151)."""
    from torch.distributions import MultivariateNormal

    risk, start, goal, thr = benchmark_problem(64, 0.5, seed=0)
    K, T, lam, seed = 3000, 25, 0.5, 17
    for sig in ([0.5, 0.25], [0.3, 0.8]):  # power-of-two sigmas: bit-equal; general: within the Cholesky's rounding
        s = make_solver(risk, 0.5, goal, thr, K, T, sig, lam, seed=seed, noise_source="torch")
        ctor_draw = s._action_noises.clone()
        p = orc.make_problem(risk, 0.5, goal.tolist(), thr)
        u_prev = torch.zeros(T, 2)
        drawn = []
        for it in range(2):
            u, opt = s.forward(start)
            eng = engine_outputs(s, u, opt)
            drawn.append(s._action_noises.clone())
            assert tuple(s._action_noises.shape) == (K, T, 2)
            ref = oracle_outputs(p, start, u_prev, s._action_noises.cpu(), sig, lam)
            assert_iteration_close(eng, ref, f"torch noise sig={sig} it{it}")
            u_prev = torch.from_numpy(eng["u_opt"])
        # the reference's construction + two forwards on the CUDA generator
        torch.manual_seed(seed)
        sigmas = torch.tensor(sig)
        dist = MultivariateNormal(loc=torch.zeros(2, device="cuda"), covariance_matrix=torch.diag(sigmas ** 2).cuda())
        want = [dist.rsample(sample_shape=torch.Size([K, T])) for _ in range(3)]
        for got, ref_draw in zip([ctor_draw] + drawn, want):
            if sig == [0.5, 0.25]:
                assert torch.equal(got, ref_draw)
            else:
                torch.testing.assert_close(got, ref_draw, rtol=2e-7, atol=0)


def test_torch_noise_shard_slice():
    """A sharded solver draws the full [K,T,2] tensor on every rank (same generator state) and keeps its own rows."""
    risk, start, goal, thr = benchmark_problem(64, 0.5, seed=0)
    K, T, sig = 1000, 10, [0.5, 0.25]
    s = make_solver(risk, 0.5, goal, thr, K, T, sig, 0.5, seed=3, noise_source="torch")
    torch.manual_seed(99)
    full = s._draw_torch_noise()
    s._sample_offset, s._local_samples = 250, 125  # what ShardInfo(rank 2 of 8) would have set
    torch.manual_seed(99)
    part = s._draw_torch_noise()
    s._sample_offset, s._local_samples = 0, K
    assert torch.equal(part, full[250:375])


# ------------------------------------------------------------------------------------------ general-angle path
@pytest.mark.parametrize("record", [True, False])
@pytest.mark.parametrize("philox", [False, True])
def test_large_turn_rates_take_the_general_angle_path(record, philox):
    """dt * max|omega| = 4 rad per step: the heading may wrap by more than pi in one step, so the branch-free step is
    not valid and the kFastAngles=false instantiations run (general fmod wrap, range-checked sin/cos)."""
    g, res, K, T, lam = 64, 0.5, 2048, 24, 0.5
    sig = [0.5, 25.0]
    risk, start, goal, thr = benchmark_problem(g, res, seed=2)
    lo, hi = (0.0, -40.0), (1.0, 40.0)
    dyn = UnicycleProblem(GridSpec(g, res), risk, min_action=lo, max_action=hi)
    obj = GoalObjectives(dyn, goal, thr)
    solver = MPPI(T, K, 3, 2, dyn, obj, torch.tensor(sig), lam, device=torch.device("cuda"), seed=4,
                  record_states=record)
    p = replace(orc.make_problem(risk, res, goal.tolist(), thr), u_min=lo, u_max=hi)
    gen = torch.Generator().manual_seed(8)
    u_prev = torch.zeros(T, 2)
    for it in range(2):
        noise = None if philox else torch.randn(K, T, 2, generator=gen) * torch.tensor(sig)
        solver._previous_action_seq.copy_(u_prev)
        u, opt = solver.forward(start, noise=noise)
        eng = engine_outputs(solver, u, opt)
        used = solver._action_noises.cpu()
        ref = oracle_outputs(p, start, u_prev, used, sig, lam)
        if not record:
            ref.pop("rec")
        # headings really do turn by more than pi per step somewhere
        if record:
            turn = np.abs(np.diff(eng["rec"][:, :, 2], axis=1))
            assert float(turn.max()) > 3.2
        assert_iteration_close(eng, ref, f"general angles record={record} philox={philox} it{it}")
        u_prev = torch.from_numpy(ref["u_opt"])


# ------------------------------------------------------------------------------------------ pre-launch + setters
def test_setters_cancel_a_prelaunched_kernel():
    """a waiting pre-launched kernel holds a by-value copy of the old parameters; set_goal_dev /
    set_terminal_goal / set_keep_mean must cancel it so that the next forward_host sees the new values."""
    risk, start, goal, thr = benchmark_problem(64, 0.5, seed=0)
    K, T = 1024, 20
    plain = make_solver(risk, 0.5, goal.tolist(), thr, K, T, [0.5, 0.5], 0.5, seed=5)
    pre = make_solver(risk, 0.5, goal.tolist(), thr, K, T, [0.5, 0.5], 0.5, seed=5)
    pre.prelaunch(True, timeout_us=200000)
    new_goal = torch.tensor([5.0, 20.0], device="cuda")
    for step in range(6):
        if step == 2:
            for s in (plain, pre):
                _cabi.check(s._lib.bnv_mppi_set_goal_dev(s._handle, new_goal.data_ptr()))
        if step == 4:
            tg = (C.c_float * 2)(9.0, 9.0)
            for s in (plain, pre):
                _cabi.check(s._lib.bnv_mppi_set_terminal_goal(s._handle, tg))
                _cabi.check(s._lib.bnv_mppi_set_keep_mean(s._handle, 0))
        u_ref, o_ref = plain.forward_host(start)
        u, o = pre.forward_host(start)
        assert torch.equal(u, u_ref) and torch.equal(o, o_ref), f"step {step}"
    pre.prelaunch(False)


def test_close_releases_the_handle_and_views_keep_it_alive():
    """no reference cycle -- dropping the solver and its views destroys the engine handle (device memory
    returns), and close() does so at once."""
    import gc

    risk, start, goal, thr = benchmark_problem(64, 0.5, seed=0)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    s = make_solver(risk, 0.5, goal, thr, 65536, 50, [0.5, 0.5], 0.5)  # ~ 45 MB of engine buffers
    s.forward(start)
    w = s._weights
    torch.cuda.synchronize()
    used = free0 - torch.cuda.mem_get_info()[0]
    assert used > 30 << 20
    del s
    gc.collect()
    assert float(w.sum()) > 0.99  # the view keeps the handle (and its buffers) alive
    del w
    gc.collect()
    torch.cuda.synchronize()
    assert free0 - torch.cuda.mem_get_info()[0] < used // 4
    s2 = make_solver(risk, 0.5, goal, thr, 4096, 20, [0.5, 0.5], 0.5)
    s2.close()
    with pytest.raises(_cabi.BnvError):
        s2.forward(start)
