"""The oracle (oracle/mppi_oracle.py) pinned against outputs of the reference itself.

Fixtures: tests/golden/*.npz, produced by tests/golden/make_golden.py from the unmodified
reference (src/planners/local_planners/mppi.py:130-240) on CPU.  fp32 oracle == reference bit for
bit; fp64 oracle sizes the tolerance used by the GPU parity tests.
"""

import numpy as np
import pytest
import torch

from oracle import mppi_oracle as orc
from tests.helpers import oracle_call, problem_from_golden, t2n

CASES = ["kat_g64_k1000_t25", "corner_wrap_g64_k512_t50", "ragged_g50_k777_t7", "cvar_g64_k512_t50",
         "tiny_g8_k33_t1", "single_sample_g16_k1_t5"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp32_is_bit_exact_with_reference(golden_cases, name):
    case = golden_cases[name]
    for i in range(int(case["n_calls"])):
        out = oracle_call(case, i)
        np.testing.assert_array_equal(t2n(out["rec"]), case[f"rec_{i}"])
        np.testing.assert_array_equal(t2n(out["weights"]), case[f"weights_{i}"])
        np.testing.assert_array_equal(t2n(out["u_opt"]), case[f"u_opt_{i}"])
        np.testing.assert_array_equal(t2n(out["opt_rec"]), case[f"opt_rec_{i}"])
        # the reference feeds u_opt back as the next call's mean sequence, unshifted (mppi.py:217)
        if i + 1 < int(case["n_calls"]):
            np.testing.assert_array_equal(case[f"u_prev_{i + 1}"], case[f"u_opt_{i}"])


def test_known_answer_vector(golden_cases):
    """SURVEY 8c known-answer values (64x64 map seed 1234, K=1000, T=25, MPPI seed 42)."""
    c = golden_cases["kat_g64_k1000_t25"]
    np.testing.assert_allclose(c["u_opt_0"][0], [0.3878429, 0.0843117], atol=1e-6)
    np.testing.assert_allclose(c["u_opt_0"][24], [0.1333414, 0.0360294], atol=1e-6)
    np.testing.assert_allclose(np.abs(c["u_opt_0"]).sum(), 9.465159, atol=2e-5)
    np.testing.assert_allclose(c["opt_rec_0"][0, 0], [8.0180492, 8.0180492, 0.7909472], atol=1e-6)  # quirk D8
    np.testing.assert_allclose(c["opt_rec_0"][0, 25], [8.3605919, 8.3484488, 0.7411563], atol=1e-6)
    np.testing.assert_allclose(c["noise_0"][0, 0], [-0.2373254, -0.1361690], atol=1e-6)
    assert int(c["weights_0"].argmax()) == 876
    np.testing.assert_allclose(c["weights_0"].max(), 0.114479, atol=1e-6)
    np.testing.assert_allclose(c["u_opt_1"][0], [0.5389564, -0.0321167], atol=1e-6)
    np.testing.assert_allclose(np.abs(c["u_opt_1"]).sum(), 10.579299, atol=2e-5)
    np.testing.assert_allclose(c["risk"][0, :3], [0.0231834, 0.3215189, 0.2078754], atol=1e-7)


@pytest.mark.parametrize("name", CASES)
def test_fp64_truth_sizes_the_tolerance(golden_cases, name):
    """|ref32 - fp64| on u* stays well inside the stated parity tolerance (2e-3)."""
    case = golden_cases[name]
    for i in range(int(case["n_calls"])):
        out64 = oracle_call(case, i, dtype=torch.float64)
        du = np.abs(t2n(out64["u_opt"]) - case[f"u_opt_{i}"]).max()
        assert du <= 1e-3, (name, i, du)
        assert abs(float(out64["weights"].sum()) - 1.0) <= 1e-12


@pytest.mark.parametrize("name", CASES)
def test_top_samples_match_reference(golden_cases, name):
    case = golden_cases[name]
    for i in range(int(case["n_calls"])):
        n = case[f"top_weights_{i}"].shape[0]
        s, w = orc.top_samples(torch.from_numpy(case[f"rec_{i}"]), torch.from_numpy(case[f"weights_{i}"]), n)
        np.testing.assert_array_equal(t2n(w), case[f"top_weights_{i}"])
        # ties (exactly equal weights, e.g. underflowed zeros) may be ordered arbitrarily: compare where unique
        uniq = np.concatenate([[True], np.diff(case[f"top_weights_{i}"]) != 0])
        uniq[:-1] &= uniq[1:]
        np.testing.assert_array_equal(t2n(s)[uniq], case[f"top_states_{i}"][uniq])


def test_shared_lookup_identity(golden_cases):
    """trav(rec[k,t]) == trav(clamped successor): the index clamp makes the stage-cost lookup at the
    un-clamped recorded position equal to the next dynamics lookup, which is what lets the
    fused kernel do T+1 lookups instead of 2T+1."""
    case = golden_cases["corner_wrap_g64_k512_t50"]
    p = problem_from_golden(case)
    rec = torch.from_numpy(case["rec_0"])
    raw = rec[:, :-1, :2].reshape(-1, 2)
    clamped = torch.stack([raw[:, 0].clamp(p.x_min, p.x_max), raw[:, 1].clamp(p.y_min, p.y_max)], dim=1)
    assert (raw != clamped).any(), "case must exercise the position clamp"
    assert torch.equal(orc.traversability(p, raw), orc.traversability(p, clamped))


def test_sharded_softmax_merge_equals_global(golden_cases):
    """LSE merge of per-shard (m, s, U) == softmax over all K (mppi.py:193-199)."""
    case = golden_cases["kat_g64_k1000_t25"]
    out = oracle_call(case, 0, dtype=torch.float64)
    lam = float(case["lam"])
    for world in (1, 2, 3, 8):
        bounds = np.linspace(0, 1000, world + 1).astype(int)
        parts = [orc.shard_partial(out["costs"][a:b], out["controls"][a:b], lam) for a, b in zip(bounds[:-1], bounds[1:])]
        _, _, u = orc.merge_partials(parts, lam)
        np.testing.assert_allclose(t2n(u), t2n(out["u_opt"]), atol=1e-12)


def test_oracle_solver_chain_reproduces_reference_sequence(golden_cases):
    """Feeding the reference's noise through OracleSolver reproduces its multi-call trajectory."""
    case = golden_cases["corner_wrap_g64_k512_t50"]
    p = problem_from_golden(case)
    s = orc.OracleSolver(p, int(case["T"]), int(case["K"]), case["sigmas"].tolist(), float(case["lam"]))
    for i in range(int(case["n_calls"])):
        u, rec = s.forward(torch.from_numpy(case[f"state_{i}"]), noise=torch.from_numpy(case[f"noise_{i}"]))
        np.testing.assert_array_equal(t2n(u), case[f"u_opt_{i}"])
        np.testing.assert_array_equal(t2n(rec), case[f"opt_rec_{i}"])
