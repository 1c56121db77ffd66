"""Drop-in boundary: the host mirror reads the REFERENCE's own objects through exactly the attributes they have.

Three layers:
  * in the build container (``/root/reference`` present) the reference's real ``GridMap`` / ``UnicycleModel`` /
    ``Objectives`` (grid_map.py:12-61, robot_model.py:15-57, objectives.py:11-27) are built for every golden case and
    ``_introspect_problem`` must return exactly what it returns for the ``benchnav_b200.problem`` carriers that the
    rest of the suite and bench.py use -- and the committed attribute snapshots must still describe those objects;
  * everywhere, bare attribute trees rebuilt from the committed snapshots (``tests/golden/ref_object_snapshots.npz``,
    written by make_golden.py from the real objects) must introspect to the same problem;
  * on the GPU, a solver constructed on those trees reproduces the golden outputs of the reference's ``forward``
    (test/test_mppi.py:155-198 is the call sequence being mirrored).
"""

import os
import sys
import types

import numpy as np
import pytest
import torch

from benchnav_b200.mppi import _introspect_problem, _slip_distribution
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem

REF_ROOT = "/root/reference"
CASES = ["kat_g64_k1000_t25", "corner_wrap_g64_k512_t50", "ragged_g50_k777_t7", "cvar_g64_k512_t50",
         "tiny_g8_k33_t1", "single_sample_g16_k1_t5"]


def _carrier_problem(case):
    grid = GridSpec(int(case["risk"].shape[0]), float(case["resolution"]))
    dyn = UnicycleProblem(grid, torch.from_numpy(case["risk"]))
    return dyn, GoalObjectives(dyn, torch.as_tensor(case["goal"].tolist()), float(case["thr"]))


def _assert_same_problem(got, want):
    risks_a, g_a, res_a, xl_a, yl_a, goal_a, thr_a, dt_a = got
    risks_b, g_b, res_b, xl_b, yl_b, goal_b, thr_b, dt_b = want
    assert torch.equal(risks_a.float(), risks_b.float())
    assert (g_a, res_a, tuple(xl_a), tuple(yl_a), thr_a, dt_a) == (g_b, res_b, tuple(xl_b), tuple(yl_b), thr_b, dt_b)
    assert torch.as_tensor(goal_a).double().tolist() == torch.as_tensor(goal_b).double().tolist()


def _tree_from_snapshot(snaps, name):
    """Bare attribute trees (types.SimpleNamespace) carrying what the reference objects carried, under their names."""

    class _Dist:  # a torch.distributions.Normal exposes .mean / .stddev
        pass

    roots = {"dynamics": types.SimpleNamespace(), "objectives": types.SimpleNamespace()}
    dt = 0.1
    for key in snaps:
        case, root, path = key.split("|")
        if case != name:
            continue
        val = snaps[key]
        if path == "transit.delta_t":
            dt = float(val)
            continue
        if val.dtype.kind == "U":
            val = None if str(val) == "__none__" else str(val)
        elif val.ndim == 0:
            val = val.item()
        elif path.endswith("_limits"):
            val = tuple(val.tolist())
        else:
            val = torch.from_numpy(val)  # dtype preserved (int64 goal)
        node = roots[root]
        parts = path.split(".")
        for part in parts[:-1]:
            if part.endswith("]"):
                attr, k = part[:-1].split("[")
                d = getattr(node, attr, None)
                if d is None:
                    d = {}
                    setattr(node, attr, d)
                node = d.setdefault(k, _Dist())
            else:
                if not hasattr(node, part):
                    setattr(node, part, types.SimpleNamespace())
                node = getattr(node, part)
        setattr(node, parts[-1], val)

    def transit(state, action, delta_t=dt):  # robot_model.py:59-60: only the default of delta_t is read
        raise NotImplementedError

    roots["dynamics"].transit = transit
    return roots["dynamics"], roots["objectives"]


@pytest.fixture(scope="module")
def snapshots():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_object_snapshots.npz")
    return dict(np.load(path))


@pytest.mark.parametrize("name", CASES)
def test_snapshot_trees_introspect_like_the_carriers(golden_cases, snapshots, name):
    dyn, obj = _tree_from_snapshot(snapshots, name)
    _assert_same_problem(_introspect_problem(dyn, obj), _introspect_problem(*_carrier_problem(golden_cases[name])))
    mean, std = _slip_distribution(dyn._grid_map)
    g = int(dyn._grid_map.grid_size)
    assert tuple(mean.shape) == (g, g) and tuple(std.shape) == (g, g)
    assert dyn.min_action.tolist() == [0.0, -1.0] and dyn.max_action.tolist() == [1.0, 1.0]  # robot_model.py:54-57
    if name == "corner_wrap_g64_k512_t50":
        assert obj._goal_pos.dtype == torch.int64  # test/test_mppi.py:133 passes an integer goal tensor


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "src")), reason="the reference tree is not on this box")
@pytest.mark.parametrize("name", CASES)
def test_real_reference_objects_introspect_like_the_carriers(golden_cases, snapshots, name):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg

    ref = mg.import_reference(REF_ROOT)
    _, dyn, obj = mg.build_objects(name, mg.CASES[name], ref)
    case = golden_cases[name]
    _assert_same_problem(_introspect_problem(dyn, obj), _introspect_problem(*_carrier_problem(case)))
    # the real Normal distribution is readable as (mean, std) for the stochastic-slip mode
    mean, std = _slip_distribution(dyn._grid_map)
    assert tuple(mean.shape) == tuple(case["risk"].shape) and float(std.min()) > 0
    # and the committed snapshots still describe these objects
    fresh = mg.snapshot_objects(name, dyn, obj)
    for key, val in fresh.items():
        np.testing.assert_array_equal(np.asarray(val), snapshots[key], err_msg=key)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat_g64_k1000_t25", "corner_wrap_g64_k512_t50", "cvar_g64_k512_t50"])
def test_solver_on_reference_shaped_objects_reproduces_golden(golden_cases, snapshots, name):
    """Tutorial 3.3's construction (test/test_mppi.py:160-169) with objects that carry the reference's attributes."""
    from benchnav_b200 import MPPI
    from tests.gpu_common import assert_iteration_close, engine_outputs

    case = golden_cases[name]
    dyn, obj = _tree_from_snapshot(snapshots, name)
    solver = MPPI(horizon=int(case["T"]), num_samples=int(case["K"]), dim_state=3, dim_control=2, dynamics=dyn,
                  objectives=obj, sigmas=torch.from_numpy(case["sigmas"]), lambda_=float(case["lam"]),
                  device=torch.device("cuda"), seed=42)
    for i in range(int(case["n_calls"])):
        solver._previous_action_seq.copy_(torch.from_numpy(case[f"u_prev_{i}"]))
        with torch.no_grad():
            u, opt = solver.forward(state=torch.from_numpy(case[f"state_{i}"]), noise=torch.from_numpy(case[f"noise_{i}"]))
        ref = {"u_opt": case[f"u_opt_{i}"], "opt_rec": case[f"opt_rec_{i}"], "rec": case[f"rec_{i}"],
               "weights": case[f"weights_{i}"]}
        assert_iteration_close(engine_outputs(solver, u, opt), ref, f"{name}[{i}] on reference-shaped objects")
    top_s, top_w = solver.get_top_samples(num_samples=min(int(case["K"]), 16))
    i = int(case["n_calls"]) - 1
    np.testing.assert_allclose(top_w.cpu().numpy(), case[f"top_weights_{i}"], atol=5e-3)
    assert tuple(top_s.shape) == tuple(case[f"top_states_{i}"].shape)
