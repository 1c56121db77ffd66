"""Host-side logic that needs no GPU: shard geometry, problem introspection, synthetic inputs."""

import numpy as np
import pytest
import torch

from benchnav_b200.dist import ShardInfo, shard_range
from benchnav_b200.mppi import _introspect_problem
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem, make_terrain


@pytest.mark.parametrize("k,w", [(16384, 1), (16384, 8), (131072, 8), (5001, 3), (8, 8), (1000, 7)])
def test_shard_ranges_partition_the_samples(k, w):
    spans = [shard_range(k, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == k
    for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
        assert a1 == b0 and a1 > a0
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_shard_range_rejects_bad_geometry():
    with pytest.raises(ValueError):
        shard_range(4, 0, 8)
    with pytest.raises(ValueError):
        shard_range(64, 3, 3)
    assert ShardInfo.from_group(None) == ShardInfo(0, 1, None)


def test_grid_limits_follow_the_reference_formula(golden_cases):
    for case in golden_cases.values():
        if "limits" not in case:
            continue
        grid = case["risk"] if "risk" in case else case["mean"]
        g = GridSpec(grid.shape[0], float(case["resolution"]))
        assert (g.x_limits[0], g.x_limits[1], g.y_limits[0], g.y_limits[1]) == tuple(case["limits"].tolist())


def test_introspection_reads_reference_shaped_objects():
    grid = GridSpec(32, 0.25)
    risk = torch.rand(32, 32)
    dyn = UnicycleProblem(grid, risk)
    obj = GoalObjectives(dyn, torch.tensor([3, 5]), 0.3)  # integer goal as in test/test_mppi.py:133
    risks, g, res, xl, yl, goal, thr, dt = _introspect_problem(dyn, obj)
    assert g == 32 and res == 0.25 and xl == (0.0, 8.0) and thr == 0.3 and dt == 0.1
    assert torch.equal(risks, risk)
    dyn._model_config.mode = "observation"
    with pytest.raises(ValueError):
        _introspect_problem(dyn, obj)
    with pytest.raises(TypeError):
        _introspect_problem(object(), obj)


def test_synthetic_problem_is_deterministic_and_sane():
    a = make_terrain(64, 0.5, seed=3)
    b = make_terrain(64, 0.5, seed=3)
    for k in a:
        assert torch.equal(a[k], b[k])
    risk, start, goal, thr = benchmark_problem(256, 0.5, seed=0)
    assert risk.shape == (256, 256) and risk.dtype == torch.float32
    assert 0.0 <= float(risk.min()) and float(risk.max()) <= 1.0
    assert float(risk[16, 16]) <= 0.2 + 1e-6  # start cell drivable
    np.testing.assert_allclose(goal.numpy(), [48.0, 48.0])
    assert 0.01 < float((1 - risk <= thr).float().mean()) < 0.5  # some, not all, cells are "stuck"


def test_product_code_never_imports_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under benchnav_b200/ (nor the example) may import it, or the
    reference package, or fall back to CPU arithmetic (static check over the package sources)."""
    import ast
    import glob
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "benchnav_b200", "*.py")) + glob.glob(os.path.join(root, "examples", "*.py"))
    assert len(files) >= 10
    for path in files:
        tree = ast.parse(open(path).read(), filename=path)
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                top = n.split(".")[0]
                assert top not in ("oracle", "src", "simulator", "planners", "environments"), f"{path} imports {n}"
    for path in glob.glob(os.path.join(root, "benchnav_b200", "csrc", "*")):
        for line in open(path):
            if line.lstrip().startswith("#include"):
                assert "oracle" not in line.lower() and "reference" not in line.lower(), (path, line)


def test_python_restatement_of_philox_matches_the_published_known_answers():
    from tests.helpers import PHILOX_KAT, philox4x32_10

    for counter, key, want in PHILOX_KAT:
        assert philox4x32_10(counter, key) == want


# ------------------------------------------------------------------------------------------ bench.py host logic
def test_bench_workloads_bytes_and_config_agree_between_arms():
    """bench.py: the SURVEY 8d byte formula reproduces the survey's per-unit figures, every workload resolves, the
    `config` object is identical in the native and the reference arm (built by one function), and a stale ncu traffic
    figure is refused."""
    import json

    import bench

    assert bench.algorithmic_bytes(16384, 50, 256) == 16_909_300      # C1
    assert bench.algorithmic_bytes(1000, 25, 64) == 532_896           # C0a
    assert bench.algorithmic_bytes(5000, 50, 64) == 5_097_396         # C0b
    assert bench.algorithmic_bytes(16384, 50, 512) == 17_695_732      # C2 per GPU
    assert bench.algorithmic_bytes(4096, 30, 64) == 2_540_132         # C3 per environment
    for name, world in (("auto", 1), ("auto", 8), ("c2", 1), ("c2", 4), ("c3", 8), ("c4", 1)):
        w = bench.resolve_workload(name, world)
        cfg = bench.config_of(w, world)
        assert json.dumps(cfg) == json.dumps(bench.config_of(bench.resolve_workload(name, world), world))
        assert cfg["num_samples_total"] == w["k_total"] and cfg["n_gpus"] == world
    assert bench.resolve_workload("auto", 8)["k_total"] == 131072 and bench.resolve_workload("auto", 8)["grid"] == 512
    assert bench.unit_factor(bench.resolve_workload("auto", 4), 4) == 4
    assert bench.unit_factor(bench.resolve_workload("c2", 4), 4) == 1
    with pytest.raises(SystemExit):
        bench.resolve_workload("c1", 2)
    with pytest.raises(SystemExit):
        bench.resolve_workload("c3", 3)
    traffic, note = bench.ncu_traffic("no_such_workload")
    assert traffic is None and "no ncu capture" in note
