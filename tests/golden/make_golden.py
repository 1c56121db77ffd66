#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py [--ref /root/reference]

It imports the reference's ``MPPI`` / ``UnicycleModel`` / ``Objectives`` / ``GridMap``
(src/planners/local_planners/mppi.py etc.) with one stub (``opensimplex.seed``, imported by
src/utils/utils.py:8 but absent from this image), drives ``MPPI.forward`` on synthetic maps and
stores, per call, the inputs (state, previous action sequence, the sigma-scaled noise the
reference drew = ``solver._action_noises``) and outputs (optimal controls, optimal state sequence,
``_state_seq_batch``, ``_weights``, top samples) in ``tests/golden/<case>.npz``.

Nothing is copied from the reference: only its numerical outputs are recorded.
"""

from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch
from torch.distributions import Normal

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference(ref_root: str):
    sys.modules.setdefault("opensimplex", types.SimpleNamespace(seed=lambda s: None))
    for p in (os.path.join(ref_root, "src"), ref_root):
        if p not in sys.path:
            sys.path.insert(0, p)
    from src.environments.grid_map import GridMap
    from src.simulator.problem_formulation.utils import ModelConfig
    from src.simulator.problem_formulation.robot_model import UnicycleModel
    from src.simulator.problem_formulation.objectives import Objectives
    from src.planners.local_planners.mppi import MPPI

    return GridMap, ModelConfig, UnicycleModel, Objectives, MPPI


CASES = {
    # SURVEY 8c known-answer vector: BASELINE config[0] as written (64x64, K=1000, T=25)
    "kat_g64_k1000_t25": dict(G=64, res=0.5, map_seed=1234, map_scale=0.8, std=0.1, metric="expected_value",
                              conf=None, goal=[24.0, 24.0], goal_int=False, thr=0.3, K=1000, T=25,
                              sigmas=[0.5, 0.5], lam=0.5, seed=42, states=[[8.0, 8.0, 0.7853982]] * 2),
    # robot in the map corner heading outwards: position clamps and the angle wrap are exercised;
    # integer goal tensor as in test/test_mppi.py:133
    "corner_wrap_g64_k512_t50": dict(G=64, res=0.5, map_seed=7, map_scale=0.6, std=0.1, metric="expected_value",
                                     conf=None, goal=[24, 24], goal_int=True, thr=0.3, K=512, T=50,
                                     sigmas=[0.5, 0.5], lam=0.5, seed=3,
                                     states=[[0.2, 0.15, 3.1], [0.05, 31.9, -3.12], [31.95, 31.97, 0.3]]),
    # ragged sizes: G not a multiple of 4, resolution not a power of two (true division), odd T, K % 32 != 0
    "ragged_g50_k777_t7": dict(G=50, res=0.3, map_seed=11, map_scale=0.9, std=0.05, metric="expected_value",
                               conf=None, goal=[11.0, 4.5], goal_int=False, thr=0.25, K=777, T=7,
                               sigmas=[0.3, 0.8], lam=1.3, seed=5, states=[[7.4, 7.6, -1.0], [7.5, 7.5, 2.0]]),
    # CVaR risk map (mppi test default metric, test/test_mppi.py:146-148): risks outside [0,1] are clamped
    "cvar_g64_k512_t50": dict(G=64, res=0.5, map_seed=23, map_scale=1.3, map_offset=-0.3, std=0.08, metric="cvar",
                              conf=0.9,
                              goal=[24.0, 24.0], goal_int=False, thr=0.3, K=512, T=50,
                              sigmas=[0.5, 0.5], lam=0.5, seed=42, states=[[8.0, 8.0, 0.0]] * 3),
    # degenerate sizes: one step horizon, tiny K
    "tiny_g8_k33_t1": dict(G=8, res=1.0, map_seed=2, map_scale=0.7, std=0.1, metric="expected_value", conf=None,
                           goal=[6.0, 6.0], goal_int=False, thr=0.3, K=33, T=1, sigmas=[0.5, 0.5], lam=0.5, seed=1,
                           states=[[1.0, 1.0, 0.5], [1.0, 1.0, 0.5]]),
    "single_sample_g16_k1_t5": dict(G=16, res=0.5, map_seed=4, map_scale=0.5, std=0.1, metric="expected_value",
                                    conf=None, goal=[6.0, 6.0], goal_int=False, thr=0.3, K=1, T=5,
                                    sigmas=[0.5, 0.5], lam=0.5, seed=9, states=[[2.0, 2.0, 0.0]] * 2),
}


# Every attribute the engine's host mirror reads through the reference's own objects (benchnav_b200/mppi.py
# _introspect_problem, _slip_distribution and the constructor), as dotted paths from (dynamics, objectives).
SNAPSHOT_PATHS = {
    "dynamics": ["min_action", "max_action", "_grid_map.grid_size", "_grid_map.resolution", "_grid_map.x_limits",
                 "_grid_map.y_limits", "_model_config.mode", "_model_config.inference_metric",
                 "_model_config.confidence_value", "_traversability_model._risks",
                 "_grid_map.distributions[predictions].mean", "_grid_map.distributions[predictions].stddev"],
    "objectives": ["_goal_pos", "_stuck_threshold"],
}


def _walk(obj, path: str):
    for part in path.split("."):
        if part.endswith("]"):
            name, key = part[:-1].split("[")
            obj = getattr(obj, name)[key]
        else:
            obj = getattr(obj, part)  # AttributeError here = the reference no longer has what the engine reads
    return obj


def snapshot_objects(name: str, dyn, obj) -> dict:
    """Attribute snapshot of the REAL reference objects of one case: values (and tensor dtypes) under the exact
    attribute names the reference uses, plus the default ``delta_t`` of ``UnicycleModel.transit``
    (robot_model.py:60).  tests/test_dropin_objects.py rebuilds bare attribute trees from it on the GPU box."""
    import inspect

    out = {}
    for root, paths in SNAPSHOT_PATHS.items():
        for path in paths:
            v = _walk(dyn if root == "dynamics" else obj, path)
            if torch.is_tensor(v):
                v = v.detach().cpu().numpy()  # dtype preserved (the goal may be int64, test/test_mppi.py:133)
            elif v is None:
                v = np.asarray("__none__")
            else:
                v = np.asarray(v)
            out[f"{name}|{root}|{path}"] = v
    out[f"{name}|dynamics|transit.delta_t"] = np.asarray(
        inspect.signature(dyn.transit).parameters["delta_t"].default, dtype=np.float64)
    return out


def build_objects(name: str, c: dict, ref):
    """The reference's own GridMap / UnicycleModel / Objectives of one case (shared by run_case and the drop-in test)."""
    GridMap, ModelConfig, UnicycleModel, Objectives, _ = ref
    g = torch.Generator().manual_seed(c["map_seed"])
    mean = torch.rand(c["G"], c["G"], generator=g) * c["map_scale"] + c.get("map_offset", 0.0)
    std = torch.full((c["G"], c["G"]), c["std"])
    dists = {"predictions": Normal(mean, std), "latent_models": Normal(mean, std)}
    gm = GridMap(c["G"], c["res"], tensors={"heights": torch.zeros(c["G"], c["G"])}, distributions=dists,
                 instance_name=name, device="cpu")
    torch.manual_seed(c["map_seed"] + 1000)  # the VaR/CVaR risk map consumes the global generator
    dyn = UnicycleModel(gm, ModelConfig("inference", c["metric"], c["conf"]), device="cpu")
    goal = torch.tensor(c["goal"]) if c["goal_int"] else torch.tensor(c["goal"], dtype=torch.float32)
    obj = Objectives(dyn, goal_pos=goal, stuck_threshold=c["thr"])
    return gm, dyn, obj


def run_case(name: str, c: dict, ref):
    GridMap, ModelConfig, UnicycleModel, Objectives, MPPI = ref
    g = torch.Generator().manual_seed(c["map_seed"])
    mean = torch.rand(c["G"], c["G"], generator=g) * c["map_scale"] + c.get("map_offset", 0.0)
    std = torch.full((c["G"], c["G"]), c["std"])
    dists = {"predictions": Normal(mean, std), "latent_models": Normal(mean, std)}
    gm = GridMap(c["G"], c["res"], tensors={"heights": torch.zeros(c["G"], c["G"])}, distributions=dists,
                 instance_name=name, device="cpu")
    torch.manual_seed(c["map_seed"] + 1000)  # the VaR/CVaR risk map consumes the global generator
    dyn = UnicycleModel(gm, ModelConfig("inference", c["metric"], c["conf"]), device="cpu")
    goal = torch.tensor(c["goal"]) if c["goal_int"] else torch.tensor(c["goal"], dtype=torch.float32)
    obj = Objectives(dyn, goal_pos=goal, stuck_threshold=c["thr"])
    solver = MPPI(c["T"], c["K"], 3, 2, dyn, obj, torch.tensor(c["sigmas"]), c["lam"],
                  device=torch.device("cpu"), seed=c["seed"])
    out = {
        "risk": dyn._traversability_model._risks.numpy().astype(np.float32),
        "resolution": np.float64(c["res"]),
        "goal": np.asarray(c["goal"], dtype=np.float64),
        "thr": np.float64(c["thr"]),
        "sigmas": np.asarray(c["sigmas"], dtype=np.float32),
        "lam": np.float64(c["lam"]),
        "K": np.int64(c["K"]),
        "T": np.int64(c["T"]),
        "n_calls": np.int64(len(c["states"])),
        "limits": np.asarray([gm.x_limits[0], gm.x_limits[1], gm.y_limits[0], gm.y_limits[1]], dtype=np.float64),
    }
    for i, st in enumerate(c["states"]):
        state = torch.tensor(st, dtype=torch.float32)
        state_in = state.clone()
        u_prev = solver._previous_action_seq.clone()
        with torch.no_grad():
            u_opt, opt_rec = solver.forward(state=state)
        assert torch.equal(state, state_in), "reference mutated the caller's state"
        n_top = min(c["K"], 16)
        top_s, top_w = solver.get_top_samples(n_top)
        out[f"state_{i}"] = state_in.numpy()
        out[f"u_prev_{i}"] = u_prev.numpy().copy()
        out[f"noise_{i}"] = solver._action_noises.numpy().copy()
        out[f"u_opt_{i}"] = u_opt.numpy().copy()
        out[f"opt_rec_{i}"] = opt_rec.numpy().copy()
        out[f"rec_{i}"] = solver._state_seq_batch.numpy().copy()
        out[f"weights_{i}"] = solver._weights.numpy().copy()
        out[f"top_states_{i}"] = top_s.numpy().copy()
        out[f"top_weights_{i}"] = top_w.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--snapshots-only", action="store_true",
                    help="only (re)write ref_object_snapshots.npz: attribute snapshots of the reference objects")
    args = ap.parse_args()
    torch.set_num_threads(1)  # results are thread-count invariant; 1 keeps the run reproducible anyway
    ref = import_reference(args.ref)
    snaps = {}
    for name, c in CASES.items():
        _, dyn, obj = build_objects(name, c, ref)
        snaps.update(snapshot_objects(name, dyn, obj))
    np.savez_compressed(os.path.join(HERE, "ref_object_snapshots.npz"), **snaps)
    print(f"ref_object_snapshots.npz: {len(snaps)} attributes of {len(CASES)} cases")
    if args.snapshots_only:
        return
    for name, c in CASES.items():
        o = run_case(name, c, ref)
        print(f"{name}: u_opt[0]={o['u_opt_0'][0]}, sum|u|={np.abs(o['u_opt_0']).sum():.6f}, "
              f"w_max={o['weights_0'].max():.6f} @ {o['weights_0'].argmax()}, risk range [{o['risk'].min():.3f}, "
              f"{o['risk'].max():.3f}], moved {np.abs(o['rec_0'][:, -1, :2] - o['state_0'][:2]).max():.3f} m")


if __name__ == "__main__":
    main()
