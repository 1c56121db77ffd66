"""Helpers for the GPU parity tests: build the engine from a golden/oracle problem, compare with tolerance.

Parity tolerance (fp32, identical injected noise; BASELINE.md section 4 / SURVEY 8c):
    |du*|_inf <= 2e-3, |d opt_state|_inf <= 2e-3,
    per-sample cost rel. err <= 2e-4 and |d recorded state| <= 1e-4 for >= 99.9 % of samples,
    |sum(w) - 1| <= 1e-5.
"""

from __future__ import annotations

import numpy as np
import torch

from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from oracle import mppi_oracle as orc

TOL_U = 2e-3
TOL_OPT = 2e-3
TOL_COST_REL = 2e-4
TOL_REC = 1e-4
FRAC_OK = 0.999
TOL_WSUM = 1e-5


def make_solver(risk: torch.Tensor, resolution: float, goal, thr: float, K: int, T: int, sigmas, lam: float,
                seed: int = 42, **kw):
    from benchnav_b200 import MPPI

    grid = GridSpec(int(risk.shape[0]), resolution)
    dyn = UnicycleProblem(grid, risk)
    obj = GoalObjectives(dyn, torch.as_tensor(goal), thr)
    return MPPI(T, K, 3, 2, dyn, obj, torch.as_tensor(sigmas, dtype=torch.float32), lam,
                device=torch.device("cuda"), seed=seed, **kw)


def solver_from_golden(case: dict, **kw):
    return make_solver(torch.from_numpy(case["risk"]), float(case["resolution"]), case["goal"].tolist(),
                       float(case["thr"]), int(case["K"]), int(case["T"]), case["sigmas"], float(case["lam"]), **kw)


def assert_iteration_close(engine: dict, ref: dict, label: str = "") -> None:
    """engine/ref: dicts of numpy arrays with u_opt, opt_rec, rec, weights, costs (costs/rec optional)."""
    du = np.abs(engine["u_opt"] - ref["u_opt"]).max()
    assert du <= TOL_U, f"{label}: |du*| = {du}"
    dopt = np.abs(engine["opt_rec"] - ref["opt_rec"]).max()
    assert dopt <= TOL_OPT, f"{label}: |d opt_rec| = {dopt}"
    wsum = float(engine["weights"].astype(np.float64).sum())
    assert abs(wsum - 1.0) <= TOL_WSUM, f"{label}: sum w = {wsum}"
    if "rec" in ref and engine.get("rec") is not None:
        bad = (np.abs(engine["rec"] - ref["rec"]).reshape(ref["rec"].shape[0], -1).max(axis=1) > TOL_REC)
        assert bad.mean() <= 1 - FRAC_OK, f"{label}: {bad.sum()} of {bad.size} samples off by > {TOL_REC} in rec"
    if "costs" in ref and engine.get("costs") is not None:
        rel = np.abs(engine["costs"] - ref["costs"]) / np.maximum(np.abs(ref["costs"]), 1e-6)
        badc = rel > TOL_COST_REL
        assert badc.mean() <= 1 - FRAC_OK, f"{label}: {badc.sum()} of {badc.size} sample costs off by > {TOL_COST_REL}"
    dw = np.abs(engine["weights"] - ref["weights"]).max()
    assert dw <= 5e-3, f"{label}: |dw| = {dw}"


def engine_outputs(solver, u_opt, opt_rec) -> dict:
    torch.cuda.synchronize()
    return {"u_opt": u_opt.cpu().numpy(), "opt_rec": opt_rec.cpu().numpy(), "weights": solver._weights.cpu().numpy(),
            "costs": solver.costs.cpu().numpy(),
            "rec": solver._state_seq_batch.cpu().numpy() if solver._state_seq_batch is not None else None}


def oracle_outputs(p: orc.Problem, state, u_prev, noise, sigmas, lam, dtype=torch.float32) -> dict:
    out = orc.mppi_iteration(p, torch.as_tensor(state), torch.as_tensor(u_prev), torch.as_tensor(noise),
                             torch.as_tensor(sigmas, dtype=torch.float32), lam, dtype=dtype)
    return {k: v.numpy() for k, v in out.items()}
