"""Shared test helpers: golden-case access and oracle problem construction (test side only)."""

from __future__ import annotations

import numpy as np
import torch

from oracle import mppi_oracle as orc


def problem_from_golden(case: dict) -> orc.Problem:
    p = orc.make_problem(torch.from_numpy(case["risk"]), float(case["resolution"]), case["goal"].tolist(),
                         float(case["thr"]))
    lim = case["limits"]
    assert (p.x_min, p.x_max, p.y_min, p.y_max) == tuple(lim.tolist())
    return p


def oracle_call(case: dict, i: int, dtype=torch.float32, u_prev=None):
    p = problem_from_golden(case)
    u_prev = torch.from_numpy(case[f"u_prev_{i}"]) if u_prev is None else u_prev
    return orc.mppi_iteration(p, torch.from_numpy(case[f"state_{i}"]), u_prev,
                              torch.from_numpy(case[f"noise_{i}"]), torch.from_numpy(case["sigmas"]),
                              float(case["lam"]), dtype=dtype)


def t2n(x: torch.Tensor) -> np.ndarray:
    return x.detach().cpu().numpy()


def ext_problem(case: dict, stochastic: bool) -> orc.Problem:
    """Problem of an extended golden case (env / dwa / stoch): `mean`+`std` maps, or a `risk` map."""
    grid = torch.from_numpy(case["mean"] if "mean" in case else case["risk"])
    goal = case["goal"].tolist() if "goal" in case else [0.0, 0.0]
    thr = float(case["thr"]) if "thr" in case else 0.0
    p = orc.make_problem(grid, float(case["resolution"]), goal, thr)
    assert (p.x_min, p.x_max, p.y_min, p.y_max) == tuple(case["limits"].tolist())
    if stochastic:
        p.slip_std = torch.from_numpy(case["std"])
    return p
