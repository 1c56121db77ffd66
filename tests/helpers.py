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


# ---- the engine's noise stream, restated from its definition (csrc/mppi_math.cuh noise_pair) ----------
_PHILOX_M0, _PHILOX_M1, _PHILOX_W0, _PHILOX_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
PHILOX_KAT = [  # Random123 known-answer vectors for philox4x32-10: (counter, key) -> output
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11) in plain Python integers."""
    c, k = list(counter), list(key)
    for _ in range(10):
        p0, p1 = _PHILOX_M0 * c[0], _PHILOX_M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + _PHILOX_W0) & 0xFFFFFFFF, (k[1] + _PHILOX_W1) & 0xFFFFFFFF]
    return tuple(c)


def engine_noise_pair(sample: int, pair: int, iteration: int, seed: int, sigma0: float, sigma1: float, env: int = 0):
    """Sigma-scaled noise of steps (2 pair, 2 pair + 1) of one sample as the engine defines it (float64 here):
    counter = (global sample, pair, iteration low word, iteration high word + (env << 16)), key = seed;
    each pair of 32-bit words (a, b) -> Box-Muller with u1 = (a + 0.5) 2^-32, angle = 2 pi ((b + 0.5) 2^-32 - 0.5)."""
    import math

    r = philox4x32_10((sample, pair, iteration & 0xFFFFFFFF, ((iteration >> 32) + (env << 16)) & 0xFFFFFFFF),
                      (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))

    def bm(a, b):
        u1 = (a + 0.5) * 2.0 ** -32
        ang = 2.0 * math.pi * ((b + 0.5) * 2.0 ** -32 - 0.5)
        rad = math.sqrt(-2.0 * math.log(u1))
        return rad * math.cos(ang), rad * math.sin(ang)

    a, b = bm(r[0], r[1]), bm(r[2], r[3])
    return sigma0 * a[0], sigma1 * a[1], sigma0 * b[0], sigma1 * b[1]
