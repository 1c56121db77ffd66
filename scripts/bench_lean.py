#!/usr/bin/env python
"""Lean mode (record_states=False) against the full contract: per-iteration time and the cost of get_top_samples(500)
(Tutorial 3.3 calls it every step), which re-rolls the selected samples when no states were recorded.  Per-step
CUDA-event pairs, 256 MiB L2 flush between steps.  One JSON object per line.  (lean numbers are reported
separately, never against the full-contract byte formula.)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchnav_b200 import MPPI  # noqa: E402
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem  # noqa: E402
from benchnav_b200.synthetic import benchmark_problem  # noqa: E402

DEV = torch.device("cuda")
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, n=300, warm=20):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for a, b in ev:
        FLUSH.fill_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n * 1e3


for K, T, g in ((16384, 50, 256), (32768, 50, 256), (131072, 50, 512)):
    risk, start, goal, thr = benchmark_problem(g, 0.5, seed=0)
    dyn = UnicycleProblem(GridSpec(g, 0.5), risk)
    row = {"case": f"G={g}, K={K}, T={T}"}
    for name, rec in (("full", True), ("lean", False)):
        s = MPPI(T, K, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 0.5]), 0.5, device=DEV, seed=1,
                 record_states=rec)
        st = start.to(DEV)
        row[f"{name}_forward_us"] = round(timed(lambda: s.forward(st)), 2)
        s.forward(st)
        row[f"{name}_top500_us"] = round(timed(lambda: s.get_top_samples(500), n=100, warm=5), 2)
        row[f"{name}_launch"] = s.launch_geometry
        s.close()
    print(json.dumps(row))
