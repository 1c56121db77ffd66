#!/usr/bin/env python
"""Small drivers for ncu captures of the widened path (run under `ncu -k regex:...`): a few launches each of
  batch   BASELINE config 3's per-GPU share (8 environments x K=4096 x T=30, one launch)
  stoch   BASELINE config 4 (256x256, K=32768, T=50, stochastic slip)
  risk    Monte-Carlo CVaR risk map (256x256, 1000 draws per cell)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchnav_b200 import MPPI, BatchedMPPI, infer_risk_map  # noqa: E402
from benchnav_b200.problem import GoalObjectives, GridSpec, SlipDistribution, UnicycleProblem  # noqa: E402
from benchnav_b200.synthetic import benchmark_problem, make_terrain  # noqa: E402

DEV = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "batch"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
if which == "batch":
    dyns, objs, states = [], [], []
    for e in range(8):
        risk, start, goal, thr = benchmark_problem(64, 0.5, seed=e)
        d = UnicycleProblem(GridSpec(64, 0.5), risk)
        dyns.append(d)
        objs.append(GoalObjectives(d, goal, thr))
        states.append(start)
    solver = BatchedMPPI(30, 4096, dyns, objs, torch.tensor([0.5, 0.5]), 0.5, device=DEV, seed=1)
    st = torch.stack(states).to(DEV)
    for _ in range(n):
        solver.forward(st)
elif which == "stoch":
    terr = make_terrain(256, 0.5, 0)
    d = SlipDistribution(terr["slip_mean"], terr["slip_std"])
    dyn = UnicycleProblem(GridSpec(256, 0.5, distributions={"predictions": d}), terr["slip_mean"])
    obj = GoalObjectives(dyn, torch.tensor([48.0, 48.0]), 0.3)
    solver = MPPI(50, 32768, 3, 2, dyn, obj, torch.tensor([0.5, 0.5]), 0.5, device=DEV, seed=1, stochastic_slip=True)
    st = torch.tensor([8.0, 8.0, 0.785398], device=DEV)
    for _ in range(n):
        solver.forward(st)
elif which == "risk":
    terr = make_terrain(256, 0.5, 0)
    mean, std = terr["slip_mean"].to(DEV), terr["slip_std"].to(DEV)
    for _ in range(n):
        infer_risk_map(mean, std, "cvar", 0.9, method="monte_carlo", num_samples=1000)
torch.cuda.synchronize()
print("done", which)
