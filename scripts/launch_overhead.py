"""Calibrate event-to-event time of tiny kernels vs our rollout kernel at small sizes (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchnav_b200 import MPPI
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem

def timeit(fn, n=2000):
    flush = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    for i in range(n):
        flush.fill_(i & 255)
        e0[i].record(); fn(); e1[i].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    return t[len(t) // 2] * 1e3

x = torch.zeros(32, device="cuda")
print("empty event pair      : %.2f us" % timeit(lambda: None))
print("tiny torch kernel     : %.2f us" % timeit(lambda: x.add_(1.0)))
risk, start, goal, thr = benchmark_problem(256, 0.5, seed=0)
dyn = UnicycleProblem(GridSpec(256, 0.5), risk)
obj = GoalObjectives(dyn, goal, thr)
st = start.cuda()
for K, T in ((128, 50), (128, 2), (16384, 2), (16384, 50), (4096, 50)):
    for rec in (True, False):
        s = MPPI(T, K, 3, 2, dyn, obj, torch.tensor([0.5, 0.5]), 0.5, device=torch.device("cuda"), record_states=rec)
        for _ in range(10): s.forward(st)
        print(f"rollout K={K:6d} T={T:3d} record={rec!s:5}: %.2f us" % timeit(lambda: s.forward(st)))
