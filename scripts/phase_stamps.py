"""Print clock64 phase stamps of the last CTA of the rollout kernel (debug aid; BNV_DEBUG_TS=1)."""
import ctypes as C
import os
import sys

os.environ["BNV_DEBUG_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from benchnav_b200 import MPPI, _cabi
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem

K, T, G = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 50, 256
STOCH = len(sys.argv) > 2 and sys.argv[2] == "stoch"  # BASELINE config 4: stochastic-slip lookups
risk, start, goal, thr = benchmark_problem(G, 0.5, seed=0)
if STOCH:
    from benchnav_b200.problem import SlipDistribution
    from benchnav_b200.synthetic import make_terrain

    terr = make_terrain(G, 0.5, 0)
    dyn = UnicycleProblem(GridSpec(G, 0.5, distributions={"predictions": SlipDistribution(terr["slip_mean"], terr["slip_std"])}),
                          terr["slip_mean"])
else:
    dyn = UnicycleProblem(GridSpec(G, 0.5), risk)
s = MPPI(T, K, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 0.5]), 0.5, device=torch.device("cuda"),
         stochastic_slip=STOCH)
print("launch", s.launch_geometry)
st = start.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
# distributed schedule (default): 4 = barrier passed, 12 = (M, S) known, 5 = weights' (M, S) in smem, 6 = u* gathered,
# 7 = optimal rollout done, 10 = exit; round-1 schedule / wave schedule: 4 = merge entered, 12..15 merge internals,
# 5 = merged, 7 = optimal rollout done
names = ["start", "loop0", "loop1", "partial", "barrier|enter", "ms_smem|merged", "u*_gathered|u_out", "opt_done", "-",
         "-", "exit", "-", "MS_known|ms_loaded", "M_sync", "S_sync", "fma_done"]
INJECT = os.environ.get("BNV_STAMPS_INJECT") == "1"  # injected noise: the T-loop without the in-loop Philox draw
nz = (torch.randn(K, T, 2, device="cuda") * 0.5) if INJECT else None
for it in range(6):
    if it >= 3:
        flush.fill_(it)
    s.forward(st, noise=nz)
    torch.cuda.synchronize()
    ts = (C.c_longlong * 24)()
    _cabi.check(s._lib.bnv_debug_timestamps(s._handle, ts))
    t0 = ts[0]
    print(("cold " if it >= 3 else "warm ") + " ".join(f"{n}={ts[i] - t0}" for i, n in enumerate(names)))
