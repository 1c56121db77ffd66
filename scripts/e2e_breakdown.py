"""Where the end-to-end microseconds go: wrapper vs raw C call vs device time (debug aid)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from benchnav_b200 import MPPI, _cabi
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem

risk, start, goal, thr = benchmark_problem(256, 0.5, seed=0)
dyn = UnicycleProblem(GridSpec(256, 0.5), risk)
s = MPPI(50, 16384, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 0.5]), 0.5, device=torch.device("cuda"))
st = start.clone().pin_memory()
out = (torch.empty(50, 2).pin_memory(), torch.empty(1, 51, 3).pin_memory())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=2000, do_flush=True):
    for _ in range(20):
        fn()
    acc = 0.0
    for i in range(n):
        if do_flush:
            flush.fill_(i & 255)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        acc += time.perf_counter() - t0
    return acc / n * 1e6


lib, h = s._lib, s._handle
sp, up, op = st.data_ptr(), out[0].data_ptr(), out[1].data_ptr()
stream = s._stream()
print("wrapper forward_host(out=)   flushed: %.2f us" % timed(lambda: s.forward_host(st, out=out)))
print("raw C bnv_mppi_forward_host  flushed: %.2f us" % timed(lambda: lib.bnv_mppi_forward_host(h, sp, None, up, op, stream)))
print("wrapper forward_host(out=)   warm L2: %.2f us" % timed(lambda: s.forward_host(st, out=out), do_flush=False))
print("raw C bnv_mppi_forward_host  warm L2: %.2f us" % timed(lambda: lib.bnv_mppi_forward_host(h, sp, None, up, op, stream), do_flush=False))
sd = start.cuda()
ud, od = torch.empty(50, 2, device="cuda"), torch.empty(51, 3, device="cuda")
def launch_only():
    lib.bnv_mppi_forward(h, sd.data_ptr(), None, ud.data_ptr(), od.data_ptr(), stream)
# host cost of an asynchronous launch (queue drained first, so the call never blocks on a full queue)
print("raw C bnv_mppi_forward (async launch only, host time): %.2f us" % timed(launch_only, do_flush=False))
def launch_sync():
    lib.bnv_mppi_forward(h, sd.data_ptr(), None, ud.data_ptr(), od.data_ptr(), stream)
    torch.cuda.synchronize()
print("raw C bnv_mppi_forward + cudaDeviceSynchronize warm L2: %.2f us" % timed(launch_sync, do_flush=False))
