"""Is the shared-memory carve-out switch between the L2-flush kernel and the rollout kernel part of the event-timed
step?  Times forward() (CUDA-event pairs, K = 16384, T = 50, G = 256) behind three different flushes of 256 MiB:
torch's fill_ (default carve-out), the library's fill with 0 bytes and with the rollout kernel's 145 KB of dynamic
shared memory."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from benchnav_b200 import MPPI, _cabi
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem

risk, start, goal, thr = benchmark_problem(256, 0.5, seed=0)
dyn = UnicycleProblem(GridSpec(256, 0.5), risk)
s = MPPI(50, 16384, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 0.5]), 0.5, device=torch.device("cuda"))
lib = _cabi.load()
st = start.cuda()
buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def run(flush, n=1500):
    for _ in range(30):
        flush(1)
        s.forward(st)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(ev):
        flush(i & 0xFF)
        a.record()
        s.forward(st)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(ts) / n * 1e3, ts[n // 2] * 1e3


# Does the flush leave L2 full of DIRTY lines whose write-back the iteration's 17 MB of stores then wait for?  The same
# step behind a flush that only READS 256 MiB (L2 left full of clean lines), and for a solver that records no states.
def read_flush(v):
    _cabi.check(lib.bnv_debug_flush(buf.data_ptr(), buf.numel(), 0x80000000, v, stream))


lean = MPPI(50, 16384, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 0.5]), 0.5, device=torch.device("cuda"),
            record_states=False)
for name, fl in (("torch fill_ (default carve-out)", lambda v: buf.fill_(v)),
                 ("library fill, 0 B dynamic smem", lambda v: _cabi.check(lib.bnv_debug_flush(buf.data_ptr(), buf.numel(), 0, v, stream))),
                 ("library fill, 145 KB dynamic smem", lambda v: _cabi.check(lib.bnv_debug_flush(buf.data_ptr(), buf.numel(), 148000, v, stream))),
                 ("no flush (back to back, per-step events)", lambda v: None)):
    mean, med = run(fl)
    print(f"{name:45s} mean {mean:6.2f} us  median {med:6.2f} us per forward()")
tiny = torch.zeros(32, device="cuda")
mean, med = run(lambda v: (buf.fill_(v), tiny.add_(1.0)))
print(f"{'torch fill_ + a 32-element kernel behind it':45s} mean {mean:6.2f} us  median {med:6.2f} us per forward()")
mean, med = run(read_flush)
print(f"{'read-only flush (sum of 256 MiB: clean L2)':45s} mean {mean:6.2f} us  median {med:6.2f} us per forward()")
full = s
s = lean
for name, fl in (("record_states=False, torch fill_", lambda v: buf.fill_(v)), ("record_states=False, read-only flush", read_flush)):
    mean, med = run(fl)
    print(f"{name:45s} mean {mean:6.2f} us  median {med:6.2f} us per forward()")
