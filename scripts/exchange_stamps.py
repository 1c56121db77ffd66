"""Wall-clock stamps of the fused multi-GPU exchange (BNV_DEBUG_TS=1), launched by torchrun, one rank per GPU.

Per step and rank, from column 0's owner (CTA 0): kernel start, first peer store issued, last peer store issued, all W
cells collected (ns, %globaltimer of that GPU).  Separates what the exchange costs on the rank that arrives LAST (stores
out -> cells in: the protocol's own latency over NVLink) from what the early ranks additionally wait (rank skew).

    BNV_DEBUG_TS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29515 scripts/exchange_stamps.py [steps]
"""
import ctypes as C
import os
import sys

os.environ["BNV_DEBUG_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from benchnav_b200 import MPPI, _cabi
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, T, K = 512, 50, 16384 * world
    risk, start, goal, thr = benchmark_problem(G, 0.5, seed=0)
    dyn = UnicycleProblem(GridSpec(G, 0.5), risk)
    solver = MPPI(T, K, 3, 2, dyn, GoalObjectives(dyn, goal, thr), torch.tensor([0.5, 0.5]), 0.5, device=dev, seed=42,
                  process_group=dist.group.WORLD)
    assert solver._fused_exchange
    st = start.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for mode in ("flushed", "back_to_back"):
        for it in range(steps + 10):
            if mode == "flushed":
                flush.fill_(it & 0xFF)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            solver.forward(st)
            ev1.record()
            torch.cuda.synchronize()
            ts = (C.c_longlong * 24)()
            _cabi.check(solver._lib.bnv_debug_timestamps(solver._handle, ts))
            if it >= 10:
                rows.append((0 if mode == "flushed" else 1, ts[16], ts[8], ts[9], ts[11], ev0.elapsed_time(ev1) * 1e6))
    mine = torch.tensor(rows, dtype=torch.float64, device=dev)
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    solver.check()
    if rank == 0:
        a = torch.stack(allr).cpu().numpy()  # [world, 2*steps, 6]
        for m, name in ((0, "L2 flushed between steps"), (1, "back to back (one synchronise per step)")):
            sel = a[:, a[0, :, 0] == m, :]
            pre = sel[:, :, 2] - sel[:, :, 1]      # kernel start -> first peer store (this rank's own work)
            send = sel[:, :, 3] - sel[:, :, 2]     # issuing the W x 3 stores
            wait = sel[:, :, 4] - sel[:, :, 3]     # stores issued -> all W cells of column 0 collected
            print(f"--- N={world}, {name}, {sel.shape[1]} steps; ns, median [p10, p90]")
            q = lambda x: f"{np.median(x):8.0f} [{np.percentile(x, 10):8.0f}, {np.percentile(x, 90):8.0f}]"  # noqa: E731
            print(f"start -> first peer store (own work), all ranks   : {q(pre)}")
            print(f"issuing the stores, all ranks                      : {q(send)}")
            print(f"stores issued -> cells collected, all ranks        : {q(wait)}")
            print(f"  ... on the rank that waited LEAST in each step   : {q(wait.min(axis=0))}   <- protocol latency (NVLink + L2)")
            print(f"  ... on the rank that waited MOST in each step    : {q(wait.max(axis=0))}   <- latency + rank skew")
            print(f"event-timed step, max over ranks                   : {q(sel[:, :, 5].max(axis=0))}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
