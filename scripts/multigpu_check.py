"""Multi-GPU parity check, launched by torchrun (one rank per GPU):
the sample-sharded solver (NCCL all-gather of the softmax partials + finalize) must reproduce the single-GPU
solver on the same global Philox stream, and get_top_samples must return the global top-n.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multigpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from benchnav_b200 import MPPI
from benchnav_b200.problem import GoalObjectives, GridSpec, UnicycleProblem
from benchnav_b200.synthetic import benchmark_problem


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, T = 512, 50
    risk, start, goal, thr = benchmark_problem(G, 0.5, seed=0)
    dyn = UnicycleProblem(GridSpec(G, 0.5), risk)
    obj = GoalObjectives(dyn, goal, thr)
    sig = torch.tensor([0.5, 0.5])
    # ragged shard sizes; the second size puts every shard on the wide variant of the rollout kernel (several waves),
    # whose last CTA does the exchange, the first on the co-resident one, whose column owners do it
    for exchange, K in (("p2p", 16384 * world + 37), ("nccl", 16384 * world + 37), ("p2p", 49152 * world + 5),
                        ("nccl", 49152 * world + 5)):
        sharded = MPPI(T, K, 3, 2, dyn, obj, sig, 0.5, device=dev, seed=7, process_group=dist.group.WORLD,
                       exchange=exchange)
        if exchange == "p2p":
            assert sharded._fused_exchange, "peer mailboxes could not be attached"
        single = MPPI(T, K, 3, 2, dyn, obj, sig, 0.5, device=dev, seed=7)  # every rank also runs the unsharded solver
        st = start.to(dev)
        for it in range(4):
            # same mean sequence on both solvers: ulp-level differences of u* would otherwise be amplified from one
            # iteration to the next by the peaked softmax (the comparison is per iteration, not of the chain)
            sharded._previous_action_seq.copy_(single._previous_action_seq)
            u_s, o_s = sharded.forward(st)
            u_1, o_1 = single.forward(st)
            torch.cuda.synchronize()
            du = float((u_s - u_1).abs().max())
            do = float((o_s - o_1).abs().max())
            a = sharded._sample_offset
            n = sharded._local_samples
            assert torch.equal(sharded._action_noises, single._action_noises[a:a + n]), "noise depends on the sharding"
            w_s, w_1 = sharded._weights.double(), single._weights[a:a + n].double()
            dw = float((w_s - w_1).abs().max())
            wsum = torch.tensor([float(w_s.sum())], device=dev, dtype=torch.float64)
            dist.all_reduce(wsum)
            assert du <= 5e-6 and do <= 5e-5 and dw <= 1e-5 and abs(float(wsum) - 1.0) <= 1e-5, (it, du, do, dw, float(wsum))
            gathered = [torch.empty_like(u_s) for _ in range(world)]
            dist.all_gather(gathered, u_s)
            for g in gathered:
                assert torch.equal(g, gathered[0])  # every rank holds the same u*
            ts, tw = sharded.get_top_samples(100)
            t1, w1 = single.get_top_samples(100)
            np.testing.assert_allclose(tw.cpu().numpy(), w1.cpu().numpy(), rtol=2e-2, atol=1e-7)
            # the merged rows are the unsharded solver's rows (compared where the weights are clearly separated)
            w_np = w1.cpu().numpy()
            sep = np.ones(100, dtype=bool)
            sep[1:] &= w_np[1:] < 0.98 * w_np[:-1]
            sep[:-1] &= w_np[1:] < 0.98 * w_np[:-1]
            if sep.any():
                assert float((ts.cpu()[torch.from_numpy(sep)] - t1.cpu()[torch.from_numpy(sep)]).abs().max()) <= 1e-4
            if rank == 0:
                print(f"{exchange} it{it}: |du*|={du:.2e} |dopt|={do:.2e} |dw|max={dw:.2e} sum(w)={float(wsum):.7f} "
                      f"shard {a}+{n} of {K}", flush=True)
        if exchange == "p2p":
            # host-driven iteration: rank 0 leads with forward_host(state), the others follow; every rank's result
            # equals the unsharded solver's
            sharded._previous_action_seq.copy_(single._previous_action_seq)
            if rank == 0:
                u_h, o_h = sharded.forward_host(start)
                u_h, o_h = u_h.to(dev), o_h.to(dev)
            else:
                u_h, o_h = sharded.forward_follow()
            u_1, o_1 = single.forward(st)
            torch.cuda.synchronize()
            assert float((u_h - u_1).abs().max()) <= 5e-6 and float((o_h - o_1).abs().max()) <= 5e-5
            if rank == 0:
                print(f"{exchange} K={K}: host-driven iteration ok |du*|={float((u_h - u_1).abs().max()):.2e}", flush=True)
            # the same in two stages: the leader returns on the first completion word (u* written), then collects the states
            sharded._previous_action_seq.copy_(single._previous_action_seq)
            if rank == 0:
                u_h = sharded.forward_action(start).to(dev)
                o_h = sharded.wait_states().to(dev)
            else:
                u_h, o_h = sharded.forward_follow()
            u_1, o_1 = single.forward(st)
            torch.cuda.synchronize()
            assert float((u_h - u_1).abs().max()) <= 5e-6 and float((o_h - o_1).abs().max()) <= 5e-5
            if rank == 0:
                print(f"{exchange} K={K}: host-driven two-stage iteration ok |du*|={float((u_h - u_1).abs().max()):.2e}", flush=True)
        sharded.check()  # no in-kernel wait timed out
        if rank == 0:
            print(f"{exchange} K={K}: launch {sharded.launch_geometry}", flush=True)
        del sharded, single
    dist.barrier()
    if rank == 0:
        print("multigpu_check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
