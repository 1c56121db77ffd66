"""world_size-2 test of the multi-rank plumbing on CPU (gloo): shard geometry + partial exchange + LSE merge.

Each rank evaluates its shard of a golden case with the oracle, the partials travel through
benchnav_b200.dist.gather_shard_partials (the call MPPI.forward makes between bnv_mppi_forward and
bnv_mppi_finalize), and the merged control sequence must equal the reference's single-process result."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q) -> None:
    import sys

    sys.path.insert(0, ROOT)
    from benchnav_b200.dist import ShardInfo, gather_shard_partials, merge_top_candidates, shard_range
    from oracle import mppi_oracle as orc
    from tests.helpers import problem_from_golden

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = dict(np.load(os.path.join(ROOT, "tests", "golden", "kat_g64_k1000_t25.npz")))
        p = problem_from_golden(case)
        lam, T = float(case["lam"]), int(case["T"])
        shard = ShardInfo.from_group(dist.group.WORLD)
        a, b = shard_range(int(case["K"]), shard.rank, shard.world_size)
        out = orc.mppi_iteration(p, torch.from_numpy(case["state_0"]), torch.from_numpy(case["u_prev_0"]),
                                 torch.from_numpy(case["noise_0"][a:b]), torch.from_numpy(case["sigmas"]), lam)
        m, s, u = orc.shard_partial(-out["costs"] * 0 + out["costs"], out["controls"], lam)
        partial = torch.cat([m.view(1), s.view(1), u.reshape(-1)]).float()
        gathered = torch.empty(world, 2 + 2 * T)
        gather_shard_partials(partial, gathered, shard)
        parts = [(gathered[r, 0], gathered[r, 1], gathered[r, 2:].view(T, 2)) for r in range(world)]
        _, _, u_opt = orc.merge_partials(parts, lam)
        # global top-n from shard-local candidate lists
        w_local = torch.softmax(-out["costs"] / lam, 0) * (s * torch.exp(-(m - torch.stack([x[0] for x in parts]).min()) / lam)
                                                            / sum(x[1] * torch.exp(-(x[0] - torch.stack([y[0] for y in parts]).min()) / lam) for x in parts))
        ts, tw = orc.top_samples(out["rec"], w_local, 8)
        gs, gw = merge_top_candidates(ts, tw, 8, shard)
        q.put((rank, u_opt.numpy(), gathered.numpy(), gw.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_partial_exchange_and_merge():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    case = dict(np.load(os.path.join(ROOT, "tests", "golden", "kat_g64_k1000_t25.npz")))
    np.testing.assert_array_equal(res[0][2], res[1][2])  # both ranks hold the same gathered partials
    np.testing.assert_array_equal(res[0][1], res[1][1])  # ... and compute the identical merge
    np.testing.assert_allclose(res[0][1], case["u_opt_0"], atol=2e-5)  # == reference single-process u*
    np.testing.assert_allclose(res[0][3], case["top_weights_0"][:8], atol=1e-5)
    np.testing.assert_array_equal(res[0][3], res[1][3])


def _env_worker(rank: int, world: int, port: int, q) -> None:
    """Batched solvers shard ENVIRONMENTS: every rank plans for its slice of the environment list and nothing is
    exchanged (the process group exists only so that the ranks agree on the partition)."""
    import sys

    sys.path.insert(0, ROOT)
    from benchnav_b200.dist import ShardInfo, shard_range
    from oracle import mppi_oracle as orc
    from tests.helpers import problem_from_golden

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = dict(np.load(os.path.join(ROOT, "tests", "golden", "tiny_g8_k33_t1.npz")))
        p = problem_from_golden(case)
        shard = ShardInfo.from_group(dist.group.WORLD)
        E = 5
        a, b = shard_range(E, shard.rank, shard.world_size)
        out = {}
        for e in range(a, b):  # environment e = the golden problem with its own goal and start state
            pe = orc.make_problem(p.risk, p.resolution, (p.goal[0] - 0.3 * e, p.goal[1] + 0.2 * e), p.stuck_threshold)
            st = torch.from_numpy(case["state_0"]) + torch.tensor([0.1 * e, 0.0, 0.05 * e])
            r = orc.mppi_iteration(pe, st, torch.from_numpy(case["u_prev_0"]), torch.from_numpy(case["noise_0"]),
                                   torch.from_numpy(case["sigmas"]), float(case["lam"]))
            out[e] = r["u_opt"].numpy()
        q.put((rank, (a, b), out))
    finally:
        dist.destroy_process_group()


def test_environment_sharding_needs_no_exchange():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_env_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0][1] == (0, 2) and res[1][1] == (2, 5)  # contiguous, disjoint, covering slices of the 5 environments
    merged = {**res[0][2], **res[1][2]}
    assert sorted(merged) == list(range(5))
    assert not np.array_equal(merged[0], merged[4])  # the environments really are different problems
