"""Sample-sharding plumbing for the multi-GPU path.

Samples are split over ranks by global index; the map, state and mean sequence are replicated.  The only
exchange per control iteration is an all-gather of each shard's softmax partial ``(m, s, U[T,2])``
(2 + 2T floats); every rank then performs the same log-sum-exp merge inside ``bnv_mppi_finalize``.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport; nothing here computes.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Global sample interval [begin, end) owned by ``rank`` -- same formula as bnv_mppi_create()."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} / world_size {world_size} invalid")
    if num_samples < world_size:
        raise ValueError("fewer samples than shards")
    return rank * num_samples // world_size, (rank + 1) * num_samples // world_size


@dataclass
class ShardInfo:
    rank: int = 0
    world_size: int = 1
    group: Optional[object] = None

    @classmethod
    def from_group(cls, group) -> "ShardInfo":
        """``None`` means a single, unsharded solver even if torch.distributed happens to be initialised."""
        if group is None:
            return cls()
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("process_group given but torch.distributed is not initialised")
        return cls(rank=dist.get_rank(group), world_size=dist.get_world_size(group), group=group)


def gather_shard_partials(partial: torch.Tensor, gathered: torch.Tensor, shard: ShardInfo) -> torch.Tensor:
    """All-gather the per-shard ``(m, s, U)`` vectors into ``gathered`` [world, 2+2T] (rank-major)."""
    if shard.world_size == 1:
        gathered.copy_(partial.view(1, -1))
        return gathered
    if gathered.is_cuda:
        dist.all_gather_into_tensor(gathered, partial, group=shard.group)
    else:  # gloo
        parts = [gathered[r] for r in range(shard.world_size)]
        dist.all_gather(parts, partial, group=shard.group)
    return gathered


def attach_peer_mailboxes(lib, handle, shard: ShardInfo) -> bool:
    """Exchange the CUDA IPC handles of the ranks' mailboxes and map the peers' (bnv_mppi_attach_peers).

    Collective: every rank of the group must call it.  Returns True when EVERY rank could map all peers -- then the
    engine exchanges the softmax partials inside the rollout kernel over NVLink peer memory; otherwise all ranks
    fall back to the all-gather + finalize pair."""
    import ctypes as C

    from . import _cabi

    mine = (C.c_ubyte * 64)()
    _cabi.check(lib.bnv_mppi_mailbox_handle(handle, mine))
    gathered = [None] * shard.world_size
    dist.all_gather_object(gathered, bytes(mine), group=shard.group)
    blob = (C.c_ubyte * (64 * shard.world_size)).from_buffer_copy(b"".join(gathered))
    ok = lib.bnv_mppi_attach_peers(handle, blob) == _cabi.BNV_OK
    votes = [None] * shard.world_size
    dist.all_gather_object(votes, bool(ok), group=shard.group)
    if not all(votes):
        if ok:
            raise RuntimeError("peer mailboxes attached on this rank but not on every rank; rebuild the solver with "
                               "exchange='nccl'")
        return False
    return True


def merge_top_candidates(states: torch.Tensor, weights: torch.Tensor, n: int, shard: ShardInfo, lib=None, handle=None,
                         stream: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global top-n from per-shard top lists (each already sorted descending); off the critical path.

    On CUDA tensors with the engine at hand (``lib``, ``handle``) and equally long lists on every rank: ONE all-gather
    of the packed candidate rows {weight, states} and the engine's own selection + gather (``bnv_mppi_merge_top``).
    Otherwise (gloo / CPU tests, ragged lists): plain torch ops."""
    w = shard.world_size
    if lib is not None and handle is not None and weights.is_cuda:
        from . import _cabi

        n_local = torch.tensor([weights.shape[0]], device=weights.device, dtype=torch.int64)
        lo, hi = n_local.clone(), n_local.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=shard.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=shard.group)
        if int(lo.item()) == int(hi.item()) and w * int(lo.item()) >= n:
            m, row_len = int(lo.item()), int(states[0].numel())
            packed = torch.empty(m, 1 + row_len, device=weights.device, dtype=torch.float32)
            packed[:, 0] = weights
            packed[:, 1:] = states.reshape(m, row_len)
            table = torch.empty(w * m, 1 + row_len, device=weights.device, dtype=torch.float32)
            dist.all_gather_into_tensor(table, packed, group=shard.group)
            out_s = torch.empty((n,) + tuple(states.shape[1:]), device=weights.device, dtype=torch.float32)
            out_w = torch.empty(n, device=weights.device, dtype=torch.float32)
            _cabi.check(lib.bnv_mppi_merge_top(handle, table.data_ptr(), w * m, 1 + row_len, n, out_s.data_ptr(),
                                               out_w.data_ptr(), stream))
            return out_s, out_w
    n_local = torch.tensor([weights.shape[0]], device=weights.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(w)]
    dist.all_gather(sizes, n_local, group=shard.group)
    cap = int(max(int(s.item()) for s in sizes))
    pad_w = torch.full((cap,), -1.0, device=weights.device, dtype=weights.dtype)
    pad_w[: weights.shape[0]] = weights
    pad_s = torch.zeros((cap,) + tuple(states.shape[1:]), device=states.device, dtype=states.dtype)
    pad_s[: states.shape[0]] = states
    all_w = [torch.empty_like(pad_w) for _ in range(w)]
    all_s = [torch.empty_like(pad_s) for _ in range(w)]
    dist.all_gather(all_w, pad_w, group=shard.group)
    dist.all_gather(all_s, pad_s, group=shard.group)
    cat_w, cat_s = torch.cat(all_w), torch.cat(all_s)
    top = torch.topk(cat_w, n)
    return cat_s[top.indices], top.values
