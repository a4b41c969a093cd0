"""Multi-GPU plumbing: index-range sharding and the per-tick concatenation of
each GPU's due list.

The record array shards by contiguous global index range, so the global
ascending due list is simply the rank-ordered concatenation of the local
lists.  Status columns never leave their owner GPU; the only exchange per tick
is (1) one count per rank and (2) the variable-length lists.  NCCL has no
allgatherv, so lists are padded to the largest count of the tick.

Backend-agnostic on purpose (torch.distributed: "nccl" on the GPUs, "gloo" in
the CPU tests); nothing here touches record data.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """[first, first+count) of the global index range owned by `rank`: sizes
    differ by at most one, lower ranks take the remainder."""
    q, r = divmod(n_total, world)
    first = rank * q + min(rank, r)
    return first, q + (1 if rank < r else 0)


def allgather_due(idx_local: torch.Tensor, act_local: torch.Tensor, count, shard_base: int,
                  group=None):
    """Concatenate every rank's (local index, action) list in rank order.

    idx_local: int32/int64 tensor of LOCAL indices, at least `count` valid
    entries; act_local: uint8 actions.  Returns (global idx int64[n_total_due],
    action uint8[n_total_due], counts list[int]) on every rank."""
    world = dist.get_world_size(group)
    dev = idx_local.device
    cnt_t = count if isinstance(count, torch.Tensor) else torch.tensor([int(count)], device=dev)
    cnt_t = cnt_t.reshape(1).to(torch.int64)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, cnt_t, group=group)
    counts = [int(c) for c in counts_t.tolist()]  # host sync: sizes the padded exchange
    maxc = max(counts) if counts else 0
    mine = counts[dist.get_rank(group)]
    if maxc == 0:
        return (torch.empty(0, dtype=torch.int64, device=dev),
                torch.empty(0, dtype=torch.uint8, device=dev), counts)
    pad_idx = torch.zeros(maxc, dtype=torch.int64, device=dev)
    pad_idx[:mine] = idx_local[:mine].to(torch.int64) + shard_base
    pad_act = torch.zeros(maxc, dtype=torch.uint8, device=dev)
    pad_act[:mine] = act_local[:mine]
    all_idx = torch.empty(world * maxc, dtype=torch.int64, device=dev)
    all_act = torch.empty(world * maxc, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(all_idx, pad_idx, group=group)
    dist.all_gather_into_tensor(all_act, pad_act, group=group)
    idx = torch.cat([all_idx[r * maxc: r * maxc + counts[r]] for r in range(world)])
    act = torch.cat([all_act[r * maxc: r * maxc + counts[r]] for r in range(world)])
    return idx, act, counts
