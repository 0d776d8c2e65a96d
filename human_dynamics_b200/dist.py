"""Multi-GPU plumbing: clips shard across ranks (one process per GPU), no data-path collective except the
final gather of per-clip outputs (BASELINE config 4; the reference itself is single-GPU).

Works with any torch.distributed backend: "nccl" over NVLink on the box, "gloo" in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_clips: int, rank: int, world: int):
    """Contiguous block partition of clip indices; never splits along T (GroupNorm couples a clip's frames)."""
    base, rem = divmod(num_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_counts(num_clips: int, world: int):
    return [shard_range(num_clips, r, world)[1] - shard_range(num_clips, r, world)[0] for r in range(world)]


def gather_outputs(local: dict, num_clips: int, dst: int = 0, group=None):
    """Gather per-clip outputs (each tensor is [local_clips, ...]) onto rank `dst` in clip order.

    Returns {key: [num_clips, ...]} on dst and None elsewhere.  Ragged shards are padded to the largest shard so
    one fixed-size gather per key suffices.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return {k: v for k, v in local.items()}
    counts = shard_counts(num_clips, world)
    cmax = max(counts)
    out = {} if rank == dst else None
    for k in sorted(local.keys()):
        v = local[k].contiguous()
        if v.shape[0] != counts[rank]:
            raise ValueError('%s: expected %d local clips, got %d' % (k, counts[rank], v.shape[0]))
        if v.shape[0] < cmax:
            pad = torch.zeros((cmax - v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            v = torch.cat([v, pad], dim=0)
        if rank == dst:
            bufs = [torch.empty_like(v) for _ in range(world)]
            dist.gather(v, gather_list=bufs, dst=dst, group=group)
            out[k] = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
        else:
            dist.gather(v, gather_list=None, dst=dst, group=group)
    return out
