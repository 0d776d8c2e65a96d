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


# cams / shapes are column slices of omegas (src/omega.py:231-235): the root re-derives them instead of receiving them
DERIVED = {'cams': ('omegas', 0, 3), 'shapes': ('omegas', 75, 85), 'shapes_delta': ('omegas_delta', 75, 85)}
# what BASELINE.json's north_star names as the per-clip outputs (85-d SMPL params, vertices, keypoints)
DEFAULT_KEYS = ('omegas', 'verts', 'kps')


class OutputGatherer(object):
    """Final gather of per-clip outputs onto `dst`, off the compute stream.

    Point-to-point (rank r -> dst) straight into the rows of one preallocated [num_clips, ...] buffer per key on dst: no
    padding for ragged shards, no concatenation pass, and the transfers of one call run on a side stream -- the caller
    keeps launching compute (the delta heads) while the dt=0 outputs travel, and joins with `wait()`.
    NCCL over NVLink on the box, gloo in the CPU tests.
    """

    def __init__(self, num_clips, dst=0, group=None):
        self.num_clips, self.dst, self.group = int(num_clips), int(dst), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.counts = shard_counts(self.num_clips, self.world)
        self.starts = [shard_range(self.num_clips, r, self.world)[0] for r in range(self.world)]
        self.full = {}              # key -> [num_clips, ...] on dst
        self.pending = []
        self.side = None

    def _buffer(self, key, like):
        shape = (self.num_clips,) + tuple(like.shape[1:])
        buf = self.full.get(key)
        if buf is None or tuple(buf.shape) != shape or buf.dtype != like.dtype or buf.device != like.device:
            buf = torch.empty(shape, dtype=like.dtype, device=like.device)
            self.full[key] = buf
        return buf

    def start(self, local: dict):
        """Begin gathering these tensors ([local_clips, ...] each).  Returns immediately; call wait() before reading."""
        if self.world == 1:
            for k, v in local.items():
                self.full[k] = v
            return
        cuda = any(v.is_cuda for v in local.values())
        ops, keep = [], []
        ctx = None
        if cuda:
            if self.side is None:
                self.side = torch.cuda.Stream()
            self.side.wait_stream(torch.cuda.current_stream())          # the producers of `local` ran on the current stream
            ctx = torch.cuda.stream(self.side)
            ctx.__enter__()
        try:
            for k in sorted(local.keys()):
                v = local[k]
                if v.shape[0] != self.counts[self.rank]:
                    raise ValueError('%s: expected %d local clips, got %d' % (k, self.counts[self.rank], v.shape[0]))
                v = v.contiguous()
                keep.append(v)
                if self.rank == self.dst:
                    buf = self._buffer(k, v)
                    a = self.starts[self.rank]
                    buf[a:a + v.shape[0]].copy_(v, non_blocking=True)
                    for r in range(self.world):
                        if r != self.dst and self.counts[r] > 0:
                            ops.append(dist.P2POp(dist.irecv, buf[self.starts[r]:self.starts[r] + self.counts[r]], r, self.group))
                elif v.shape[0] > 0:
                    ops.append(dist.P2POp(dist.isend, v, self.dst, self.group))
            reqs = dist.batch_isend_irecv(ops) if ops else []
        finally:
            if ctx is not None:
                ctx.__exit__(None, None, None)
        self.pending.append((reqs, keep))

    def wait(self):
        """Join every transfer started so far.  Returns {key: [num_clips, ...]} on dst (derived keys filled in), None elsewhere."""
        for reqs, _ in self.pending:
            for r in reqs:
                r.wait()
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        self.pending = []
        if self.rank != self.dst:
            return None
        out = dict(self.full)
        for k, (src, a, b) in DERIVED.items():
            if k not in out and src in out:
                out[k] = out[src][..., a:b]
        if 'cams_delta' not in out and 'omegas' in out and 'omegas_delta' in out:
            D = out['omegas_delta'].shape[2]
            out['cams_delta'] = out['omegas'][:, :, None, 0:3].expand(-1, -1, D, -1)      # set_cams, tester.py:210-213
        return out


TRANSFER_KEYS_ALL = ('omegas', 'verts', 'kps', 'joints', 'poses', 'omegas_delta', 'verts_delta', 'kps_delta', 'joints_delta', 'poses_delta')
