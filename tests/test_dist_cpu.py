"""CPU: the N>1 host logic (clip sharding + final gather) on the gloo backend, world_size 2 and 3 (ragged)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from human_dynamics_b200.dist import OutputGatherer, gather_outputs, shard_counts, shard_range


def test_shard_range_partitions_clips():
    for clips, world in ((256, 8), (7, 3), (2, 4), (32, 1)):
        seen = []
        for r in range(world):
            a, b = shard_range(clips, r, world)
            seen.extend(range(a, b))
        assert seen == list(range(clips))
        assert sum(shard_counts(clips, world)) == clips
    assert shard_counts(256, 8) == [32] * 8               # BASELINE config 4: 32 clips per GPU = config 3 per rank


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, clips, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    a, b = shard_range(clips, rank, world)
    full = {'omegas': torch.arange(clips * 20 * 85, dtype=torch.float32).reshape(clips, 20, 85),
            'kps': torch.arange(clips * 20 * 50, dtype=torch.float32).reshape(clips, 20, 25, 2) * 0.5}
    local = {k: v[a:b].clone() for k, v in full.items()}
    got = gather_outputs(local, clips, dst=0)
    ok = True
    if rank == 0:
        ok = all(torch.equal(got[k], full[k]) for k in full)       # bit-identical, clip order preserved
    else:
        ok = got is None
    # point-to-point gatherer: two start() calls (dt=0 outputs first, deltas later), one wait; cams / shapes derived on dst
    full['omegas_delta'] = torch.arange(clips * 20 * 2 * 85, dtype=torch.float32).reshape(clips, 20, 2, 85) + 0.25
    g = OutputGatherer(clips, dst=0)
    g.start({k: full[k][a:b].clone() for k in ('omegas', 'kps')})
    g.start({'omegas_delta': full['omegas_delta'][a:b].clone()})
    got2 = g.wait()
    if rank == 0:
        ok = ok and all(torch.equal(got2[k], full[k]) for k in ('omegas', 'kps', 'omegas_delta'))
        ok = ok and torch.equal(got2['cams'], full['omegas'][..., 0:3]) and torch.equal(got2['shapes_delta'], full['omegas_delta'][..., 75:85])
        ok = ok and torch.equal(got2['cams_delta'][:, :, 1], full['omegas'][..., 0:3])
    else:
        ok = ok and got2 is None
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize('world,clips', [(2, 8), (3, 7)])
def test_gather_outputs_gloo(world, clips):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_single_process_gather_is_identity():
    x = {'verts': torch.zeros(4, 2, 3)}
    assert gather_outputs(x, 4)['verts'] is x['verts']
