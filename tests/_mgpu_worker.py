"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one process per GPU).

Every rank runs its contiguous shard of clips; rank 0 additionally runs ALL clips alone.  Pure data parallelism must be
invisible: the gathered N-GPU result has to equal the 1-GPU result bit for bit, for every
fetch key of tester.py:217-255.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    from human_dynamics_b200 import synthetic, HMMRConfig
    from human_dynamics_b200.engine import HMMREngine
    from human_dynamics_b200.dist import shard_range, gather_outputs, OutputGatherer, TRANSFER_KEYS_ALL
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    size = int(os.environ.get('HD_MGPU_SIZE', '224'))
    clips, T = int(os.environ.get('HD_MGPU_CLIPS', '5')), 20            # 5 clips over 2 ranks: ragged shards (3 + 2)
    w = synthetic.make_synthetic_weights(seed=1)
    smpl = synthetic.make_synthetic_smpl(seed=2)
    img = synthetic.make_images(clips * T, seed=77, size=size).reshape(clips, T, size, size, 3)     # same on every rank
    lo, hi = shard_range(clips, rank, world)
    eng = HMMREngine(w, smpl, HMMRConfig(batch_size=hi - lo, sequence_length=T, img_size=size))
    out = eng.predict(torch.from_numpy(img[lo:hi]).cuda())
    keys = [k for k in out if not k.startswith('_')]
    local_out = {k: out[k].contiguous() for k in keys}
    gathered = gather_outputs(local_out, clips, dst=0)
    # the overlapped point-to-point gatherer bench.py uses: dt=0 keys start while the delta heads run; 10 keys travel, 4 are derived
    g = OutputGatherer(clips, dst=0)
    out2 = eng.predict(torch.from_numpy(img[lo:hi]).cuda(),
                       on_main_ready=lambda o: g.start({k: o[k] for k in TRANSFER_KEYS_ALL if k in o}))
    g.start({k: out2[k] for k in TRANSFER_KEYS_ALL if k.endswith('_delta')})
    gathered2 = g.wait()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        assert gathered is not None and len(keys) == 14
        eng1 = HMMREngine(w, smpl, HMMRConfig(batch_size=clips, sequence_length=T, img_size=size))
        full = eng1.predict(torch.from_numpy(img).cuda())
        torch.cuda.synchronize()
        for k in keys:
            same = torch.equal(gathered[k], full[k].contiguous())
            same2 = torch.equal(gathered2[k].contiguous(), full[k].contiguous())
            print('%-14s %s %s / p2p %s' % (k, tuple(gathered[k].shape), 'bit-identical' if same else 'DIFFERS', 'bit-identical' if same2 else 'DIFFERS'))
            ok = ok and same and same2
    else:
        assert gathered is None
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        print('MULTI-GPU MISMATCH on rank', rank)
        return 1
    if rank == 0:
        print('multi-gpu ok: %d ranks, %d clips, %d keys bit-identical to the 1-GPU run' % (world, clips, len(keys)))
    return 0


if __name__ == '__main__':
    sys.exit(main())
