"""Device-resident C3 step time vs the two trunk stage sizes (frame_chunk for root + blocks 1-2, late_chunk for blocks 3-4).
    python tools/sweep_chunks.py "32,640 64,640 160,640 160,320"        (CUDA-graph replay, 3 warm + 6 timed steps each)"""
import sys

import torch

sys.path.insert(0, '.')
from human_dynamics_b200 import synthetic, HMMRConfig          # noqa: E402
from human_dynamics_b200.engine import HMMREngine              # noqa: E402


def main():
    pairs = [tuple(int(v) for v in p.split(',')) for p in (sys.argv[1] if len(sys.argv) > 1 else '160,640').split()]
    dev = torch.device('cuda', 0)
    w = synthetic.make_synthetic_weights(seed=1)
    smpl = synthetic.make_synthetic_smpl(seed=2)
    B, T = 32, 20
    img = torch.from_numpy(synthetic.make_images(B * T, seed=100)).view(B, T, 224, 224, 3).to(dev)
    ref = None
    for fc, lc in pairs:
        eng = HMMREngine(w, smpl, HMMRConfig(batch_size=B, sequence_length=T, frame_chunk=fc, late_chunk=lc), device=dev)
        for _ in range(3):
            out, nodes = eng.predict_graphed(img)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            out, nodes = eng.predict_graphed(img)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        v = out['verts'].clone()
        same = '' if ref is None else ('  verts identical to first config: %s' % bool(torch.equal(v, ref)))
        ref = v if ref is None else ref
        print('frame_chunk %4d late_chunk %4d : %7.3f ms/step  %8.0f frames/s  (%d graph nodes)%s' % (fc, lc, ms, B * T / ms * 1e3, nodes, same), flush=True)
        del eng
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
