"""Diagnostics (GPU): per-stage relative errors of the SIMT and tcgen05 paths against the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from human_dynamics_b200 import synthetic, HMMRConfig
from human_dynamics_b200.engine import HMMREngine
from human_dynamics_b200.nets import PackedConv
from oracle import nets_ref


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())


def fc_err(K, impl, M=512, N=256, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.normal(0, 1, size=(M, K)).astype(np.float32)
    w = (rng.normal(0, 1, size=(K, N)) / np.sqrt(K)).astype(np.float32)
    pc = PackedConv(w, torch.device('cuda'), tc=(impl if impl != 'simt' else False))
    out = torch.empty((M, N), device='cuda')
    pc.bind(torch.from_numpy(x).cuda(), M, 1, 1, out, impl=impl).run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = x.astype(np.float64) @ w.astype(np.float64)
    return rel(out.cpu().numpy(), ref)


for K in (64, 256, 1024, 2304, 4608, 6144, 16384):
    print('K=%5d  simt %.2e  tc3 %.2e  tc3h %.2e  tc1 %.2e' % (K, fc_err(K, 'simt'), fc_err(K, 'tc3'), fc_err(K, 'tc3h'), fc_err(K, 'tc1')), flush=True)

w = synthetic.make_synthetic_weights(seed=1)
smpl = synthetic.make_synthetic_smpl(seed=2)
B, T = 2, 20
img = synthetic.make_images(B * T, seed=0).reshape(B, T, 224, 224, 3)
ref64 = nets_ref.hmmr_predict(img, w, smpl, dtype=torch.float64)
ref32 = nets_ref.hmmr_predict(img, w, smpl)
keys = ['_phi', '_movie_strips', 'omegas', 'verts', 'joints', 'kps', 'omegas_delta', 'verts_delta']
print('oracle f32 vs f64: ' + '  '.join('%s %.1e' % (k, rel(ref32[k], ref64[k])) for k in keys))
for impl in ('simt', 'tc3', 'tc3h'):
    eng = HMMREngine(w, smpl, HMMRConfig(batch_size=B, sequence_length=T, frame_chunk=16), impl=impl)
    got = eng.predict(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()
    print('%-5s vs f64: ' % impl + '  '.join('%s %.1e' % (k, rel(got[k].cpu().numpy(), ref64[k])) for k in keys))
    print('%-5s vs f32: ' % impl + '  '.join('%s %.1e' % (k, rel(got[k].cpu().numpy(), ref32[k])) for k in keys), flush=True)
