"""Per-stage CUDA-event timing of the staged SMPL path at C5 size (65536 poses): pose / blend GEMM / LBS / joints."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from human_dynamics_b200 import synthetic
from human_dynamics_b200.smpl import SMPLConstants
from human_dynamics_b200._lib import lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
smpl = synthetic.make_synthetic_smpl(seed=2)
c = SMPLConstants(smpl)
beta, theta = synthetic.make_smpl_inputs(N, seed=0)
b, t = torch.from_numpy(beta).cuda(), torch.from_numpy(theta).cuda()
cam = torch.ones((N, 3), device='cuda')
out = c.forward(b, t, cam=cam)
torch.cuda.synchronize()
names = []
orig = {}
evs = []
for fn in ('hd_smpl_pose', 'hd_conv_gemm', 'hd_smpl_lbs', 'hd_smpl_joints', 'hd_smpl_forward'):
    f = getattr(lib, fn)
    orig[fn] = f
    def wrap(*a, _f=f, _n=fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = _f(*a); e1.record()
        evs.append((_n, e0, e1))
        return r
    setattr(lib, fn, wrap)
import human_dynamics_b200.smpl as S, human_dynamics_b200.nets as Nn
for rep in range(3):
    evs.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c.forward(b, t, cam=cam, out=out); e1.record()
    torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
print('N=%d total %.3f ms  -> %.2f M poses/s, %.1f GB/s algorithmic (84384 B/pose)' % (N, tot, N / tot / 1e3, N * 84384 / tot / 1e6))
for n, a, bb in evs:
    print('  %-18s %.3f ms' % (n, a.elapsed_time(bb)))
