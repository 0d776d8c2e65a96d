"""Profiling driver: two ResNet passes over one frame chunk (first = warm-up).  55 launches per pass.
Usage: python tools/prof_resnet.py [chunk] [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_dynamics_b200 import synthetic
from human_dynamics_b200.nets import PackedResNet, ResNetPlan
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = sys.argv[2] if len(sys.argv) > 2 else 'tc3'
w = synthetic.make_resnet_weights(seed=1)
dev = torch.device('cuda')
plan = ResNetPlan(PackedResNet(w, dev, tc=(mode if mode != 'simt' else False)), chunk, 224, mode)
x = torch.from_numpy(synthetic.make_images(chunk, seed=0)).to(dev)
phi = torch.empty((chunk, 2048), device=dev)
for _ in range(2):
    plan.run(x, phi)
torch.cuda.synchronize()
for i, op in enumerate(plan.ops):
    d = op.d
    print('op %2d: M=%6d K=%5d N=%5d k%dx%d s%d impl=%d' % (i, d.n_img * d.Ho * d.Wo, d.KH * d.KW * d.Cin, d.Cout, d.KH, d.KW, d.stride, d.impl))
