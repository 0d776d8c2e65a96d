"""Runs the ResNet root conv1 (7x7/2 from the padded fp16 planes, 160 frames) a few times -- target of an ncu capture."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                     # noqa: E402
from human_dynamics_b200 import synthetic                         # noqa: E402
from human_dynamics_b200._lib import lib, check                   # noqa: E402
from human_dynamics_b200.nets import PackedResNet, ResNetPlan     # noqa: E402

dev = torch.device('cuda')
w = synthetic.make_resnet_weights(seed=1)
packed = PackedResNet(w, dev, tc='auto')
nxt = packed.units[1]
plan = ResNetPlan(packed, 160, 224, 'auto', units=(0, 1), root=True, tail=False, next_pre=nxt['pre'], next_has_shortcut=False)
img = torch.from_numpy(synthetic.make_images(160, seed=3)).to(dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
check(lib.hd_pack_conv1_planes(C.c_void_p(img.data_ptr()), C.c_void_p(plan.planes[0].data_ptr()), C.c_void_p(plan.planes[1].data_ptr()),
                               160, 224, 224, plan.planes[0].shape[2], st), 'pack')
for _ in range(4):
    plan.conv1_op.run(st)
torch.cuda.synchronize()
