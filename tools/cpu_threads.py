"""How many host threads should the reference arm (torch-CPU port) use on this box?  Sweep and print frames/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_dynamics_b200 import synthetic
from oracle import nets_ref
w = synthetic.make_resnet_weights(seed=1)
img = synthetic.make_images(20, seed=0)
for t in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64, os.cpu_count()]:
    torch.set_num_threads(t)
    nets_ref.encoder_resnet(img[:4], w)
    t0 = time.time(); nets_ref.encoder_resnet(img, w); dt = time.time() - t0
    print('threads %3d  resnet %.1f frames/s  (os.cpu_count=%d)' % (t, 20 / dt, os.cpu_count()), flush=True)
