"""Profiling driver: N full HMMR steps (B=32, T=20) on device-resident input.  333 launches per step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_dynamics_b200 import synthetic, HMMRConfig
from human_dynamics_b200.engine import HMMREngine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, T = 32, 20
w = synthetic.make_synthetic_weights(seed=1)
smpl = synthetic.make_synthetic_smpl(seed=2)
eng = HMMREngine(w, smpl, HMMRConfig(batch_size=B, sequence_length=T))
img = torch.from_numpy(synthetic.make_images(B * T, seed=0)).cuda().view(B, T, 224, 224, 3)
for _ in range(steps):
    eng.predict(img)
torch.cuda.synchronize()
