"""Prints max |err| / max |ref| of the CUDA path against the reference-source vectors (tests/golden/ref_exec_v1.npz), per Tester key."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_dynamics_b200 import synthetic, HMMRConfig          # noqa: E402
from src.evaluation.tester import Tester                       # noqa: E402

g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'ref_exec_v1.npz')))
ids = g['vert_ids']
w = synthetic.make_synthetic_weights(seed=1)
smpl = synthetic.make_synthetic_smpl(seed=2)
t = Tester(HMMRConfig(batch_size=2, sequence_length=20, weights=w, smpl_model=smpl, pred_mode='pred'))
r = t.predict(synthetic.make_images(40, seed=21, size=224).reshape(2, 20, 224, 224, 3))
for k in sorted(r):
    a = np.asarray(r[k], np.float64)
    if k == 'verts':
        a = a[:, :, ids]
    if k == 'verts_delta':
        a = a[:, :, :, ids]
    b = g['tester_' + k].astype(np.float64)
    print('%-14s %.2e' % (k, np.abs(a - b).max() / np.abs(b).max()))
