"""Generates tests/golden/hmmr_golden_v1.npz from the CPU oracle on seeded synthetic inputs.

REGRESSION PIN OF THE ORACLE (the reference-derived vectors are ref_exec_v1.npz, made by make_ref_exec_golden.py): these vectors
come from the oracle restatement (float64 variant), not from the reference's TF1
graph -- TensorFlow 1.8 cannot be imported in this environment.  They freeze the oracle so that a
change to it is noticed, and give the GPU tests a fixture that needs no oracle run.  Inputs are NOT stored: they are
regenerated from seeds by human_dynamics_b200.synthetic (weights seed 1, SMPL seed 2, images seed 11, SMPL inputs seed 12).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from human_dynamics_b200 import synthetic          # noqa: E402
from oracle import nets_ref, smpl_ref              # noqa: E402

VERT_IDS = np.arange(0, 6890, 53)                   # 130 sampled vertices


def build():
    w = synthetic.make_synthetic_weights(seed=1)
    smpl = synthetic.make_synthetic_smpl(seed=2)
    out = {}
    # SMPL-only (config 1 style + posed)
    beta, theta = synthetic.make_smpl_inputs(5, seed=12)
    theta[0] = 0
    cam = np.tile(np.array([[0.9, 0.1, -0.2]], np.float32), (5, 1))
    s = smpl_ref.SMPLRef(smpl, dtype=np.float64)
    v, j, Rs = s(beta, theta, get_skin=True)
    out['smpl_verts'] = v[:, VERT_IDS]
    out['smpl_joints'] = j
    out['smpl_Rs'] = Rs
    out['smpl_Jtr'] = s.J_transformed
    out['smpl_kps'] = smpl_ref.batch_orth_proj_idrot(j, cam, np.float64)
    # full window, B=1, T=4, 64x64 frames
    B, T, S = 1, 4, 64
    img = synthetic.make_images(B * T, seed=11, size=S).reshape(B, T, S, S, 3)
    r = nets_ref.hmmr_predict(img, w, smpl, dtype=torch.float64)
    out['hmmr_phi'] = r['_phi']
    out['hmmr_movie_strips'] = r['_movie_strips']
    for k in ('omegas', 'kps', 'joints', 'omegas_delta', 'kps_delta'):
        out['hmmr_' + k] = r[k]
    out['hmmr_verts'] = r['verts'][:, :, VERT_IDS]
    out['hmmr_verts_delta'] = r['verts_delta'][:, :, :, VERT_IDS]
    out['vert_ids'] = VERT_IDS
    return {k: np.asarray(v, np.float64 if v.dtype.kind == 'f' else v.dtype) for k, v in out.items()}


if __name__ == '__main__':
    g = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hmmr_golden_v1.npz')
    np.savez_compressed(path, **g)
    print('wrote', path, os.path.getsize(path), 'bytes')
