"""Golden fixtures (tests/golden/hmmr_golden_v1.npz, made by tests/golden/make_golden.py from the float64 oracle).

CPU: the float32 oracle reproduces them (regression pin of the oracle).  GPU: the CUDA path reproduces them
through the C-ABI without running the oracle at all."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hmmr_golden_v1.npz')


def rel_err(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.fixture(scope='module')
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def _inputs():
    from human_dynamics_b200 import synthetic
    beta, theta = synthetic.make_smpl_inputs(5, seed=12)
    theta[0] = 0
    cam = np.tile(np.array([[0.9, 0.1, -0.2]], np.float32), (5, 1))
    img = synthetic.make_images(4, seed=11, size=64).reshape(1, 4, 64, 64, 3)
    return beta, theta, cam, img


def test_oracle_f32_reproduces_golden(gold, weights, smpl_model):
    from oracle import nets_ref, smpl_ref
    beta, theta, cam, img = _inputs()
    ids = gold['vert_ids']
    s = smpl_ref.SMPLRef(smpl_model)
    v, j, Rs = s(beta, theta, get_skin=True)
    assert rel_err(v[:, ids], gold['smpl_verts']) < 1e-5
    assert rel_err(j, gold['smpl_joints']) < 1e-5 and rel_err(Rs, gold['smpl_Rs']) < 1e-5
    assert rel_err(s.J_transformed, gold['smpl_Jtr']) < 1e-5
    assert rel_err(smpl_ref.batch_orth_proj_idrot(j, cam), gold['smpl_kps']) < 1e-5
    w = {k: v for k, v in weights.items() if not k.startswith('fc2_res')}
    r = nets_ref.hmmr_predict(img, w, smpl_model)
    assert rel_err(r['_phi'], gold['hmmr_phi']) < 1e-5
    assert rel_err(r['_movie_strips'], gold['hmmr_movie_strips']) < 1e-5
    for k in ('omegas', 'kps', 'joints', 'omegas_delta', 'kps_delta'):
        assert rel_err(r[k], gold['hmmr_' + k]) < 3e-5, k
    assert rel_err(r['verts'][:, :, ids], gold['hmmr_verts']) < 3e-5
    assert rel_err(r['verts_delta'][:, :, :, ids], gold['hmmr_verts_delta']) < 5e-5


@pytest.mark.gpu
def test_cuda_path_reproduces_golden(gold, weights, smpl_model):
    from human_dynamics_b200 import HMMRConfig
    from human_dynamics_b200.engine import HMMREngine
    beta, theta, cam, img = _inputs()
    ids = torch.from_numpy(gold['vert_ids']).cuda()
    eng = HMMREngine(weights, smpl_model, HMMRConfig(batch_size=1, sequence_length=4, img_size=64))
    o = eng.smpl.forward(torch.from_numpy(beta).cuda(), torch.from_numpy(theta).cuda(), cam=torch.from_numpy(cam).cuda())
    assert rel_err(o['verts'][:, ids].cpu().numpy(), gold['smpl_verts']) < 1e-4
    assert rel_err(o['joints'].cpu().numpy(), gold['smpl_joints']) < 1e-4
    assert rel_err(o['kps'].cpu().numpy(), gold['smpl_kps']) < 1e-4
    assert np.array_equal(o['Rs'][0].cpu().numpy(), np.tile(np.eye(3, dtype=np.float32), (24, 1, 1)))
    r = eng.predict(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()
    assert rel_err(r['_phi'].cpu().numpy(), gold['hmmr_phi']) < 1e-4
    for k in ('omegas', 'kps', 'joints', 'omegas_delta', 'kps_delta'):
        assert rel_err(r[k].cpu().numpy(), gold['hmmr_' + k]) < 1e-4, k
    assert rel_err(r['verts'][:, :, ids].cpu().numpy(), gold['hmmr_verts']) < 1e-4
    assert rel_err(r['verts_delta'][:, :, :, ids].cpu().numpy(), gold['hmmr_verts_delta']) < 1e-4
