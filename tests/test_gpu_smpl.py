"""GPU parity: CUDA SMPL path (through the C-ABI) vs the CPU oracle restatement of src/tf_smpl/*."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-4      # north_star tolerance: 1e-4 relative FP32


def rel_err(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


def _run(model, beta, theta, cam=None, joint_type='cocoplus', tc=True, lbs_tc_min=None):
    from human_dynamics_b200.smpl import SMPLConstants
    c = SMPLConstants(model, joint_type=joint_type, tc=tc)
    if lbs_tc_min is not None:
        c.lbs_tc_min_batch = lbs_tc_min
    o = c.forward(torch.from_numpy(beta).cuda(), torch.from_numpy(theta).cuda(),
                  cam=None if cam is None else torch.from_numpy(cam).cuda())
    torch.cuda.synchronize()
    return {k: (None if v is None else v.cpu().numpy()) for k, v in o.items()}


def _oracle(model, beta, theta, cam=None, joint_type='cocoplus', dtype=np.float32):
    from oracle.smpl_ref import SMPLRef, batch_orth_proj_idrot
    s = SMPLRef(model, joint_type=joint_type, dtype=dtype)
    verts, joints, Rs = s(beta, theta, get_skin=True)
    out = {'verts': verts, 'joints': joints, 'Rs': Rs, 'Jtr': s.J_transformed}
    if cam is not None:
        out['kps'] = batch_orth_proj_idrot(joints, cam, dtype)
    return out


@pytest.mark.parametrize('n,zero_pose', [(4, True), (1, False), (37, False), (300, False)])
def test_smpl_forward_matches_oracle(smpl_model, n, zero_pose):
    from human_dynamics_b200 import synthetic
    beta, theta = synthetic.make_smpl_inputs(n, seed=n, zero_pose=zero_pose)
    cam = np.random.RandomState(n).uniform(0.5, 1.5, size=(n, 3)).astype(np.float32)
    got = _run(smpl_model, beta, theta, cam)
    ref = _oracle(smpl_model, beta, theta, cam)
    ref64 = _oracle(smpl_model, beta, theta, cam, dtype=np.float64)
    for k in ('verts', 'joints', 'Rs', 'Jtr', 'kps'):
        assert rel_err(ref[k], ref64[k]) < 1e-5, 'oracle f32 vs f64 ill-conditioned on ' + k
        assert rel_err(got[k], ref[k]) < REL, k
    if zero_pose:       # C1: theta = 0 => Rs = I exactly, verts = v_shaped
        assert np.array_equal(got['Rs'], np.tile(np.eye(3, dtype=np.float32), (n, 24, 1, 1)))


@pytest.mark.parametrize('lbs_tc', [False, True])
@pytest.mark.parametrize('n', [256, 777])
def test_smpl_tensor_core_blend_path(smpl_model, smpl_model_dense, n, lbs_tc):
    """N >= 256 takes the staged path (pose -> tcgen05 blend GEMM -> skinning -> keypoints); must agree with the oracle
    and, to rounding, with the fused FP32 kernel.  lbs_tc: skinning on the tensor cores (smpl_lbs_tc.cu; by default only for
    batches that fill the chip) vs the CUDA-core skinning kernel."""
    from human_dynamics_b200 import synthetic
    for model, jt in ((smpl_model, 'cocoplus'), (smpl_model_dense, 'lsp')):
        beta, theta = synthetic.make_smpl_inputs(n, seed=n)
        cam = np.random.RandomState(n).uniform(0.5, 1.5, size=(n, 3)).astype(np.float32)
        got = _run(model, beta, theta, cam, joint_type=jt, lbs_tc_min=0 if lbs_tc else 1 << 30)
        ref = _oracle(model, beta, theta, cam, joint_type=jt)
        fused = _run(model, beta, theta, cam, joint_type=jt, tc=False)
        for k in ('verts', 'joints', 'Rs', 'Jtr', 'kps'):
            assert rel_err(got[k], ref[k]) < REL, k
            assert rel_err(got[k], fused[k]) < 2e-5, k


def test_smpl_dense_weights_lsp(smpl_model_dense):
    from human_dynamics_b200 import synthetic
    beta, theta = synthetic.make_smpl_inputs(9, seed=5)
    got = _run(smpl_model_dense, beta, theta, joint_type='lsp')
    ref = _oracle(smpl_model_dense, beta, theta, joint_type='lsp')
    assert got['joints'].shape == (9, 14, 3)
    for k in ('verts', 'joints', 'Rs', 'Jtr'):
        assert rel_err(got[k], ref[k]) < REL, k


def test_smpl_strided_omega_and_slots(smpl_model):
    """beta/theta/cam as column views of one [N,85] omega buffer + interleaved output slots."""
    from human_dynamics_b200.smpl import SMPLConstants
    n, D = 21, 2
    rng = np.random.RandomState(0)
    omega = rng.normal(0, 0.3, size=(n, 85)).astype(np.float32)
    om = torch.from_numpy(omega).cuda()
    c = SMPLConstants(smpl_model)
    V, K = c.num_verts, c.num_kps
    out = {'verts': torch.zeros((n * D, V, 3), device='cuda'), 'joints': torch.zeros((n * D, K, 3), device='cuda'),
           'Rs': torch.zeros((n * D, 24, 3, 3), device='cuda'), 'Jtr': torch.zeros((n * D, 24, 3), device='cuda'),
           'kps': torch.zeros((n * D, K, 2), device='cuda')}
    c.forward(om[:, 75:85], om[:, 3:75], cam=om[:, 0:3], out=out, slot=(D, 1))
    torch.cuda.synchronize()
    ref = _oracle(smpl_model, omega[:, 75:85], omega[:, 3:75], omega[:, :3])
    v = out['verts'].cpu().numpy().reshape(n, D, V, 3)
    assert rel_err(v[:, 1], ref['verts']) < REL
    assert np.all(v[:, 0] == 0)
    assert rel_err(out['kps'].cpu().numpy().reshape(n, D, K, 2)[:, 1], ref['kps']) < REL


def test_reference_named_surface(smpl_model):
    """src.tf_smpl.* entry points with the reference's names and return structure."""
    from src.tf_smpl.batch_smpl import SMPL
    from src.tf_smpl.batch_lbs import batch_rodrigues, batch_global_rigid_transformation
    from src.tf_smpl.projection import batch_orth_proj_idrot
    from src.ops import batch_orth_proj_idrot as proj_alias
    from oracle import smpl_ref
    from human_dynamics_b200 import synthetic
    beta, theta = synthetic.make_smpl_inputs(6, seed=11)
    smpl = SMPL(smpl_model)
    joints = smpl(torch.from_numpy(beta).cuda(), torch.from_numpy(theta).cuda())
    verts, joints2, Rs = smpl(torch.from_numpy(beta).cuda(), torch.from_numpy(theta.reshape(6, 24, 3)).cuda(), get_skin=True)
    ref = _oracle(smpl_model, beta, theta)
    assert rel_err(joints.cpu().numpy(), ref['joints']) < REL
    assert rel_err(verts.cpu().numpy(), ref['verts']) < REL
    assert rel_err(smpl.J_transformed.cpu().numpy(), ref['Jtr']) < REL
    R = batch_rodrigues(torch.from_numpy(theta.reshape(-1, 3)).cuda())
    assert rel_err(R.cpu().numpy(), smpl_ref.batch_rodrigues(theta.reshape(-1, 3))) < REL
    Js = np.random.RandomState(1).normal(0, 0.3, size=(6, 24, 3)).astype(np.float32)
    for rb in (False, True):
        nj, A = batch_global_rigid_transformation(Rs, torch.from_numpy(Js).cuda(), smpl.parents, rotate_base=rb)
        nj_ref, A_ref = smpl_ref.batch_global_rigid_transformation(ref['Rs'], Js, smpl.parents, rotate_base=rb)
        assert rel_err(nj.cpu().numpy(), nj_ref) < REL and rel_err(A.cpu().numpy(), A_ref) < REL
    cam = np.random.RandomState(2).uniform(0.5, 1.5, size=(6, 3)).astype(np.float32)
    k = batch_orth_proj_idrot(joints2, torch.from_numpy(cam).cuda())
    assert proj_alias is batch_orth_proj_idrot
    assert rel_err(k.cpu().numpy(), smpl_ref.batch_orth_proj_idrot(ref['joints'], cam)) < REL
    with pytest.raises(RuntimeError):
        batch_rodrigues(torch.zeros(3, 3))          # CPU tensor is an error: no fallback


def test_smpl_large_batch_properties(smpl_model):
    """Size-independent checks at C5 scale (N=65536 would not fit the oracle's time budget): root-rotation
    equivariance v' = R0 (v - J0) + J0 and joints = verts . regressor."""
    from human_dynamics_b200.smpl import SMPLConstants
    from human_dynamics_b200 import synthetic
    from oracle import smpl_ref
    n = 8192
    beta, theta = synthetic.make_smpl_inputs(n, seed=3)
    c = SMPLConstants(smpl_model)
    b, t = torch.from_numpy(beta).cuda(), torch.from_numpy(theta).cuda()
    o1 = c.forward(b, t)
    v1, j1 = o1['verts'].clone(), o1['joints'].clone()
    t0 = t.clone(); t0[:, :3] = 0
    o0 = c.forward(b, t0)
    torch.cuda.synchronize()
    R0 = torch.from_numpy(smpl_ref.batch_rodrigues(theta[:, :3])).cuda()
    J0 = o0['Jtr'][:, 0:1]
    v_expect = torch.einsum('nij,nvj->nvi', R0, o0['verts'] - J0) + J0
    assert float((v_expect - v1).abs().max()) < 2e-5
    reg = torch.from_numpy(smpl_model['cocoplus_regressor'].astype(np.float32)).cuda()       # (K,V)
    j_expect = torch.einsum('kv,nvc->nkc', reg, v1)
    assert float((j_expect - j1).abs().max()) < 2e-5
    # sampled rows against the oracle
    idx = np.arange(0, n, 1024)
    ref = _oracle(smpl_model, beta[idx], theta[idx])
    assert rel_err(v1[idx].cpu().numpy(), ref['verts']) < REL


def test_batch_rot2aa_matches_oracle():
    """src.tf_smpl.batch_lbs.batch_rot2aa (batch_lbs.py:63-105) on the GPU vs the numpy restatement, incl. identity rows."""
    from src.tf_smpl.batch_lbs import batch_rot2aa, batch_rodrigues
    from oracle import smpl_ref
    rng = np.random.RandomState(2)
    th = rng.normal(0, 0.8, size=(500, 3)).astype(np.float32)
    th[::50] = 0.0
    R = batch_rodrigues(torch.from_numpy(th).cuda())
    aa = batch_rot2aa(R).cpu().numpy()
    ref = smpl_ref.batch_rot2aa(R.cpu().numpy().astype(np.float64), np.float64)
    assert np.abs(aa - ref).max() < 5e-4          # acos near 1 amplifies fp32 rounding of the trace
    nrm = np.linalg.norm(th, axis=1)
    big = (nrm > 0.3) & (nrm < 3.0)                  # beyond pi the axis-angle vector of the same rotation is a different one
    assert np.abs(aa[big] - th[big]).max() < 2e-5
    assert np.all(aa[::50] == 0.0)


def test_smpl_tensor_core_skinning_into_delta_slots(smpl_model):
    """smpl_lbs_tc with the [B,T,D,...] in-place stacking (pose n -> slot n*D + d, odd slots are only 8-byte aligned) and strided
    omega views, ragged last batch / last vertex tile."""
    from human_dynamics_b200.smpl import SMPLConstants
    n, D = 300, 3
    rng = np.random.RandomState(1)
    omega = rng.normal(0, 0.3, size=(n, 85)).astype(np.float32)
    om = torch.from_numpy(omega).cuda()
    c = SMPLConstants(smpl_model)
    c.lbs_tc_min_batch = 0
    V, K = c.num_verts, c.num_kps
    out = {'verts': torch.zeros((n * D, V, 3), device='cuda'), 'joints': torch.zeros((n * D, K, 3), device='cuda'),
           'Rs': torch.zeros((n * D, 24, 3, 3), device='cuda'), 'Jtr': torch.zeros((n * D, 24, 3), device='cuda'),
           'kps': torch.zeros((n * D, K, 2), device='cuda')}
    ref = _oracle(smpl_model, omega[:, 75:85], omega[:, 3:75], omega[:, :3])
    for d in (1, 2):
        c.forward(om[:, 75:85], om[:, 3:75], cam=om[:, 0:3], out=out, slot=(D, d))
    torch.cuda.synchronize()
    v = out['verts'].cpu().numpy().reshape(n, D, V, 3)
    assert np.all(v[:, 0] == 0)
    for d in (1, 2):
        assert rel_err(v[:, d], ref['verts']) < REL
    assert np.array_equal(v[:, 1], v[:, 2])
