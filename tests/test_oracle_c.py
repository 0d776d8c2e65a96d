"""CPU: the plain-C SMPL oracle (oracle/smpl_ref.c) against the independent numpy oracle (oracle/smpl_ref.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib(kind):
    path = os.path.join(ROOT, 'oracle', 'liboracle_smpl_%s.so' % kind)
    if not os.path.exists(path):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')])
    return C.CDLL(path)


@pytest.mark.parametrize('kind,dtype,tol', [('f64', np.float64, 1e-12), ('f32', np.float32, 2e-5)])
def test_c_oracle_matches_numpy_oracle(smpl_model, kind, dtype, tol):
    from oracle.smpl_ref import SMPLRef, batch_orth_proj_idrot
    from human_dynamics_b200 import synthetic
    lib = _lib(kind)
    ref = SMPLRef(smpl_model, dtype=dtype)
    N, V, K = 3, ref.size[0], ref.joint_regressor.shape[1]
    beta, theta = synthetic.make_smpl_inputs(N, seed=4)
    theta[0] = 0
    cam = np.array([[0.9, 0.1, -0.2]] * N, dtype)

    def arr(a, dt=dtype):
        return np.ascontiguousarray(a, dtype=dt)
    args = [arr(ref.v_template.reshape(-1)), arr(ref.shapedirs), arr(ref.posedirs), arr(ref.J_regressor), arr(ref.weights),
            arr(ref.joint_regressor), arr(np.maximum(ref.parents, 0), np.int32), arr(beta), arr(theta)]
    verts = np.zeros((N, V, 3), dtype); joints = np.zeros((N, K, 3), dtype)
    Rs = np.zeros((N, 24, 9), dtype); Jtr = np.zeros((N, 24, 3), dtype)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.oracle_smpl_forward(N, V, K, *[ptr(a) for a in args], ptr(verts), ptr(joints), ptr(Rs), ptr(Jtr))
    v, j, R = ref(beta, theta, get_skin=True)
    scale = lambda a: max(np.abs(a).max(), 1e-12)
    assert np.abs(verts - v).max() / scale(v) < tol
    assert np.abs(joints - j).max() / scale(j) < tol
    assert np.abs(Rs.reshape(N, 24, 3, 3) - R).max() < tol * 10
    assert np.abs(Jtr - ref.J_transformed).max() / scale(ref.J_transformed) < tol
    kps = np.zeros((N, K, 2), dtype)
    lib.oracle_orth_proj(N, K, ptr(joints), ptr(arr(cam)), ptr(kps))
    assert np.abs(kps - batch_orth_proj_idrot(j, cam, dtype)).max() < tol * 10
    if kind == 'f32':
        assert np.array_equal(Rs[0].reshape(24, 3, 3), np.tile(np.eye(3, dtype=np.float32), (24, 1, 1)))   # theta = 0 => R = I exactly
