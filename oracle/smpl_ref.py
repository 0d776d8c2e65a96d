"""ORACLE (test infrastructure, not product): CPU restatement of the reference SMPL path.

PARITY STATUS: pinned to the reference's SOURCE, unpinned against TensorFlow's KERNELS.  The reference ships
no tests / golden vectors for this path and TF 1.8 cannot be imported here; instead
src/tf_smpl/{batch_lbs,batch_smpl,projection}.py are executed unmodified over a numpy TensorFlow stand-in
(oracle/ref_exec/) and this module is checked against their output (tests/golden/ref_exec_v1.npz,
tests/test_ref_exec.py: verts / joints / Rs / FK / projection within 2e-5, regenerated live where
/root/reference exists).  The SMPL path uses only core tf ops (matmul, reshape, tile, pad, scatter_nd, ...),
no tf.contrib layer.  Every function follows the cited reference lines op by op (same operation order, so
float32 rounding is comparable); `dtype` selects the float64 "truth" or the float32 "TF-faithful" variant.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product path (human_dynamics_b200/, src/) never does.
"""
from __future__ import annotations

import numpy as np


def batch_skew(vec, dtype=np.float32):
    """src/tf_smpl/batch_lbs.py:15-39 -- scatter [-z, y, z, -x, -y, x] into flat idx [1,2,3,5,6,7]."""
    vec = np.asarray(vec, dtype)
    n = vec.shape[0]
    res = np.zeros((n, 9), dtype)
    res[:, 1] = -vec[:, 2]
    res[:, 2] = vec[:, 1]
    res[:, 3] = vec[:, 2]
    res[:, 5] = -vec[:, 0]
    res[:, 6] = -vec[:, 1]
    res[:, 7] = vec[:, 0]
    return res.reshape(n, 3, 3)


def batch_rodrigues(theta, dtype=np.float32):
    """src/tf_smpl/batch_lbs.py:42-60.  theta [M,3] -> R [M,3,3].

    angle = ||theta + 1e-8|| (eps added to every component BEFORE the norm, :48);
    r = theta / angle (un-shifted theta, :49); R = cos*I + (1-cos)*r r^T + sin*skew(r).
    """
    theta = np.asarray(theta, dtype)
    eps = dtype(1e-8)
    shifted = theta + eps
    angle = np.sqrt(np.sum(shifted * shifted, axis=1, dtype=dtype))[:, None]       # tf.norm
    r = (theta / angle)[:, :, None]                                                # [M,3,1]
    angle = angle[:, :, None]
    cos = np.cos(angle).astype(dtype)
    sin = np.sin(angle).astype(dtype)
    outer = np.matmul(r, r.transpose(0, 2, 1))
    eyes = np.tile(np.eye(3, dtype=dtype)[None], (theta.shape[0], 1, 1))
    R = cos * eyes + (dtype(1) - cos) * outer + sin * batch_skew(r[:, :, 0], dtype)
    return R.astype(dtype)


def batch_rot2aa(Rs, dtype=np.float32):
    """src/tf_smpl/batch_lbs.py:63-105."""
    Rs = np.asarray(Rs, dtype)
    cos = 0.5 * (np.trace(Rs, axis1=1, axis2=2) - 1)                                 # :91
    cos = np.clip(cos, -1, 1)                                                        # :92
    theta = np.arccos(cos)                                                           # :94
    m21 = Rs[:, 2, 1] - Rs[:, 1, 2]
    m02 = Rs[:, 0, 2] - Rs[:, 2, 0]
    m10 = Rs[:, 1, 0] - Rs[:, 0, 1]
    denom = np.sqrt(m21 * m21 + m02 * m02 + m10 * m10)
    small = np.abs(theta) < 0.00001
    with np.errstate(divide='ignore', invalid='ignore'):
        axis0 = np.where(small, m21, m21 / denom)                                    # :101-103
        axis1 = np.where(small, m02, m02 / denom)
        axis2 = np.where(small, m10, m10 / denom)
    return (theta[:, None] * np.stack([axis0, axis1, axis2], 1)).astype(dtype)       # :105


def batch_global_rigid_transformation(Rs, Js, parent, rotate_base=False, dtype=np.float32):
    """src/tf_smpl/batch_lbs.py:133-194.  Rs [N,24,3,3], Js [N,24,3] -> new_J [N,24,3], A [N,24,4,4]."""
    Rs = np.asarray(Rs, dtype)
    Js = np.asarray(Js, dtype)
    N = Rs.shape[0]
    nj = len(parent)
    if rotate_base:  # :151-156 (no caller enables it)
        rot_x = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype)
        root_rotation = np.matmul(Rs[:, 0], rot_x)
    else:
        root_rotation = Rs[:, 0]
    Js_e = Js[..., None]                                                            # [N,24,3,1]

    def make_A(R, t):                                                               # :163-168
        R_homo = np.concatenate([R, np.zeros((N, 1, 3), dtype)], axis=1)            # [N,4,3]
        t_homo = np.concatenate([t, np.ones((N, 1, 1), dtype)], axis=1)             # [N,4,1]
        return np.concatenate([R_homo, t_homo], axis=2)

    results = [make_A(root_rotation, Js_e[:, 0])]
    for i in range(1, nj):                                                          # :172-177
        j_here = Js_e[:, i] - Js_e[:, parent[i]]
        A_here = make_A(Rs[:, i], j_here)
        results.append(np.matmul(results[parent[i]], A_here).astype(dtype))
    results = np.stack(results, axis=1)                                             # [N,24,4,4]
    new_J = results[:, :, :3, 3]
    Js_w0 = np.concatenate([Js_e, np.zeros((N, nj, 1, 1), dtype)], axis=2)          # [N,24,4,1]
    init_bone = np.matmul(results, Js_w0).astype(dtype)                             # [N,24,4,1]
    init_bone = np.concatenate([np.zeros((N, nj, 4, 3), dtype), init_bone], axis=3)
    A = results - init_bone
    return new_J.astype(dtype), A.astype(dtype)


class SMPLRef(object):
    """src/tf_smpl/batch_smpl.py:26-162 restated on numpy.

    `model` is the un-pickled dict (dense ndarrays accepted in place of chumpy / scipy-sparse).
    """

    def __init__(self, model: dict, joint_type='cocoplus', dtype=np.float32):
        self.dtype = dtype
        dd = model
        self.v_template = np.asarray(dd['v_template'], dtype)                        # :35
        self.size = [self.v_template.shape[0], 3]
        self.num_betas = dd['shapedirs'].shape[-1]
        self.shapedirs = np.reshape(np.asarray(dd['shapedirs']), [-1, self.num_betas]).T.astype(dtype)  # :45-48
        self.J_regressor = _dense(dd['J_regressor']).T.astype(dtype)                 # (V,24) :51-55
        nb = dd['posedirs'].shape[-1]
        self.posedirs = np.reshape(np.asarray(dd['posedirs']), [-1, nb]).T.astype(dtype)  # (207,V*3) :60-63
        self.parents = np.asarray(dd['kintree_table'])[0].astype(np.int32)           # :66
        self.weights = np.asarray(dd['weights'], dtype)                              # :69-73
        self.joint_regressor = _dense(dd['cocoplus_regressor']).T.astype(dtype)      # (V,K) :76-80
        if joint_type == 'lsp':
            self.joint_regressor = self.joint_regressor[:, :14]                      # :81-82
        if joint_type not in ('cocoplus', 'lsp'):
            raise ValueError('Unknown joint type: %s' % joint_type)
        self.J_transformed = None

    def __call__(self, beta, theta, get_skin=False):
        dt = self.dtype
        beta = np.asarray(beta, dt)
        theta = np.asarray(theta, dt)
        N = beta.shape[0]
        V = self.size[0]
        v_shaped = np.matmul(beta, self.shapedirs).reshape(-1, V, 3) + self.v_template          # :110-112
        Jx = np.matmul(v_shaped[:, :, 0], self.J_regressor)                                      # :115-118
        Jy = np.matmul(v_shaped[:, :, 1], self.J_regressor)
        Jz = np.matmul(v_shaped[:, :, 2], self.J_regressor)
        J = np.stack([Jx, Jy, Jz], axis=2)
        Rs = batch_rodrigues(theta.reshape(-1, 3), dt).reshape(-1, 24, 3, 3)                     # :123-124
        pose_feature = (Rs[:, 1:] - np.eye(3, dtype=dt)).reshape(-1, 207)                        # :127-128
        v_posed = np.matmul(pose_feature, self.posedirs).reshape(-1, V, 3) + v_shaped            # :131-133
        self.J_transformed, A = batch_global_rigid_transformation(Rs, J, self.parents, dtype=dt)  # :136-137
        W = np.tile(self.weights, (N, 1)).reshape(N, -1, 24)                                     # :141-142
        T = np.matmul(W, A.reshape(N, 24, 16)).reshape(N, -1, 4, 4).astype(dt)                   # :144-146
        v_posed_homo = np.concatenate([v_posed, np.ones((N, V, 1), dt)], axis=2)                 # :147-148
        v_homo = np.matmul(T, v_posed_homo[..., None]).astype(dt)                                # :149
        verts = v_homo[:, :, :3, 0]                                                              # :151
        jx = np.matmul(verts[:, :, 0], self.joint_regressor)                                     # :154-157
        jy = np.matmul(verts[:, :, 1], self.joint_regressor)
        jz = np.matmul(verts[:, :, 2], self.joint_regressor)
        joints = np.stack([jx, jy, jz], axis=2).astype(dt)
        if get_skin:
            return verts.astype(dt), joints, Rs
        return joints


def batch_orth_proj_idrot(X, camera, dtype=np.float32):
    """src/tf_smpl/projection.py:16-29: [s*(x+tx), s*(y+ty)]."""
    X = np.asarray(X, dtype)
    camera = np.asarray(camera, dtype).reshape(-1, 1, 3)
    X_trans = X[:, :, :2] + camera[:, :, 1:]
    shape = X_trans.shape
    return (camera[:, :, 0] * X_trans.reshape(shape[0], -1)).reshape(shape).astype(dtype)


def _dense(m):
    return np.asarray(m.todense()) if hasattr(m, 'todense') else np.asarray(m)
