"""Host side of the batched SMPL forward: packs the model constants once, then one C-ABI call per batch.

Mirrors `SMPL.__init__` / `SMPL.__call__` of the reference (src/tf_smpl/batch_smpl.py:27-162); the
reference-named class lives in src/tf_smpl/batch_smpl.py and delegates here.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, fptr, dptr, current_stream


def _dense(m):
    m = m.r if (hasattr(m, 'r') and not isinstance(m, np.ndarray) and not hasattr(m, 'todense')) else m   # undo_chumpy, batch_smpl.py:22-23
    return np.asarray(m.todense()) if hasattr(m, 'todense') else np.asarray(m)


class _ChStub(object):
    """Stand-in for chumpy objects inside the official SMPL pickles (chumpy is not a dependency here).

    chumpy.Ch pickles as (class, state-dict) with the wrapped ndarray under 'x'; the reference reads it through `.r`
    (`undo_chumpy`, batch_smpl.py:22-23).  Any other attribute of the state is kept but unused."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {'x': state})

    @property
    def r(self):
        return np.asarray(self.__dict__['x'])


class _SMPLUnpickler(pickle.Unpickler):
    """pickle.Unpickler that maps every class from the `chumpy` package to `_ChStub` (scipy.sparse / numpy load normally)."""

    def find_class(self, module, name):
        if module == 'chumpy' or module.startswith('chumpy.'):
            return _ChStub
        return super().find_class(module, name)


def load_smpl_model(pkl_path_or_dict):
    """The un-pickled SMPL model dict (batch_smpl.py:31-32: `pickle.load(f, encoding='latin1')`), without needing chumpy."""
    if isinstance(pkl_path_or_dict, dict):
        return pkl_path_or_dict
    with open(pkl_path_or_dict, 'rb') as f:
        return _SMPLUnpickler(f, encoding='latin1').load()


class SMPLConstants(object):
    """Device-resident SMPL constants in the layout the kernels want (hd_smpl_consts)."""

    def __init__(self, model, joint_type='cocoplus', device=None, tc=True):
        if joint_type not in ('cocoplus', 'lsp'):
            raise ValueError('BAD!! Unknown joint type: %s, it must be either "cocoplus" or "lsp"' % joint_type)
        dd = load_smpl_model(model)
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        v_template = _dense(dd['v_template']).astype(np.float64)               # (V,3)
        V = v_template.shape[0]
        shapedirs = _dense(dd['shapedirs']).astype(np.float64)                 # (V,3,10)
        nb = shapedirs.shape[-1]
        posedirs = _dense(dd['posedirs']).astype(np.float64)                   # (V,3,207)
        if nb != 10 or posedirs.shape[-1] != 207:
            raise ValueError('expected 10 betas and 207 pose-blend bases, got %d / %d' % (nb, posedirs.shape[-1]))
        Jreg = _dense(dd['J_regressor']).astype(np.float64)                    # (24,V)
        if Jreg.shape != (24, V):
            raise ValueError('J_regressor must be (24, V)')
        weights = _dense(dd['weights']).astype(np.float64)                     # (V,24)
        kreg = _dense(dd['cocoplus_regressor']).astype(np.float64)             # (K,V)
        if joint_type == 'lsp':
            kreg = kreg[:14]                                                   # batch_smpl.py:81-82
        parents = np.asarray(dd['kintree_table'])[0].astype(np.int64).astype(np.int32)   # uint32(-1) -> -1, :66
        self.parents = parents.copy()
        self.num_verts = V
        self.num_kps = kreg.shape[0]
        self.size = [V, 3]
        self.num_betas = nb

        sd = shapedirs.reshape(-1, nb).T                                       # (10, V*3)   :45-48
        pd = posedirs.reshape(-1, 207).T                                       # (207, V*3)  :60-63
        dirs = np.concatenate([sd, pd], axis=0).astype(np.float32)
        # J = (beta.shapedirs + v_template).J_regressor is linear in beta: precompose (float64) once.
        J_template = (Jreg @ v_template).astype(np.float32)                    # (24,3)
        J_shapedirs = np.einsum('jv,vcb->bjc', Jreg, shapedirs).reshape(nb, 72).astype(np.float32)

        nnz = int(max(1, (weights != 0).sum(axis=1).max()))
        if nnz <= 4:
            nnz = 4
        elif nnz < 24:
            nnz = min(24, (nnz + 3) // 4 * 4)
        mask = weights != 0
        order = np.argsort(~mask, axis=1, kind='stable')[:, :nnz]              # non-zero joints first, ascending
        lbs_w = np.take_along_axis(weights, order, 1)
        lbs_idx = np.where(lbs_w != 0, order, 0)

        kp_ptr = [0]
        kp_vidx, kp_w = [], []
        for k in range(kreg.shape[0]):
            nzv = np.nonzero(kreg[k])[0]
            kp_vidx.append(nzv)
            kp_w.append(kreg[k, nzv])
            kp_ptr.append(kp_ptr[-1] + len(nzv))
        kp_vidx = np.concatenate(kp_vidx) if kp_vidx else np.zeros(0, np.int64)
        kp_w = np.concatenate(kp_w) if kp_w else np.zeros(0)

        def f32(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)

        def i32(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)

        self.v_template = f32(v_template.reshape(-1))
        self.dirs = f32(dirs)
        self.J_template = f32(J_template.reshape(-1))
        self.J_shapedirs = f32(J_shapedirs)
        self.lbs_idx = i32(lbs_idx)
        self.lbs_w = f32(lbs_w)
        self.kp_ptr = i32(np.asarray(kp_ptr))
        self.kp_vidx = i32(kp_vidx if len(kp_vidx) else np.zeros(1))
        self.kp_w = f32(kp_w if len(kp_w) else np.zeros(1))
        self.lbs_nnz = nnz

        c = _lib.SmplConsts()
        c.num_verts, c.num_kps, c.lbs_nnz, c.kp_nnz_total = V, self.num_kps, nnz, int(kp_ptr[-1])
        c.v_template = self.v_template.data_ptr()
        c.dirs = self.dirs.data_ptr()
        c.J_template = self.J_template.data_ptr()
        c.J_shapedirs = self.J_shapedirs.data_ptr()
        c.lbs_idx = self.lbs_idx.data_ptr()
        c.lbs_w = self.lbs_w.data_ptr()
        c.kp_ptr = self.kp_ptr.data_ptr()
        c.kp_vidx = self.kp_vidx.data_ptr()
        c.kp_w = self.kp_w.data_ptr()
        for i in range(24):
            c.parents[i] = int(parents[i])
        self.c = c
        self._ws = None
        # Tensor-core blend for large batches: v_posed = [beta | R-I] . dirs + v_template as one [N,256] x [256, V*3] GEMM
        # (fp16 head/remainder split, FP32-class), then HBM-shaped skinning.  Small batches use the fused SIMT kernel.
        self.tc_min_batch = 256
        self.blend = None
        self._tc_bufs = {}
        if tc:
            from .nets import PackedConv
            self.vp_ld = (V * 3 + 3) // 4 * 4
            wb = np.zeros((256, self.vp_ld), np.float32)
            wb[:217, :V * 3] = dirs
            bias = np.zeros(self.vp_ld, np.float32)
            bias[:V * 3] = v_template.reshape(-1)
            self.blend = PackedConv(wb, dev, post_shift=bias, tc='tc3h')
            # dense skinning weights as the A operand of the tensor-core skinning GEMM: [roundup128(V), 32] fp16 head + unscaled remainder
            wd = np.zeros(((V + 127) // 128 * 128, 32), np.float32)
            wd[:V, :24] = weights
            w_hi = wd.astype(np.float16)
            self.w_hi = torch.from_numpy(w_hi).to(dev)
            self.w_lo = torch.from_numpy((wd - w_hi.astype(np.float32)).astype(np.float16)).to(dev)
        self.lbs_tc = bool(tc) and os.environ.get('HD_LBS_TC', '1') != '0' and (V * 3 * 4) % 8 == 0
        self.lbs_tc_min_batch = int(os.environ.get('HD_LBS_TC_MIN', '2368'))         # 148 SMs x 16 poses

    def workspace(self, N):
        need = int(lib.hd_smpl_workspace_bytes(N))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, beta, theta, cam=None, want_joints=True, want_Rs=True, want_Jtr=True, out=None, slot=(1, 0)):
        """beta (N,10)-view, theta (N,72)-view, cam (N,3)-view: float32 CUDA, unit inner stride, any row stride.

        `out` may hold pre-allocated output tensors; with slot=(D, d) pose n is written to row n*D+d of each
        of them (they must then have N*D leading rows) -- the in-place [B,T,D,...] stacking of tester.py:252.
        Returns dict(verts, joints, Rs, Jtr, kps).
        """
        mul, off = slot
        for name, t, w in (('beta', beta, 10), ('theta', theta, 72)) + ((('cam', cam, 3),) if cam is not None else ()):
            if not t.is_cuda or t.dtype != torch.float32:
                raise _lib.HDError('SMPL %s must be a float32 CUDA tensor (no CPU fallback exists)' % name)
            if t.dim() != 2 or t.shape[1] != w or t.stride(1) != 1:
                raise _lib.HDError('SMPL %s must be (N,%d) with unit inner stride, got %s' % (name, w, tuple(t.shape)))
        N = beta.shape[0]
        if theta.shape[0] != N or (cam is not None and cam.shape[0] != N):
            raise _lib.HDError('SMPL batch mismatch')
        V, K = self.num_verts, self.num_kps
        dev = beta.device
        o = out or {}
        verts = o.get('verts') if 'verts' in o else torch.empty((N * mul, V, 3), dtype=torch.float32, device=dev)
        joints = o.get('joints') if 'joints' in o else (torch.empty((N * mul, K, 3), dtype=torch.float32, device=dev) if want_joints else None)
        Rs = o.get('Rs') if 'Rs' in o else (torch.empty((N * mul, 24, 3, 3), dtype=torch.float32, device=dev) if want_Rs else None)
        Jtr = o.get('Jtr') if 'Jtr' in o else (torch.empty((N * mul, 24, 3), dtype=torch.float32, device=dev) if want_Jtr else None)
        kps = None
        if cam is not None:
            kps = o.get('kps') if 'kps' in o else torch.empty((N * mul, K, 2), dtype=torch.float32, device=dev)
        if N >= self.tc_min_batch and self.blend is not None and self.blend.tc:
            self._forward_tc(beta, theta, cam, N, verts, joints, Rs, Jtr, kps, int(mul), int(off))
        elif N > 0:
            ws = self.workspace(N)
            rc = lib.hd_smpl_forward(C.byref(self.c), fptr(beta), beta.stride(0), fptr(theta), theta.stride(0), N,
                                     fptr(verts), fptr(joints), fptr(Rs), fptr(Jtr),
                                     fptr(cam) if cam is not None else None, cam.stride(0) if cam is not None else 0,
                                     fptr(kps) if kps is not None else None, int(mul), int(off),
                                     dptr(ws), ws.numel(), current_stream())
            check(rc, 'hd_smpl_forward')
        return {'verts': verts, 'joints': joints, 'Rs': Rs, 'Jtr': Jtr, 'kps': kps}

    def _forward_tc(self, beta, theta, cam, N, verts, joints, Rs, Jtr, kps, mul, off):
        """pose -> tensor-core blend GEMM -> skinning -> keypoints (hd_smpl_pose / hd_conv_gemm / hd_smpl_lbs / hd_smpl_joints)."""
        dev = beta.device
        if N not in self._tc_bufs:
            while len(self._tc_bufs) >= 3:                     # bounded cache: an entry holds N * 83 KB of v_posed
                self._tc_bufs.pop(next(iter(self._tc_bufs)))
            f32 = dict(dtype=torch.float32, device=dev)
            coef = (torch.empty((N, 256), dtype=torch.float16, device=dev), torch.empty((N, 256), dtype=torch.float16, device=dev))
            vpos = torch.empty((N, self.vp_ld), **f32)
            a12 = torch.empty((N, 288), **f32)
            rsw = torch.empty((N, 216), **f32)
            # tensor-core skinning walks 16-pose batches, one CTA per batch at a time: below ~148 batches the CUDA-core kernel (one
            # CTA per 16 poses x 128 vertices) fills the chip better (640 poses: 56 us vs 100 us)
            a12t = (torch.empty((N, 12, 32), dtype=torch.float16, device=dev), torch.empty((N, 12, 32), dtype=torch.float16, device=dev)) \
                if (self.lbs_tc and N >= self.lbs_tc_min_batch) else None
            # operand rows arrive pre-split from the pose kernel: cp.async producer + TMA-store epilogue (K = 256)
            op = self.blend.bind(None, N, 1, 1, vpos, inp_split=coef, impl='tc3h')
            self._tc_bufs[N] = (coef, vpos, a12, rsw, op, a12t)
        coef, vpos, a12, rsw, op, a12t = self._tc_bufs[N]
        st = current_stream()
        check(lib.hd_smpl_pose(C.byref(self.c), fptr(beta), beta.stride(0), fptr(theta), theta.stride(0), N, fptr(Rs), fptr(Jtr),
                               fptr(a12), None, 256, C.c_void_p(coef[0].data_ptr()), C.c_void_p(coef[1].data_ptr()),
                               C.c_void_p(a12t[0].data_ptr()) if a12t else None, C.c_void_p(a12t[1].data_ptr()) if a12t else None, mul, off,
                               dptr(rsw), rsw.numel() * 4, st), 'hd_smpl_pose')
        op.run(st)
        if a12t is not None:       # T = W . A on the tensor cores, applied from TMEM (smpl_lbs_tc.cu)
            check(lib.hd_smpl_lbs_tc(C.c_void_p(self.w_hi.data_ptr()), C.c_void_p(self.w_lo.data_ptr()), C.c_void_p(a12t[0].data_ptr()),
                                     C.c_void_p(a12t[1].data_ptr()), fptr(vpos), self.vp_ld, fptr(verts), N, self.num_verts, mul, off, st),
                  'hd_smpl_lbs_tc')
        else:
            check(lib.hd_smpl_lbs(C.byref(self.c), fptr(vpos), self.vp_ld, fptr(a12), fptr(verts), N, mul, off, st), 'hd_smpl_lbs')
        if (joints is not None or kps is not None) and self.num_kps > 0:
            check(lib.hd_smpl_joints(C.byref(self.c), fptr(verts), fptr(cam) if cam is not None else None,
                                     cam.stride(0) if cam is not None else 0, fptr(joints), fptr(kps) if kps is not None else None,
                                     N, mul, off, st), 'hd_smpl_joints')


def _noop():
    pass


def batch_rodrigues(theta):
    """theta (M,3) float32 CUDA -> (M,3,3).  src/tf_smpl/batch_lbs.py:42-60."""
    if not theta.is_cuda:
        raise _lib.HDError('batch_rodrigues: CUDA tensor required (no CPU fallback exists)')
    theta = theta.contiguous().float()
    if theta.dim() != 2 or theta.shape[1] != 3:
        raise _lib.HDError('batch_rodrigues: theta must be (M,3)')
    M = theta.shape[0]
    R = torch.empty((M, 3, 3), dtype=torch.float32, device=theta.device)
    check(lib.hd_rodrigues(fptr(theta), fptr(R), M, current_stream()), 'hd_rodrigues')
    return R


def batch_rot2aa(Rs):
    """Rs (B,3,3) float32 CUDA -> axis-angle (B,3).  src/tf_smpl/batch_lbs.py:63-105."""
    if not Rs.is_cuda:
        raise _lib.HDError('batch_rot2aa: CUDA tensor required (no CPU fallback exists)')
    Rs = Rs.contiguous().float()
    if Rs.dim() != 3 or tuple(Rs.shape[1:]) != (3, 3):
        raise _lib.HDError('batch_rot2aa: Rs must be (B,3,3)')
    aa = torch.empty((Rs.shape[0], 3), dtype=torch.float32, device=Rs.device)
    check(lib.hd_rot2aa(fptr(Rs), fptr(aa), Rs.shape[0], current_stream()), 'hd_rot2aa')
    return aa


def batch_global_rigid_transformation(Rs, Js, parent, rotate_base=False):
    """Rs (N,24,3,3), Js (N,24,3), parent int[24] -> (new_J (N,24,3), A (N,24,4,4)).  batch_lbs.py:133-194."""
    if not (Rs.is_cuda and Js.is_cuda):
        raise _lib.HDError('batch_global_rigid_transformation: CUDA tensors required (no CPU fallback exists)')
    Rs = Rs.contiguous().float()
    Js = Js.contiguous().float()
    N = Rs.shape[0]
    if tuple(Rs.shape[1:]) != (24, 3, 3) or tuple(Js.shape) != (N, 24, 3):
        raise _lib.HDError('batch_global_rigid_transformation: expected Rs (N,24,3,3), Js (N,24,3)')
    par = (C.c_int * 24)(*[(-1 if (int(p) < 0 or int(p) >= 2 ** 31) else int(p)) for p in np.asarray(parent).tolist()])
    new_J = torch.empty((N, 24, 3), dtype=torch.float32, device=Rs.device)
    A = torch.empty((N, 24, 4, 4), dtype=torch.float32, device=Rs.device)
    check(lib.hd_global_rigid(fptr(Rs), fptr(Js), par, fptr(new_J), fptr(A), N, int(bool(rotate_base)), current_stream()),
          'hd_global_rigid')
    return new_J, A


def batch_orth_proj_idrot(X, camera):
    """X (N,P,3), camera (N,3) -> (N,P,2).  src/tf_smpl/projection.py:16-29."""
    if not (X.is_cuda and camera.is_cuda):
        raise _lib.HDError('batch_orth_proj_idrot: CUDA tensors required (no CPU fallback exists)')
    X = X.contiguous().float()
    camera = camera.reshape(-1, 3).contiguous().float()
    N, P = X.shape[0], X.shape[1]
    if X.shape[2] != 3 or camera.shape[0] != N:
        raise _lib.HDError('batch_orth_proj_idrot: expected X (N,P,3) and camera (N,3)')
    out = torch.empty((N, P, 2), dtype=torch.float32, device=X.device)
    check(lib.hd_orth_proj(fptr(X), fptr(camera), fptr(out), N, P, current_stream()), 'hd_orth_proj')
    return out
