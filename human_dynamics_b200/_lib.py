"""ctypes binding of libhd_b200.so (the C-ABI declared in include/hd_b200.h).

There is NO CPU fallback: if the shared library is missing this module raises at import,
and every op raises if it is handed a non-CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libhd_b200.so')

HD_IMPL_SIMT, HD_IMPL_TC_3XTF32, HD_IMPL_TC_1XTF32, HD_IMPL_TC_3XF16 = 0, 1, 2, 3
HD_CONV_NO_TMA_EPILOGUE = 1
HD_CONV_INPUT_PLANES = 2
IMPL_BY_NAME = {'simt': HD_IMPL_SIMT, 'tc3': HD_IMPL_TC_3XTF32, 'tc1': HD_IMPL_TC_1XTF32, 'tc3h': HD_IMPL_TC_3XF16}


class ConvDesc(C.Structure):
    """Mirror of hd_conv_desc (include/hd_b200.h)."""
    _fields_ = [
        ('in_', C.c_void_p), ('in_ld', C.c_longlong),
        ('n_img', C.c_int), ('H', C.c_int), ('W', C.c_int), ('Cin', C.c_int),
        ('Ho', C.c_int), ('Wo', C.c_int), ('KH', C.c_int), ('KW', C.c_int),
        ('stride', C.c_int), ('pad_t', C.c_int), ('pad_l', C.c_int),
        ('w_kn', C.c_void_p), ('w_nk_hi', C.c_void_p), ('w_nk_lo', C.c_void_p),
        ('Cout', C.c_int), ('K_pad', C.c_int),
        ('pre_scale', C.c_void_p), ('pre_shift', C.c_void_p), ('pre_img_stride', C.c_int), ('pre_relu', C.c_int),
        ('post_scale', C.c_void_p), ('post_shift', C.c_void_p), ('post_relu', C.c_int),
        ('res', C.c_void_p), ('res_ld', C.c_longlong), ('res_H', C.c_int), ('res_W', C.c_int), ('res_stride', C.c_int),
        ('out', C.c_void_p), ('out_ld', C.c_longlong),
        ('impl', C.c_int),
        ('tmap_hi', C.c_void_p), ('tmap_lo', C.c_void_p),
        ('in_hi', C.c_void_p), ('in_lo', C.c_void_p),
        ('out_hi', C.c_void_p), ('out_lo', C.c_void_p), ('out2_ld', C.c_longlong),
        ('post2_scale', C.c_void_p), ('post2_shift', C.c_void_p), ('post2_relu', C.c_int),
        ('tmap_res', C.c_void_p), ('tmap_out', C.c_void_p), ('tmap_out_hi', C.c_void_p), ('tmap_out_lo', C.c_void_p),
        ('flags', C.c_int),
        ('tmap_hi_n64', C.c_void_p), ('tmap_lo_n64', C.c_void_p),
        ('out_subsample', C.c_int),
    ]


WEIGHT_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_longlong))      # hd_weight_fn


class SmplConsts(C.Structure):
    """Mirror of hd_smpl_consts."""
    _fields_ = [
        ('num_verts', C.c_int), ('num_kps', C.c_int), ('lbs_nnz', C.c_int), ('kp_nnz_total', C.c_int),
        ('v_template', C.c_void_p), ('dirs', C.c_void_p), ('J_template', C.c_void_p), ('J_shapedirs', C.c_void_p),
        ('lbs_idx', C.c_void_p), ('lbs_w', C.c_void_p),
        ('kp_ptr', C.c_void_p), ('kp_vidx', C.c_void_p), ('kp_w', C.c_void_p),
        ('parents', C.c_int * 24),
    ]


# name -> (restype, argtypes); must list every symbol include/hd_b200.h declares.
_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
SIGNATURES = {
    'hd_version': (_i, []),
    'hd_status_string': (C.c_char_p, [_i]),
    'hd_last_error': (C.c_char_p, []),
    'hd_launch_count': (_ll, []),
    'hd_launch_count_reset': (None, []),
    'hd_conv_gemm': (_i, [C.POINTER(ConvDesc), _vp]),
    'hd_conv_gemm_profile': (_i, [C.POINTER(ConvDesc), _vp, _vp]),
    'hd_make_weight_tmap': (_i, [_vp, _i, _i, _i, _i, _vp]),
    'hd_make_act_tmap': (_i, [_vp, _ll, _i, _ll, _i, _vp]),
    'hd_pack_conv1_planes': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'hd_conv1_7x7s2': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'hd_maxpool3x3s2_same': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'hd_subsample': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'hd_bnrelu_avgpool': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'hd_process_image': (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    'hd_crop_geometry': (_i, [_i, _i, _vp, _i, _vp, _vp, _vp]),
    'hd_groupnorm_stats': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'hd_groupnorm_relu_split': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'hd_split_f16': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'hd_ief_fc1_theta': (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _vp]),
    'hd_ief_fc3': (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    'hd_ief_delta_init': (_i, [_vp, _vp, _i, _i, _vp]),
    'hd_net_destroy': (None, [_vp]),
    'hd_net_error': (C.c_char_p, [_vp]),
    'hd_net_num_launches': (_ll, [_vp]),
    'hd_resnet50_create': (_i, [_vp, _vp, _i, _i, C.POINTER(_vp)]),
    'hd_resnet50_forward': (_i, [_vp, _vp, _vp, _vp]),
    'hd_fmovie_create': (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_vp)]),
    'hd_fmovie_forward': (_i, [_vp, _vp, _vp, _vp]),
    'hd_ief_create': (_i, [_vp, _vp, _i, C.POINTER(_i), _i, C.POINTER(_vp)]),
    'hd_ief_forward': (_i, [_vp, _vp, _vp, _vp, _vp]),
    'hd_smpl_pack_sizes': (_i, [_i, _i, _vp, _vp, C.POINTER(_i), C.POINTER(_i)]),
    'hd_smpl_pack': (_i, [_i, _i] + [_vp] * 13 + [_i] + [_vp] * 4),
    'hd_smpl_workspace_bytes': (_sz, [_i]),
    'hd_smpl_forward': (_i, [C.POINTER(SmplConsts), _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _sz, _vp]),
    'hd_smpl_pose': (_i, [C.POINTER(SmplConsts), _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    'hd_smpl_lbs_tc': (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _vp, _i, _i, _i, _i, _vp]),
    'hd_smpl_lbs': (_i, [C.POINTER(SmplConsts), _vp, _ll, _vp, _vp, _i, _i, _i, _vp]),
    'hd_smpl_joints': (_i, [C.POINTER(SmplConsts), _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    'hd_rodrigues': (_i, [_vp, _vp, _i, _vp]),
    'hd_rot2aa': (_i, [_vp, _vp, _i, _vp]),
    'hd_global_rigid': (_i, [_vp, _vp, C.POINTER(C.c_int), _vp, _vp, _i, _i, _vp]),
    'hd_orth_proj': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'human_dynamics_b200: %s is missing -- build it with `make -C human_dynamics_b200/csrc` '
            '(or `python -c "import __graft_entry__ as g; g.build()"`). There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class HDError(RuntimeError):
    pass


def check(rc: int, what: str = ''):
    if rc != 0:
        raise HDError('%s failed: %s [%s]' % (what or 'libhd_b200 call', lib.hd_status_string(rc).decode(),
                                              lib.hd_last_error().decode()))


def dptr(t, dtype=None):
    """Device pointer of a CUDA tensor (contiguity is the caller's business: strided views are allowed)."""
    import torch
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HDError('expected a CUDA torch.Tensor (no CPU fallback exists), got %r' % (type(t),))
    if dtype is not None and t.dtype != dtype:
        raise HDError('expected dtype %s, got %s' % (dtype, t.dtype))
    return C.c_void_p(t.data_ptr())


def fptr(t):
    import torch
    return dptr(t, torch.float32)


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
