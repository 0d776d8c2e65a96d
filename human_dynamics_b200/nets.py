"""Host-side layer plans for the three networks on the path: slim ResNet-v2-50, f_movie, IEF.

A *plan* is a list of pre-filled C descriptors (hd_conv_desc) over pre-allocated device buffers, so a
forward pass is a sequence of ctypes calls with no Python-side tensor math and no allocation.
Weights come in as a dict of numpy arrays keyed by TF variable names.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, fptr, current_stream, ConvDesc

RESNET_BLOCKS = ((64, 3, 2), (128, 4, 2), (256, 6, 2), (512, 3, 1))
BN_EPS = 1e-5        # resnet_arg_scope batch_norm_epsilon [TF-ext]
GN_EPS = 1e-6        # tf.contrib.layers.group_norm epsilon [TF-ext]
GN_GROUPS = 32


def _dev(a, device, dtype=np.float32):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(device)


def fold_bn(w, prefix):
    """BN inference -> (scale, shift): y = x*scale + shift  (A.5)."""
    g = w[prefix + '/gamma'].astype(np.float64)
    b = w[prefix + '/beta'].astype(np.float64)
    m = w[prefix + '/moving_mean'].astype(np.float64)
    v = w[prefix + '/moving_variance'].astype(np.float64)
    s = g / np.sqrt(v + BN_EPS)
    return s.astype(np.float32), (b - m * s).astype(np.float32)


def tf32_split(w):
    """w (float32) -> (hi, lo), both exactly representable in TF32 (low 13 mantissa bits zero, which is all the
    tensor core reads): hi = w rounded to nearest TF32, lo = (w - hi) rounded to nearest TF32.  Rounding (not
    truncating) keeps the representation error zero-mean, so it does not build up over the 53 layers."""
    def rn_tf32(x):
        b = np.ascontiguousarray(x, np.float32).view(np.uint32)
        return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    w = np.ascontiguousarray(w, np.float32)
    hi = rn_tf32(w)
    lo = rn_tf32((w - hi).astype(np.float32))
    return hi, lo


def f16_split(w):
    """w (float32) -> (hi, lo) float16: hi = RN_f16(w), lo = RN_f16((w - hi) * 2^11).  Same 11+11 significant bits as the
    TF32 split at twice the tensor-core rate; the 2^11 scale keeps the remainder of small weights out of the fp16
    subnormal range (the kernel accumulates the scaled cross terms separately and rescales once)."""
    w = np.ascontiguousarray(w, np.float32)
    hi = w.astype(np.float16)
    lo = ((w - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return hi, lo


class PackedConv(object):
    """Device-resident weights (+ epilogue vectors) of one conv / FC layer."""

    def __init__(self, w_hwio, device, post_scale=None, post_shift=None, post_relu=False, stride=1, pad=(0, 0),
                 tc=False):
        w_hwio = np.asarray(w_hwio, np.float32)
        if w_hwio.ndim == 2:
            w_hwio = w_hwio[None, None]
        self.KH, self.KW, self.Cin, self.Cout = w_hwio.shape
        self.K = self.KH * self.KW * self.Cin
        self.stride = stride
        self.pad_t, self.pad_l = pad
        self.device = device
        w_kn = w_hwio.reshape(self.K, self.Cout)
        self.w_kn = _dev(w_kn, device)
        self.post_scale = _dev(post_scale, device) if post_scale is not None else None
        self.post_shift = _dev(post_shift, device) if post_shift is not None else None
        self.post_relu = bool(post_relu)
        self.tc = False            # False | 'f16' | 'tf32': which tensor-core packing this layer carries
        self.K_pad = 0
        want = {True: 'f16', 'auto': 'f16', 'tc3h': 'f16', 'tc3': 'tf32', 'tc1': 'tf32'}.get(tc, tc)
        gather = want == 'f16' and self.Cin % 32 != 0 and self.Cout <= 64   # conv1: row-segment gather producer
        if want == 'f16' and self.Cin % 64 != 0 and not gather:
            want = 'tf32'
        self.gather = bool(gather)
        if want in ('f16', 'tf32') and (self.Cin % 32 == 0 or gather):
            seg = self.KW * self.Cin
            segp = (seg + 7) // 8 * 8                        # gather layout: each kernel row's KW*Cin floats padded to x8
            self.K_pad = (self.KH * segp + 63) // 64 * 64 if gather else self.K
            box = 64 if self.Cout <= 64 else 128              # must equal the kernel's N tile (conv_tc.cu)
            rows = (self.Cout + box - 1) // box * box
            w_nk = np.zeros((rows, self.K_pad), np.float32)
            if gather:                                       # K index = ky*segp + (kx*Cin + ci)   (conv_tc.cu GATHER producer)
                wg = w_hwio.reshape(self.KH, seg, self.Cout)
                for ky in range(self.KH):
                    w_nk[:self.Cout, ky * segp:ky * segp + seg] = wg[ky].T
            else:
                w_nk[:self.Cout, :self.K] = w_kn.T
            if want == 'f16':
                hi, lo = f16_split(w_nk)
                self.w_nk_hi = torch.from_numpy(hi).to(device)
                self.w_nk_lo = torch.from_numpy(lo).to(device)
                eb = 2
            else:
                hi, lo = tf32_split(w_nk)
                self.w_nk_hi = _dev(hi, device)
                self.w_nk_lo = _dev(lo, device)
                eb = 4
            self.tmap_hi = (C.c_ubyte * 128)()
            self.tmap_lo = (C.c_ubyte * 128)()
            check(lib.hd_make_weight_tmap(C.c_void_p(self.w_nk_hi.data_ptr()), rows, self.K_pad, box, eb, C.cast(self.tmap_hi, C.c_void_p)),
                  'hd_make_weight_tmap')
            check(lib.hd_make_weight_tmap(C.c_void_p(self.w_nk_lo.data_ptr()), rows, self.K_pad, box, eb, C.cast(self.tmap_lo, C.c_void_p)),
                  'hd_make_weight_tmap')
            self.tmap_hi64 = self.tmap_lo64 = None
            if want == 'f16' and box == 128 and self.Cout % 64 == 0:      # 64-row boxes: few-tile GEMMs run on 64-wide tiles
                self.tmap_hi64, self.tmap_lo64 = (C.c_ubyte * 128)(), (C.c_ubyte * 128)()
                for t, m in ((self.w_nk_hi, self.tmap_hi64), (self.w_nk_lo, self.tmap_lo64)):
                    check(lib.hd_make_weight_tmap(C.c_void_p(t.data_ptr()), rows, self.K_pad, 64, eb, C.cast(m, C.c_void_p)), 'hd_make_weight_tmap')
            self.tc = want

    def bind(self, inp, n_img, H, W, out, in_ld=None, out_ld=None, pre=None, res=None, res_geom=None, impl='auto',
             inp_split=None, out_split=None, post2=None, out_subsample=0):
        """Fill a descriptor.  inp/out/res: CUDA float32 tensors (only their data_ptr is used).

        pre = (scale, shift, img_stride, relu); res_geom = (res_ld, res_H, res_W, res_stride).
        out_subsample = s > 1: `out` is the dense [n, ceil(Ho/s), ceil(Wo/s), Cout] tensor x[:, ::s, ::s] (hd_b200.h);
        inp_split = (hi, lo) fp16 tensors: pre-activated, pre-split A operand (then `inp` may be None);
        out_split = (hi, lo) fp16 tensors + post2 = (scale|None, shift|None, relu): second output (then `out` may be None).
        """
        d = ConvDesc()
        Ho = (H + 2 * self.pad_t - self.KH) // self.stride + 1 if self.KH > 1 else (H - 1) // self.stride + 1
        Wo = (W + 2 * self.pad_l - self.KW) // self.stride + 1 if self.KW > 1 else (W - 1) // self.stride + 1
        d.in_ = inp.data_ptr() if inp is not None else None
        d.in_ld = self.Cin if in_ld is None else in_ld
        if inp_split is not None:
            d.in_hi, d.in_lo = inp_split[0].data_ptr(), inp_split[1].data_ptr()
        d.n_img, d.H, d.W, d.Cin = n_img, H, W, self.Cin
        d.Ho, d.Wo, d.KH, d.KW = Ho, Wo, self.KH, self.KW
        d.stride, d.pad_t, d.pad_l = self.stride, self.pad_t, self.pad_l
        d.w_kn = self.w_kn.data_ptr()
        d.Cout = self.Cout
        d.K_pad = self.K_pad
        if pre is not None:
            d.pre_scale, d.pre_shift = pre[0].data_ptr(), pre[1].data_ptr()
            d.pre_img_stride, d.pre_relu = int(pre[2]), int(pre[3])
        if self.post_scale is not None:
            d.post_scale = self.post_scale.data_ptr()
        if self.post_shift is not None:
            d.post_shift = self.post_shift.data_ptr()
        d.post_relu = int(self.post_relu)
        if res is not None:
            d.res = res.data_ptr()
            if res_geom is None:
                res_geom = (self.Cout, Ho, Wo, 1)
            d.res_ld, d.res_H, d.res_W, d.res_stride = res_geom
        d.out = out.data_ptr() if out is not None else None
        d.out_ld = self.Cout if out_ld is None else out_ld
        d.out_subsample = int(out_subsample) if out is not None else 0
        if out_split is not None:
            d.out_hi, d.out_lo, d.out2_ld = out_split[0].data_ptr(), out_split[1].data_ptr(), self.Cout
            if post2 is not None:
                if post2[0] is not None:
                    d.post2_scale = post2[0].data_ptr()
                if post2[1] is not None:
                    d.post2_shift = post2[1].data_ptr()
                d.post2_relu = int(post2[2])
        # ragged layers (Cin % 32 != 0, unaligned views) always run on the exact-FP32 SIMT kernel
        use_tc = bool(self.tc) and impl in ('auto', 'tc3', 'tc1', 'tc3h') and \
            (self.gather or inp_split is not None or ((d.in_ld % 4 == 0) and (inp.data_ptr() % 16 == 0)))
        if (inp_split is not None or out_split is not None) and not (use_tc and self.tc == 'f16'):
            raise _lib.HDError('pre-split activations need the fp16 tensor-core packing (impl auto / tc3h, Cin % 64 == 0)')
        if use_tc:
            if self.tc == 'f16':
                d.impl = _lib.HD_IMPL_TC_3XF16
            else:
                d.impl = _lib.HD_IMPL_TC_1XTF32 if impl == 'tc1' else _lib.HD_IMPL_TC_3XTF32
            d.w_nk_hi, d.w_nk_lo = self.w_nk_hi.data_ptr(), self.w_nk_lo.data_ptr()
            d.tmap_hi = C.cast(self.tmap_hi, C.c_void_p)
            d.tmap_lo = C.cast(self.tmap_lo, C.c_void_p)
            if getattr(self, 'tmap_hi64', None) is not None:
                d.tmap_hi_n64 = C.cast(self.tmap_hi64, C.c_void_p)
                d.tmap_lo_n64 = C.cast(self.tmap_lo64, C.c_void_p)
        else:
            d.impl = _lib.HD_IMPL_SIMT
        op = ConvOp(d, (self, inp, out, pre, res, inp_split, out_split, post2), (Ho, Wo))
        if not TMA_EPILOGUE:
            d.flags |= _lib.HD_CONV_NO_TMA_EPILOGUE
        op.encode_act_maps()
        return op


class SubsampleOp(object):
    """x[:, ::s, ::s, :] into a dense buffer (slim's max_pool2d(1x1, stride) shortcut): one op in a plan's op list."""
    __slots__ = ('src', 'dst', 'geom', 'd')

    def __init__(self, src, dst, n, H, C, stride):
        self.src, self.dst, self.geom, self.d = src, dst, (n, H, H, C, stride), None

    def rebind(self, field, tensor):
        self.src = tensor

    def encode_act_maps(self):
        pass

    def run(self, stream):
        n, H, W, Cc, s = self.geom
        check(lib.hd_subsample(fptr(self.src), fptr(self.dst), n, H, W, Cc, s, stream), 'hd_subsample')


SUBSAMPLE_RES = os.environ.get('HD_SUBSAMPLE_RES', '1') != '0'    # A/B switch: strided identity shortcuts via hd_subsample + plain residual
SUBSAMPLE_EPI = os.environ.get('HD_SUBSAMPLE_EPI', '1') != '0'    # A/B switch: the unit in front of a strided identity unit writes x[:, ::s, ::s] itself
TMA_EPILOGUE = os.environ.get('HD_TMA_EPILOGUE', '1') != '0'     # A/B switch for the K <= 256 layers (results identical)


class ConvOp(object):
    __slots__ = ('d', 'keep', 'out_hw', 'ref', 'dyn', 'maps')

    def __init__(self, d, keep, out_hw):
        self.d, self.keep, self.out_hw = d, keep, out_hw
        self.ref = C.byref(d)
        self.dyn = {}           # descriptor field -> tensor it currently points at (rebinding overwrites, never appends)
        self.maps = None

    def encode_act_maps(self):
        """(Re-)encode the activation tensor maps of the TMA epilogue for the pointers currently in the descriptor.
        Only layers the kernel can run that way get maps (conv_tc.cu launch_conv_tc): fp16-split input, Cout % 32 == 0,
        residual row == output row; everything else keeps the per-thread epilogue."""
        d = self.d
        K = d.KH * d.KW * d.Cin
        ok = (d.impl == _lib.HD_IMPL_TC_3XF16 and d.in_hi and d.Cout % 32 == 0 and
              (not d.res or (d.res_stride == 1 and d.res_H == d.Ho and d.res_W == d.Wo)))
        for f in ('tmap_res', 'tmap_out', 'tmap_out_hi', 'tmap_out_lo'):
            setattr(d, f, None)
        if not ok:
            return
        M = d.n_img * d.Ho * d.Wo
        if self.maps is None:
            self.maps = {f: (C.c_ubyte * 128)() for f in ('res', 'out', 'out_hi', 'out_lo')}
        for f, ptr, ld, eb in (('res', d.res, d.res_ld, 4), ('out', d.out, d.out_ld, 4), ('out_hi', d.out_hi, d.out2_ld, 2),
                               ('out_lo', d.out_lo, d.out2_ld, 2)):
            if not ptr or (f == 'out' and d.out_subsample > 1):
                continue
            if ptr % 16 or (ld * eb) % 16:
                for g in ('tmap_res', 'tmap_out', 'tmap_out_hi', 'tmap_out_lo'):
                    setattr(d, g, None)
                return
            check(lib.hd_make_act_tmap(C.c_void_p(ptr), M, d.Cout, ld, eb, C.cast(self.maps[f], C.c_void_p)), 'hd_make_act_tmap')
            setattr(d, 'tmap_' + f, C.cast(self.maps[f], C.c_void_p))

    def rebind(self, field, tensor):
        """Point one descriptor field at another tensor (stage input / output of a cached plan)."""
        setattr(self.d, field, tensor.data_ptr())
        self.dyn[field] = tensor

    def run(self, stream):
        rc = lib.hd_conv_gemm(self.ref, stream)
        if rc:
            check(rc, 'hd_conv_gemm')


DROP_DEAD_FP32 = os.environ.get('HD_DROP_DEAD_FP32', '1') != '0'   # A/B switch: skip fp32 block outputs nobody reads
FAST_HEADS = os.environ.get('HD_FAST_HEADS', '1') != '0'          # A/B switch: f_movie / IEF through the pre-split + small-GEMM kernels
CONV1_PLANES = os.environ.get('HD_CONV1_PLANES', '1') != '0'      # A/B switch: conv1 from padded fp16 planes vs fp32 row-segment gather


class PackedConv1Planes(object):
    """ResNet root conv1 (7x7 stride 2, explicit pad 3+3, bias; A.2) on the tensor cores, reading its input as two padded
    RGBX fp16 planes [n, S+6, WP, 4] (head / remainder, hd_pack_conv1_planes): every (output pixel, kernel row) needs 8
    consecutive pixels = 64 contiguous, 16-byte-aligned bytes per plane, which four cp.async move straight into the swizzled
    A tile.  GEMM view: K = 8 kernel rows x 8 pixels x 4 channels = 256 (the 8th row / pixel / channel carry zero weights)."""

    def __init__(self, w_hwio, bias, device):
        w = np.asarray(w_hwio, np.float32)
        assert w.shape == (7, 7, 3, 64), w.shape
        self.device = device
        self.Cout, self.K = 64, 256
        w_nk = np.zeros((64, 8, 8, 4), np.float32)                 # [co, ky, kx, c]
        w_nk[:, :7, :7, :3] = w.transpose(3, 0, 1, 2)
        hi, lo = f16_split(w_nk.reshape(64, 256))
        self.w_nk_hi = torch.from_numpy(hi).to(device)
        self.w_nk_lo = torch.from_numpy(lo).to(device)
        self.bias = _dev(bias, device)
        self.tmap_hi = (C.c_ubyte * 128)()
        self.tmap_lo = (C.c_ubyte * 128)()
        for t, m in ((self.w_nk_hi, self.tmap_hi), (self.w_nk_lo, self.tmap_lo)):
            check(lib.hd_make_weight_tmap(C.c_void_p(t.data_ptr()), 64, 256, 64, 2, C.cast(m, C.c_void_p)), 'hd_make_weight_tmap')

    @staticmethod
    def plane_width(size):
        return (size + 8 + 1) // 2 * 2

    def alloc_planes(self, n, size):
        """Zero-initialised planes: the 3-pixel border (and the spare columns) must be zero and is never written again."""
        shape = (n, size + 6, self.plane_width(size), 4)
        return (torch.zeros(shape, dtype=torch.float16, device=self.device), torch.zeros(shape, dtype=torch.float16, device=self.device))

    def bind(self, planes, n, size, out):
        d = ConvDesc()
        WP = self.plane_width(size)
        d.in_hi, d.in_lo = planes[0].data_ptr(), planes[1].data_ptr()
        d.in_ld = 4
        d.n_img, d.H, d.W, d.Cin = n, size + 6, WP, 32
        d.Ho = d.Wo = size // 2
        d.KH, d.KW, d.stride, d.pad_t, d.pad_l = 8, 1, 2, 0, 0
        d.Cout, d.K_pad = 64, 256
        d.post_shift = self.bias.data_ptr()
        d.out, d.out_ld = out.data_ptr(), 64
        d.impl = _lib.HD_IMPL_TC_3XF16
        d.w_nk_hi, d.w_nk_lo = self.w_nk_hi.data_ptr(), self.w_nk_lo.data_ptr()
        d.tmap_hi, d.tmap_lo = C.cast(self.tmap_hi, C.c_void_p), C.cast(self.tmap_lo, C.c_void_p)
        d.flags = _lib.HD_CONV_INPUT_PLANES | (0 if TMA_EPILOGUE else _lib.HD_CONV_NO_TMA_EPILOGUE)
        op = ConvOp(d, (self, planes, out), (size // 2, size // 2))
        op.encode_act_maps()
        return op


# ------------------------------------------------------------------------------------------------
# ResNet-v2-50 (slim)  -- src/models.py:50-77
# ------------------------------------------------------------------------------------------------
class PackedResNet(object):
    def __init__(self, w, device, tc=False, blocks=RESNET_BLOCKS):
        p = 'resnet_v2_50'
        self.device = device
        self.blocks = blocks
        self.conv1_w = _dev(np.asarray(w[p + '/conv1/weights'], np.float32).reshape(147, 64), device)
        self.conv1_b = _dev(w[p + '/conv1/biases'], device)
        # conv2d_same(7x7, stride 2): explicit pad 3+3 then VALID (A.2); tensor-core path gathers the ragged K=147 element-wise
        self.conv1 = PackedConv(w[p + '/conv1/weights'], device, post_shift=w[p + '/conv1/biases'], stride=2, pad=(3, 3), tc=tc)
        want = {True: 'f16', 'auto': 'f16', 'tc3h': 'f16'}.get(tc, None)
        self.conv1_planes = PackedConv1Planes(w[p + '/conv1/weights'], w[p + '/conv1/biases'], device) if want == 'f16' else None
        self.units = []
        d_in = 64
        for b, (base, units, bstride) in enumerate(blocks, start=1):
            depth = 4 * base
            for u in range(1, units + 1):
                q = '%s/block%d/unit_%d/bottleneck_v2' % (p, b, u)
                stride = bstride if u == units else 1
                ps, pb = fold_bn(w, q + '/preact')
                unit = {'stride': stride, 'base': base, 'depth': depth, 'd_in': d_in,
                        'pre': (_dev(ps, device), _dev(pb, device))}
                if d_in != depth:
                    unit['shortcut'] = PackedConv(w[q + '/shortcut/weights'], device,
                                                  post_shift=w[q + '/shortcut/biases'], stride=stride, tc=tc)
                s1, b1 = fold_bn(w, q + '/conv1/BatchNorm')
                unit['conv1'] = PackedConv(w[q + '/conv1/weights'], device, s1, b1, True, tc=tc)
                s2, b2 = fold_bn(w, q + '/conv2/BatchNorm')
                # conv2d_same: stride 1 -> SAME (pad 1); stride 2 -> explicit pad 1+1 then VALID  (A.2)
                unit['conv2'] = PackedConv(w[q + '/conv2/weights'], device, s2, b2, True, stride=stride, pad=(1, 1), tc=tc)
                unit['conv3'] = PackedConv(w[q + '/conv3/weights'], device, post_shift=w[q + '/conv3/biases'], tc=tc)
                self.units.append(unit)
                d_in = depth
        s, b = fold_bn(w, p + '/postnorm')
        self.post = (_dev(s, device), _dev(b, device))
        self.out_dim = d_in


class ResNetPlan(object):
    """Forward plan for a fixed number of frames n (activation buffers are reused across chunks).

    `units=(lo, hi)` restricts the plan to bottleneck units [lo, hi) so the trunk can be run in two stages with
    different frame counts: the early blocks have thousands of tiles per layer at any batch, the late blocks
    (14x14 / 7x7 maps) only fill the 148 SMs when many frames are batched (wave quantisation).
    root=True prepends conv1 + pool1; tail=True appends postnorm + global mean.
    """

    def __init__(self, packed: PackedResNet, n, size=224, impl='auto', units=None, root=True, tail=True, next_pre=None,
                 next_has_shortcut=False):
        self.p = packed
        self.n = n
        self.size = size
        self.root, self.tail = root, tail
        dev = packed.device
        lo, hi = units if units is not None else (0, len(packed.units))
        H1 = size // 2                        # conv1 output (explicit pad 3, stride 2)
        H2 = (H1 + 1) // 2                    # pool1 SAME
        self.H1, self.H2 = H1, H2
        # spatial size / depth entering unit `lo`
        H, d_in = H2, 64
        for unit in packed.units[:lo]:
            H = (H - 1) // unit['stride'] + 1
            d_in = unit['depth']
        self.in_hw, self.in_depth = H, d_in
        # buffer sizes (floats per frame) needed by units [lo, hi)
        mx_io, mx_r = H * H * d_in, 0
        h = H
        for unit in packed.units[lo:hi]:
            ho = (h - 1) // unit['stride'] + 1
            mx_io = max(mx_io, h * h * unit['depth'] if 'shortcut' in unit else 0, ho * ho * unit['depth'])
            mx_r = max(mx_r, h * h * unit['base'])
            h = ho
        if root:
            mx_io = max(mx_io, H1 * H1 * 64)
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        # split mode: every conv reads its A operand as a pre-activated fp16 head/remainder pair written by the producing
        # epilogue (cp.async straight into the swizzled tile); fp32 copies exist only where a residual needs them
        self.split = impl in ('auto', 'tc3h') and all(
            c.tc == 'f16' and not c.gather for u in packed.units[lo:hi] for k, c in u.items() if isinstance(c, PackedConv))
        self.bufA = torch.empty(n * mx_io, **f32)
        self.bufB = torch.empty(n * mx_io, **f32)
        self.bufS = torch.empty(n * mx_io, **f32)
        self.ops = []
        self.conv1_op = None
        self.planes = None
        if root and packed.conv1_planes is not None and impl in ('auto', 'tc3h') and CONV1_PLANES and size % 2 == 0:
            self.planes = packed.conv1_planes.alloc_planes(n, size)
            self.conv1_op = packed.conv1_planes.bind(self.planes, n, size, self.bufS)
        elif root and packed.conv1.tc and impl != 'simt':
            self.conv1_op = packed.conv1.bind(self.bufS, n, size, size, self.bufS, in_ld=3, impl=impl)   # `in_` is set per run
        self.in_refs = []                     # (op, field) descriptor fields that read the stage input
        self.pool_split = None
        self.pool_f32_dead = False
        x, y = self.bufA, self.bufB
        units = packed.units[lo:hi]
        if self.split:
            def pair(count):
                return (torch.empty(max(1, count), **f16), torch.empty(max(1, count), **f16))
            xs, ys = pair(n * mx_io), pair(n * mx_io)
            r1, r2 = pair(n * mx_r), pair(n * mx_r)
            self.in_split = xs
            if root:
                self.pool_split = (units[0]['pre'][0], units[0]['pre'][1], xs)
                self.pool_f32_dead = DROP_DEAD_FP32 and 'shortcut' in units[0]
            self.out_split = None
            sub_ready = False                 # bufS already holds x[:, ::s, ::s] of this unit's input (written by the previous conv3)
            for ui, unit in enumerate(units):
                s = unit['stride']
                Ho = (H - 1) // s + 1
                if 'shortcut' in unit:
                    self.ops.append(unit['shortcut'].bind(None, n, H, H, self.bufS, inp_split=xs, impl=impl))
                    if ui == 0:
                        self.in_refs += [(self.ops[-1], 'in_hi', 0), (self.ops[-1], 'in_lo', 1)]
                    res, res_geom = self.bufS, (unit['depth'], Ho, Ho, 1)
                elif s > 1 and SUBSAMPLE_RES:
                    # strided identity shortcut: a dense subsampled copy so conv3's residual is row-aligned (TMA slab loads) -- written by
                    # the previous unit's conv3 epilogue when that unit is in this plan, else by one hd_subsample pass
                    if not sub_ready:
                        self.ops.append(SubsampleOp(x, self.bufS[:n * Ho * Ho * unit['depth']], n, H, unit['depth'], s))
                        if ui == 0:
                            self.in_refs.append((self.ops[-1], 'res', 2))
                    res, res_geom = self.bufS, (unit['depth'], Ho, Ho, 1)
                else:
                    res, res_geom = x, (unit['depth'], H, H, s)
                self.ops.append(unit['conv1'].bind(None, n, H, H, None, inp_split=xs, out_split=r1, impl=impl))
                if ui == 0:
                    self.in_refs += [(self.ops[-1], 'in_hi', 0), (self.ops[-1], 'in_lo', 1)]
                self.ops.append(unit['conv2'].bind(None, n, H, H, None, inp_split=r1, out_split=r2, impl=impl))
                last = ui == len(units) - 1
                nxt = units[ui + 1]['pre'] if not last else next_pre        # the next unit's pre-activation BN (+ReLU)
                osplit = ys if nxt is not None else None
                # the fp32 block output only feeds an IDENTITY shortcut: when the next unit changes depth (first unit of a block) its
                # shortcut is a conv of the pre-activation, and the fp32 copy would be written for nobody
                nxt_conv_shortcut = ('shortcut' in units[ui + 1]) if not last else bool(next_has_shortcut)
                y_out = None if (osplit is not None and nxt_conv_shortcut and DROP_DEAD_FP32) else y
                # the next unit is a strided identity unit: the only reader of this unit's fp32 output is that shortcut, x[:, ::s, ::s]
                # -- write just those pixels, densely, into bufS (free here: this unit's own residual is not in bufS)
                s_next = units[ui + 1]['stride'] if not last else 1
                sub_ready = bool(SUBSAMPLE_EPI and SUBSAMPLE_RES and not last and s_next > 1 and 'shortcut' not in units[ui + 1] and
                                 'shortcut' not in unit and s == 1 and y_out is not None and osplit is not None)
                if sub_ready:
                    y_out = self.bufS
                self.ops.append(unit['conv3'].bind(None, n, Ho, Ho, y_out, inp_split=r2, res=res, res_geom=res_geom, impl=impl,
                                                   out_split=osplit, post2=(nxt[0], nxt[1], 1) if nxt is not None else None,
                                                   out_subsample=s_next if sub_ready else 0))
                if ui == 0 and 'shortcut' not in unit and not (s > 1 and SUBSAMPLE_RES):
                    self.in_refs.append((self.ops[-1], 'res', 2))
                if last:
                    self.out_split = osplit
                x, y = y, x
                xs, ys = ys, xs
                H = Ho
                d_in = unit['depth']
        else:
            self.bufR1 = torch.empty(max(1, n * mx_r), **f32)
            self.bufR2 = torch.empty(max(1, n * mx_r), **f32)
            for ui, unit in enumerate(units):
                s = unit['stride']
                Ho = (H - 1) // s + 1
                pre = (unit['pre'][0], unit['pre'][1], 0, 1)
                if 'shortcut' in unit:
                    self.ops.append(unit['shortcut'].bind(x, n, H, H, self.bufS, pre=pre, impl=impl))
                    if ui == 0:
                        self.in_refs.append((self.ops[-1], 'in_', 2))
                    res, res_geom = self.bufS, (unit['depth'], Ho, Ho, 1)
                else:
                    res, res_geom = x, (unit['depth'], H, H, s)      # identity, or max_pool2d(1x1, stride) = subsample
                self.ops.append(unit['conv1'].bind(x, n, H, H, self.bufR1, pre=pre, impl=impl))
                if ui == 0:
                    self.in_refs.append((self.ops[-1], 'in_', 2))
                self.ops.append(unit['conv2'].bind(self.bufR1, n, H, H, self.bufR2, impl=impl))
                self.ops.append(unit['conv3'].bind(self.bufR2, n, Ho, Ho, y, res=res, res_geom=res_geom, impl=impl))
                if ui == 0 and 'shortcut' not in unit:
                    self.in_refs.append((self.ops[-1], 'res', 2))
                x, y = y, x
                H = Ho
                d_in = unit['depth']
        self.final = x
        self.final_hw = H * H
        self.out_hw, self.out_depth = H, d_in
        self.in_buf = self.bufA               # stage input when root=False

    def set_input(self, t, t_split=None):
        """Point the stage at an external input feature map [n, in_hw, in_hw, in_depth] (fp32 `t`, and in split mode its
        pre-activated fp16 pair `t_split`); no copy."""
        srcs = (t_split[0] if t_split else None, t_split[1] if t_split else None, t)
        for op, field, which in self.in_refs:
            src = srcs[which]
            if src is None:
                raise _lib.HDError('stage input %s missing' % field)
            op.rebind(field, src)
        for op in {id(o): o for o, _, _ in self.in_refs}.values():
            op.encode_act_maps()

    def set_output(self, t, t_split=None):
        """Let the last unit write its output feature map [n, out_hw, out_hw, out_depth] straight into `t` (and, in split
        mode, the next stage's pre-activated pair into `t_split`)."""
        op = self.ops[-1]
        if op.d.out:                          # (no fp32 output when the consumer's shortcut is a conv)
            op.rebind('out', t)
        if t_split is not None:
            op.rebind('out_hi', t_split[0])
            op.rebind('out_lo', t_split[1])
        op.encode_act_maps()
        self.final = t

    def run(self, images, out, stream=None):
        """root=True: images (n,size,size,3) contiguous float32 CUDA view; else `images` is ignored and the stage
        input must already be in `in_buf` ([n, in_hw, in_hw, in_depth]).  tail=True: out = phi (n,2048) view;
        else the stage output feature map is left in `self.final`."""
        st = current_stream() if stream is None else stream
        n, p = self.n, self.p
        if self.root:
            if self.planes is not None:
                if images is not None:        # None: the planes were filled by the caller (uint8 frames through hd_process_image_planes)
                    check(lib.hd_pack_conv1_planes(fptr(images), C.c_void_p(self.planes[0].data_ptr()), C.c_void_p(self.planes[1].data_ptr()),
                                                   n, self.size, self.size, self.planes[0].shape[2], st), 'hd_pack_conv1_planes')
                self.conv1_op.run(st)
            elif self.conv1_op is not None:
                self.conv1_op.d.in_ = images.data_ptr()
                self.conv1_op.run(st)
            else:
                check(lib.hd_conv1_7x7s2(fptr(images), fptr(p.conv1_w), fptr(p.conv1_b), fptr(self.bufS), n, self.size, self.size, st),
                      'hd_conv1_7x7s2')
            ps = self.pool_split
            check(lib.hd_maxpool3x3s2_same(fptr(self.bufS), None if self.pool_f32_dead else fptr(self.bufA), n, self.H1, self.H1, 64,
                                           fptr(ps[0]) if ps else None, fptr(ps[1]) if ps else None,
                                           C.c_void_p(ps[2][0].data_ptr()) if ps else None,
                                           C.c_void_p(ps[2][1].data_ptr()) if ps else None, st), 'hd_maxpool3x3s2_same')
        for op in self.ops:
            op.run(st)
        if self.tail:
            check(lib.hd_bnrelu_avgpool(fptr(self.final), fptr(p.post[0]), fptr(p.post[1]), fptr(out), n, self.final_hw,
                                        p.out_dim, st), 'hd_bnrelu_avgpool')

    @property
    def num_launches(self):
        return ((3 if self.planes is not None else 2) if self.root else 0) + len(self.ops) + (1 if self.tail else 0)


# ------------------------------------------------------------------------------------------------
# f_movie temporal encoder -- src/models.py:121-228
# ------------------------------------------------------------------------------------------------
class PackedFMovie(object):
    def __init__(self, w, device, num_conv_layers=3, tc=False):
        self.device = device
        self.blocks = []
        for i in range(num_conv_layers):
            name = 'block_%d' % i
            blk = {}
            for k in (1, 2):
                blk['gn%d' % k] = (_dev(w['AZ_FC_block_preact_gn%d%s/gamma' % (k, name)], device),
                                   _dev(w['AZ_FC_block_preact_gn%d%s/beta' % (k, name)], device))
                blk['conv%d' % k] = PackedConv(w['AZ_FC_block2_conv%d%s/weights' % (k, name)], device,
                                               post_shift=w['AZ_FC_block2_conv%d%s/biases' % (k, name)], pad=(1, 0), tc=tc)
            self.blocks.append(blk)
        self.C = self.blocks[0]['conv1'].Cin if self.blocks else 2048


class FMoviePlan(object):
    def __init__(self, packed: PackedFMovie, B, T, impl='auto'):
        self.p, self.B, self.T = packed, B, T
        dev, Cc = packed.device, packed.C
        self.gain = torch.empty((B, Cc), dtype=torch.float32, device=dev)
        self.offset = torch.empty((B, Cc), dtype=torch.float32, device=dev)
        self.mid = torch.empty((B, T, Cc), dtype=torch.float32, device=dev)
        self.bufs = [torch.empty((B, T, Cc), dtype=torch.float32, device=dev) for _ in range(2)]
        self.impl = impl
        self._bound_for = None

    def _bind_fast(self, x):
        """impl auto / tc3h: GroupNorm + ReLU + split in one kernel (hd_groupnorm_relu_split), then the temporal conv reads its A
        operand as the pre-split pair (cp.async producer, TMA epilogue) -- the register-staged GN prologue path ran at ~1/4 of
        that speed (profiles/r02_launches_step_c3_before_heads.csv: 248 us per conv at 640 rows)."""
        dev, Cc = self.p.device, self.p.C
        B, T = self.B, self.T
        if not hasattr(self, 'act'):
            self.act = (torch.empty((B * T, Cc), dtype=torch.float16, device=dev), torch.empty((B * T, Cc), dtype=torch.float16, device=dev))
        steps, cur = [], x
        for i, blk in enumerate(self.p.blocks):
            out = self.bufs[i % 2]
            steps.append(('gns', cur, blk['gn1']))
            steps.append(('conv', blk['conv1'].bind(None, B, T, 1, self.mid, inp_split=self.act, impl=self.impl)))
            steps.append(('gns', self.mid, blk['gn2']))
            steps.append(('conv', blk['conv2'].bind(None, B, T, 1, out, inp_split=self.act, res=cur, res_geom=(Cc, T, 1, 1), impl=self.impl)))
            cur = out
        self.steps, self.out = steps, cur
        self._bound_for = x.data_ptr()

    def _bind(self, x):
        if self.impl in ('auto', 'tc3h') and self.p.blocks and all(b[k].tc == 'f16' for b in self.p.blocks for k in ('conv1', 'conv2')) \
                and self.T * (self.p.C // GN_GROUPS) <= 1280 and FAST_HEADS:
            return self._bind_fast(x)
        steps = []
        cur = x
        pre = (self.gain, self.offset, self.p.C, 1)
        B, T = self.B, self.T
        for i, blk in enumerate(self.p.blocks):
            out = self.bufs[i % 2]
            steps.append(('gn', cur, blk['gn1']))
            steps.append(('conv', blk['conv1'].bind(cur, B, T, 1, self.mid, pre=pre, impl=self.impl)))
            steps.append(('gn', self.mid, blk['gn2']))
            steps.append(('conv', blk['conv2'].bind(self.mid, B, T, 1, out, pre=pre, res=cur,
                                                    res_geom=(self.p.C, T, 1, 1), impl=self.impl)))
            cur = out
        self.steps, self.out = steps, cur
        self._bound_for = x.data_ptr()

    def run(self, x, stream=None):
        """x (B,T,C) contiguous float32 CUDA -> (B,T,C) (a plan-owned buffer; x itself if there are no blocks)."""
        st = current_stream() if stream is None else stream
        if self._bound_for != x.data_ptr():
            self._bind(x)
        for s in self.steps:
            if s[0] == 'gn':
                check(lib.hd_groupnorm_stats(fptr(s[1]), fptr(s[2][0]), fptr(s[2][1]), fptr(self.gain), fptr(self.offset),
                                             self.B, self.T, self.p.C, GN_GROUPS, GN_EPS, st), 'hd_groupnorm_stats')
            elif s[0] == 'gns':
                check(lib.hd_groupnorm_relu_split(fptr(s[1]), fptr(s[2][0]), fptr(s[2][1]), C.c_void_p(self.act[0].data_ptr()),
                                                  C.c_void_p(self.act[1].data_ptr()), self.B, self.T, self.p.C, GN_GROUPS, GN_EPS, st),
                      'hd_groupnorm_relu_split')
            else:
                s[1].run(st)
        return self.out

    @property
    def num_launches(self):
        return 4 * len(self.p.blocks)


# ------------------------------------------------------------------------------------------------
# IEF regressor -- src/models.py:80-116, 299-415
# ------------------------------------------------------------------------------------------------
class PackedIEFHead(object):
    def __init__(self, w, scope, device, feat=2048, tc=False):
        q = scope + '/3D_module'
        W1 = np.asarray(w[q + '/fc1/weights'], np.float32)
        self.d = W1.shape[0] - feat
        self.feat = feat
        # state = concat[phi, theta] (models.py:402): split fc1 so phi.W1[:feat] is computed once per window
        self.fc1_phi = PackedConv(W1[:feat], device, post_shift=w[q + '/fc1/biases'], tc=tc)
        self.fc1_theta = PackedConv(W1[feat:], device, post_relu=True)
        self.fc2 = PackedConv(w[q + '/fc2/weights'], device, post_shift=w[q + '/fc2/biases'], post_relu=True, tc=tc)
        self.fc3 = PackedConv(w[q + '/fc3/weights'], device, post_shift=w[q + '/fc3/biases'], tc=tc)   # ragged N: scalar epilogue path


class PackedIEF(object):
    def __init__(self, w, device, scope='single_view_ief', delta_t_values=(-5, 5), tc=False):
        self.device = device
        self.main = PackedIEFHead(w, scope, device, tc=tc)
        self.deltas = {}
        for dt in delta_t_values:
            dt = int(dt)
            if dt == 0:
                continue
            sc = scope + ('_future%d' % dt if dt > 0 else '_past%d' % abs(dt))
            self.deltas[dt] = PackedIEFHead(w, sc, device, tc=tc)
        self.mean_param = _dev(np.asarray(w['mean_param'], np.float32).reshape(1, 85), device)


class IEFPlan(object):
    """call_hmr_ief for N rows: main 85-d head + 72-d delta heads started from the main prediction
    (use_delta_from_pred=True, use_optcam=True as wired by tester.py:196-207)."""

    def __init__(self, packed: PackedIEF, N, num_stage=3, delta_keys=None, impl='auto'):
        self.p, self.N, self.num_stage = packed, N, num_stage
        dev = packed.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.P = torch.empty((N, 1024), **f32)
        self.h1 = torch.empty((N, 1024), **f32)
        self.h2 = torch.empty((N, 1024), **f32)
        self.theta = torch.empty((N, 85), **f32)
        self.delta_keys = sorted(packed.deltas.keys()) if delta_keys is None else [k for k in delta_keys if k != 0]
        D = max(1, len(self.delta_keys))
        self.delta_all = torch.empty((N, D, 85), **f32)          # [N, D, 85]: the stacking of tester.py:252-253
        self.delta_out = {dt: self.delta_all[:, i, :] for i, dt in enumerate(self.delta_keys)}
        self.impl = impl
        self._bound_for = None
        heads = [packed.main] + [packed.deltas[k] for k in self.delta_keys]
        self.fast = FAST_HEADS and impl in ('auto', 'tc3h') and all(h.fc1_phi.tc == 'f16' and h.fc2.tc == 'f16' for h in heads)
        if self.fast:
            f16 = dict(dtype=torch.float16, device=dev)
            self.phi_split = (torch.empty((N, heads[0].feat), **f16), torch.empty((N, heads[0].feat), **f16))
            self.h1_split = (torch.empty((N, 1024), **f16), torch.empty((N, 1024), **f16))

    def _head_ops_fast(self, head, start_view, state_view, ld):
        """impl auto / tc3h.  phi arrives once as a pre-split pair (hd_split_f16); per stage: hd_ief_fc1_theta (K = 85 / 72, writes h1
        pre-split) -> fc2 on the tensor cores (cp.async producer, TMA epilogue) -> hd_ief_fc3 (D = 85 / 72 + the IEF update).  The
        generic kernels ran fc1-theta on 40 SIMT blocks (50 us) and fc3 as ONE 128-row tile per 128 poses on 5 CTAs (80 us)."""
        N = self.N
        ops = [('conv', head.fc1_phi.bind(None, N, 1, 1, self.P, inp_split=self.phi_split, impl=self.impl))]
        for s in range(self.num_stage):
            prev = start_view if s == 0 else state_view
            prev_ld = start_view.stride(0) if s == 0 else ld
            ops.append(('fc1t', prev, prev_ld, head))
            ops.append(('conv', head.fc2.bind(None, N, 1, 1, self.h2, inp_split=self.h1_split, impl=self.impl)))
            ops.append(('fc3', prev, prev_ld, state_view, ld, head))
        return ops

    def _run_ops(self, ops, st):
        N = self.N
        for op in ops:
            kind = op[0] if isinstance(op, tuple) else None
            if kind is None:
                op.run(st)
            elif kind == 'conv':
                op[1].run(st)
            elif kind == 'fc1t':
                _, prev, prev_ld, head = op
                check(lib.hd_ief_fc1_theta(fptr(self.P), fptr(prev), prev_ld, fptr(head.fc1_theta.w_kn), head.d, 1024,
                                           C.c_void_p(self.h1_split[0].data_ptr()), C.c_void_p(self.h1_split[1].data_ptr()), None, N, st),
                      'hd_ief_fc1_theta')
            else:
                _, prev, prev_ld, out, out_ld, head = op
                check(lib.hd_ief_fc3(fptr(self.h2), fptr(head.fc3.w_kn), fptr(head.fc3.post_shift), fptr(prev), prev_ld, fptr(out), out_ld, N,
                                     1024, head.d, st), 'hd_ief_fc3')

    def _head_ops(self, head, phi, start_view, state_view, ld):
        """ops for one hmr_ief: start_view = theta_prev of stage 0, state_view = in-place theta afterwards."""
        if self.fast:
            return self._head_ops_fast(head, start_view, state_view, ld)
        N = self.N
        ops = [head.fc1_phi.bind(phi, N, 1, 1, self.P, impl=self.impl)]
        for s in range(self.num_stage):
            prev = start_view if s == 0 else state_view
            prev_ld = start_view.stride(0) if s == 0 else ld
            ops.append(head.fc1_theta.bind(prev, N, 1, 1, self.h1, in_ld=prev_ld, res=self.P, res_geom=(1024, 1, 1, 1), impl='simt'))
            ops.append(head.fc2.bind(self.h1, N, 1, 1, self.h2, impl=self.impl))
            ops.append(head.fc3.bind(self.h2, N, 1, 1, state_view, out_ld=ld, res=prev, res_geom=(prev_ld, 1, 1, 1), impl=self.impl))
        return ops

    def _bind(self, phi, theta0):
        self.main_ops = self._head_ops(self.p.main, phi, theta0, self.theta, 85)
        self.delta_ops = {}
        for dt in self.delta_keys:
            view = self.delta_out[dt][:, 3:75]
            self.delta_ops[dt] = self._head_ops(self.p.deltas[dt], phi, view, view, view.stride(0))
        self._bound_for = (phi.data_ptr(), theta0.data_ptr())

    def run_main(self, phi, theta0, stream=None):
        """Main 85-d head only: phi (N,2048), theta0 (N,85) contiguous -> theta (N,85)."""
        st = current_stream() if stream is None else stream
        if self._bound_for != (phi.data_ptr(), theta0.data_ptr()):
            self._bind(phi, theta0)
        if self.fast:
            check(lib.hd_split_f16(fptr(phi), C.c_void_p(self.phi_split[0].data_ptr()), C.c_void_p(self.phi_split[1].data_ptr()),
                                   phi.numel(), st), 'hd_split_f16')
        self._run_ops(self.main_ops, st)
        return self.theta

    def run_deltas(self, stream=None):
        """Delta heads, started from the main prediction (run_main must have run): {dt: (N,85) view of delta_all[:, i]}."""
        st = current_stream() if stream is None else stream
        for dt in self.delta_keys:
            check(lib.hd_ief_delta_init(fptr(self.theta), fptr(self.delta_out[dt]), self.delta_out[dt].stride(0), self.N, st),
                  'hd_ief_delta_init')
            self._run_ops(self.delta_ops[dt], st)
        return self.delta_out

    def run(self, phi, theta0, stream=None):
        """phi (N,2048), theta0 (N,85) contiguous -> (theta (N,85), {dt: (N,85) view of delta_all[:, i]})."""
        theta = self.run_main(phi, theta0, stream)
        return theta, self.run_deltas(stream)

    @property
    def num_launches(self):
        per = 1 + 3 * self.num_stage
        return per + len(self.delta_keys) * (per + 1) + (1 if self.fast else 0)


def run_ief_head(head: PackedIEFHead, phi, start, num_stage=3, impl='auto', stream=None, out=None):
    """hmr_ief for one head from an arbitrary start: phi (N,feat), start (N,d) (unit inner stride) -> (N,d).

    Generic (binds descriptors on the fly); the Tester path uses the cached IEFPlan instead.
    """
    st = current_stream() if stream is None else stream
    N, d = phi.shape[0], head.d
    if start.shape[0] != N or start.shape[1] != d or start.stride(1) != 1:
        raise _lib.HDError('hmr_ief: omega_start must be (N,%d) with unit inner stride' % d)
    f32 = dict(dtype=torch.float32, device=phi.device)
    P, h1, h2 = torch.empty((N, 1024), **f32), torch.empty((N, 1024), **f32), torch.empty((N, 1024), **f32)
    theta = torch.empty((N, d), **f32) if out is None else out
    ld = theta.stride(0)
    head.fc1_phi.bind(phi, N, 1, 1, P, impl=impl).run(st)
    for s in range(num_stage):
        prev = start if s == 0 else theta
        pld = prev.stride(0)
        head.fc1_theta.bind(prev, N, 1, 1, h1, in_ld=pld, res=P, res_geom=(1024, 1, 1, 1), impl='simt').run(st)
        head.fc2.bind(h1, N, 1, 1, h2, impl=impl).run(st)
        head.fc3.bind(h2, N, 1, 1, theta, out_ld=ld, res=prev, res_geom=(pld, 1, 1, 1), impl=impl).run(st)
    return theta
