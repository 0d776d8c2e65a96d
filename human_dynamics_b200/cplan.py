"""ctypes face of the library-owned network plans (include/hd_b200.h: hd_resnet50_* / hd_fmovie_* / hd_ief_*).

These are the entry points a C / C++ consumer of libhd_b200.so calls instead of re-implementing the layer plans of nets.py
(INTEGRATION.md B).  The Python engine keeps its own plans (two trunk stages, CUDA-graph capture, delta-head overlap); this
module exists so that the C path is exercised by the test suite and stays bit-identical to them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, current_stream


class _WeightSource(object):
    """hd_weight_fn over a dict of TF-named numpy arrays; keeps every array it handed out alive."""

    def __init__(self, weights):
        self.weights = weights
        self.alive = {}
        self.missing = []

        def get(user, name, numel):
            key = name.decode()
            arr = self.weights.get(key)
            if arr is None:
                self.missing.append(key)
                return None
            a = self.alive.get(key)
            if a is None:
                a = self.alive[key] = np.ascontiguousarray(arr, np.float32)
            numel[0] = a.size
            return a.ctypes.data
        self.fn = _lib.WEIGHT_FN(get)


class _CNet(object):
    def __init__(self):
        self.handle = C.c_void_p()

    def _created(self, rc, what, src):
        if rc:
            msg = lib.hd_last_error().decode()
            raise _lib.HDError('%s failed: %s [%s]' % (what, lib.hd_status_string(rc).decode(), msg))
        src.alive.clear()                      # the plan copied what it needed

    @property
    def num_launches(self):
        return int(lib.hd_net_num_launches(self.handle))

    def close(self):
        if self.handle:
            lib.hd_net_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CResNet50(_CNet):
    """encoder_resnet (src/models.py:50-77) for a fixed frame count."""

    def __init__(self, weights, n_frames, size=224):
        super().__init__()
        src = _WeightSource(weights)
        self._created(lib.hd_resnet50_create(C.cast(src.fn, C.c_void_p), None, n_frames, size, C.byref(self.handle)), 'hd_resnet50_create', src)
        self.n, self.size = n_frames, size

    def __call__(self, images, out=None):
        if tuple(images.shape) != (self.n, self.size, self.size, 3) or not images.is_contiguous():
            raise _lib.HDError('images must be a contiguous (%d,%d,%d,3) tensor' % (self.n, self.size, self.size))
        phi = torch.empty((self.n, 2048), dtype=torch.float32, device=images.device) if out is None else out
        check(lib.hd_resnet50_forward(self.handle, _lib.fptr(images), _lib.fptr(phi), current_stream()), 'hd_resnet50_forward')
        return phi


class CFMovie(_CNet):
    """az_fc2_groupnorm (src/models.py:121-228)."""

    def __init__(self, weights, B, T, num_conv_layers=3):
        super().__init__()
        src = _WeightSource(weights)
        self._created(lib.hd_fmovie_create(C.cast(src.fn, C.c_void_p), None, B, T, num_conv_layers, C.byref(self.handle)), 'hd_fmovie_create', src)
        self.B, self.T = B, T

    def __call__(self, phi, out=None):
        if tuple(phi.shape) != (self.B, self.T, 2048) or not phi.is_contiguous():
            raise _lib.HDError('phi must be a contiguous (%d,%d,2048) tensor' % (self.B, self.T))
        y = torch.empty_like(phi) if out is None else out
        check(lib.hd_fmovie_forward(self.handle, _lib.fptr(phi), _lib.fptr(y), current_stream()), 'hd_fmovie_forward')
        return y


class CIEF(_CNet):
    """call_hmr_ief (src/models.py:299-415) as wired by tester.py:196-207."""

    def __init__(self, weights, N, delta_t_values=(-5, 5)):
        super().__init__()
        src = _WeightSource(weights)
        dts = [int(d) for d in delta_t_values if int(d) != 0]
        arr = (C.c_int * max(1, len(dts)))(*dts)
        self._created(lib.hd_ief_create(C.cast(src.fn, C.c_void_p), None, N, arr, len(dts), C.byref(self.handle)), 'hd_ief_create', src)
        self.N, self.delta_keys = N, sorted(dts)

    def __call__(self, phi):
        if tuple(phi.shape) != (self.N, 2048) or not phi.is_contiguous():
            raise _lib.HDError('phi must be a contiguous (%d,2048) tensor' % self.N)
        theta = torch.empty((self.N, 85), dtype=torch.float32, device=phi.device)
        D = len(self.delta_keys)
        deltas = torch.empty((self.N, max(1, D), 85), dtype=torch.float32, device=phi.device)
        check(lib.hd_ief_forward(self.handle, _lib.fptr(phi), _lib.fptr(theta), _lib.fptr(deltas) if D else None, current_stream()),
              'hd_ief_forward')
        return theta, {dt: deltas[:, i] for i, dt in enumerate(self.delta_keys)}
