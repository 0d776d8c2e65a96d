"""GPU crop pre-processing: the step immediately before the path.

`process_image` of the reference (src/evaluation/run_video.py:56-107, with `resize_img`, src/util/common.py:7-14) turns a
uint8 video frame and a bbox [cx, cy, scale] into the 224x224x3 float crop in [-1, 1] that `Tester.predict` consumes:
scale to [-1,1] -> cv2.resize (bilinear) by `scale` -> edge-pad 224 -> crop around the scaled centre.  Here the integer
geometry is computed on the host exactly as the reference computes it (float64 numpy), and ONE kernel produces the crop
straight from the uint8 frame (hd_process_image): the scaled / padded intermediates never exist, and the host uploads
1 byte per sample instead of 4.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, current_stream

IMG_SIZE = 224


def crop_geometry(im_shape, bbox_param, img_size=IMG_SIZE):
    """Host half of process_image: everything that is integer bookkeeping (run_video.py:69-100).

    Returns dict(new_size [h,w], scale_factors [fy,fx], center (after crop), start_pt (in the PADDED scaled image, like the
    reference), origin (x0,y0) = top-left of the crop in UNPADDED scaled-image coordinates, scale)."""
    bbox_param = np.asarray(bbox_param, np.float64)
    center, scale = bbox_param[:2], bbox_param[2]
    shape = np.array(im_shape[0:2])
    new_size = (np.floor(shape * scale)).astype(int)                               # common.py:8
    if new_size[0] < 1 or new_size[1] < 1:
        raise ValueError('bbox scale %g leaves an empty image' % scale)
    factors = [new_size[0] / float(shape[0]), new_size[1] / float(shape[1])]       # [y, x]  common.py:11-13
    center_scaled = np.round(center * factors).astype(int)                         # run_video.py:75 (x*fy, y*fx as the reference does)
    center_scaled = center_scaled + img_size
    margin = img_size // 2
    start_pt = (center_scaled - margin).astype(int)
    end_pt = (center_scaled + margin).astype(int)
    width, height = new_size[1] + 2 * img_size, new_size[0] + 2 * img_size
    if start_pt[0] < 0 or start_pt[1] < 0 or end_pt[0] > width or end_pt[1] > height:
        # the reference would return a crop smaller than img_size here (run_video.py:91-94), which its fixed-shape graph cannot
        # consume; we keep the static shape and say so
        raise ValueError('bbox centre %s is more than one crop away from the frame: the reference yields a ragged crop' % (center,))
    return {'new_size': new_size, 'scale_factors': factors, 'center': center_scaled - start_pt, 'start_pt': start_pt,
            'origin': (int(start_pt[0] - img_size), int(start_pt[1] - img_size)), 'scale': scale,
            'im_shape': [img_size, img_size]}


def process_images(frames, bbox_params, img_size=IMG_SIZE, out=None, split_out=None):
    """frames: (N,H,W,3) uint8 (CUDA tensor, or host tensor / ndarray -> uploaded as uint8); bbox_params (N,3).

    -> (crops (N,S,S,3) float32 CUDA in [-1,1], [geometry dict per frame])."""
    if isinstance(frames, np.ndarray):
        frames = torch.from_numpy(np.ascontiguousarray(frames))
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
        raise _lib.HDError('process_images: frames must be (N,H,W,3) uint8')
    if not torch.cuda.is_available():
        raise _lib.HDError('process_images needs a CUDA device: there is no CPU fallback')
    if not frames.is_cuda:
        frames = frames.cuda(non_blocking=True)
    frames = frames.contiguous()
    N, H, W = frames.shape[0], frames.shape[1], frames.shape[2]
    bbox_params = np.asarray(bbox_params, np.float64).reshape(N, 3)
    geoms = [crop_geometry((H, W), bbox_params[i], img_size) for i in range(N)]
    g = np.array([[q['new_size'][0], q['new_size'][1], q['origin'][0], q['origin'][1]] for q in geoms], np.int32)
    g_dev = torch.from_numpy(g).to(frames.device, non_blocking=True)
    if out is None:
        out = torch.empty((N, img_size, img_size, 3), dtype=torch.float32, device=frames.device)
    elif tuple(out.shape) != (N, img_size, img_size, 3) or out.dtype != torch.float32 or not out.is_contiguous():
        raise _lib.HDError('process_images: out must be a contiguous float32 (N,%d,%d,3) CUDA tensor' % (img_size, img_size))
    check(lib.hd_process_image(C.c_void_p(frames.data_ptr()), N, H, W, C.c_void_p(g_dev.data_ptr()), C.c_void_p(out.data_ptr()),
                               img_size, None, None, 0, current_stream()), 'hd_process_image')
    return out, geoms


def geometry_table(im_shape, bbox_params, img_size=IMG_SIZE, with_infos=False):
    """int32 [N,4] = {Hs, Ws, x0, y0} rows for hd_process_image -- `crop_geometry` for a whole track of same-size frames in one
    numpy pass (same float64 formulas, same integers) -- and, on request, the per-frame info dicts."""
    b = np.asarray(bbox_params, np.float64).reshape(-1, 3)
    shape = np.array(im_shape[0:2])
    new_size = np.floor(shape[None, :] * b[:, 2:3]).astype(int)                    # (N, [h, w])
    if (new_size < 1).any():
        raise ValueError('a bbox scale leaves an empty image')
    factors = new_size / shape[None, :].astype(np.float64)                         # [fy, fx]
    center_scaled = np.round(b[:, :2] * factors).astype(int) + img_size            # (x*fy, y*fx) like run_video.py:75
    start = center_scaled - img_size // 2
    end = center_scaled + img_size // 2
    if (start < 0).any() or (end[:, 0] > new_size[:, 1] + 2 * img_size).any() or (end[:, 1] > new_size[:, 0] + 2 * img_size).any():
        raise ValueError('a bbox centre is more than one crop away from the frame: the reference yields a ragged crop')
    g = np.concatenate([new_size, start - img_size], axis=1).astype(np.int32)      # Hs, Ws, x0, y0
    infos = [crop_geometry(im_shape, bb, img_size) for bb in b] if with_infos else None
    return g, infos
