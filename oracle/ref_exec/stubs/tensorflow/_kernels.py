"""Dense kernels behind the stand-in's NN ops (float32, NHWC): torch-CPU conv, numpy for the rest."""
import numpy as np


def same_pads(size, k, stride):
    """TF 'SAME': out = ceil(size / stride); pad_total = max((out-1)*stride + k - size, 0); pad_before = pad_total // 2."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def conv2d_nhwc(x, w, stride, padding):
    import torch
    import torch.nn.functional as F
    KH, KW = w.shape[0], w.shape[1]
    xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).permute(0, 3, 1, 2)
    if padding == 'SAME':
        pt, pb = same_pads(x.shape[1], KH, stride)
        pl, pr = same_pads(x.shape[2], KW, stride)
        xt = F.pad(xt, (pl, pr, pt, pb))
    elif padding != 'VALID':
        raise ValueError(padding)
    wt = torch.from_numpy(np.ascontiguousarray(w, np.float32)).permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(xt.contiguous(), wt, stride=stride)
    return np.ascontiguousarray(y.permute(0, 2, 3, 1).numpy())


def max_pool_nhwc(x, k, stride, padding):
    KH, KW = k
    if padding == 'SAME':
        pt, pb = same_pads(x.shape[1], KH, stride)
        pl, pr = same_pads(x.shape[2], KW, stride)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)), mode='constant', constant_values=-np.inf)
    Ho = (x.shape[1] - KH) // stride + 1
    Wo = (x.shape[2] - KW) // stride + 1
    out = None
    for ky in range(KH):
        for kx in range(KW):
            v = x[:, ky:ky + (Ho - 1) * stride + 1:stride, kx:kx + (Wo - 1) * stride + 1:stride, :]
            out = v.copy() if out is None else np.maximum(out, v)
    return out
