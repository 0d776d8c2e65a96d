"""ORACLE (test infrastructure, not product): CPU restatement of the HMMR networks.

PARITY STATUS: pinned to the reference's SOURCE, unpinned against TensorFlow's KERNELS.  TensorFlow 1.8
(requirements.txt:11) cannot be imported here and the reference holds no golden vectors for this path
(SURVEY.md 8c).  The first-party wiring (src/models.py, omega.py, tester.py) is checked against vectors made
by executing those reference files themselves over a numpy TensorFlow stand-in (oracle/ref_exec/,
tests/golden/ref_exec_v1.npz, tests/test_ref_exec.py: all 14 Tester.predict keys within 8.5e-6).  The
arithmetic that lives inside tf.contrib.slim / tf.contrib.layers of TF 1.8 [TF-ext] is not part of
/root/reference; it is restated from its published semantics here AND in the stand-in, so it stays an
assumption.  Each such ASSUMPTION is listed here and has its own focused unit test in tests/test_oracle_nets.py:

  A1  SAME padding: out=ceil(in/s); pad_total=max((out-1)*s+k-in,0); extra pad at the END.
  A2  resnet_utils.conv2d_same: stride 1 -> SAME; stride>1 -> explicit pad (k-1)//2 each side + VALID.
  A3  resnet_v2_50 topology: blocks (64,3,s2),(128,4,s2),(256,6,s2),(512,3,s1); stride on LAST unit;
      root conv1 7x7/2 with bias, no norm/act; pool1 3x3/2 SAME; tail postnorm BN+ReLU, mean over H,W.
  A4  bottleneck_v2: preact=relu(BN(x)); shortcut = x | x[:, ::s, ::s] | conv1x1(preact)+bias;
      conv1 1x1 (no bias)+BN+ReLU; conv2 3x3 conv2d_same (no bias)+BN+ReLU; conv3 1x1 + bias.
  A5  BatchNorm inference: gamma*(x-mean)/sqrt(var+1e-5)+beta.
  A7  group_norm: groups=32, eps=1e-6, biased variance over (T,1,C/32) per (clip, group), per-channel
      gamma/beta:  gain=rsqrt(var+eps)*gamma; offset=beta-mean*gain; y=x*gain+offset.
  A8  slim.fully_connected = relu(xW+b) unless activation_fn=None; dropout(is_training=False)=identity;
      tf.contrib.layers.conv2d(activation_fn=None) = conv + bias.

All functions take / return NHWC (channels-last) tensors like the reference; `weights` is a
dict of numpy arrays keyed by TF variable names (HWIO convs, [in,out] FCs).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

RESNET_BLOCKS = ((64, 3, 2), (128, 4, 2), (256, 6, 2), (512, 3, 1))
BN_EPS = 1e-5      # resnet_arg_scope default batch_norm_epsilon
GN_EPS = 1e-6      # tf.contrib.layers.group_norm default epsilon
GN_GROUPS = 32


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a)).to(dtype)


def same_pad(in_size, k, s):
    """A1: TF SAME padding -> (pad_before, pad_after, out)."""
    out = int(math.ceil(in_size / s))
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2, out


def conv2d_nhwc(x, w_hwio, bias=None, stride=1, padding='SAME'):
    """tf conv2d on NHWC input with HWIO filter.  padding: 'SAME' | 'VALID' | ((pt,pb),(pl,pr))."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    if padding == 'SAME':
        pt, pb, _ = same_pad(x.shape[1], kh, stride)
        pl, pr, _ = same_pad(x.shape[2], kw, stride)
    elif padding == 'VALID':
        pt = pb = pl = pr = 0
    else:
        (pt, pb), (pl, pr) = padding
    xc = x.permute(0, 3, 1, 2)
    if pt or pb or pl or pr:
        xc = F.pad(xc, (pl, pr, pt, pb))
    wc = w_hwio.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(xc, wc, bias=bias, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d_same(x, w_hwio, stride, bias=None):
    """A2: resnet_utils.conv2d_same."""
    k = w_hwio.shape[0]
    if stride == 1:
        return conv2d_nhwc(x, w_hwio, bias, 1, 'SAME')
    pad_total = k - 1
    pb = pad_total // 2
    pe = pad_total - pb
    return conv2d_nhwc(x, w_hwio, bias, stride, ((pb, pe), (pb, pe)))


def max_pool_same_3x3_s2(x):
    """pool1: slim.max_pool2d(3, stride=2, padding='SAME'); padded cells are ignored (-inf)."""
    pt, pb, _ = same_pad(x.shape[1], 3, 2)
    pl, pr, _ = same_pad(x.shape[2], 3, 2)
    xc = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float('-inf'))
    return F.max_pool2d(xc, 3, 2).permute(0, 2, 3, 1).contiguous()


def batch_norm_inf(x, w, prefix, dtype):
    """A5."""
    g = _t(w[prefix + '/gamma'], dtype)
    b = _t(w[prefix + '/beta'], dtype)
    m = _t(w[prefix + '/moving_mean'], dtype)
    v = _t(w[prefix + '/moving_variance'], dtype)
    return g * (x - m) / torch.sqrt(v + BN_EPS) + b


def bottleneck_v2(x, w, q, base, stride, dtype):
    """A4.  q = '.../bottleneck_v2' scope prefix."""
    depth = 4 * base
    d_in = x.shape[-1]
    preact = torch.relu(batch_norm_inf(x, w, q + '/preact', dtype))
    if depth == d_in:
        shortcut = x if stride == 1 else x[:, ::stride, ::stride, :]
    else:
        shortcut = conv2d_nhwc(preact, _t(w[q + '/shortcut/weights'], dtype),
                               _t(w[q + '/shortcut/biases'], dtype), stride, 'VALID')
    r = conv2d_nhwc(preact, _t(w[q + '/conv1/weights'], dtype), None, 1, 'SAME')
    r = torch.relu(batch_norm_inf(r, w, q + '/conv1/BatchNorm', dtype))
    r = conv2d_same(r, _t(w[q + '/conv2/weights'], dtype), stride)
    r = torch.relu(batch_norm_inf(r, w, q + '/conv2/BatchNorm', dtype))
    r = conv2d_nhwc(r, _t(w[q + '/conv3/weights'], dtype), _t(w[q + '/conv3/biases'], dtype), 1, 'SAME')
    return shortcut + r


def encoder_resnet(x, weights, dtype=torch.float32, blocks=RESNET_BLOCKS, return_endpoints=False):
    """src/models.py:50-77 -> slim resnet_v2_50(num_classes=None) + squeeze.  x NHWC -> (N, 2048)."""
    x = _t(x, dtype)
    p = 'resnet_v2_50'
    ends = {}
    net = conv2d_same(x, _t(weights[p + '/conv1/weights'], dtype), 2, _t(weights[p + '/conv1/biases'], dtype))
    ends['conv1'] = net
    net = max_pool_same_3x3_s2(net)
    ends['pool1'] = net
    for b, (base, units, bstride) in enumerate(blocks, start=1):
        for u in range(1, units + 1):
            s = bstride if u == units else 1
            net = bottleneck_v2(net, weights, '%s/block%d/unit_%d/bottleneck_v2' % (p, b, u), base, s, dtype)
        ends['block%d' % b] = net
    net = torch.relu(batch_norm_inf(net, weights, p + '/postnorm', dtype))
    net = net.mean(dim=(1, 2))                       # global pool (keep_dims) + squeeze models.py:75
    if return_endpoints:
        return net, ends
    return net


def group_norm_tf(x4, gamma, beta, groups=GN_GROUPS, eps=GN_EPS):
    """A7: tf.contrib.layers.group_norm(x[N,T,1,C], channels_axis=-1, reduction_axes=(-3,-2))."""
    n, t, one, c = x4.shape
    xg = x4.reshape(n, t, one, groups, c // groups)
    mean = xg.mean(dim=(1, 2, 4), keepdim=True)
    var = ((xg - mean) ** 2).mean(dim=(1, 2, 4), keepdim=True)      # biased, two-pass
    gain = torch.rsqrt(var + eps) * gamma.reshape(1, 1, 1, groups, c // groups)
    offset = beta.reshape(1, 1, 1, groups, c // groups) - mean * gain
    return (xg * gain + offset).reshape(n, t, one, c)


def az_fc_block2(net_input, weights, name, dtype):
    """src/models.py:144-228 with use_groupnorm=True.  net_input (B,T,C)."""
    x4 = net_input[:, :, None, :]                                                        # :152
    y = group_norm_tf(x4, _t(weights['AZ_FC_block_preact_gn1%s/gamma' % name], dtype),
                      _t(weights['AZ_FC_block_preact_gn1%s/beta' % name], dtype))        # :155-161
    y = torch.relu(y)                                                                    # :171
    y = conv2d_nhwc(y, _t(weights['AZ_FC_block2_conv1%s/weights' % name], dtype),
                    _t(weights['AZ_FC_block2_conv1%s/biases' % name], dtype), 1, 'SAME')  # :173-184
    y = group_norm_tf(y, _t(weights['AZ_FC_block_preact_gn2%s/gamma' % name], dtype),
                      _t(weights['AZ_FC_block_preact_gn2%s/beta' % name], dtype))        # :188-194
    y = torch.relu(y)                                                                    # :204
    y = conv2d_nhwc(y, _t(weights['AZ_FC_block2_conv2%s/weights' % name], dtype),
                    _t(weights['AZ_FC_block2_conv2%s/biases' % name], dtype), 1, 'SAME')  # :209-221
    return y[:, :, 0, :] + net_input                                                     # :224-226


def az_fc2_groupnorm(net, weights, num_conv_layers, dtype=torch.float32):
    """src/models.py:121-141 ("f_movie")."""
    net = _t(net, dtype)
    for i in range(num_conv_layers):
        net = az_fc_block2(net, weights, 'block_%d' % i, dtype)
    return net


def fc2_res(phi, weights, dtype=torch.float32, name='fc2_res'):
    """src/models.py:270-296 (pred_mode='hal')."""
    phi = _t(phi, dtype)
    net = torch.relu(phi @ _t(weights[name + '/fc1/weights'], dtype) + _t(weights[name + '/fc1/biases'], dtype))
    net = torch.relu(net @ _t(weights[name + '/fc2/weights'], dtype) + _t(weights[name + '/fc2/biases'], dtype))
    net = net @ _t(weights[name + '/fc3/weights'], dtype) + _t(weights[name + '/fc3/biases'], dtype)
    return net + phi


def encoder_fc3_dropout(x, weights, scope, dtype):
    """src/models.py:80-116 at inference (dropout = identity).  scope = '<ief scope>/3D_module'."""
    net = torch.relu(x @ _t(weights[scope + '/fc1/weights'], dtype) + _t(weights[scope + '/fc1/biases'], dtype))
    net = torch.relu(net @ _t(weights[scope + '/fc2/weights'], dtype) + _t(weights[scope + '/fc2/biases'], dtype))
    return net @ _t(weights[scope + '/fc3/weights'], dtype) + _t(weights[scope + '/fc3/biases'], dtype)


def hmr_ief(phi, omega_start, weights, scope, num_stage=3, dtype=torch.float32):
    """src/models.py:380-415."""
    theta_prev = omega_start
    theta_here = None
    for _ in range(num_stage):
        state = torch.cat([phi, theta_prev], dim=1)                                      # :402
        delta_theta = encoder_fc3_dropout(state, weights, scope + '/3D_module', dtype)
        theta_here = theta_prev + delta_theta                                            # :410
        theta_prev = theta_here
    return theta_here


def call_hmr_ief(phi, omega_start, weights, scope, num_output=85, num_stage=3,
                 predict_delta_keys=(), use_delta_from_pred=False, use_optcam=True, dtype=torch.float32):
    """src/models.py:299-377."""
    phi = _t(phi, dtype)
    omega_start = _t(omega_start, dtype)
    theta_here = hmr_ief(phi, omega_start, weights, scope, num_stage, dtype)
    num_output_delta = 72 if use_optcam else 3 + 72
    deltas = {}
    for delta_t in predict_delta_keys:
        if delta_t == 0:
            continue
        scope_delta = scope + ('_future%d' % delta_t if delta_t > 0 else '_past%d' % abs(delta_t))
        start = theta_here if use_delta_from_pred else omega_start
        beta = start[:, -10:]
        start = start[:, 3:3 + num_output_delta] if use_optcam else start[:, :num_output_delta]
        delta_pred = hmr_ief(phi, start, weights, scope_delta, num_stage, dtype)
        if use_optcam:
            n = delta_pred.shape[0]
            delta_pred = torch.cat([torch.ones(n, 1, dtype=dtype), torch.zeros(n, 2, dtype=dtype),
                                    delta_pred, beta], dim=1)                            # :367-371
        else:
            delta_pred = torch.cat([delta_pred[:, :75], beta], dim=1)
        deltas[delta_t] = delta_pred
    return theta_here, deltas


def batch_pred_omega(input_features, batch_size, weights, num_output, omega_mean, sequence_length, scope,
                     predict_delta_keys=(), use_delta_from_pred=False, use_optcam=False, dtype=torch.float32):
    """src/models.py:233-267."""
    feats = _t(input_features, dtype).reshape(batch_size * sequence_length, -1)
    omega_pred, deltas = call_hmr_ief(feats, omega_mean, weights, scope, num_output, 3, predict_delta_keys,
                                      use_delta_from_pred, use_optcam, dtype)
    omega_pred = omega_pred.reshape(batch_size, sequence_length, num_output)
    return omega_pred, {k: v.reshape(batch_size, sequence_length, num_output) for k, v in deltas.items()}


def hmmr_predict(images, weights, smpl_model, num_conv_layers=3, delta_t_values=(-5, 5), pred_mode='pred',
                 dtype=torch.float32, num_kps=25):
    """Tester.build_test_model + predict, src/evaluation/tester.py:169-258 (+ omega.py:263-304).

    images (B,T,224,224,3) -> dict with the 14 fetch keys of tester.py:217-255.
    """
    from .smpl_ref import SMPLRef, batch_orth_proj_idrot
    npdt = np.float64 if dtype == torch.float64 else np.float32
    images = np.asarray(images)
    B, T = images.shape[:2]
    I_t = images.reshape((B * T,) + images.shape[2:])                                    # :171-174
    img_feat = encoder_resnet(I_t, weights, dtype)                                       # :175-179
    img_feat_full = img_feat.reshape(B, T, -1)                                           # :180
    theta_mean = np.tile(np.asarray(weights['mean_param']).reshape(1, 85), (B, 1))       # :78-83
    omega_mean = np.tile(theta_mean, (T, 1))                                             # :181
    if pred_mode == 'pred':
        movie_strips = az_fc2_groupnorm(img_feat_full, weights, num_conv_layers, dtype)  # :184-188
    elif pred_mode == 'hal':
        movie_strips = fc2_res(img_feat_full, weights, dtype)                            # :189-190
    else:
        raise Exception('Pred mode {} not recognized'.format(pred_mode))
    keys = [0] + [int(d) for d in delta_t_values]
    omegas_raw, deltas_pred = batch_pred_omega(movie_strips, B, weights, 85, omega_mean, T, 'single_view_ief',
                                               keys, use_delta_from_pred=True, use_optcam=True, dtype=dtype)
    smpl = SMPLRef(smpl_model, dtype=npdt)

    def compute(raw, cams_override=None):                                                # omega.py:231-304
        raw = raw.numpy().astype(npdt)
        cams = raw[:, :, :3] if cams_override is None else cams_override
        poses_aa = raw[:, :, 3:75]
        shapes = raw[:, :, 75:85]
        verts, joints, Rs = smpl(shapes.reshape(B * T, 10), poses_aa.reshape(B * T, 24, 3), get_skin=True)
        kps = batch_orth_proj_idrot(joints, cams.reshape(B * T, 3), npdt)
        return {'cams': cams, 'joints': joints.reshape(B, T, -1, 3), 'kps': kps.reshape(B, T, -1, 2),
                'poses': Rs.reshape(B, T, 24, 3, 3), 'shapes': shapes, 'verts': verts.reshape(B, T, -1, 3),
                'omegas': raw, 'Jtr': smpl.J_transformed.reshape(B, T, 24, 3)}

    out = compute(omegas_raw)
    cams0 = out['cams']
    result = {k: v for k, v in out.items() if k != 'Jtr'}
    per_delta = [compute(deltas_pred[dt], cams0) for dt in sorted(deltas_pred.keys())]   # tester.py:244-255
    if per_delta:
        for k in ('cams', 'joints', 'kps', 'poses', 'shapes', 'verts', 'omegas'):
            result[k + '_delta'] = np.stack([d[k] for d in per_delta], axis=2)
    result['_phi'] = img_feat_full.numpy()
    result['_movie_strips'] = movie_strips.numpy()
    return result


def single_frame_predict(images, weights, smpl_model, dtype=torch.float32):
    """BASELINE config 2: ResNet -> call_hmr_ief('single_view_ief', 85, 3) -> SMPL + projection,
    as wired in trainer_sequence_fc.py:531-545 (use_hmr_only branch) minus losses."""
    from .smpl_ref import SMPLRef, batch_orth_proj_idrot
    npdt = np.float64 if dtype == torch.float64 else np.float32
    N = images.shape[0]
    phi = encoder_resnet(images, weights, dtype)
    theta_mean = np.tile(np.asarray(weights['mean_param']).reshape(1, 85), (N, 1))
    theta, _ = call_hmr_ief(phi, theta_mean, weights, 'single_view_ief', 85, 3, (), dtype=dtype)
    raw = theta.numpy().astype(npdt)
    smpl = SMPLRef(smpl_model, dtype=npdt)
    verts, joints, Rs = smpl(raw[:, 75:85], raw[:, 3:75].reshape(N, 24, 3), get_skin=True)
    kps = batch_orth_proj_idrot(joints, raw[:, :3], npdt)
    return {'phi': phi.numpy(), 'omegas': raw, 'verts': verts, 'joints': joints, 'poses': Rs, 'kps': kps,
            'cams': raw[:, :3], 'shapes': raw[:, 75:85]}
