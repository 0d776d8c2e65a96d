"""CPU: one focused test per [TF-ext] ASSUMPTION listed in oracle/nets_ref.py (A1-A8), plus wiring checks of the
first-party graph code (models.py / tester.py).  Each can be flipped individually once a real TF 1.8 run exists."""
import numpy as np
import pytest
import torch

from oracle import nets_ref as R
from human_dynamics_b200 import synthetic


def test_A1_same_padding_rule():
    assert R.same_pad(112, 3, 2) == (0, 1, 56)        # pool1: extra pad at the END, not torch's padding=1
    assert R.same_pad(20, 3, 1) == (1, 1, 20)         # temporal conv k=3, T=20
    assert R.same_pad(7, 3, 2) == (1, 1, 4)
    assert R.same_pad(8, 1, 2) == (0, 0, 4)


def test_A1_pool1_ignores_padding_and_pads_bottom_right():
    x = torch.arange(16, dtype=torch.float32).reshape(1, 4, 4, 1) - 100.0        # all negative: zero padding would win
    y = R.max_pool_same_3x3_s2(x)
    assert y.shape == (1, 2, 2, 1)
    assert y[0, 0, 0, 0] == x[0, :3, :3, 0].max()            # window starts at (0,0): no top/left pad
    assert y[0, 1, 1, 0] == x[0, 2:, 2:, 0].max()            # last window hangs over the bottom/right edge


def test_A2_conv2d_same_strided_is_symmetric_explicit_pad():
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.normal(size=(1, 8, 8, 2)).astype(np.float32))
    w = torch.from_numpy(rng.normal(size=(3, 3, 2, 3)).astype(np.float32))
    y = R.conv2d_same(x, w, 2)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=2, padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(y, ref, atol=1e-6)                 # == torch Conv2d(padding=k//2)
    y_same = R.conv2d_nhwc(x, w, None, 2, 'SAME')            # plain SAME would pad (0,1): different result
    assert not torch.allclose(y, y_same, atol=1e-3)
    w7 = torch.from_numpy(rng.normal(size=(7, 7, 2, 1)).astype(np.float32))
    assert R.conv2d_same(torch.zeros(1, 224, 224, 2), w7, 2).shape == (1, 112, 112, 1)      # conv1: pad 3+3, VALID


def test_A3_topology_shapes_and_stride_on_last_unit(weights):
    img = synthetic.make_images(1, seed=0, size=64)
    phi, ends = R.encoder_resnet(img, weights, return_endpoints=True)
    assert phi.shape == (1, 2048)
    assert ends['conv1'].shape == (1, 32, 32, 64) and ends['pool1'].shape == (1, 16, 16, 64)
    assert ends['block1'].shape == (1, 8, 8, 256)            # stride applied by the LAST unit of block1
    assert ends['block2'].shape == (1, 4, 4, 512) and ends['block3'].shape == (1, 2, 2, 1024)
    assert ends['block4'].shape == (1, 2, 2, 2048)           # block4 stride 1
    # root conv1 has bias but no norm / activation: negative values survive
    assert ends['conv1'].min() < 0


def test_A4_bottleneck_unit_by_hand(weights):
    q = 'resnet_v2_50/block1/unit_3/bottleneck_v2'           # identity shortcut + stride 2: shortcut = x[:, ::2, ::2]
    rng = np.random.RandomState(1)
    x = torch.from_numpy(rng.normal(size=(1, 6, 6, 256)).astype(np.float64))
    y = R.bottleneck_v2(x, weights, q, 64, 2, torch.float64)
    w = {k: torch.from_numpy(np.asarray(v, np.float64)) for k, v in weights.items() if k.startswith(q)}

    def bn(t, p):
        return w[p + '/gamma'] * (t - w[p + '/moving_mean']) / torch.sqrt(w[p + '/moving_variance'] + 1e-5) + w[p + '/beta']
    pre = torch.relu(bn(x, q + '/preact'))
    r = torch.relu(bn(torch.einsum('nhwc,cd->nhwd', pre, w[q + '/conv1/weights'][0, 0]), q + '/conv1/BatchNorm'))
    rp = torch.nn.functional.pad(r.permute(0, 3, 1, 2), (1, 1, 1, 1))
    r = torch.nn.functional.conv2d(rp, w[q + '/conv2/weights'].permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1)
    r = torch.relu(bn(r, q + '/conv2/BatchNorm'))
    r = torch.einsum('nhwc,cd->nhwd', r, w[q + '/conv3/weights'][0, 0]) + w[q + '/conv3/biases']
    assert torch.allclose(y, x[:, ::2, ::2] + r, atol=1e-10)
    # projection shortcut acts on the PRE-ACTIVATION, with bias, no BN/ReLU (unit_1)
    q1 = 'resnet_v2_50/block1/unit_1/bottleneck_v2'
    x1 = torch.from_numpy(rng.normal(size=(1, 4, 4, 64)).astype(np.float64))
    y1 = R.bottleneck_v2(x1, weights, q1, 64, 1, torch.float64)
    w1 = {k: torch.from_numpy(np.asarray(v, np.float64)) for k, v in weights.items() if k.startswith(q1)}
    pre1 = torch.relu(w1[q1 + '/preact/gamma'] * (x1 - w1[q1 + '/preact/moving_mean'])
                      / torch.sqrt(w1[q1 + '/preact/moving_variance'] + 1e-5) + w1[q1 + '/preact/beta'])
    sc = torch.einsum('nhwc,cd->nhwd', pre1, w1[q1 + '/shortcut/weights'][0, 0]) + w1[q1 + '/shortcut/biases']
    x_id = R.bottleneck_v2(x1, {**weights, q1 + '/shortcut/weights': np.zeros((1, 1, 64, 256), np.float32),
                                q1 + '/shortcut/biases': np.zeros(256, np.float32)}, q1, 64, 1, torch.float64)
    assert torch.allclose(y1 - x_id, sc, atol=1e-10)


def test_A5_batchnorm_inference_formula():
    w = {'p/gamma': np.array([2.0], np.float32), 'p/beta': np.array([0.5], np.float32),
         'p/moving_mean': np.array([1.0], np.float32), 'p/moving_variance': np.array([4.0], np.float32)}
    y = R.batch_norm_inf(torch.tensor([[3.0]]), w, 'p', torch.float64)
    assert abs(float(y) - (2.0 * (3.0 - 1.0) / np.sqrt(4.0 + 1e-5) + 0.5)) < 1e-12


def test_A7_group_norm_statistics():
    rng = np.random.RandomState(2)
    x = torch.from_numpy(rng.normal(2.0, 3.0, size=(2, 20, 1, 2048)))
    y = R.group_norm_tf(x, torch.ones(2048, dtype=torch.float64), torch.zeros(2048, dtype=torch.float64))
    yg = y.reshape(2, 20, 1, 32, 64)
    assert torch.allclose(yg.mean(dim=(1, 2, 4)), torch.zeros(2, 32, dtype=torch.float64), atol=1e-10)     # per (clip, group)
    assert torch.allclose(yg.var(dim=(1, 2, 4), unbiased=False), torch.ones(2, 32, dtype=torch.float64), atol=1e-5)
    ref = torch.nn.functional.group_norm(x[:, :, 0].permute(0, 2, 1), 32, eps=1e-6).permute(0, 2, 1)[:, :, None]
    assert torch.allclose(y, ref, atol=1e-9)                 # == torch GroupNorm(32, eps=1e-6) over (T, C/32)
    # window coupling: changing one frame changes every frame's output
    x2 = x.clone(); x2[0, 0] += 1.0
    y2 = R.group_norm_tf(x2, torch.ones(2048, dtype=torch.float64), torch.zeros(2048, dtype=torch.float64))
    assert (y2[0, 19] - y[0, 19]).abs().max() > 1e-6 and torch.equal(y2[1], y[1])


def test_A1_temporal_conv_zero_pads_window_edges(weights):
    x = torch.zeros(1, 20, 1, 2048, dtype=torch.float64); x[0, 0] = 1.0
    w = torch.from_numpy(np.asarray(weights['AZ_FC_block2_conv1block_0/weights'], np.float64))
    y = R.conv2d_nhwc(x, w, None, 1, 'SAME')
    assert y.shape == (1, 20, 1, 2048)
    assert torch.allclose(y[0, 0, 0], w[1, 0].sum(0), atol=1e-9) and torch.allclose(y[0, 1, 0], w[0, 0].sum(0), atol=1e-9)
    assert y[0, 2:].abs().max() == 0


def test_A8_fc_semantics_and_ief_wiring(weights):
    N = 3
    phi = np.random.RandomState(3).normal(size=(N, 2048)).astype(np.float64)
    om = np.tile(weights['mean_param'].reshape(1, 85).astype(np.float64), (N, 1))
    th, deltas = R.call_hmr_ief(phi, om, weights, 'single_view_ief', 85, 3, (0, -5, 5), True, True, torch.float64)
    q = 'single_view_ief/3D_module'
    W = {k: torch.from_numpy(np.asarray(weights[q + k], np.float64)) for k in
         ('/fc1/weights', '/fc1/biases', '/fc2/weights', '/fc2/biases', '/fc3/weights', '/fc3/biases')}
    t = torch.from_numpy(om)
    for _ in range(3):                                     # state = concat[phi, theta]; relu, relu, linear; theta += delta
        s = torch.cat([torch.from_numpy(phi), t], 1)
        h = torch.relu(s @ W['/fc1/weights'] + W['/fc1/biases'])
        h = torch.relu(h @ W['/fc2/weights'] + W['/fc2/biases'])
        t = t + h @ W['/fc3/weights'] + W['/fc3/biases']
    assert torch.allclose(th, t, atol=1e-10)
    for dt in (-5, 5):                                     # delta heads: [1, 0, 0, pose72, beta of the main prediction]
        d = deltas[dt]
        assert d.shape == (N, 85)
        assert torch.equal(d[:, 0], torch.ones(N, dtype=torch.float64)) and torch.equal(d[:, 1:3], torch.zeros(N, 2, dtype=torch.float64))
        assert torch.equal(d[:, 75:], th[:, 75:])
    assert set(deltas.keys()) == {-5, 5}


def test_predict_output_contract(weights, smpl_model):
    """tester.py:217-255: 7 keys + 7 '_delta' keys stacked over sorted delta_t on axis 2; cams of deltas = cams of dt=0."""
    B, T, S = 1, 3, 64
    img = synthetic.make_images(B * T, seed=5, size=S).reshape(B, T, S, S, 3)
    out = R.hmmr_predict(img, weights, smpl_model)
    shapes = {'cams': (B, T, 3), 'joints': (B, T, 25, 3), 'kps': (B, T, 25, 2), 'poses': (B, T, 24, 3, 3),
              'shapes': (B, T, 10), 'verts': (B, T, 6890, 3), 'omegas': (B, T, 85)}
    for k, shp in shapes.items():
        assert out[k].shape == shp
        assert out[k + '_delta'].shape == shp[:2] + (2,) + shp[2:]
    assert np.array_equal(out['cams_delta'][:, :, 0], out['cams']) and np.array_equal(out['cams_delta'][:, :, 1], out['cams'])
    assert np.array_equal(out['omegas_delta'][..., 0], np.ones((B, T, 2), np.float32))
    assert np.array_equal(out['shapes_delta'][:, :, 0], out['shapes'])
    with pytest.raises(Exception):
        R.hmmr_predict(img, weights, smpl_model, pred_mode='nope')
