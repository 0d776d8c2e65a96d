"""Network-level C entries (include/hd_b200.h: hd_resnet50_* / hd_fmovie_* / hd_ief_*, csrc/net_plan.cu) vs the Python host plans
of nets.py: same kernels, same descriptors, same buffers' roles -> the outputs must be identical bit for bit; plus the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-4


def rel_err(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize('n,size', [(3, 64), (2, 224)])
def test_c_resnet50_equals_python_plan_bit_for_bit(weights, n, size):
    from human_dynamics_b200 import synthetic
    from human_dynamics_b200.cplan import CResNet50
    from human_dynamics_b200.nets import PackedResNet, ResNetPlan
    from oracle import nets_ref
    dev = torch.device('cuda')
    img_h = synthetic.make_images(n, seed=n, size=size)
    img = torch.from_numpy(img_h).to(dev)
    plan = ResNetPlan(PackedResNet(weights, dev, tc='auto'), n, size, 'auto')
    assert plan.split
    phi_py = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    plan.run(img, phi_py)
    net = CResNet50(weights, n, size)
    assert net.num_launches == plan.num_launches
    phi_c = net(img)
    torch.cuda.synchronize()
    assert torch.equal(phi_c, phi_py)
    assert rel_err(phi_c.cpu().numpy(), nets_ref.encoder_resnet(img_h, weights).numpy()) < REL
    phi_c2 = net(img)                                  # plans are reusable; buffers carry no state between calls
    torch.cuda.synchronize()
    assert torch.equal(phi_c2, phi_py)
    net.close()


def test_c_fmovie_equals_python_plan_bit_for_bit(weights):
    from human_dynamics_b200.cplan import CFMovie
    from human_dynamics_b200.nets import PackedFMovie, FMoviePlan
    from oracle import nets_ref
    dev = torch.device('cuda')
    B, T = 2, 20
    x_h = np.random.RandomState(5).normal(0, 1, size=(B, T, 2048)).astype(np.float32)
    x = torch.from_numpy(x_h).to(dev)
    plan = FMoviePlan(PackedFMovie(weights, dev, 3, tc='auto'), B, T, 'auto')
    y_py = plan.run(x).clone()
    net = CFMovie(weights, B, T, 3)
    y_c = net(x)
    y_c_other = net(x, out=torch.empty_like(x))        # other output pointer: descriptors are re-encoded
    torch.cuda.synchronize()
    assert net.num_launches == plan.num_launches
    assert torch.equal(y_c, y_py) and torch.equal(y_c_other, y_py)
    assert rel_err(y_c.cpu().numpy(), nets_ref.az_fc2_groupnorm(torch.from_numpy(x_h), weights, 3).numpy()) < REL


def test_c_ief_equals_python_plan_bit_for_bit(weights):
    from human_dynamics_b200.cplan import CIEF
    from human_dynamics_b200.nets import PackedIEF, IEFPlan
    from oracle import nets_ref
    dev = torch.device('cuda')
    N = 40
    phi_h = np.random.RandomState(6).normal(0, 1, size=(N, 2048)).astype(np.float32)
    phi = torch.from_numpy(phi_h).to(dev)
    packed = PackedIEF(weights, dev, tc='auto')
    plan = IEFPlan(packed, N, 3, None, 'auto')
    assert plan.fast
    theta0 = packed.mean_param.expand(N, 85).contiguous()
    th_py, d_py = plan.run(phi, theta0)
    net = CIEF(weights, N, (-5, 5))
    th_c, d_c = net(phi)
    torch.cuda.synchronize()
    assert net.num_launches == plan.num_launches
    assert torch.equal(th_c, th_py)
    assert sorted(d_c) == sorted(d_py) == [-5, 5]
    for dt in d_c:
        assert torch.equal(d_c[dt], d_py[dt]), dt
    om = np.tile(np.asarray(weights['mean_param'], np.float32).reshape(1, 85), (N, 1))
    ref, dref = nets_ref.call_hmr_ief(torch.from_numpy(phi_h), om, weights, 'single_view_ief', 85, 3, [0, -5, 5],
                                      use_delta_from_pred=True, use_optcam=True)
    assert rel_err(th_c.cpu().numpy(), ref.numpy()) < REL
    for dt in (-5, 5):
        assert rel_err(d_c[dt].cpu().numpy(), dref[dt].numpy()) < REL


def test_c_plan_names_the_missing_variable(weights):
    from human_dynamics_b200 import _lib
    from human_dynamics_b200.cplan import CFMovie
    w = {k: v for k, v in weights.items() if k != 'AZ_FC_block2_conv2block_1/biases'}
    with pytest.raises(_lib.HDError, match='AZ_FC_block2_conv2block_1/biases'):
        CFMovie(w, 2, 20, 3)
