"""Host-side wiring of the ResNet layer plan checked WITHOUT a GPU: device allocations and library calls are stubbed, the descriptors
the plan fills are real.  Catches plan-logic mistakes (buffer roles, dead outputs, epilogue subsampling, launch counts) on the CPU."""
import numpy as np
import pytest
import torch


@pytest.fixture
def fake_device(monkeypatch):
    from human_dynamics_b200 import nets

    class FakeLib(object):
        def __getattr__(self, name):
            return lambda *a, **k: 0
    monkeypatch.setattr(nets, 'lib', FakeLib())
    monkeypatch.setattr(nets, '_dev', lambda a, device, dtype=np.float32: torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)))
    monkeypatch.setattr(torch.Tensor, 'to', lambda self, *a, **k: self)
    e, z = torch.empty, torch.zeros
    monkeypatch.setattr(torch, 'empty', lambda *a, **k: e(*a, **{kk: v for kk, v in k.items() if kk != 'device'}))
    monkeypatch.setattr(torch, 'zeros', lambda *a, **k: z(*a, **{kk: v for kk, v in k.items() if kk != 'device'}))
    return nets


def _convs(plan):
    return [o for o in plan.ops if getattr(o, 'd', None) is not None]


def test_resnet_plan_wiring(fake_device, monkeypatch):
    nets = fake_device
    from human_dynamics_b200 import synthetic
    packed = nets.PackedResNet(synthetic.make_resnet_weights(seed=1), 'cpu', tc='auto')
    assert len(packed.units) == 16
    for epi in (True, False):
        monkeypatch.setattr(nets, 'SUBSAMPLE_EPI', epi)
        plan = nets.ResNetPlan(packed, 2, 64, 'auto')
        assert plan.split and plan.pool_f32_dead
        convs = _convs(plan)
        assert len(convs) == 52                                            # 16 units x 3 + 4 shortcut convs (conv1 of the root is separate)
        subs = [o for o in plan.ops if isinstance(o, nets.SubsampleOp)]
        epis = [(o.d.Cout, o.d.Ho, o.d.out_subsample) for o in convs if o.d.out_subsample > 1]
        # the units in front of the three strided identity units write x[:, ::2, ::2] themselves (maps 16 -> 8 -> 4 -> 2 at size 64)
        assert (len(subs), epis) == ((0, [(256, 16, 2), (512, 8, 2), (1024, 4, 2)]) if epi else (3, []))
        assert sum(1 for o in convs if not o.d.out and o.d.res) == 3       # fp32 outputs in front of a conv shortcut are never written
        assert plan.num_launches == 3 + len(plan.ops) + 1
        for o in convs:
            d = o.d
            assert d.in_hi and d.in_lo and not d.in_ and d.impl == 3       # every trunk conv reads a pre-split pair
            if d.res:
                assert (d.res_stride, d.res_H, d.res_W, d.res_ld) == (1, d.Ho, d.Wo, d.Cout)   # residual rows == output rows (TMA slabs)
            if d.out_subsample > 1:
                assert d.out and not d.tmap_out and d.tmap_out_hi and d.res                      # dense subsample: no fp32 TMA store
        # the residual of a strided identity unit is the dense subsampled buffer in both variants
        strided = [o for o in convs if o.d.KH == 1 and o.d.res and o.d.res == plan.bufS.data_ptr() and o.d.Cout in (256, 512, 1024)]
        assert len(strided) >= 3


def test_stage_plans_keep_the_pairs_inside_a_stage(fake_device):
    """The engine cuts the trunk after unit 7 (block 2): every (unit, strided identity unit) pair lies inside one stage, so neither
    stage needs an hd_subsample pass; the stage-A output in front of block 3's conv shortcut has no fp32 copy."""
    nets = fake_device
    from human_dynamics_b200 import synthetic
    packed = nets.PackedResNet(synthetic.make_resnet_weights(seed=1), 'cpu', tc='auto')
    nxt = packed.units[7]
    pa = nets.ResNetPlan(packed, 2, 64, 'auto', units=(0, 7), root=True, tail=False, next_pre=nxt['pre'], next_has_shortcut='shortcut' in nxt)
    pb = nets.ResNetPlan(packed, 2, 64, 'auto', units=(7, 16), root=False, tail=True)
    assert not any(isinstance(o, nets.SubsampleOp) for o in pa.ops + pb.ops)
    assert sum(1 for o in _convs(pa) + _convs(pb) if o.d.out_subsample > 1) == 3
    last_a = _convs(pa)[-1].d
    assert not last_a.out and last_a.out_hi and last_a.post2_relu == 1
    assert (pb.in_hw, pb.in_depth) == (pa.out_hw, pa.out_depth) == (4, 512)


def test_fmovie_and_ief_plan_wiring(fake_device):
    """f_movie fast path: per block GN+ReLU+split -> conv (k=3 over T, pad 1) -> GN+ReLU+split -> conv + residual; IEF fast path: phi split
    once, per head one hoisted phi.W1 GEMM, per stage fc1-theta / fc2 (tensor cores) / fc3; delta heads write into the [N, D, 85] stack."""
    nets = fake_device
    from human_dynamics_b200 import synthetic
    w = synthetic.make_synthetic_weights(seed=1)
    B, T = 2, 20
    fm = nets.FMoviePlan(nets.PackedFMovie(w, 'cpu', 3, tc='auto'), B, T, 'auto')
    x = torch.zeros((B, T, 2048))
    fm._bind(x)
    assert [s[0] for s in fm.steps] == ['gns', 'conv'] * 6 and fm.num_launches == 12
    convs = [s[1].d for s in fm.steps if s[0] == 'conv']
    for i, d in enumerate(convs):
        assert (d.n_img, d.H, d.W, d.KH, d.KW, d.pad_t, d.pad_l, d.Ho, d.Wo, d.Cin, d.Cout) == (B, T, 1, 3, 1, 1, 0, T, 1, 2048, 2048)
        assert d.in_hi == fm.act[0].data_ptr() and d.impl == 3 and d.post_shift and not d.post_relu
        assert bool(d.res) == (i % 2 == 1)                                   # the block's second conv adds the block input
    assert convs[1].res == x.data_ptr() and convs[3].res == convs[1].out and convs[5].res == convs[3].out
    assert fm.out.data_ptr() == convs[5].out

    N = B * T
    ief = nets.IEFPlan(nets.PackedIEF(w, 'cpu', tc='auto'), N, 3, None, 'auto')
    assert ief.fast and ief.delta_keys == [-5, 5] and ief.num_launches == 33
    phi, theta0 = torch.zeros((N, 2048)), torch.zeros((N, 85))
    ief._bind(phi, theta0)
    kinds = [op[0] for op in ief.main_ops]
    assert kinds == ['conv'] + ['fc1t', 'conv', 'fc3'] * 3
    # stage 0 starts from theta0, later stages from the running theta; the delta heads run in place on columns 3:75 of their slot
    assert ief.main_ops[1][1].data_ptr() == theta0.data_ptr() and ief.main_ops[4][1].data_ptr() == ief.theta.data_ptr()
    for i, dt in enumerate(ief.delta_keys):
        ops = ief.delta_ops[dt]
        view = ops[1][1]
        assert view.data_ptr() == ief.delta_all.data_ptr() + (i * 85 + 3) * 4 and ops[1][2] == 2 * 85 and ops[3][4] == 2 * 85
        assert ops[3][5].d == 72
    assert ief.main_ops[3][5].d == 85
