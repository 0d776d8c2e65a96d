"""GPU parity at the sizes BASELINE.json names (the configurations bench.py times), plus the drop-in surface the
round-1 tests never called (src.models.batch_pred_omega / encoder_fc3_dropout, FeatureExtractor's ragged tail).

The oracle cannot run 640 frames in the test budget, and it does not have to: clips (and, on the single-frame path,
frames) are independent, so the full-size GPU run is checked on the first and last clip against a 2-clip oracle run.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-4          # BASELINE.json north_star: 1e-4 rel FP32 on omegas / verts / kps

KEYS = tuple(a + b for b in ('', '_delta') for a in ('cams', 'joints', 'kps', 'poses', 'shapes', 'verts', 'omegas'))


def rel_err(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.fixture(scope='module')
def c3(weights, smpl_model):
    """One C3-sized run (32 clips x T=20, default 160/640-frame trunk stages) shared by the tests below."""
    from human_dynamics_b200 import synthetic, HMMRConfig
    from human_dynamics_b200.engine import HMMREngine
    B, T = 32, 20
    img = synthetic.make_images(B * T, seed=11).reshape(B, T, 224, 224, 3)
    eng = HMMREngine(weights, smpl_model, HMMRConfig(batch_size=B, sequence_length=T))
    assert (eng.config.frame_chunk, eng.config.late_chunk) == (160, 640)
    dev_out = eng.predict(torch.from_numpy(img).cuda())
    torch.cuda.synchronize()
    dev_out = {k: v.clone() for k, v in dev_out.items() if not k.startswith('_')}
    return eng, img, dev_out


def test_c3_full_size_first_and_last_clip_match_oracle(c3, weights, smpl_model):
    """BASELINE configs[2] as timed by bench.py: B=32, T=20, frame_chunk=160, late_chunk=640."""
    from oracle import nets_ref
    eng, img, out = c3
    sel = [0, img.shape[0] - 1]
    ref = nets_ref.hmmr_predict(img[sel], weights, smpl_model)
    assert set(KEYS) <= set(out.keys())
    for k in KEYS:
        g = out[k][sel].cpu().numpy()
        assert g.shape == ref[k].shape, (k, g.shape, ref[k].shape)
        assert rel_err(g, ref[k]) < REL, (k, rel_err(g, ref[k]))


def test_c3_predict_host_equals_device_run_bit_for_bit(c3):
    """The end-to-end leg (pinned host frames, 32-frame first pass, chunked H2D, D2H of all 14 tensors) runs the same
    kernels on the same rows: its results must equal the device-resident run exactly."""
    eng, img, out = c3
    host, h2d, d2h = eng.predict_host(torch.from_numpy(img).pin_memory())
    torch.cuda.current_stream().synchronize()
    assert h2d == img.size * 4 and d2h == sum(out[k].numel() * 4 for k in KEYS)
    for k in KEYS:
        assert torch.equal(host[k], out[k].cpu()), k


def test_c3_chunking_does_not_change_results(c3, weights, smpl_model):
    """Trunk stage sizes are a scheduling choice: 16-frame passes give bit-identical per-clip outputs."""
    from human_dynamics_b200 import HMMRConfig
    from human_dynamics_b200.engine import HMMREngine
    eng, img, out = c3
    B, T = 4, img.shape[1]
    eng2 = HMMREngine(weights, smpl_model, HMMRConfig(batch_size=B, sequence_length=T, frame_chunk=16, late_chunk=32))
    o2 = eng2.predict(torch.from_numpy(img[:B]).cuda())
    torch.cuda.synchronize()
    for k in ('omegas', 'omegas_delta', 'cams', 'shapes'):          # everything up to the IEF output: bit-identical
        assert torch.equal(o2[k], out[k][:B]), k
    for k in ('verts', 'kps', 'verts_delta'):                       # SMPL picks its kernel by batch size (80 vs 640 poses): same numbers to 1e-5
        assert rel_err(o2[k].cpu().numpy(), out[k][:B].cpu().numpy()) < 2e-5, k


def test_c2_single_frame_batch64(weights, smpl_model):
    """BASELINE configs[1]: ResNet-50 + 3-iter IEF + SMPL at batch 64; frames are independent -> check frames {0, 63}."""
    from human_dynamics_b200 import synthetic, HMMRConfig
    from human_dynamics_b200.engine import HMMREngine
    from oracle import nets_ref
    n = 64
    img = synthetic.make_images(n, seed=21)
    eng = HMMREngine(weights, smpl_model, HMMRConfig(batch_size=n, sequence_length=1))
    got = eng.predict(torch.from_numpy(img).cuda().view(n, 1, 224, 224, 3), single_frame=True)
    torch.cuda.synchronize()
    sel = [0, n - 1]
    ref = nets_ref.single_frame_predict(img[sel], weights, smpl_model)
    for k in ('omegas', 'verts', 'joints', 'kps', 'poses'):
        g = got[k][sel].cpu().numpy().reshape(ref[k].shape)
        assert rel_err(g, ref[k]) < REL, (k, rel_err(g, ref[k]))


def test_c5_smpl_65536_poses(smpl_model):
    """BASELINE configs[4] at full size: sampled rows against the oracle + size-independent invariants on all rows."""
    from human_dynamics_b200.smpl import SMPLConstants
    from human_dynamics_b200 import synthetic
    from oracle import smpl_ref
    n = 65536
    beta, theta = synthetic.make_smpl_inputs(n, seed=5)
    c = SMPLConstants(smpl_model)
    b, t = torch.from_numpy(beta).cuda(), torch.from_numpy(theta).cuda()
    cam = torch.from_numpy(np.random.RandomState(1).uniform(0.5, 1.5, size=(n, 3)).astype(np.float32)).cuda()
    o = c.forward(b, t, cam=cam)
    torch.cuda.synchronize()
    idx = np.concatenate([np.arange(0, n, 4099), [n - 1]])
    ref = smpl_ref.SMPLRef(smpl_model, dtype=np.float64)
    v, j, Rs = ref(beta[idx], theta[idx].reshape(-1, 24, 3), get_skin=True)
    kp = smpl_ref.batch_orth_proj_idrot(j, cam[idx].cpu().numpy(), dtype=np.float64)
    assert rel_err(o['verts'][idx].cpu().numpy(), v) < REL
    assert rel_err(o['joints'][idx].cpu().numpy(), j) < REL
    assert rel_err(o['Rs'][idx].cpu().numpy().reshape(Rs.shape), Rs) < REL
    assert rel_err(o['kps'][idx].cpu().numpy(), kp) < REL
    # invariants over ALL rows: joints = verts . regressor, kps = s (xy + t), rotations orthonormal
    reg = torch.from_numpy(smpl_model['cocoplus_regressor'].astype(np.float32)).cuda()
    for lo in range(0, n, 8192):
        vv = o['verts'][lo:lo + 8192]
        je = torch.einsum('kv,nvc->nkc', reg, vv)
        assert float((je - o['joints'][lo:lo + 8192]).abs().max()) < 2e-5
    ke = cam[:, None, 0:1] * (o['joints'][:, :, :2] + cam[:, None, 1:3])
    assert float((ke - o['kps']).abs().max()) < 1e-5
    R = o['Rs'].view(-1, 3, 3)
    eye = torch.eye(3, device=R.device)
    assert float((R @ R.transpose(1, 2) - eye).abs().max()) < 1e-5
    assert torch.isfinite(o['verts']).all()


def test_models_surface_batch_pred_omega_and_fc3(weights, smpl_model):
    """src.models.batch_pred_omega (models.py:233-267; cached-plan and generic paths) and encoder_fc3_dropout (:80-116)."""
    from human_dynamics_b200 import HMMRConfig
    from human_dynamics_b200.engine import HMMREngine
    from human_dynamics_b200 import runtime as rt
    import src.models as M
    from oracle import nets_ref
    B, T = 3, 7
    eng = HMMREngine(weights, smpl_model, HMMRConfig(batch_size=B, sequence_length=T))
    rt.set_default_engine(eng)
    rng = np.random.RandomState(4)
    feats = rng.normal(0, 0.5, size=(B, T, 2048)).astype(np.float32)
    om = np.tile(weights['mean_param'].reshape(1, 85), (B * T, 1)).astype(np.float32)
    ft, omt = torch.from_numpy(feats).cuda(), torch.from_numpy(om).cuda()
    r_om, r_d = nets_ref.batch_pred_omega(feats, B, weights, 85, om, T, 'single_view_ief', predict_delta_keys=(0, -5, 5),
                                          use_delta_from_pred=True, use_optcam=True)
    # Tester wiring (tester.py:196-207): cached plan
    g_om, g_d = M.batch_pred_omega(ft, B, False, 85, omt, T, 'single_view_ief', predict_delta_keys=(0, -5, 5),
                                   use_delta_from_pred=True, use_optcam=True)
    assert g_om.shape == (B, T, 85) and sorted(g_d.keys()) == [-5, 5]
    assert rel_err(g_om.cpu().numpy(), r_om.numpy()) < REL
    for dt in (-5, 5):
        assert g_d[dt].shape == (B, T, 85)
        assert rel_err(g_d[dt].cpu().numpy(), r_d[dt].numpy()) < REL
    # generic path: deltas start from omega_mean (use_delta_from_pred=False), models.py:349
    r_om2, r_d2 = nets_ref.batch_pred_omega(feats, B, weights, 85, om, T, 'single_view_ief', predict_delta_keys=(5,),
                                            use_delta_from_pred=False, use_optcam=True)
    g_om2, g_d2 = M.batch_pred_omega(ft, B, False, 85, omt, T, 'single_view_ief', predict_delta_keys=(5,),
                                     use_delta_from_pred=False, use_optcam=True)
    assert rel_err(g_om2.cpu().numpy(), r_om2.numpy()) < REL
    assert rel_err(g_d2[5].cpu().numpy(), r_d2[5].numpy()) < REL
    # one regressor evaluation: x = concat[phi, theta] -> delta theta  (models.py:101-113)
    x = np.concatenate([feats.reshape(B * T, -1), om], axis=1)
    delta, _ = M.encoder_fc3_dropout(torch.from_numpy(x).cuda(), num_output=85, is_training=False, scope='single_view_ief')
    r_delta = nets_ref.encoder_fc3_dropout(torch.from_numpy(x), weights, 'single_view_ief/3D_module', torch.float32)
    assert rel_err(delta.cpu().numpy(), r_delta.numpy()) < REL
    with pytest.raises(NotImplementedError):
        M.encoder_fc3_dropout(torch.from_numpy(x).cuda(), is_training=True)


def test_feature_extractor_ragged_tail(weights):
    """FeatureExtractor.compute_all_phis (resnet_extractor.py:74-98): last partial batch is zero-padded, output trimmed."""
    from human_dynamics_b200 import synthetic
    from src.datasets.resnet_extractor import FeatureExtractor
    from oracle import nets_ref
    S, bs, T = 64, 4, 9                       # 2 full batches + a tail of 1
    fx = FeatureExtractor(weights, img_size=S, batch_size=bs)
    frames = synthetic.make_images(T, seed=13, size=S)
    phis = fx.compute_all_phis(frames)
    assert phis.shape == (T, 2048)
    ref = nets_ref.encoder_resnet(frames, weights).numpy()
    assert rel_err(phis, ref) < REL
    one = fx.compute_phis(np.concatenate([frames[8:9], np.zeros((bs - 1, S, S, 3), np.float32)]))
    assert np.array_equal(one[0], phis[8])
    with pytest.raises(ValueError):
        fx.compute_phis(frames[:3])


def test_fp16_split_saturates_instead_of_nan():
    """activations beyond the fp16 range used to give hi=inf, (y-inf)*2048=NaN.  The split now clamps the
    head to +-65504: the pair stays finite (saturated), the fp32 output is exact."""
    from human_dynamics_b200.nets import PackedConv
    dev = torch.device('cuda')
    rng = np.random.RandomState(0)
    n, H, Cin, Cout = 2, 14, 64, 256
    x = np.maximum(rng.normal(0, 1, size=(n, H, H, Cin)), 0).astype(np.float32)
    w = (rng.normal(0, 1, size=(1, 1, Cin, Cout)) / np.sqrt(Cin)).astype(np.float32)
    bias = np.full(Cout, 1.0e5, np.float32)                  # pushes every output past 65504
    pc = PackedConv(w, dev, post_shift=bias, tc='tc3h')
    xt = torch.from_numpy(x).to(dev)
    hi = xt.half(); lo = ((xt - hi.float()) * 2048).half()
    out = torch.zeros((n, H, H, Cout), device=dev)
    oh = torch.zeros((n, H, H, Cout), dtype=torch.float16, device=dev); ol = torch.zeros_like(oh)
    op = pc.bind(None, n, H, H, out, inp_split=(hi, lo), out_split=(oh, ol), post2=(None, None, 1), impl='tc3h')
    op.run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and float(out.min()) > 65504
    assert torch.isfinite(oh).all() and torch.isfinite(ol).all()
    assert float(oh.float().min()) == 65504.0 and float(ol.float().abs().max()) == 0.0


def test_cuda_graph_replay_equals_eager(weights, smpl_model):
    """predict_graphed: one graph launch per window; results bit-identical to the eager launches, and new frames written
    into the same input buffer are picked up by the replay."""
    from human_dynamics_b200 import synthetic, HMMRConfig, _lib
    from human_dynamics_b200.engine import HMMREngine
    B, T, S = 2, 20, 64
    eng = HMMREngine(weights, smpl_model, HMMRConfig(batch_size=B, sequence_length=T, img_size=S))
    a = torch.from_numpy(synthetic.make_images(B * T, seed=1, size=S)).cuda().view(B, T, S, S, 3)
    b = torch.from_numpy(synthetic.make_images(B * T, seed=2, size=S)).cuda().view(B, T, S, S, 3)
    buf = a.clone()
    eager_a = {k: v.clone() for k, v in eng.predict(a).items() if not k.startswith('_')}
    eager_b = {k: v.clone() for k, v in eng.predict(b).items() if not k.startswith('_')}
    out, nodes = eng.predict_graphed(buf)
    torch.cuda.synchronize()
    assert nodes > 100
    for k in eager_a:
        assert torch.equal(out[k], eager_a[k]), k
    buf.copy_(b)
    _lib.lib.hd_launch_count_reset()
    out, _ = eng.predict_graphed(buf)
    torch.cuda.synchronize()
    assert _lib.lib.hd_launch_count() == 0          # nothing was launched from the host: the graph did it
    for k in eager_b:
        assert torch.equal(out[k], eager_b[k]), k


def test_tester_host_paths_numpy_and_uint8_frames(weights, smpl_model):
    """Tester.predict on a plain numpy array (page-locked in place) and Tester.predict_frames on uint8 video frames
    (GPU process_image feeding conv1 directly) against the device path and the cv2 + torch oracle."""
    pytest.importorskip('cv2')
    from human_dynamics_b200 import HMMRConfig
    from human_dynamics_b200.preprocess import process_images
    from src.evaluation.tester import Tester
    from oracle import nets_ref, preproc_ref
    B, T, H, W = 2, 20, 150, 200
    cfg = HMMRConfig(batch_size=B, sequence_length=T, weights=weights, smpl_model=smpl_model)
    tester = Tester(cfg)
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:H, 0:W]
    frames = np.stack([np.clip((120 + 90 * np.sin(xx / (5.0 + i) + i) * np.cos(yy / 6.0))[..., None] + rng.randint(-30, 30, size=(H, W, 3)), 0, 255)
                       for i in range(B * T)]).astype(np.uint8).reshape(B, T, H, W, 3)
    boxes = np.stack([rng.uniform(60, 140, B * T), rng.uniform(40, 110, B * T), rng.uniform(0.8, 1.6, B * T)], axis=1).reshape(B, T, 3)
    # (1) uint8 frames: one crossing
    got_u8 = {k: v.copy() for k, v in tester.predict_frames(frames, boxes).items()}
    # (2) the two-step route: GPU crops -> predict on the device
    crops, _ = process_images(frames.reshape(B * T, H, W, 3), boxes.reshape(-1, 3))
    dev = tester.predict(crops.view(B, T, 224, 224, 3), as_numpy=True)
    for k in dev:
        assert np.array_equal(got_u8[k], dev[k]), k
    # (3) a plain (pageable) numpy array of crops through Tester.predict: page-locked in place, same numbers
    crops_np = crops.cpu().numpy().reshape(B, T, 224, 224, 3).copy()
    n_reg = len(tester._registered)                    # (the uint8 frame array of step (1) is registered already)
    got_np = tester.predict(crops_np)
    assert len(tester._registered) == n_reg + 1
    for k in dev:
        assert np.array_equal(got_np[k], dev[k]), k
    again = tester.predict(crops_np)                   # same buffer: no second registration, next ring slot
    assert len(tester._registered) == n_reg + 1 and again['verts'] is not got_np['verts']
    assert np.array_equal(again['verts'], got_np['verts'])
    # (4) against the reference restatement end to end (cv2 process_image -> oracle graph), clip 0
    ref_crops = np.stack([preproc_ref.process_image(frames[0, t], boxes[0, t])['image'] for t in range(T)]).astype(np.float32)
    ref = nets_ref.hmmr_predict(ref_crops[None], weights, smpl_model)
    for k in ('omegas', 'verts', 'kps', 'verts_delta'):
        assert rel_err(got_u8[k][:1], ref[k]) < REL, (k, rel_err(got_u8[k][:1], ref[k]))


def test_predict_stream_matches_predict(weights, smpl_model):
    """Tester.predict_stream: windows pipelined two deep (upload / compute of window i+1 under the device->host copies of window i)
    give exactly the per-window results of the synchronous predict, in order, for float crops and for uint8 frames."""
    from human_dynamics_b200 import synthetic, HMMRConfig
    from src.evaluation.tester import Tester
    B, T, S = 2, 20, 64
    tester = Tester(HMMRConfig(batch_size=B, sequence_length=T, img_size=S, weights=weights, smpl_model=smpl_model))
    wins = [synthetic.make_images(B * T, seed=40 + i, size=S).reshape(B, T, S, S, 3) for i in range(4)]
    want = [{k: v.copy() for k, v in tester.predict(w).items()} for w in wins]
    got = [{k: v.copy() for k, v in r.items()} for r in tester.predict_stream(wins)]
    assert len(got) == len(wins)
    for g, w in zip(got, want):
        for k in w:
            assert np.array_equal(g[k], w[k]), k
    rng = np.random.RandomState(0)
    frames = [rng.randint(0, 256, size=(B, T, 96, 128, 3), dtype=np.uint8) for _ in range(3)]
    boxes = [np.stack([rng.uniform(40, 90, B * T), rng.uniform(30, 70, B * T), rng.uniform(0.8, 1.3, B * T)], 1).reshape(B, T, 3) for _ in range(3)]
    want8 = [{k: v.copy() for k, v in tester.predict_frames(f, b).items()} for f, b in zip(frames, boxes)]
    got8 = [{k: v.copy() for k, v in r.items()} for r in tester.predict_stream(frames, bbox_params=boxes)]
    for g, w in zip(got8, want8):
        for k in ('omegas', 'verts', 'kps_delta'):
            assert np.array_equal(g[k], w[k]), k


def test_split_graph_replay_equals_eager(weights, smpl_model):
    """predict_graphed_split: graph A (up to the dt=0 outputs), host hook, graph B (delta heads) == eager predict, hook sees dt=0."""
    from human_dynamics_b200 import synthetic, HMMRConfig
    from human_dynamics_b200.engine import HMMREngine
    B, T, S = 2, 20, 64
    eng = HMMREngine(weights, smpl_model, HMMRConfig(batch_size=B, sequence_length=T, img_size=S))
    a = torch.from_numpy(synthetic.make_images(B * T, seed=5, size=S)).cuda().view(B, T, S, S, 3)
    eager = {k: v.clone() for k, v in eng.predict(a).items() if not k.startswith('_')}
    seen = []
    buf = a.clone()
    for _ in range(2):
        out, nodes = eng.predict_graphed_split(buf, lambda o: seen.append(sorted(o.keys())))
    torch.cuda.synchronize()
    assert nodes > 100 and len(seen) == 2 and 'verts' in seen[-1] and 'verts_delta' not in seen[-1]      # one hook call per replay
    for k in eager:
        assert torch.equal(out[k], eager[k]), k
