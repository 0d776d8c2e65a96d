"""process_image (src/evaluation/run_video.py:56-107): host geometry on CPU, the CUDA crop kernel on the GPU, both against
the cv2 restatement (oracle/preproc_ref.py) and the committed golden fixture."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'process_image_golden_v1.npz')


def _gen():
    spec = importlib.util.spec_from_file_location('make_preproc_golden', os.path.join(ROOT, 'tests', 'golden', 'make_preproc_golden.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_reproduces_golden_fixture():
    cv2 = pytest.importorskip('cv2')
    from oracle import preproc_ref
    gen = _gen()
    with np.load(GOLD) as z:
        for i, (H, W, cx, cy, s) in enumerate(z['cases']):
            r = preproc_ref.process_image(gen.frame(i, int(H), int(W)), [cx, cy, s])
            assert np.allclose(r['image'][::7, ::7], z['img_%d' % i], atol=1e-6)
            assert list(r['center']) + list(r['start_pt']) + list(r['im_shape']) == list(z['meta_%d' % i])


def test_crop_geometry_matches_reference_bookkeeping():
    """center / start_pt / im_shape (run_video.py:75-100) for the golden cases and a random sweep."""
    pytest.importorskip('cv2')
    from oracle import preproc_ref
    from human_dynamics_b200.preprocess import crop_geometry
    gen = _gen()
    with np.load(GOLD) as z:
        for i, (H, W, cx, cy, s) in enumerate(z['cases']):
            g = crop_geometry((int(H), int(W)), [cx, cy, s])
            assert list(g['center']) + list(g['start_pt']) + list(g['im_shape']) == list(z['meta_%d' % i])
    rng = np.random.RandomState(0)
    for _ in range(200):
        H, W = int(rng.randint(60, 400)), int(rng.randint(60, 400))
        s = float(rng.uniform(0.3, 2.0))
        cx, cy = float(rng.uniform(0, W)), float(rng.uniform(0, H))
        r = preproc_ref.process_image(np.zeros((H, W, 3), np.uint8), [cx, cy, s])
        g = crop_geometry((H, W), [cx, cy, s])
        assert list(g['center']) == list(r['center']) and list(g['start_pt']) == list(r['start_pt'])
        assert g['im_shape'] == r['im_shape'] == [224, 224]
    with pytest.raises(ValueError):
        crop_geometry((100, 100), [900.0, 50.0, 1.0])          # more than one crop away: the reference returns a ragged crop


@pytest.mark.gpu
def test_process_images_kernel_matches_cv2_restatement():
    pytest.importorskip('cv2')
    import torch
    from oracle import preproc_ref
    from human_dynamics_b200.preprocess import process_images
    gen = _gen()
    with np.load(GOLD) as z:
        for i, (H, W, cx, cy, s) in enumerate(z['cases']):
            f = gen.frame(i, int(H), int(W))
            crops, geoms = process_images(f[None], [[cx, cy, s]])
            got = crops[0].cpu().numpy()
            assert np.abs(got[::7, ::7] - z['img_%d' % i]).max() < 2e-6, i                  # committed fixture
            ref = preproc_ref.process_image(f, [cx, cy, s])
            assert np.abs(got - ref['image']).max() < 2e-6, i                                # every pixel, live cv2
            assert list(geoms[0]['center']) == list(ref['center'])
    # a batch of same-size frames with different boxes (the video case), random sweep incl. heavy down / up scaling
    rng = np.random.RandomState(5)
    H, W, N = 270, 480, 12
    frames = np.stack([gen.frame(20 + i, H, W) for i in range(N)])
    boxes = np.stack([rng.uniform(0, W, N), rng.uniform(0, H, N), rng.uniform(0.35, 2.2, N)], axis=1)
    crops, _ = process_images(torch.from_numpy(frames), boxes)
    got = crops.cpu().numpy()
    for i in range(N):
        ref = preproc_ref.process_image(frames[i], boxes[i])
        assert np.abs(got[i] - ref['image']).max() < 2e-6, i
    assert got.min() >= -1.0 and got.max() <= 1.0


@pytest.mark.gpu
def test_run_video_process_image_surface():
    """src.evaluation.run_video.process_image: same dict as the reference (run_video.py:99-107)."""
    pytest.importorskip('cv2')
    from oracle import preproc_ref
    from src.evaluation.run_video import process_image
    gen = _gen()
    f = gen.frame(3, 200, 300)
    got = process_image(f, np.array([150.2, 90.9, 0.8]))
    ref = preproc_ref.process_image(f, [150.2, 90.9, 0.8])
    assert set(got) == {'image', 'im_path', 'im_shape', 'center', 'scale', 'start_pt'}
    assert got['image'].shape == (224, 224, 3) and got['image'].dtype == np.float32
    assert np.abs(got['image'] - ref['image']).max() < 2e-6
    assert list(got['center']) == list(ref['center']) and list(got['start_pt']) == list(ref['start_pt'])
    assert got['im_shape'] == ref['im_shape'] and got['scale'] == ref['scale']


def test_geometry_table_equals_per_frame_crop_geometry():
    """The vectorised host table hd_process_image consumes ({Hs, Ws, x0, y0} per frame, engine.predict_host / predict_frames) against the
    per-frame bookkeeping that is itself checked against the reference's process_image (same float64 formulas -> same integers)."""
    from human_dynamics_b200.preprocess import crop_geometry, geometry_table
    rng = np.random.RandomState(5)
    for _ in range(50):
        H, W = int(rng.randint(60, 500)), int(rng.randint(60, 500))
        n = int(rng.randint(1, 9))
        boxes = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n), rng.uniform(0.3, 2.0, n)], axis=1)
        boxes[0] = [W / 2.0, H / 2.0, 1.0]
        boxes[-1, :2] = np.round(boxes[-1, :2]) + 0.5                      # exact .5 centres: np.round's half-to-even on both sides
        table, infos = geometry_table((H, W), boxes, with_infos=True)
        assert table.dtype == np.int32 and table.shape == (n, 4) and len(infos) == n
        for i in range(n):
            g = crop_geometry((H, W), boxes[i])
            assert list(table[i]) == [g['new_size'][0], g['new_size'][1], g['origin'][0], g['origin'][1]], (H, W, boxes[i])
            assert list(infos[i]['start_pt']) == list(g['start_pt']) and list(infos[i]['center']) == list(g['center'])
    with pytest.raises(ValueError):
        geometry_table((100, 100), [[50.0, 50.0, 1.0], [900.0, 50.0, 1.0]])   # one frame of the track is out of reach
    with pytest.raises(ValueError):
        geometry_table((100, 100), [[50.0, 50.0, 0.001]])                      # scale leaves an empty image


def test_c_crop_geometry_equals_python_bookkeeping():
    """hd_crop_geometry (host-only C entry for consumers without Python) vs crop_geometry on random boxes incl. exact .5 products
    (round-half-to-even) and the refused out-of-reach case."""
    import ctypes as C
    from human_dynamics_b200 import _lib
    from human_dynamics_b200.preprocess import crop_geometry
    rng = np.random.RandomState(11)
    cases = [(rng.randint(60, 500), rng.randint(60, 500), rng.uniform(0, 500), rng.uniform(0, 500), rng.uniform(0.3, 2.0)) for _ in range(300)]
    cases += [(200, 300, 100.5, 50.5, 1.0), (200, 300, 101.5, 51.5, 1.0), (100, 100, 0.5, 99.5, 2.0), (224, 224, 112.0, 112.0, 1.0)]
    ok = 0
    for H, W, cx, cy, s in cases:
        H, W = int(H), int(W)
        cx, cy = min(cx, W), min(cy, H)
        bbox = np.array([cx, cy, s], np.float64)
        geom, cen, sp = np.zeros(4, np.int32), np.zeros(2, np.int32), np.zeros(2, np.int32)
        rc = _lib.lib.hd_crop_geometry(H, W, bbox.ctypes.data_as(C.c_void_p), 224, geom.ctypes.data_as(C.c_void_p),
                                       cen.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p))
        try:
            g = crop_geometry((H, W), bbox)
        except ValueError:
            assert rc == 1
            continue
        assert rc == 0
        ok += 1
        assert geom.tolist() == [g['new_size'][0], g['new_size'][1], g['origin'][0], g['origin'][1]], (H, W, cx, cy, s)
        assert cen.tolist() == list(g['center']) and sp.tolist() == list(g['start_pt'])
    assert ok > 250
    bbox = np.array([900.0, 50.0, 1.0])
    assert _lib.lib.hd_crop_geometry(100, 100, bbox.ctypes.data_as(C.c_void_p), 224, geom.ctypes.data_as(C.c_void_p), None, None) == 1
