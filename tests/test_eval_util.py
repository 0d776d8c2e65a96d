"""src/evaluation/eval_util.py (drop-in for the reference's eval_util.py:14-344) against independent formulations:
explicit loops, scipy's orthogonal Procrustes, cv2.Rodrigues, and hand-made known answers."""
import numpy as np
import pytest

from src.evaluation import eval_util as E


def test_accel_and_accel_error_against_loops():
    rng = np.random.RandomState(0)
    j = rng.normal(size=(12, 25, 3))
    acc = E.compute_accel(j)
    ref = [np.mean([np.linalg.norm((j[t + 2, k] - j[t + 1, k]) - (j[t + 1, k] - j[t, k])) for k in range(25)]) for t in range(10)]
    assert np.allclose(acc, ref)
    g, p = rng.normal(size=(9, 14, 3)), rng.normal(size=(9, 14, 3))
    vis = np.array([1, 1, 1, 0, 1, 1, 1, 1, 1], bool)
    err = E.compute_error_accel(g, p, vis)
    keep = [t for t in range(7) if vis[t] and vis[t + 1] and vis[t + 2]]
    assert keep == [0, 4, 5, 6]
    ref = [np.mean([np.linalg.norm((p[t] - 2 * p[t + 1] + p[t + 2])[k] - (g[t] - 2 * g[t + 1] + g[t + 2])[k]) for k in range(14)]) for t in keep]
    assert np.allclose(err, ref)
    assert len(E.compute_error_accel(g, p)) == 7
    # a constant-velocity track has zero acceleration
    lin = np.arange(6)[:, None, None] * np.ones((1, 14, 3))
    assert np.allclose(E.compute_accel(lin), 0)


def test_vertex_error_and_pelvis_alignment():
    rng = np.random.RandomState(1)
    a = rng.normal(size=(3, 6890, 3))
    assert np.allclose(E.compute_error_verts(a, a + np.array([0.3, 0.0, 0.4])), 0.5)
    j = rng.normal(size=(14, 3))
    al, pel = E.align_by_pelvis(j, get_pelvis=True)
    assert np.allclose(pel, (j[2] + j[3]) / 2) and np.allclose((al[2] + al[3]) / 2, 0)
    assert np.allclose(E.align_by_pelvis(np.stack([j, j + 1]))[1], al)


def test_similarity_transform_recovers_known_transform_and_matches_scipy():
    from scipy.linalg import orthogonal_procrustes
    from scipy.spatial.transform import Rotation
    rng = np.random.RandomState(2)
    X = rng.normal(size=(14, 3))
    R = Rotation.from_rotvec([0.3, -1.1, 0.7]).as_matrix()
    Y = 1.7 * X @ R.T + np.array([0.5, -2.0, 3.0])
    assert np.allclose(E.compute_similarity_transform(X, Y), Y, atol=1e-9)            # N x 3 in, N x 3 out
    assert np.allclose(E.compute_similarity_transform(X.T, Y.T), Y.T, atol=1e-9)      # 3 x N like the reference docstring
    # noisy target: optimal rotation equals scipy's, residual is not larger than for the true transform
    Yn = Y + 0.05 * rng.normal(size=Y.shape)
    out = E.compute_similarity_transform(X, Yn)
    Xc, Yc = X - X.mean(0), Yn - Yn.mean(0)
    Rs, sca = orthogonal_procrustes(Xc, Yc)
    s = sca / (Xc ** 2).sum()
    assert np.allclose(out, s * Xc @ Rs + Yn.mean(0), atol=1e-8)
    # reflections are not allowed: det(R) = +1 even if a mirror image would fit better
    M = X * np.array([1, 1, -1.0])
    o = E.compute_similarity_transform(X, M)
    Rfit = np.linalg.lstsq(X - X.mean(0), o - o.mean(0), rcond=None)[0]
    assert np.linalg.det(Rfit) > 0


def test_error_3d_and_keypoint_errors():
    from scipy.spatial.transform import Rotation
    rng = np.random.RandomState(3)
    gt = rng.normal(size=(5, 14, 3))
    R = Rotation.from_rotvec([0.2, 0.1, -0.4]).as_matrix()
    pred = 0.8 * gt @ R.T + np.array([1.0, 2.0, 3.0])
    e, e_pa = E.compute_error_3d(gt, pred)
    assert len(e) == 5 and np.allclose(e_pa, 0, atol=1e-9) and min(e) > 0.05
    e2, e2_pa = E.compute_error_3d(gt, pred, vis=[1, 0, 0, 1, 0])
    assert np.allclose(e2, [e[0], e[3]])
    # per-frame loop version of the same metric
    for i in range(5):
        g0 = gt[i] - (gt[i, 2] + gt[i, 3]) / 2
        p0 = pred[i] - (pred[i, 2] + pred[i, 3]) / 2
        assert np.isclose(e[i], np.mean(np.linalg.norm(g0 - p0, axis=1)))
    # 2D: prediction = ground truth seen through another weak-perspective camera -> zero error after alignment, PCK 1
    kp = rng.uniform(-1, 1, size=(4, 25, 2))
    vis = np.ones((4, 25, 1)); vis[0, :22] = 0                      # frame 0: only 3 visible < min_visible -> NaN
    vis[1, ::2] = 0
    gt2 = np.concatenate([kp, vis], axis=2)
    pr2 = (kp / 1.3) - np.array([0.2, -0.1])
    a, b, c = E.compute_error_kp(gt2, pr2)
    assert np.isnan(a[0]) and np.isnan(b[0]) and np.isnan(c[0])
    assert np.allclose(b[1:], 0, atol=1e-4) and np.allclose(c[1:], 1.0) and min(a[1:]) > 0.05
    aligned, cam = E.compute_opt_cam_with_vis(pr2[2], kp[2], np.ones(25, bool))
    assert np.allclose(aligned, kp[2], atol=1e-4) and np.isclose(cam[0], 1.3, atol=1e-4)


def test_rotation_conversions_match_cv2():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.RandomState(4)
    pose = rng.normal(0, 0.7, size=72)
    pose[3:6] = 0.0
    pose[6:9] = np.array([np.pi, 0, 0]) * 0.999999
    R = E.axis_angle_to_rot_mat(pose)
    ref = np.array([cv2.Rodrigues(p)[0] for p in pose.reshape(-1, 3)])
    assert R.shape == (24, 3, 3) and np.allclose(R, ref, atol=1e-9)
    back = E.rot_mat_to_axis_angle(R)
    ref_back = np.array([cv2.Rodrigues(r)[0] for r in ref]).reshape(72)
    assert back.shape == (72,)
    keep = np.ones(72, bool); keep[6:9] = False                    # the almost-pi joint: acos is ill-conditioned there, compare it with the input
    assert np.allclose(back[keep], ref_back[keep], atol=1e-6) and np.allclose(back[6:9], pose[6:9], atol=2e-5)
    exact_pi = E.rot_mat_to_axis_angle(np.diag([1.0, -1.0, -1.0])[None])
    assert np.allclose(np.abs(exact_pi), [np.pi, 0, 0], atol=1e-9)


def test_accumulators():
    acc = {}
    E.update_dict_entries(acc, {'a': 1.0, 'b': [1, 2]})
    E.update_dict_entries(acc, {'a': 3.0, 'b': [3]})
    assert acc == {'a': [1.0, 3.0], 'b': [[1, 2], [3]]}
    ext = {}
    E.extend_dict_entries(ext, {'x': [1, 2], 'y': 5})
    E.extend_dict_entries(ext, {'x': [3], 'y': 6})
    assert ext == {'x': [1, 2, 3], 'y': [5, 6]}
    d = {'m': [[1.0, 3.0], [np.nan, 5.0]]}
    E.mean_of_dict_values(d)
    assert d['m'] == 3.5
    c = {'v': [np.ones(2), np.zeros(3)]}
    E.concat_dict_entries(c)
    assert c['v'].shape == (5,)
