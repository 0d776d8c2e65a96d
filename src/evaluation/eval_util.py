"""Evaluation metrics on predicted joints / vertices / keypoints -- drop-in for the reference's
src/evaluation/eval_util.py (:14-153 error metrics, :156-254 alignment helpers, :257-344 bookkeeping / conversions).

Same function names, arguments and return values; the bodies are batched numpy (one SVD call for all frames instead of a
Python loop per frame, closed-form Rodrigues instead of a cv2 call per joint).  This is host-side bookkeeping after the hot
path: it consumes the arrays `Tester.predict` returns and never touches the device.
"""
import numpy as np


# ----------------------------------------------------------------------------------------------- error metrics
def compute_accel(joints):
    """joints (N x K x 3) -> mean joint acceleration per frame triple (N-2)   (eval_util.py:14-27)."""
    j = np.asarray(joints)
    second_diff = j[2:] - 2 * j[1:-1] + j[:-2]                   # (v[t+1] - v[t]) with v = first difference
    return np.linalg.norm(second_diff, axis=2).mean(axis=1)


def compute_error_accel(joints_gt, joints_pred, vis=None):
    """Acceleration error per frame triple; a triple counts only if all three frames are visible (eval_util.py:63-96)."""
    gt, pr = np.asarray(joints_gt), np.asarray(joints_pred)
    diff = (pr[:-2] - 2 * pr[1:-1] + pr[2:]) - (gt[:-2] - 2 * gt[1:-1] + gt[2:])
    normed = np.linalg.norm(diff, axis=2)
    if vis is None:
        keep = np.ones(len(normed), dtype=bool)
    else:
        v = np.asarray(vis).astype(bool)
        keep = v[:-2] & v[1:-1] & v[2:]
    return normed[keep].mean(axis=1)


def compute_error_verts(verts_gt, verts_pred):
    """Mean per-vertex Euclidean error per frame (N)   (eval_util.py:140-153)."""
    a, b = np.asarray(verts_gt), np.asarray(verts_pred)
    assert len(a) == len(b)
    return np.linalg.norm(a - b, axis=2).mean(axis=1)


def align_by_pelvis(joints, get_pelvis=False):
    """Translate so that the hip midpoint (LSP joints 3 and 2) is the origin; (14 x 3) or batched (... x 14 x 3)   (:156-173)."""
    j = np.asarray(joints)
    pelvis = (j[..., 3, :] + j[..., 2, :]) / 2.0
    out = j - pelvis[..., None, :]
    return (out, pelvis) if get_pelvis else out


def _procrustes_batch(X, Y):
    """X, Y (B x N x D): similarity transform of every X[b] onto Y[b] (optimal scale, rotation with det +1, translation)."""
    mx, my = X.mean(axis=1, keepdims=True), Y.mean(axis=1, keepdims=True)
    Xc, Yc = X - mx, Y - my
    var = (Xc ** 2).sum(axis=(1, 2))
    K = np.einsum('bnd,bne->bde', Xc, Yc)                        # D x D cross-covariance per item
    U, _, Vh = np.linalg.svd(K)
    V = np.swapaxes(Vh, 1, 2)
    sign = np.sign(np.linalg.det(np.einsum('bij,bkj->bik', U, V)))
    Z = np.tile(np.eye(X.shape[2]), (len(X), 1, 1))
    Z[:, -1, -1] = sign
    R = V @ Z @ np.swapaxes(U, 1, 2)
    scale = np.einsum('bij,bji->b', R, K) / var
    t = np.swapaxes(my, 1, 2) - scale[:, None, None] * (R @ np.swapaxes(mx, 1, 2))
    return scale[:, None, None] * np.einsum('bij,bnj->bni', R, X) + np.swapaxes(t, 1, 2)


def compute_similarity_transform(S1, S2):
    """Orthogonal Procrustes: S1 after the best similarity transform onto S2.  3 x N (or 2 x N) like the reference, or N x 3
    (then the result is N x 3 too)   (eval_util.py:176-233)."""
    S1, S2 = np.asarray(S1, np.float64), np.asarray(S2, np.float64)
    transposed = S1.shape[0] != 3 and S1.shape[0] != 2
    X, Y = (S1, S2) if transposed else (S1.T, S2.T)
    assert X.shape[0] == Y.shape[0]
    out = _procrustes_batch(X[None], Y[None])[0]
    return out if transposed else out.T


def compute_error_3d(gt3ds, preds, vis=None):
    """MPJPE after pelvis alignment and after Procrustes alignment, per visible frame (two lists)   (eval_util.py:30-60)."""
    assert len(gt3ds) == len(preds)
    gt = np.asarray(gt3ds, np.float64).reshape(len(gt3ds), -1, 3)
    pr = np.asarray(preds, np.float64).reshape(len(preds), -1, 3)
    if vis is not None:
        keep = np.asarray(vis).astype(bool)
        gt, pr = gt[keep], pr[keep]
    if len(gt) == 0:
        return [], []
    gt, pr = align_by_pelvis(gt), align_by_pelvis(pr)
    errors = np.linalg.norm(gt - pr, axis=2).mean(axis=1)
    errors_pa = np.linalg.norm(gt - _procrustes_batch(pr, gt), axis=2).mean(axis=1)
    return list(errors), list(errors_pa)


def compute_opt_cam_with_vis(got, want, vis):
    """Optimal weak-perspective camera [scale, tx, ty] mapping 2D keypoints `got` onto `want` over the visible ones   (:236-264)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    vis = np.asarray(vis).astype(bool)
    w = vis[:, None].astype(np.float64)
    n = vis.sum()
    mu1, mu2 = (got * w).sum(axis=0) / n, (want * w).sum(axis=0) / n
    x, y = w * (got - mu1), w * (want - mu2)
    a_inv = np.linalg.inv(x.T @ x + 1e-6 * np.identity(2))
    scale = np.trace(a_inv @ (x.T @ y)) / 2.0
    trans = mu2 / scale - mu1
    return scale * (got + trans), np.hstack((scale, trans.ravel()))


def compute_error_kp(kps_gt, kps_pred, alpha=0.05, min_visible=6):
    """2D keypoint error, error after optimal-camera alignment and PCK@alpha; NaN where fewer than `min_visible` keypoints are
    annotated   (eval_util.py:99-137)."""
    assert len(kps_gt) == len(kps_pred)
    e_kp, e_pa, e_pck = [], [], []
    for kp_gt, kp_pred in zip(np.asarray(kps_gt), np.asarray(kps_pred)):
        vis = kp_gt[:, 2].astype(bool)
        if vis.sum() < max(min_visible, 1):
            e_kp.append(np.nan); e_pa.append(np.nan); e_pck.append(np.nan)
            continue
        gt2 = kp_gt[:, :2]
        aligned, _ = compute_opt_cam_with_vis(got=kp_pred, want=gt2, vis=vis)
        d = np.linalg.norm(gt2[vis] - kp_pred[vis], axis=1)
        d_pa = np.linalg.norm(gt2[vis] - aligned[vis], axis=1)
        e_kp.append(d.mean()); e_pa.append(d_pa.mean()); e_pck.append((d_pa < alpha).mean())
    return e_kp, e_pa, e_pck


# ----------------------------------------------------------------------------------------------- accumulators (:267-314)
def concat_dict_entries(dictionary):
    for k in list(dictionary):
        dictionary[k] = np.concatenate(dictionary[k])


def extend_dict_entries(accumulator, appender):
    for k, v in appender.items():
        dst = accumulator.setdefault(k, [])
        if hasattr(v, '__iter__'):
            dst.extend(v)
        else:
            dst.append(v)


def mean_of_dict_values(dictionary):
    for k in list(dictionary):
        dictionary[k] = float(round(np.nanmean([np.nanmean(values) for values in dictionary[k]]), 5))


def update_dict_entries(accumulator, appender):
    for k, v in appender.items():
        accumulator.setdefault(k, []).append(v)


# ----------------------------------------------------------------------------------------------- conversions (:319-344)
def axis_angle_to_rot_mat(poses_aa):
    """poses_aa (72) -> (24 x 3 x 3): Rodrigues' formula for all joints at once (the reference calls cv2.Rodrigues per joint)."""
    aa = np.asarray(poses_aa, np.float64).reshape(-1, 3)
    theta = np.linalg.norm(aa, axis=1)
    safe = np.where(theta < 1e-12, 1.0, theta)
    k = aa / safe[:, None]
    Kx = np.zeros((len(aa), 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    s, c = np.sin(theta)[:, None, None], np.cos(theta)[:, None, None]
    R = np.eye(3)[None] + s * Kx + (1 - c) * (Kx @ Kx)
    R[theta < 1e-12] = np.eye(3)
    return R


def rot_mat_to_axis_angle(rot_matrices):
    """(24 x 3 x 3) -> (72).  Angle from the trace, axis from the antisymmetric part; near pi the axis comes from the diagonal."""
    R = np.asarray(rot_matrices, np.float64).reshape(-1, 3, 3)
    cos = np.clip((np.trace(R, axis1=1, axis2=2) - 1) / 2.0, -1.0, 1.0)
    theta = np.arccos(cos)
    w = np.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], axis=1)
    norm = np.linalg.norm(w, axis=1)
    out = np.zeros((len(R), 3))
    ok = norm > 1e-8
    out[ok] = w[ok] / norm[ok, None] * theta[ok, None]
    near_pi = (~ok) & (theta > 1.0)                      # sin(theta) ~ 0 with theta ~ pi: R = 2 a a^T - I
    for i in np.nonzero(near_pi)[0]:
        a = np.sqrt(np.maximum((np.diag(R[i]) + 1) / 2.0, 0.0))
        j = int(np.argmax(a))
        sgn = np.sign(R[i][j] + (np.arange(3) == j))     # fix the relative signs from row j of (R + I) / 2
        out[i] = theta[i] * a * np.where(sgn == 0, 1.0, sgn)
    return out.reshape(-1)
