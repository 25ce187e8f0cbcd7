"""Shared helpers for the parity tests (golden loading, seeded inputs/weights, pose metrics)."""
import hashlib
import os

import numpy as np

from pwclonet_pylidarslam_b200 import synthetic as syn

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances stated by BASELINE.json north_star
TOL_FEATURE_REL = 1e-4
TOL_TRANSLATION_M = 1e-4
TOL_ROTATION_RAD = 1e-5


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def load_golden(tag):
    g = dict(np.load(os.path.join(GOLD_DIR, tag + ".npz")))
    first_pair, n_pairs, n_points, weight_seed = [int(v) for v in g["meta"]]
    x1, x2, gt = syn.make_batch(first_pair, n_pairs, n_points)
    assert sha(x1, x2) == str(g["input_sha"]), "synthetic generator drifted from the golden inputs"
    return g, x1, x2, weight_seed


def weights_for(shapes, seed, g=None):
    w = syn.make_state_dict(shapes, seed=seed)
    if g is not None:
        assert sha(*[w[k] for k in sorted(w)]) == str(g["weights_sha"]), "weight generator drifted"
    return w


def model_shapes():
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    return {k: tuple(v.shape) for k, v in PWCLONet({"device": "cpu"}).state_dict().items()}


def quat_angle(qa, qb):
    """rotation angle (rad) between unit quaternions, sign-invariant"""
    qa = qa / np.linalg.norm(qa, axis=-1, keepdims=True)
    qb = qb / np.linalg.norm(qb, axis=-1, keepdims=True)
    d = np.abs(np.sum(qa.astype(np.float64) * qb.astype(np.float64), axis=-1))
    # 2*acos(d) loses precision near d=1: use the chord |qa -+ qb|
    chord = np.minimum(np.linalg.norm(qa.astype(np.float64) - qb, axis=-1), np.linalg.norm(qa.astype(np.float64) + qb, axis=-1))
    return 2.0 * np.arcsin(np.clip(chord / 2.0, 0, 1)) * 2.0 / 2.0 if d is not None else None


def pose_errors(pose, ref):
    """pose, ref: [B,4,7] = (t, q) -> (max translation error [m], max rotation error [rad])"""
    te = np.abs(pose[..., :3].astype(np.float64) - ref[..., :3]).max()
    re = quat_angle(pose[..., 3:], ref[..., 3:]).max()
    return float(te), float(re)


def rel_err(a, b):
    """max |a-b| / max|b| (per-tensor relative error, the 'features within 1e-4 relative' metric)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def knn_rows_match(idx, g, i):
    """compare an index tensor with golden kNN call i: row sums everywhere, sorted rows on the stored
    subset; returns the number of rows that differ (ties / near-ties)"""
    idx = np.asarray(idx)
    rs = idx.sum(-1).astype(np.int32)
    bad = int((rs != g[f"knn_{i}_rowsum"]).sum())
    stride = max(1, idx.shape[1] // 32)
    sub = np.sort(idx[:, ::stride], axis=-1)
    bad_rows = int((sub != g[f"knn_{i}_rows"]).any(-1).sum())
    return bad, bad_rows
