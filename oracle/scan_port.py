"""oracle/scan_port.py -- TEST INFRASTRUCTURE ONLY (checker for the input-pipeline row N3).

numpy restatement of the reference's per-frame preprocessing,
slam/dataset/kitti_odometry_dataset.py:375-397 (float64 `Tr` transform of the float32 scan) and
`filter_pcd` :149-172 (ground / range mask, `npoints` random survivors), with the random subset
defined by the counter-based generator the CUDA path uses (philox4x32-10) instead of numpy's global
MT19937 stream: the sample is the npoints survivors with the smallest (key, index), in that order.

Pinned: `philox4x32_10` against the Random123 known-answer vectors; the transform + mask against the
literal reference expression (np.matmul with the homogeneous column) -- tests/test_scan_cpu.py; the
selection against its definition (a uniformly random subset without replacement: every survivor is
equally likely, checked statistically).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised over numpy uint32 arrays (counter words) with a scalar key; returns 4 uint32 arrays"""
    c = [np.asarray(x, np.uint32).copy() for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c


def point_keys(n, scan, seed):
    """31-bit selection key of every point of scan `scan`: point i uses Philox counter 32*(i//128) + i%32 and
    output word (i//32)%4 (one call per lane of the warp that streams a 128-point block)"""
    i = np.arange(n, dtype=np.int64)
    ctr = (((i >> 7) << 5) | (i & 31)).astype(np.uint32)
    r = philox4x32_10(ctr, np.uint32(scan), np.uint32(0), np.uint32(0), seed & 0xFFFFFFFF, seed >> 32)
    word = (i >> 5) & 3
    k = np.choose(word, r)
    return (k >> np.uint32(1)).astype(np.uint32)


def draws(count, scan, seed):
    g = np.arange((count + 3) // 4, dtype=np.uint32)
    r = philox4x32_10(g, np.uint32(scan), np.uint32(1), np.uint32(0), seed & 0xFFFFFFFF, seed >> 32)
    return np.stack(r, axis=1).reshape(-1)[:count]


def _two_prod(a, b):
    """error-free product (Veltkamp / Dekker): a*b = p + e exactly"""
    p = a * b
    c = 134217729.0                      # 2^27 + 1
    ah = c * a
    ah = ah - (ah - a)
    al = a - ah
    bh = c * b
    bh = bh - (bh - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def fma(a, b, c):
    """round(a*b + c) with one rounding, vectorised: TwoProduct + TwoSum, then one add of the error terms
    (differs from a hardware FMA only when a*b+c lies within 2^-106 relative of a rounding boundary)"""
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64))
    p, e = _two_prod(a, b)
    s = p + c
    bb = s - p
    t = (p - (s - bb)) + (c - bb)
    return s + (t + e)


def affine(T, xyz):
    """rows of T (3x4, float64) applied as fma(T3, 1, fma(T2, z, fma(T1, y, T0*x))): the order in which
    np.matmul(Tr, [x y z 1]^T) accumulates its 4-term dot products (one multiply, then fused multiply-adds in
    storage order) -- checked bit for bit against np.matmul in tests/test_scan_cpu.py"""
    T = np.asarray(T, np.float64).reshape(3, 4)
    x, y, z = (xyz[:, i].astype(np.float64) for i in range(3))
    return np.stack([fma(T[r, 2], z, fma(T[r, 1], y, T[r, 0] * x)) + T[r, 3] for r in range(3)], axis=1)


KITTI_ODOMETRY_CROP = (1, +1, 1.1, 0, 2, 30.0)


def keep_mask(P, crop=KITTI_ODOMETRY_CROP):
    """filter_pcd on the transformed float64 points: kitti_odometry_dataset.py:151-159 by default; any crop
    (ground_axis, ground_sign, ground_thr, near_a, near_b, near_thr), e.g. kitti_360_dataset_2.py:113-123"""
    ga, gs, gt, a, b, nt = crop
    ground = P[:, ga] > gt if gs > 0 else P[:, ga] < gt
    near = np.logical_and(np.logical_and(P[:, a] < nt, P[:, a] > -nt), np.logical_and(P[:, b] < nt, P[:, b] > -nt))
    return np.logical_and(np.logical_not(ground), near)


def reference_mask_kitti360(points_f32, near_treshold=30.0):
    """the literal reference expression of kitti_360_dataset_2.py:113-123 on the float32 velodyne points, for pinning"""
    wheel_axis_z = -(1.73 - 0.3)
    is_ground = points_f32[:, 2] < wheel_axis_z
    near_x = np.logical_and(points_f32[:, 0] < near_treshold, points_f32[:, 0] > -near_treshold)
    near_y = np.logical_and(points_f32[:, 1] < near_treshold, points_f32[:, 1] > -near_treshold)
    return np.logical_and(np.logical_not(is_ground), np.logical_and(near_x, near_y))


def select(mask, scan, seed, npoints):
    """indices of the sample, in output order"""
    n = mask.shape[0]
    idx = np.where(mask)[0]
    M = idx.shape[0]
    if M >= npoints:
        k = point_keys(n, scan, seed)[idx].astype(np.uint64)
        comp = (k << np.uint64(32)) | idx.astype(np.uint64)
        return idx[np.argsort(comp, kind="stable")[:npoints]].astype(np.int32), M
    u = draws(npoints - M, scan, seed).astype(np.uint64)
    if M > 0:
        extra = idx[((u * np.uint64(M)) >> np.uint64(32)).astype(np.int64)]
        return np.concatenate([idx, extra]).astype(np.int32), M
    return ((u * np.uint64(n)) >> np.uint64(32)).astype(np.int32), 0


def prepare_scan(raw, Tr, scan, seed, npoints, post=None, crop=KITTI_ODOMETRY_CROP):
    """raw float32[n,4] -> (float32[npoints,3], int32[npoints] source rows, survivors)"""
    P = affine(Tr, raw[:, :3])
    sel, M = select(keep_mask(P, crop), scan, seed, npoints)
    Q = P[sel]
    if post is not None:
        Q = affine(post, Q)
    return Q.astype(np.float32), sel, M


def reference_transform_and_mask(raw, Tr4):
    """the literal reference code path (kitti_odometry_dataset.py:383-394 + :151-159), for pinning"""
    n = raw.shape[0]
    pts = np.concatenate([raw[:n, :3], np.ones((n, 1))], axis=-1)
    pts = np.matmul(Tr4, pts.T).T[:, :3]
    is_ground = pts[:, 1] > 1.1
    near_x = np.logical_and(pts[:, 0] < 30, pts[:, 0] > -30)
    near_y = np.logical_and(pts[:, 2] < 30, pts[:, 2] > -30)
    return pts, np.logical_and(np.logical_not(is_ground), np.logical_and(near_x, near_y))
