"""Deterministic synthetic inputs for the PWCLO-Net hot path.

`make_pair(seed, n_points)` ray-casts a 64-beam x 2048-azimuth spinning LiDAR (elevations
+3..-24 deg, reference config/dataset/kitti_odometry.yaml:4-7) into a box-world, from two ego
poses, expresses the returns in the reference's camera-style frame (x right, y down, z forward --
what `Tr` produces, slam/dataset/kitti_odometry_dataset.py:386-394), applies the reference's
crop (`filter_pcd`, kitti_odometry_dataset.py:149-172: drop y > 1.1, keep |x|,|z| < 30) and
draws `n_points` returns without replacement.  No duplicate rows, no point inside the FPS origin
ball (||p||^2 <= 1e-3).

`make_state_dict(shapes, seed)` fills a PWCLO-Net state dict (reference layout) with
seeded values: xavier-uniform-range conv weights, non-trivial BatchNorm affine + running stats so
that BN folding is exercised.
"""
import numpy as np

SENSOR_HEIGHT = 1.73
N_BEAMS = 64
N_AZIMUTH = 2048
MAX_RANGE = 80.0
SEED_BASE = 20240000


def _scene(rng):
    n_box = 60
    wl = rng.uniform(1.0, 12.0, size=(n_box, 2))
    h = rng.uniform(1.5, 15.0, size=n_box)
    c = np.empty((0, 2))
    while c.shape[0] < n_box:
        cand = rng.uniform(-45.0, 45.0, size=(2 * n_box, 2))
        cand = cand[np.hypot(cand[:, 0], cand[:, 1]) > 3.0 + 9.0]  # keep the ego disc free of box bodies
        c = np.concatenate([c, cand])[:n_box]
    lo = np.stack([c[:, 0] - wl[:, 0] / 2, SENSOR_HEIGHT - h, c[:, 1] - wl[:, 1] / 2], 1)
    hi = np.stack([c[:, 0] + wl[:, 0] / 2, np.full(n_box, SENSOR_HEIGHT), c[:, 1] + wl[:, 1] / 2], 1)
    # four long walls at +-U[15,35] m
    d = rng.uniform(15.0, 35.0, size=4)
    wh = rng.uniform(3.0, 6.0, size=4)
    walls_lo = np.array([[d[0], 0, -60.0], [-d[1] - 0.5, 0, -60.0], [-60.0, 0, d[2]], [-60.0, 0, -d[3] - 0.5]])
    walls_hi = np.array([[d[0] + 0.5, 0, 60.0], [-d[1], 0, 60.0], [60.0, 0, d[2] + 0.5], [60.0, 0, -d[3]]])
    walls_lo[:, 1] = SENSOR_HEIGHT - wh
    walls_hi[:, 1] = SENSOR_HEIGHT
    return np.concatenate([lo, walls_lo]).astype(np.float64), np.concatenate([hi, walls_hi]).astype(np.float64)


def _local_dirs():
    el = np.deg2rad(np.linspace(3.0, -24.0, N_BEAMS))
    az = np.linspace(0.0, 2.0 * np.pi, N_AZIMUTH, endpoint=False)
    el, az = np.meshgrid(el, az, indexing="ij")
    d = np.stack([np.sin(az) * np.cos(el), -np.sin(el), np.cos(az) * np.cos(el)], -1)
    return d.reshape(-1, 3)


def _scan(lo, hi, origin, rot, rng):
    """First-hit ranges of all rays; returns points in the sensor's own frame."""
    dl = _local_dirs()
    dw = dl @ rot.T
    t_best = np.full(dw.shape[0], np.inf)
    # ground plane y = SENSOR_HEIGHT (world)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (SENSOR_HEIGHT - origin[1]) / dw[:, 1]
    tg[~(dw[:, 1] > 1e-9)] = np.inf
    t_best = np.minimum(t_best, tg)
    # axis-aligned boxes, slab method (fp32, per-axis to keep temporaries at [rays, boxes])
    dws = np.where(np.abs(dw) < 1e-9, 1e-9, dw).astype(np.float32)
    inv = (1.0 / dws)
    lo32 = (lo - origin).astype(np.float32)
    hi32 = (hi - origin).astype(np.float32)
    tn = np.full((dw.shape[0], lo.shape[0]), -np.inf, np.float32)
    tf = np.full((dw.shape[0], lo.shape[0]), np.inf, np.float32)
    for a in range(3):
        t0 = inv[:, a:a + 1] * lo32[None, :, a]
        t1 = inv[:, a:a + 1] * hi32[None, :, a]
        np.maximum(tn, np.minimum(t0, t1), out=tn)
        np.minimum(tf, np.maximum(t0, t1), out=tf)
    hit = (tf >= tn) & (tn > 0.0)
    tn[~hit] = np.inf
    t_best = np.minimum(t_best, tn.min(axis=1).astype(np.float64))
    ok = np.isfinite(t_best) & (t_best <= MAX_RANGE)
    r = t_best[ok] + rng.normal(0.0, 0.02, size=int(ok.sum()))
    return dl[ok] * r[:, None]


def _crop_and_sample(pts, n_points, rng):
    keep = (pts[:, 1] <= 1.1) & (np.abs(pts[:, 0]) < 30.0) & (np.abs(pts[:, 2]) < 30.0)
    idx = np.nonzero(keep)[0]
    if idx.size < n_points:
        return None
    sel = rng.choice(idx, n_points, replace=False)
    out = pts[sel].astype(np.float32)
    if np.unique(out, axis=0).shape[0] != n_points:
        return None
    if np.any((out.astype(np.float64) ** 2).sum(1) <= 2e-3):
        return None
    return out


def make_pair(seed, n_points=8192):
    """Returns dict(pc1[N,3] f32, pc2[N,3] f32, q[4] (w,x,y,z), t[3]) with p2 = R(q) p1 + t."""
    for attempt in range(16):
        rng = np.random.Generator(np.random.PCG64([int(seed), attempt]))
        lo, hi = _scene(rng)
        yaw = np.deg2rad(rng.normal(0.0, 1.0))
        trans = np.array([rng.normal(0.0, 0.05), rng.normal(0.0, 0.02), 1.0 + rng.normal(0.0, 0.3)])
        c, s = np.cos(yaw), np.sin(yaw)
        rot = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])  # sensor-2 axes in world (yaw about y)
        p1 = _crop_and_sample(_scan(lo, hi, np.zeros(3), np.eye(3), rng), n_points, rng)
        p2 = _crop_and_sample(_scan(lo, hi, trans, rot, rng), n_points, rng)
        if p1 is None or p2 is None:
            continue
        r21 = rot.T
        t21 = -rot.T @ trans
        q = np.array([np.cos(-yaw / 2.0), 0.0, np.sin(-yaw / 2.0), 0.0])
        return {"pc1": p1, "pc2": p2, "q": q.astype(np.float32), "t": t21.astype(np.float32), "R": r21}
    raise RuntimeError(f"synthetic scene for seed {seed} never produced {n_points} returns")


def make_batch(first_pair, n_pairs, n_points=8192):
    """[B,3,N] fp32 arrays in the layout PWCLONet.forward takes (reference pwclo_net.py:109-126)."""
    pairs = [make_pair(SEED_BASE + first_pair + i, n_points) for i in range(n_pairs)]
    xyz1 = np.stack([p["pc1"].T for p in pairs]).astype(np.float32)
    xyz2 = np.stack([p["pc2"].T for p in pairs]).astype(np.float32)
    gt = np.stack([np.concatenate([p["t"], p["q"]]) for p in pairs]).astype(np.float32)
    return np.ascontiguousarray(xyz1), np.ascontiguousarray(xyz2), gt


def make_state_dict(shapes, seed=1):
    """shapes: ordered mapping name -> tuple.  Values are a pure function of (name order, seed)."""
    rng = np.random.Generator(np.random.PCG64(int(seed)))
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros(shp, np.int64)
        elif name.endswith("running_var") or name.endswith("bn.weight"):
            out[name] = rng.uniform(0.5, 1.5, size=shp).astype(np.float32)
        elif name.endswith("running_mean") or name.endswith("bn.bias"):
            out[name] = rng.normal(0.0, 0.1, size=shp).astype(np.float32)
        elif name.endswith("conv.bias"):
            out[name] = rng.normal(0.0, 0.05, size=shp).astype(np.float32)
        elif name.endswith("conv.weight"):
            fan_out, fan_in = shp[0], int(np.prod(shp[1:]))
            a = np.sqrt(6.0 / (fan_in + fan_out))
            out[name] = rng.uniform(-a, a, size=shp).astype(np.float32)
        else:
            raise KeyError(f"unexpected parameter name {name}")
    return out


# KITTI odometry sequence-00 style velodyne -> camera calibration (rows 0..2 of the 4x4 `Tr`,
# slam/dataset/kitti_odometry_dataset.py:353-355): x_cam = -y_velo, y_cam = -z_velo, z_cam = x_velo up to a
# sub-degree rotation and a few centimetres of lever arm.
KITTI_TR = np.array([[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
                     [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
                     [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01]], np.float64)


def make_raw_scan(seed):
    """A synthetic KITTI `.bin` payload: float32 [n,4] = (x, y, z, reflectance) in the VELODYNE frame, i.e. what
    `np.fromfile(..., float32).reshape(-1, 4)` returns (kitti_odometry_dataset.py:375-376), such that
    KITTI_TR maps it to the camera-style frame of `make_pair`.  All ~10^5 returns of one revolution, unfiltered,
    in beam-major firing order."""
    rng = np.random.Generator(np.random.PCG64([int(seed), 77]))
    lo, hi = _scene(rng)
    cam = _scan(lo, hi, np.zeros(3), np.eye(3), rng)
    T = np.vstack([KITTI_TR, [0.0, 0.0, 0.0, 1.0]])
    Ti = np.linalg.inv(T)
    velo = cam @ Ti[:3, :3].T + Ti[:3, 3]
    refl = rng.uniform(0.0, 1.0, size=(velo.shape[0], 1))
    return np.ascontiguousarray(np.concatenate([velo, refl], axis=1), np.float32)
