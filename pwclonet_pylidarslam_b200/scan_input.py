"""GPU input pipeline: raw KITTI scans -> the [B,N,3] float32 clouds PWCLO-Net consumes.

Replaces, per frame, the host numpy code of slam/dataset/kitti_odometry_dataset.py:375-397 (`Tr`
transform in float64) and `filter_pcd` :149-172 (ground / range crop + `npoints` random survivors), and
the optional augmentation transform :404-446, by `pwclo_prepare_scans` (two launches) over a packed batch
of scans.  There is no CPU path: tensors must live on the GPU.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def pack_scans(scans, pin=True):
    """list of float32 [n_i,4] arrays (what np.fromfile(bin, float32).reshape(-1, 4) returns) ->
    (packed float32 [sum n_i, 4] host tensor, int64 [len+1] offsets, max n_i).  For a frame PAIR the reference
    truncates both scans to the shorter one first (kitti_odometry_dataset.py:378-384): see `pack_pairs`."""
    sizes = [int(s.shape[0]) for s in scans]
    off = np.zeros(len(scans) + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    buf = torch.empty((int(off[-1]), 4), dtype=torch.float32)
    if pin and torch.cuda.is_available():
        buf = buf.pin_memory()
    for s, o in zip(scans, off[:-1]):
        buf[o:o + s.shape[0]] = torch.from_numpy(np.ascontiguousarray(s, np.float32).reshape(-1, 4))
    return buf, torch.from_numpy(off), max(sizes) if sizes else 0


def pack_pairs(scans1, scans2, pin=True):
    """frame pairs: both scans of a pair are cut to the shorter length (reference :378-384); returns the
    packing of [pair0.scan1, pair0.scan2, pair1.scan1, ...]"""
    seq = []
    for a, b in zip(scans1, scans2):
        n = min(a.shape[0], b.shape[0])
        seq += [a[:n], b[:n]]
    return pack_scans(seq, pin)


# crop specifications (ground_axis, ground_sign, ground_thr, near_axis_a, near_axis_b, near_thr), see include/pwclo_b200.h
KITTI_ODOMETRY_CROP = (1, +1, 1.1, 0, 2, 30.0)                 # kitti_odometry_dataset.py:151-159 (camera-style frame)


def kitti360_crop(near_treshold=30.0, velodyne_height=1.73, wheel_axis_height=0.3):
    """slam/dataset/kitti_360_dataset_2.py:113-123: velodyne frame (pass the identity as Tr), ground = z below the wheel
    axis, comparisons in float32 -- the thresholds are rounded to float32 so that the float64 comparisons of the kernel
    decide exactly as the reference's float32 ones do on float32 coordinates."""
    ground = float(np.float32(-(velodyne_height - wheel_axis_height)))
    return (2, -1, ground, 0, 1, float(np.float32(near_treshold)))


IDENTITY_TR = np.ascontiguousarray(np.eye(4)[:3], np.float64)


MAX_SCAN_POINTS = 1 << 20     # the select kernel's candidate slack (2048 beyond npoints) covers scans up to ~2M points


def prepare_scans(raw, offsets, Tr, npoints, seed, post=None, max_points=None, return_index=False, crop=KITTI_ODOMETRY_CROP,
                  check=False):
    """(`max_points` is accepted for compatibility and ignored.)
    raw float32 [total,4] CUDA, offsets int64 [S+1] CUDA, Tr float64 [3,4] (or [S,3,4]) CUDA ->
    clouds float32 [S,npoints,3] (+ int32 [S,npoints] source rows, int32 [S] survivors).

    Scans with fewer than `npoints` survivors are completed by draws with replacement, as the reference does
    (kitti_odometry_dataset.py:164-165).  A scan so large that the select kernel's candidate / boundary lists overflow
    (> ~2M points; KITTI scans have ~1.2e5) is reported as survivors[scan] = -1 and the rows it could not fill stay
    zero, never uninitialised memory.  `check=True` reads the flags back (one device->host sync) and raises instead."""
    for t, name, dt in ((raw, "raw", torch.float32), (offsets, "offsets", torch.int64), (Tr, "Tr", torch.float64)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dt and t.is_contiguous()):
            raise RuntimeError(f"{name} must be a contiguous {dt} CUDA tensor (there is no CPU path)")
    S = offsets.numel() - 1
    per_scan = Tr.dim() == 3
    if Tr.numel() != (12 * S if per_scan else 12):
        raise RuntimeError("Tr must be [3,4] or [S,3,4]")
    if post is not None and not (post.is_cuda and post.dtype == torch.float64 and post.is_contiguous() and post.numel() == 12 * S):
        raise RuntimeError("post must be a contiguous float64 CUDA tensor [S,3,4]")
    total = int(raw.shape[0])
    L = _lib.lib()
    ws = torch.empty(max(16, L.pwclo_prepare_scans_workspace_bytes(total, S)), dtype=torch.uint8, device=raw.device)
    if S == 1 and total > 2 * MAX_SCAN_POINTS:
        raise RuntimeError(f"a scan of {total} points is outside what the select kernel supports ({2 * MAX_SCAN_POINTS})")
    out = torch.zeros((S, npoints, 3), dtype=torch.float32, device=raw.device)
    idx = torch.zeros((S, npoints), dtype=torch.int32, device=raw.device)
    surv = torch.empty((S,), dtype=torch.int32, device=raw.device)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(raw.device):
        _lib.check(L.pwclo_prepare_scans_crop(p(raw), p(offsets), S, total, p(Tr), 1 if per_scan else 0, p(post),
                                              ctypes.c_ulonglong(int(seed) & (2 ** 64 - 1)), int(npoints), int(crop[0]),
                                              int(crop[1]), float(crop[2]), int(crop[3]), int(crop[4]), float(crop[5]),
                                              p(out), p(idx), p(surv), p(ws), ws.numel(), _lib.stream_ptr()), "prepare_scans")
    if check:
        bad = torch.nonzero(surv < 0).flatten().tolist()
        if bad:
            raise RuntimeError(f"prepare_scans: scans {bad[:8]} overflow the selection lists of the kernel (too many points)")
    return (out, idx, surv) if return_index else out


def load_kitti_bin(path):
    """a KITTI / KITTI-360 velodyne scan file -> float32 [n,4] (x, y, z, reflectance), exactly
    `np.fromfile(path, dtype=np.float32).reshape(-1, 4)` (kitti_odometry_dataset.py:375-376)"""
    a = np.fromfile(path, dtype=np.float32)
    if a.size % 4 != 0:
        raise RuntimeError(f"{path}: {a.size} floats is not a whole number of (x, y, z, reflectance) records")
    return a.reshape(-1, 4)


def prepare_pairs_from_files(files_current, files_previous, Tr, npoints, seed, device="cuda:0", crop=KITTI_ODOMETRY_CROP, post=None):
    """frame pairs from scan files to network inputs: read, cut each pair to its shorter scan, pack into one pinned
    buffer, one H2D copy, `prepare_scans`.  Returns (xyz_f1 [P,npoints,3], xyz_f2 [P,npoints,3]) on `device` with
    frame 1 = current and frame 2 = previous scan, the order the KITTI dataset feeds PWCLO-Net
    (kitti_odometry_dataset.py:330-343, 462-463).  Tr: [3,4] / [4,4] array (identity rows for KITTI-360)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("prepare_pairs_from_files needs a CUDA device (there is no CPU path)")
    cur = [load_kitti_bin(p) for p in files_current]
    prev = [load_kitti_bin(p) for p in files_previous]
    buf, off, _ = pack_pairs(cur, prev)
    biggest = int((off[1:] - off[:-1]).max()) if off.numel() > 1 else 0
    if biggest > 2 * MAX_SCAN_POINTS:       # offsets are still on the host here: refuse before any kernel runs
        raise RuntimeError(f"a scan of {biggest} points is outside what the select kernel supports ({2 * MAX_SCAN_POINTS})")
    Tr = np.ascontiguousarray(np.asarray(Tr, np.float64)[:3, :4])
    clouds = prepare_scans(buf.to(dev, non_blocking=True), off.to(dev), torch.from_numpy(Tr).to(dev), npoints, seed,
                           post=post, crop=crop)
    return clouds[0::2].contiguous(), clouds[1::2].contiguous()
