"""Drop-in for the reference's `pointnet2_ops._ext` extension module.

Same nine functions, argument orders, dtypes, shapes and allocation behaviour as the pybind
module built from the reference's _ext-src (EXT/src/bindings.cpp:7-18, argument checks
EXT/include/utils.h:5-25): inputs must be contiguous float32 / int32 CUDA tensors, outputs are
freshly allocated by the callee, work is enqueued on the current torch CUDA stream without any
host synchronisation.  Differences, on purpose: errors raise RuntimeError instead of killing the
process (cuda_utils.h:30-39 calls exit(-1)), the tensor's device is made current for the launch,
and there is no `temp` buffer in FPS.

`register_as_pointnet2_ops_ext()` installs this module as `pointnet2_ops._ext` so that the
reference's own pointnet2_utils.py:7-8 picks it up instead of JIT-compiling its kernels.
"""
import ctypes
import sys

import torch

from . import _lib

_FPS_FLAGS = 1  # PWCLO_FPS_ORIGIN_SKIP: the built reference kernel skips |p|^2 <= 1e-3


def _chk(t, name, dtype):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (CPU not supported)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a{'n int' if dtype == torch.int32 else ' float'} tensor")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def furthest_point_sampling(points, nsamples, flags=_FPS_FLAGS, tie_in=None, return_tie=False):
    """The reference signature is (points, nsamples).  Extension (not in the reference): `return_tie=True` also returns
    an int32[B] flag per cloud, 1 when the run saw two candidates tie for an arg-max; handing that flag as `tie_in` to
    the sampling of the SELECTED points (the next pyramid level, P2/pointnet2_modules.py:200-206) lets tie-free clouds
    return 0..nsamples-1 without the dependent rounds -- bit-identical (include/pwclo_b200.h)."""
    _chk(points, "points", torch.float32)
    B, N, _ = points.shape
    out = torch.zeros((B, nsamples), dtype=torch.int32, device=points.device)
    if tie_in is None and not return_tie:
        with torch.cuda.device(points.device):
            _lib.check(_lib.lib().pwclo_furthest_point_sampling(_p(points), B, N, int(nsamples), int(flags), _p(out),
                                                                _lib.stream_ptr()), "furthest_point_sampling")
        return out
    if tie_in is not None:
        _chk(tie_in, "tie_in", torch.int32)
        if tie_in.numel() != B:
            raise RuntimeError("tie_in must hold one int32 flag per cloud")
    tie = torch.empty(B, dtype=torch.int32, device=points.device) if return_tie else None
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().pwclo_furthest_point_sampling_prefix(
            _p(points), B, N, int(nsamples), int(flags), _p(out), _p(tie_in) if tie_in is not None else None,
            _p(tie) if tie is not None else None, _lib.stream_ptr()), "furthest_point_sampling_prefix")
    return (out, tie) if return_tie else out


def gather_points(points, idx):
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().pwclo_gather_points(_p(points), _p(idx), B, C, N, M, _p(out), _lib.stream_ptr()),
                   "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, M = grad_out.shape
    out = torch.zeros((B, C, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.lib().pwclo_gather_points_grad(_p(grad_out), _p(idx), B, C, int(n), M, _p(out),
                                                       _lib.stream_ptr()), "gather_points_grad")
    return out


def group_points(points, idx):
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, N = points.shape
    _, S, K = idx.shape
    out = torch.empty((B, C, S, K), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().pwclo_group_points(_p(points), _p(idx), B, C, N, S, K, _p(out), _lib.stream_ptr()),
                   "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    B, C, S, K = grad_out.shape
    out = torch.zeros((B, C, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.lib().pwclo_group_points_grad(_p(grad_out), _p(idx), B, C, int(n), S, K, _p(out),
                                                      _lib.stream_ptr()), "group_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    _chk(new_xyz, "new_xyz", torch.float32)
    _chk(xyz, "xyz", torch.float32)
    B, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = torch.empty((B, m, nsample), dtype=torch.int32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        _lib.check(_lib.lib().pwclo_ball_query(_p(new_xyz), _p(xyz), B, n, m, float(radius), int(nsample), _p(out),
                                               _lib.stream_ptr()), "ball_query")
    return out


def three_nn(unknown, known):
    _chk(unknown, "unknown", torch.float32)
    _chk(known, "known", torch.float32)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        _lib.check(_lib.lib().pwclo_three_nn(_p(unknown), _p(known), B, n, m, _p(dist2), _p(idx), _lib.stream_ptr()),
                   "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    _chk(points, "points", torch.float32)
    _chk(idx, "idx", torch.int32)
    _chk(weight, "weight", torch.float32)
    B, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, c, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().pwclo_three_interpolate(_p(points), _p(idx), _p(weight), B, c, m, n, _p(out),
                                                      _lib.stream_ptr()), "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    _chk(grad_out, "grad_out", torch.float32)
    _chk(idx, "idx", torch.int32)
    _chk(weight, "weight", torch.float32)
    B, c, n = grad_out.shape
    out = torch.zeros((B, c, m), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.lib().pwclo_three_interpolate_grad(_p(grad_out), _p(idx), _p(weight), B, c, n, int(m), _p(out),
                                                           _lib.stream_ptr()), "three_interpolate_grad")
    return out


# ---- not part of the reference extension: the kNN the reference does in PyTorch -----------------
KNN_SORTED = True          # exact sorted-slab search (pwclo_knn_sorted) instead of brute force
KNN_SORTED_MIN_N = 64      # below this many reference points brute force is used
KNN_SUM_ORDER = 1  # PWCLO_KNN_SUM_XZ_Y: torch's CUDA reduction order (see DESIGN.md "kNN formulation")


def knn(xyz, new_xyz, k, sum_order=None, warp_qt=None, return_warped=False, return_dist=False):
    _chk(xyz, "xyz", torch.float32)
    _chk(new_xyz, "new_xyz", torch.float32)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty((B, S, k), dtype=torch.int32, device=xyz.device)
    dist = torch.empty((B, S, k), dtype=torch.float32, device=xyz.device) if return_dist else None
    warped = None
    if warp_qt is not None:
        _chk(warp_qt, "warp_qt", torch.float32)
        if return_warped:
            warped = torch.empty_like(new_xyz)
    so = KNN_SUM_ORDER if sum_order is None else int(sum_order)
    with torch.cuda.device(xyz.device):
        L = _lib.lib()
        wq = _p(warp_qt) if warp_qt is not None else None
        wo = _p(warped) if warped is not None else None
        dp = _p(dist) if dist is not None else None
        ws_bytes = L.pwclo_knn_workspace_bytes(B, N, S) if KNN_SORTED and N >= KNN_SORTED_MIN_N else 0
        if ws_bytes:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device)
            _lib.check(L.pwclo_knn_sorted(_p(xyz), _p(new_xyz), B, N, S, int(k), so, wq, wo, _p(idx), dp, _p(ws), ws_bytes,
                                          _lib.stream_ptr()), "knn_sorted")
        else:
            _lib.check(L.pwclo_knn(_p(xyz), _p(new_xyz), B, N, S, int(k), so, wq, wo, _p(idx), dp, _lib.stream_ptr()), "knn")
    out = [idx]
    if return_dist:
        out.append(dist)
    if return_warped:
        out.append(warped)
    return out[0] if len(out) == 1 else tuple(out)


def register_as_pointnet2_ops_ext():
    """Make `import pointnet2_ops._ext` (reference P2/pointnet2_utils.py:7-8) resolve to this module."""
    mod = sys.modules[__name__]
    sys.modules["pointnet2_ops._ext"] = mod
    pkg = sys.modules.get("pointnet2_ops")
    if pkg is not None:
        pkg._ext = mod
    return mod
