"""oracle/cpu_ops.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/pointnet2_cpu.c (the plain-C CPU restatement of the reference's
pointnet2 operators + knn_point).  Inputs/outputs are numpy arrays (or CPU torch tensors through
the `torch_ext` adapter, which has the 9-function surface of the reference's `pointnet2_ops._ext`,
EXT/src/bindings.cpp:7-18).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pointnet2_cpu.c")
_SO = os.path.join(_HERE, "liboracle_pwclo.so")
_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off: no silent contraction, fmaf() only where the oracle asks for it."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ref_block_threads(work, cap=512):
    return lib().oracle_ref_block_threads(int(work), int(cap))


def fps(xyz, m, origin_skip=True, thread_cap=512):
    xyz, px = _f(xyz)
    B, n, _ = xyz.shape
    out = np.zeros((B, m), np.int32)
    lib().oracle_fps(px, B, n, int(m), int(bool(origin_skip)), int(thread_cap), _p(out))
    return out


def gather_points(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.zeros((B, C, M), np.float32)
    lib().oracle_gather_points(pp, pi, B, C, N, M, _p(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, M = grad_out.shape
    out = np.zeros((B, C, n), np.float32)
    lib().oracle_gather_points_grad(pg, pi, B, C, int(n), M, _p(out))
    return out


def group_points(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    B, C, N = points.shape
    _, S, K = idx.shape
    out = np.zeros((B, C, S, K), np.float32)
    lib().oracle_group_points(pp, pi, B, C, N, S, K, _p(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, S, K = grad_out.shape
    out = np.zeros((B, C, n), np.float32)
    lib().oracle_group_points_grad(pg, pi, B, C, int(n), S, K, _p(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, pq = _f(new_xyz)
    xyz, px = _f(xyz)
    B, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = np.zeros((B, m, nsample), np.int32)
    lib().oracle_ball_query(pq, px, B, n, m, ctypes.c_float(radius), int(nsample), _p(out))
    return out


def three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.zeros((B, n, 3), np.float32)
    idx = np.zeros((B, n, 3), np.int32)
    lib().oracle_three_nn(pu, pk, B, n, m, _p(d2), _p(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, pp = _f(points)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, c, m = points.shape
    n = idx.shape[1]
    out = np.zeros((B, c, n), np.float32)
    lib().oracle_three_interpolate(pp, pi, pw, B, c, m, n, _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, c, n = grad_out.shape
    out = np.zeros((B, c, m), np.float32)
    lib().oracle_three_interpolate_grad(pg, pi, pw, B, c, n, int(m), _p(out))
    return out


def knn(xyz, new_xyz, k, sum_order=0, return_dist=False):
    """P2/pytorch_utils.py:32-49 with (distance, index) ascending tie rule."""
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = np.zeros((B, S, k), np.int32)
    dist = np.zeros((B, S, k), np.float32) if return_dist else None
    lib().oracle_knn(px, pq, B, N, S, int(k), int(sum_order), _p(idx), _p(dist) if return_dist else None)
    return (idx, dist) if return_dist else idx


def knn_distances(xyz, new_xyz, sum_order=0):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.zeros((B, S, N), np.float32)
    lib().oracle_knn_distances(px, pq, B, N, S, int(sum_order), _p(out))
    return out


class _TorchExt:
    """CPU stand-in with the exact surface of the reference's `pointnet2_ops._ext`
    (EXT/src/bindings.cpp:7-18; argument orders from EXT/src/*.cpp), torch tensors in and out."""

    @staticmethod
    def _t(a):
        import torch
        return torch.from_numpy(a)

    def furthest_point_sampling(self, points, nsamples):
        return self._t(fps(points.detach().numpy(), nsamples))

    def gather_points(self, points, idx):
        return self._t(gather_points(points.detach().numpy(), idx.numpy()))

    def gather_points_grad(self, grad_out, idx, n):
        return self._t(gather_points_grad(grad_out.detach().numpy(), idx.numpy(), n))

    def group_points(self, points, idx):
        return self._t(group_points(points.detach().numpy(), idx.numpy()))

    def group_points_grad(self, grad_out, idx, n):
        return self._t(group_points_grad(grad_out.detach().numpy(), idx.numpy(), n))

    def ball_query(self, new_xyz, xyz, radius, nsample):
        return self._t(ball_query(new_xyz.detach().numpy(), xyz.detach().numpy(), radius, nsample))

    def three_nn(self, unknown, known):
        d2, idx = three_nn(unknown.detach().numpy(), known.detach().numpy())
        return [self._t(d2), self._t(idx)]

    def three_interpolate(self, points, idx, weight):
        return self._t(three_interpolate(points.detach().numpy(), idx.numpy(), weight.detach().numpy()))

    def three_interpolate_grad(self, grad_out, idx, weight, m):
        return self._t(three_interpolate_grad(grad_out.detach().numpy(), idx.numpy(), weight.detach().numpy(), m))


torch_ext = _TorchExt()
