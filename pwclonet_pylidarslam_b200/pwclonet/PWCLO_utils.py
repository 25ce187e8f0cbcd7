"""Quaternion helpers with the reference's names and term order (PW/PWCLO_utils.py:31-132).
Quaternions are scalar-first; like the reference, the `scalar_last` arguments are accepted and
ignored."""
import os

import torch


def switch_quat(q, scalar_last: bool = False):
    """numpy quaternion(s) [4] or [B,4]: scalar-first -> scalar-last (scalar_last=True) or back (PW/PWCLO_utils.py:5-28;
    not used by the network, kept for the module's surface)"""
    import numpy as np
    q = np.asarray(q)
    if q.ndim not in (1, 2):
        raise RuntimeError(f"[switch_quat] Unrecognized shape of quaternions: {q.shape}")
    return np.roll(q, -1 if scalar_last else 1, axis=-1)


def inv_q(q, device=None, scalar_last: bool = False):
    """q^-1 = conj(q) / (|q|^2 + 1e-10); q [B,4]"""
    q_2 = torch.sum(q * q, dim=-1, keepdim=True) + 1e-10
    # conj(q) built on the device (a host-side constant would be a pageable H2D copy: not CUDA-graph capturable);
    # q * 1 and q * -1 are exact, so this equals the reference's q * [1,-1,-1,-1]
    conj = torch.cat((q[..., :1], -q[..., 1:]), dim=-1)
    return conj / q_2


_HAMILTON_CONST = {}
HAMILTON_FEW = int(os.environ.get("PWCLO_HAMILTON_FEW", "16"))


def _hamilton_const(device):
    """(PERM, SIGN) [4,4]: term k of output component i is SIGN[k][i] * a_k * b_PERM[k][i], PERM[k][i] = i xor k"""
    c = _HAMILTON_CONST.get(device)
    if c is None:
        if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            return None                       # a host constant cannot be uploaded inside a capture: one op per term
        ar = torch.arange(4, device=device)
        sign = torch.tensor([[1, 1, 1, 1], [-1, 1, -1, 1], [-1, 1, 1, -1], [-1, -1, 1, 1]], dtype=torch.float32, device=device)
        c = _HAMILTON_CONST[device] = (torch.bitwise_xor(ar[:, None], ar[None, :]), sign)
    return c


def _hamilton(a, b):
    """a (x) b over the last axis, the reference's term order (PW/PWCLO_utils.py:62-100: every component is
    ((a0 b? +- a1 b?) +- a2 b?) +- a3 b?).  For a handful of quaternions (the pose composition of a refinement level: 8)
    the 16 products are one gather + two multiplies + three adds instead of 28 one-element launches (and ~170 in the
    backward): the same roundings in the same order, since a sign flip is exact."""
    if a.shape[:-1].numel() * b.shape[:-1].numel() <= HAMILTON_FEW * HAMILTON_FEW and a.dtype == torch.float32:
        c = _hamilton_const(a.device)
        if c is not None:
            perm, sign = c
            p = a.unsqueeze(-1) * (b[..., perm] * sign)            # [..., k, i]
            return ((p[..., 0, :] + p[..., 1, :]) + p[..., 2, :]) + p[..., 3, :]
    a0, a1, a2, a3 = a.unbind(-1)
    b0, b1, b2, b3 = b.unbind(-1)
    return torch.stack((a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
                        a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
                        a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0), dim=-1)


def mul_q_point(q, points, scalar_last: bool = False):
    """q [B,4,1] or [B,4], points [B,4,N] -> q (x) points [B,4,N]"""
    B = points.size(0)
    return _hamilton(q.reshape(B, 1, 4), points.permute(0, 2, 1)).permute(0, 2, 1).contiguous()


def mul_point_q(points, q, scalar_last: bool = False):
    """points [B,4,N], q [B,4,1] or [B,4] -> points (x) q [B,4,N]"""
    B = points.size(0)
    return _hamilton(points.permute(0, 2, 1), q.reshape(B, 1, 4)).permute(0, 2, 1).contiguous()


class FusedWarp(torch.autograd.Function):
    """warp() as one sm_100a launch forward and one backward (pwclo_warp_fwd / pwclo_warp_bwd)"""

    @staticmethod
    def forward(ctx, xyz, q, t):
        import ctypes
        from .. import _lib
        B, _, N = xyz.shape
        xyz, q4, t3 = xyz.contiguous(), q.reshape(B, 4).contiguous(), t.reshape(B, 3).contiguous()
        out = torch.empty_like(xyz)
        p = lambda a: ctypes.c_void_p(a.data_ptr())
        with torch.cuda.device(xyz.device):
            _lib.check(_lib.lib().pwclo_warp_fwd(p(xyz), p(q4), p(t3), B, N, p(out), _lib.stream_ptr()), "warp_fwd")
        ctx.save_for_backward(xyz, q4)
        ctx.shapes = (q.shape, t.shape, ctx.needs_input_grad[0])
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes
        from .. import _lib
        xyz, q4 = ctx.saved_tensors
        qs, ts, need_x = ctx.shapes
        B, _, N = xyz.shape
        g = g.contiguous()
        dq = torch.empty(B, 4, dtype=xyz.dtype, device=xyz.device)
        dt = torch.empty(B, 3, dtype=xyz.dtype, device=xyz.device)
        dx = torch.empty_like(xyz) if need_x else None
        p = lambda a: ctypes.c_void_p(a.data_ptr()) if a is not None else None
        with torch.cuda.device(xyz.device):
            _lib.check(_lib.lib().pwclo_warp_bwd(p(xyz), p(q4), p(g), B, N, p(dx), p(dq), p(dt), _lib.stream_ptr()), "warp_bwd")
        return dx, dq.reshape(qs), dt.reshape(ts)


FUSED_WARP = os.environ.get("PWCLO_FUSED_WARP", "1") != "0"


def warp(xyz, q, t, device=None, scalar_last: bool = False):
    """xyz [B,3,N], q [B,4,1], t [B,3,1] -> q (x) [0,xyz] (x) q^-1 + t   [B,3,N]"""
    if (FUSED_WARP and xyz.is_cuda and xyz.dtype == torch.float32 and torch.is_grad_enabled()
            and (q.requires_grad or t.requires_grad or xyz.requires_grad)):
        return FusedWarp.apply(xyz, q, t)
    return warp_composed(xyz, q, t)


def warp_composed(xyz, q, t):
    """the reference's expression tree, one torch op per term (used without autograd and as the parity reference)"""
    B, _, N = xyz.size()
    q_inv = inv_q(torch.squeeze(q, dim=2))
    xyz_ = torch.cat((torch.zeros([B, 1, N], device=xyz.device, dtype=xyz.dtype), xyz), dim=1)
    out = mul_point_q(mul_q_point(q, xyz_), q_inv)
    return out[:, 1:, :] + t
