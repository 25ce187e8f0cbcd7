"""Quaternion helpers with the reference's names and term order (PW/PWCLO_utils.py:31-132).
Quaternions are scalar-first; like the reference, the `scalar_last` arguments are accepted and
ignored."""
import torch


def inv_q(q, device=None, scalar_last: bool = False):
    """q^-1 = conj(q) / (|q|^2 + 1e-10); q [B,4]"""
    q_2 = torch.sum(q * q, dim=-1, keepdim=True) + 1e-10
    # conj(q) built on the device (a host-side constant would be a pageable H2D copy: not CUDA-graph capturable);
    # q * 1 and q * -1 are exact, so this equals the reference's q * [1,-1,-1,-1]
    conj = torch.cat((q[..., :1], -q[..., 1:]), dim=-1)
    return conj / q_2


def _hamilton(a, b):
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


def warp(xyz, q, t, device=None, scalar_last: bool = False):
    """xyz [B,3,N], q [B,4,1], t [B,3,1] -> q (x) [0,xyz] (x) q^-1 + t   [B,3,N]"""
    B, _, N = xyz.size()
    q_inv = inv_q(torch.squeeze(q, dim=2))
    xyz_ = torch.cat((torch.zeros([B, 1, N], device=xyz.device, dtype=xyz.dtype), xyz), dim=1)
    out = mul_point_q(mul_q_point(q, xyz_), q_inv)
    return out[:, 1:, :] + t
