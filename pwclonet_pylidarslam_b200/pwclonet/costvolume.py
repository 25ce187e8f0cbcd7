"""Attentive point-to-patch cost volume, reference PW/costvolume.py:19-190 (same constructor,
parameter names and forward signature).  This is the autograd / training composition on top of the
sm_100a kNN + grouping kernels; inference goes through pwclonet_pylidarslam_b200.fused."""
import os

import torch
import torch.nn as nn

from .. import pointnet2_utils as pointutils
from .. import pytorch_utils as pt_utils


class CostGeometry(torch.autograd.Function):
    """(center [B,3,S], grouped [B,3,S,K]) -> [B,10,S,K] = (p, q, q - p, |q - p|) on the sm_100a kernels
    (`pwclo_cost_geometry_fwd/_bwd`), one launch per direction; PWCLO_COST_GEO=0 keeps the torch composition"""

    @staticmethod
    def forward(ctx, center, grouped):
        import ctypes
        from .. import _lib
        center, grouped = center.contiguous(), grouped.contiguous()
        B, _, S, K = grouped.shape
        out = torch.empty((B, 10, S, K), dtype=grouped.dtype, device=grouped.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(grouped.device):
            _lib.check(_lib.lib().pwclo_cost_geometry_fwd(p(center), p(grouped), B, S, K, p(out), _lib.stream_ptr()),
                       "cost_geometry_fwd")
        ctx.save_for_backward(center, grouped)
        return out

    @staticmethod
    def backward(ctx, go):
        import ctypes
        from .. import _lib
        center, grouped = ctx.saved_tensors
        B, _, S, K = grouped.shape
        go = go.contiguous()
        gc = torch.empty_like(center) if ctx.needs_input_grad[0] else None
        gg = torch.empty_like(grouped) if ctx.needs_input_grad[1] else None
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        with torch.cuda.device(grouped.device):
            _lib.check(_lib.lib().pwclo_cost_geometry_bwd(p(center), p(grouped), p(go), B, S, K, p(gc), p(gg), _lib.stream_ptr()),
                       "cost_geometry_bwd")
        return gc, gg


class CostVolume(nn.Module):
    def __init__(self, nsample, nsample_q, in_channel1, in_channel2, mlp1, mlp2):
        super().__init__()
        self.nsample, self.nsample_q = nsample, nsample_q
        self.in_channel = [in_channel1, in_channel2, 10]
        xu = torch.nn.init.xavier_uniform_
        self.mlp_convs = pt_utils.SharedMLP([in_channel1 + in_channel2 + 10] + list(mlp1), bn=True, init=xu)
        self.mlp_conv_xyz_1 = pt_utils.SharedMLP([10, mlp1[-1]], bn=True, init=xu)
        self.mlp_conv_xyz_2 = pt_utils.SharedMLP([10, mlp1[-1]], bn=True, init=xu)
        self.mlp2_convs = pt_utils.SharedMLP([mlp1[-1] * 2] + list(mlp2), bn=True, init=xu)
        self.mlp3_convs = pt_utils.SharedMLP([mlp1[-1] * 2 + in_channel1] + list(mlp2), bn=True, init=xu)
        self.out_channel = mlp2[-1]

    @staticmethod
    def _geometry(center, grouped):
        """10 channels (p, q, q-p, |q-p|) of costvolume.py:94-105 / :159-169"""
        if (grouped.is_cuda and grouped.dtype == torch.float32 and torch.is_grad_enabled()
                and os.environ.get("PWCLO_COST_GEO", "1") != "0"):
            return CostGeometry.apply(center, grouped)
        k = grouped.size(3)
        p = center.unsqueeze(3).expand(-1, -1, -1, k)
        d = grouped - p
        euc = torch.sqrt(torch.sum(torch.square(d), dim=1, keepdim=True) + 1e-20)
        return torch.cat((p, grouped, d, euc), dim=1)

    def forward(self, warped_xyz, warped_points, f2_xyz, f2_points):
        """warped_xyz (B,3,S), warped_points (B,C,S), f2_xyz (B,3,N), f2_points (B,C,N) -> (B,mlp[-1],S)"""
        w_t = warped_xyz.permute(0, 2, 1).contiguous()
        f2_t = f2_xyz.permute(0, 2, 1).contiguous()
        # stage 2's neighbour search and geometry encoding need the warped cloud only: in training they run on the branch
        # stream beside stage 1
        branch = pt_utils.on_branch(pt_utils.branch_stream(warped_xyz.device, 1) if self.training else None)
        with branch:
            _, idx = pt_utils.knn_point(self.nsample, w_t, w_t)
            geo2 = self._geometry(warped_xyz, pointutils.grouping_operation(warped_xyz.contiguous(), idx))
            enc2 = self.mlp_conv_xyz_2(geo2)
        _, idx_q = pt_utils.knn_point(self.nsample_q, f2_t, w_t)
        geo = self._geometry(warped_xyz, pointutils.grouping_operation(f2_xyz.contiguous(), idx_q))
        p_f = warped_points.unsqueeze(3).expand(-1, -1, -1, self.nsample_q)
        x = pt_utils.cat_for(self.mlp_convs, (geo, p_f, pointutils.grouping_operation(f2_points.contiguous(), idx_q)))
        x = self.mlp_convs(x)
        e1 = pt_utils.softmax_pool(self.mlp2_convs(torch.cat((self.mlp_conv_xyz_1(geo), x), dim=1)), x)
        branch.join(idx, enc2)
        c_e = pointutils.grouping_operation(e1.contiguous(), idx)
        n_f = warped_points.unsqueeze(3).expand(-1, -1, -1, self.nsample)
        return pt_utils.softmax_pool(self.mlp3_convs(torch.cat((enc2, n_f, c_e), dim=1)), c_e)
