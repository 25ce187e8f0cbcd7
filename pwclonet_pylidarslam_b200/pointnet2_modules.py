"""Set-conv / set-upconv modules with the reference's names, constructor arguments, parameter
layout and forward signatures (P2/pointnet2_modules.py:159-245 and :410-515).

Two execution paths, one set of parameters:
  * training / autograd path (`self.training` or grad enabled inputs): the op-by-op composition of
    the reference, every sampling / neighbour / grouping op backed by the sm_100a kernels
    (pointnet2_utils, pytorch_utils.knn_point), convolutions by torch;
  * inference path: `pwclonet_pylidarslam_b200.fused` runs the whole layer as fused kernels with
    BatchNorm folded (used by PWCLONet.forward in eval mode).
"""
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from . import pytorch_utils as pt_utils


class PointnetSAModulePWCLONet(nn.Module):
    def __init__(self, mlp: List[int], npoint: int, nsample: int, bn: bool = True):
        super().__init__()
        self.npoint, self.nsample = npoint, nsample
        spec = list(mlp)
        if spec[0] == 0:          # no input features: the grouped absolute xyz takes their place
            spec[0] += 3
        spec[0] += 3              # xyz_diff is always concatenated
        self.mlp_spec = spec
        self.mlp_module = pt_utils.SharedMLP(spec, bn=bn, init=torch.nn.init.xavier_uniform_)

    def geometry(self, xyz: torch.Tensor, tie_in=None, return_tie=False):
        """the non-differentiable, coordinates-only part of forward: (FPS indices [B,npoint], new_xyz [B,npoint,3],
        neighbour indices [B,npoint,nsample]).  A caller that processes several clouds with this module (the two
        frames of a pair) can compute it for all of them in one batch and hand slices to forward(geom=...).
        tie_in / return_tie: the per-cloud tie flags of `_ext.furthest_point_sampling` -- a pyramid passes the flag of
        the level that produced `xyz` so that tie-free clouds skip the sampling rounds (same indices)."""
        if tie_in is None and not return_tie:
            fidx, tie = pointnet2_utils.furthest_point_sample(xyz, self.npoint), None
        else:
            from . import _ext
            fidx, tie = _ext.furthest_point_sampling(xyz.contiguous(), self.npoint, tie_in=tie_in, return_tie=True)
        new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
        _, idx = pt_utils.knn_point(self.nsample, xyz, new_xyz)
        return (fidx, new_xyz, idx, tie) if return_tie else (fidx, new_xyz, idx)

    def forward(self, xyz: torch.Tensor, features: Optional[torch.Tensor], geom=None):
        """xyz (B,N,3), features (B,C,N) or None -> new_xyz (B,npoint,3), new_features (B,C',npoint)"""
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        fidx, new_xyz, idx = self.geometry(xyz) if geom is None else geom
        grouped_xyz = pointnet2_utils.grouping_operation(xyz_flipped, idx)
        xyz_diff = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            x = pt_utils.cat_for(self.mlp_module, (xyz_diff, pointnet2_utils.grouping_operation(features, idx)))
        else:
            x = pt_utils.cat_for(self.mlp_module, (xyz_diff, grouped_xyz))
        x = self.mlp_module(x)
        return new_xyz, pt_utils.max_over_neighbours(x)


class PointnetFPModulePWCLONet(nn.Module):
    def __init__(self, *, mlp: List[int], radius: float, nsample: int, post_mlp: List[int], bn: bool = True,
                 use_xyz: bool = True, knn: bool = False, sample_uniformly: bool = False):
        super().__init__()
        self.nsample, self.knn, self.use_xyz, self.radius = nsample, knn, use_xyz, radius
        spec = list(mlp)
        if use_xyz:
            spec[0] += 3
        self.mlp = pt_utils.SharedMLP(spec, bn=bn, init=torch.nn.init.xavier_uniform_)
        self.post_mlp = pt_utils.SharedMLP(list(post_mlp), bn=bn, init=torch.nn.init.xavier_uniform_)
        self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz)

    def forward(self, xyz2, xyz1, features2, features1):
        """xyz2 (B,N2,3) dense, xyz1 (B,N1,3) coarse, features2 (B,C2,N2), features1 (B,C1,N1) -> (B,C',N2)"""
        if self.knn:
            _, idx = pt_utils.knn_point(self.nsample, xyz1, xyz2)
            x = pointnet2_utils.grouping_operation(features1, idx)
            g_xyz = pointnet2_utils.grouping_operation(xyz1.transpose(1, 2).contiguous(), idx)
            xyz_diff = g_xyz - xyz2.transpose(1, 2).unsqueeze(-1)
            if self.use_xyz:
                x = pt_utils.cat_for(self.mlp, (x, xyz_diff))
        else:
            x = self.grouper(xyz1, xyz2, features1)
        x = self.mlp(x)
        x = pt_utils.max_over_neighbours(x)
        if features2 is not None:
            x = torch.cat([x, features2], dim=1)
        return self.post_mlp(x.unsqueeze(-1)).squeeze(-1)
