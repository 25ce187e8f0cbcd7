"""reference PW/pose_warp_refinement.py:25-158 (training / autograd composition)"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import PWCLO_utils as pwclo
from .. import pytorch_utils as pt_utils
from .costvolume import CostVolume
from .flowpredictor import FlowPredictor
from .pose_calculator import PoseCalculator
from ..pointnet2_modules import PointnetFPModulePWCLONet


class PoseWarpRefinement(nn.Module):
    def __init__(self, in_channel_f1, in_channel_f2, in_channel_f1_prev, in_channel_mask, knn=False, radius=0.0,
                 last_pose_estimation=False, pose=None, device="cpu", scalar_last=True):
        super().__init__()
        if (not knn) and radius == 0.0:
            raise RuntimeError("PoseWarpRefinement: when `knn` is set to False, `radius` should be precised.")
        self.pose, self.device, self.scalar_last = pose, device, scalar_last
        self.last_pose_estimation = last_pose_estimation
        self.in_channel = [in_channel_f1, in_channel_f2, in_channel_f1_prev]
        self.setupconv_features = PointnetFPModulePWCLONet(nsample=8, mlp=[in_channel_f1_prev, 128, 64],
                                                           post_mlp=[64 + in_channel_f1, 64], radius=radius * 0.2,
                                                           knn=True, use_xyz=True, bn=True)
        self.setupconv_mask = PointnetFPModulePWCLONet(nsample=8, mlp=[in_channel_mask, 128, 64],
                                                       post_mlp=[64 + in_channel_f1, 64], radius=radius * 0.2,
                                                       knn=True, use_xyz=True, bn=True)
        self.cost_volume = CostVolume(nsample=4, nsample_q=6, in_channel1=in_channel_f1, in_channel2=in_channel_f2,
                                      mlp1=[128, 64, 64], mlp2=[128, 64])
        self.flow_predictor_features = FlowPredictor(in_channel=in_channel_f1 + 64 + 64, mlp=[128, 64])
        if not last_pose_estimation:
            self.flow_predictor_mask = FlowPredictor(in_channel=in_channel_f1 + 64 + 64, mlp=[128, 64])
        self.pose_calculator = PoseCalculator(in_channel=64, out_channel=256, kernel_size=1, padding="valid",
                                              activation=None, pose=pose, squeeze=False)
        self.out_channel = [4, 3, 64]

    def forward(self, xyz_f1, points_f1, xyz_f2, points_f2, xyz_f1_prev, points_f1_prev, embedding_mask_prev,
                q_prev, t_prev):
        B = xyz_f1.size(0)
        q_coarse = torch.reshape(q_prev, [B, 4, 1])
        t_coarse = torch.reshape(t_prev, [B, 3, 1])
        xyz_f1_t = xyz_f1.permute(0, 2, 1).contiguous()
        xyz_prev_t = xyz_f1_prev.permute(0, 2, 1).contiguous()
        # the mask set-upconv is independent of the feature set-upconv, the warp and the cost volume: in training on the GPU
        # it runs on a second stream (forward and, through autograd, backward) and joins before the mask predictor
        branch = pt_utils.on_branch(pt_utils.branch_stream(xyz_f1.device) if self.training else None)
        with branch:
            coarse_m = self.setupconv_mask(xyz_f1_t, xyz_prev_t, points_f1, embedding_mask_prev)
        coarse_f = self.setupconv_features(xyz_f1_t, xyz_prev_t, points_f1, points_f1_prev)
        warped = pwclo.warp(xyz_f1, q_coarse, t_coarse)
        residual = self.cost_volume(warped, points_f1, xyz_f2, points_f2)
        emb = self.flow_predictor_features(points_f1, residual, coarse_f)
        branch.join(coarse_m)
        mask = coarse_m if self.last_pose_estimation else self.flow_predictor_mask(coarse_m, emb, points_f1)
        q_det, t_det = self.pose_calculator(emb, F.softmax(mask, dim=2))
        q = torch.squeeze(pwclo.mul_point_q(q_det, q_coarse), dim=2)
        t = torch.squeeze(pwclo.warp(t_coarse, q_det, t_det), dim=2)
        return q, t, emb, mask
