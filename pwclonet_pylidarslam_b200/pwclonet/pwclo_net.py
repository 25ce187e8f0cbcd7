"""PWCLONet with the reference's constructor, attributes, parameter layout (510 state-dict
tensors) and forward contract (PW/pwclo_net.py:32-218):

    pose_params[B,4,7], log_dict = net(xyz_f1[B,3,N], None, xyz_f2[B,3,N], None)

Execution:
  * `net.eval()` + no grad  -> the fused sm_100a inference engine (`..fused.FusedPWCLONet`): BN folded,
    one kernel per layer, no grouped tensor ever materialised, no forced device->host sync (the
    reference's log_dict lives on the CPU, pwclo_net.py:186-193; ours is produced lazily);
  * otherwise                -> the autograd composition below (same op order as the reference),
    all sampling / neighbour / grouping ops on the sm_100a kernels.
"""
import warnings
from enum import Enum

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..pointnet2_modules import PointnetSAModulePWCLONet
from .costvolume import CostVolume
from .flowpredictor import FlowPredictor
from .pose_calculator import PoseCalculator
from .pose_warp_refinement import PoseWarpRefinement


class _Pose:
    """Stand-in for slam.common.pose.Pose("quaternions") (only num_rot_params() is used on this path)."""

    def __init__(self, pose_type="quaternions"):
        self.pose_type = pose_type

    def num_rot_params(self):
        return 4 if self.pose_type == "quaternions" else 3


class _Config(dict):
    __getattr__ = dict.get


class LazyLog(dict):
    """log_dict whose CPU tensors are only materialised when read (keeps the forward asynchronous)."""

    def __init__(self, thunks):
        super().__init__()
        self._thunks = dict(thunks)

    def __missing__(self, key):
        v = self._thunks[key]()
        self[key] = v
        return v

    def keys(self):
        return self._thunks.keys()

    def __contains__(self, key):
        return key in self._thunks


class PWCLONet(nn.Module):
    def __init__(self, config=None, pose=None):
        super().__init__()
        config = _Config(config or {})
        config.setdefault("num_input_channels", 3)
        config.setdefault("sequence_len", 2)
        config.setdefault("device", "cpu")
        config.setdefault("scalar_last", False)
        self.config = config
        self.pose = pose if pose is not None else _Pose("quaternions")
        self.num_out_poses = config.get("num_out_poses", 1)
        self.num_input_channels = config["num_input_channels"]
        self.sequence_len = config["sequence_len"]
        self.device = torch.device(config["device"])
        if self.num_out_poses != 1:
            warnings.warn("current version of PWCLONet allows predicting only one pose")
        self.nb_levels = config.get("num_out_poses", 4)
        self.use_fused = config.get("use_fused", True)
        self.lazy_log = config.get("lazy_log", True)

        self.psa_1 = PointnetSAModulePWCLONet(npoint=2048, nsample=32, mlp=[0, 8, 8, 16], bn=True)
        self.psa_2 = PointnetSAModulePWCLONet(npoint=1024, nsample=32, mlp=[16, 16, 16, 32], bn=True)
        self.psa_3 = PointnetSAModulePWCLONet(npoint=256, nsample=16, mlp=[32, 32, 32, 64], bn=True)
        self.psa_4 = PointnetSAModulePWCLONet(npoint=64, nsample=16, mlp=[64, 64, 64, 128], bn=True)
        self.cost_volume = CostVolume(nsample=4, nsample_q=32, in_channel1=64, in_channel2=64, mlp1=[128, 64, 64],
                                      mlp2=[128, 64])
        self.flow_feature_encoding = PointnetSAModulePWCLONet(npoint=64, nsample=16, mlp=[64, 128, 64, 64], bn=True)
        self.l4_flow_predictor = FlowPredictor(in_channel=128 + 64, mlp=[128, 64])
        self.pose_calculator_4 = PoseCalculator(in_channel=64, out_channel=256, kernel_size=1, padding="valid",
                                                activation=None, squeeze=True)
        dev, sl = config["device"], config["scalar_last"]
        self.pose_warp_refinement_3 = PoseWarpRefinement(64, 64, 64, 64, radius=2.0, last_pose_estimation=False,
                                                         device=dev, scalar_last=sl)
        self.pose_warp_refinement_2 = PoseWarpRefinement(32, 32, 64, 64, radius=1.0, last_pose_estimation=False,
                                                         device=dev, scalar_last=sl)
        self.pose_warp_refinement_1 = PoseWarpRefinement(16, 16, 64, 64, radius=0.5, last_pose_estimation=True,
                                                         device=dev, scalar_last=sl)
        self._fused = None

    # -- fused-engine plumbing ------------------------------------------------------------------
    def train(self, mode=True):
        self._fused = None          # parameters / BN statistics may change: re-fold on next eval forward
        return super().train(mode)

    def load_state_dict(self, *a, **kw):
        self._fused = None
        return super().load_state_dict(*a, **kw)

    def fused_engine(self):
        if self._fused is None:
            from ..fused import FusedPWCLONet
            self._fused = FusedPWCLONet(self)
        return self._fused

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, xyz_f1, points_f1, xyz_f2, points_f2, bn_decay=None):
        fused_ok = (self.use_fused and not self.training and not torch.is_grad_enabled()
                    and points_f1 is None and points_f2 is None and xyz_f1.is_cuda)
        if fused_ok:
            pose, mask1, xyz1_l1 = self.fused_engine().forward(xyz_f1, xyz_f2)
        else:
            pose, mask1, xyz1_l1 = self._forward_composed(xyz_f1, points_f1, xyz_f2, points_f2)
        thunks = {
            "embedding_mask": lambda: torch.linalg.norm(
                F.softmax(mask1.detach().cpu(), dim=2).permute(0, 2, 1), dim=-1, ord=2),
            "point_cloud": lambda: xyz1_l1.detach().cpu(),
        }
        log = LazyLog(thunks)
        if not self.lazy_log:
            log = {k: log[k] for k in thunks}
        return pose, log

    def _forward_composed(self, xyz_f1, points_f1, xyz_f2, points_f2):
        xyz_f1_t = xyz_f1.permute(0, 2, 1).contiguous()
        xyz_f2_t = xyz_f2.permute(0, 2, 1).contiguous()
        x1, f1, x2, f2 = [xyz_f1_t], [points_f1], [xyz_f2_t], [points_f2]
        # FPS / gather / kNN depend on coordinates only and are not differentiable: one batch of 2B clouds per level
        # for both frames (per cloud the same indices as the reference's per-frame calls; with 8 pairs per GPU the
        # 16384 -> 2048 sampling is a 2047-round chain on 8 CTAs, twice, otherwise).  The layers keep their per-frame
        # calls: train-mode BatchNorm statistics must be those of one frame's batch, as in the reference.
        Bp = xyz_f1_t.shape[0]
        geoms, cur = [], torch.cat((xyz_f1_t, xyz_f2_t), dim=0).detach()
        for psa in (self.psa_1, self.psa_2, self.psa_3, self.psa_4):
            g = psa.geometry(cur)
            geoms.append(g)
            cur = g[1]
        for psa, g in zip((self.psa_1, self.psa_2, self.psa_3, self.psa_4), geoms):
            a, b = psa(x1[-1], f1[-1], geom=tuple(t[:Bp] for t in g))
            x1.append(a), f1.append(b)
        for psa, g in zip((self.psa_1, self.psa_2, self.psa_3, self.psa_4), geoms):
            a, b = psa(x2[-1], f2[-1], geom=tuple(t[Bp:] for t in g))
            x2.append(a), f2.append(b)
        X1 = [None] + [x.permute(0, 2, 1).contiguous() for x in x1[1:]]   # [B,3,S] per level 1..4
        X2 = [None] + [x.permute(0, 2, 1).contiguous() for x in x2[1:]]

        flow_embedding = self.cost_volume(X1[3], f1[3], X2[3], f2[3])
        # flow_feature_encoding samples and groups xyz1 of level 3 exactly as psa_4 did for frame 1 (same npoint / nsample)
        ffe = self.flow_feature_encoding
        same = ffe.npoint == self.psa_4.npoint and ffe.nsample == self.psa_4.nsample
        xyz_f1_4_t, emb_4 = ffe(x1[3], flow_embedding, geom=tuple(t[:Bp] for t in geoms[3]) if same else None)
        new_xyz_f1_4 = xyz_f1_4_t.permute(0, 2, 1).contiguous()
        mask_4 = self.l4_flow_predictor(f1[4], emb_4)
        q_4, t_4 = self.pose_calculator_4(emb_4, F.softmax(mask_4, dim=2))

        q_3, t_3, emb_3, mask_3 = self.pose_warp_refinement_3(X1[3], f1[3], X2[3], f2[3], new_xyz_f1_4, emb_4, mask_4,
                                                              q_4, t_4)
        q_2, t_2, emb_2, mask_2 = self.pose_warp_refinement_2(X1[2], f1[2], X2[2], f2[2], X1[3], emb_3, mask_3, q_3,
                                                              t_3)
        q_1, t_1, emb_1, mask_1 = self.pose_warp_refinement_1(X1[1], f1[1], X2[1], f2[1], X1[2], emb_2, mask_2, q_2,
                                                              t_2)
        rows = []
        for q, t in ((q_1, t_1), (q_2, t_2), (q_3, t_3), (q_4, t_4)):
            qn = q / (torch.sqrt(torch.sum(q * q, dim=-1, keepdim=True) + 1e-10) + 1e-10)
            rows.append(torch.cat((t, qn), dim=-1).reshape(-1, 1, 7))
        return torch.cat(rows, dim=1), mask_1, x1[1]


class PWCLONET(Enum):
    pwclonet = PWCLONet

    @staticmethod
    def load(config, pose=None):
        assert "type" in config and config["type"] in PWCLONET.__members__
        return PWCLONET.__members__[config["type"]].value(config)
