"""PWCLONet with the reference's constructor, attributes, parameter layout (510 state-dict
tensors) and forward contract (PW/pwclo_net.py:32-218):

    pose_params[B,4,7], log_dict = net(xyz_f1[B,3,N], None, xyz_f2[B,3,N], None)

Execution:
  * `net.eval()`, xyz-only input that does not require grad (with or without `torch.no_grad()`: the
    reference's own test path, train.py:348-359 / :806-808, calls the model in eval() with autograd on)
                             -> the fused sm_100a inference engine (`..fused.FusedPWCLONet`): BN folded,
    one kernel per layer, no grouped tensor ever materialised, no forced device->host sync (the
    reference's log_dict lives on the CPU, pwclo_net.py:186-193; ours is produced lazily).  Its output
    carries no autograd graph; config `eval_autograd=True` keeps the composed path in eval mode for
    whoever differentiates through an eval() forward;
  * otherwise                -> the autograd composition below (same op order as the reference),
    all sampling / neighbour / grouping ops on the sm_100a kernels.
"""
import contextlib
import operator
import warnings
from collections.abc import Mapping
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


class LazyLog(Mapping):
    """log_dict whose CPU tensors are only materialised when read (keeps the forward asynchronous).  A read-only
    Mapping over the reference's keys (PW/pwclo_net.py:186-193): `[]`, `.get`, `.items()`, `.values()`, iteration,
    `len`, `dict(log)` and `**log` all see every key; a value is computed (one device->host copy) the first
    time it is read and cached."""

    def __init__(self, thunks):
        self._thunks = dict(thunks)
        self._values = {}

    def __getitem__(self, key):
        if key not in self._values:
            self._values[key] = self._thunks[key]()
        return self._values[key]

    def __contains__(self, key):           # Mapping's default would evaluate the thunk
        return key in self._thunks

    def __iter__(self):
        return iter(self._thunks)

    def __len__(self):
        return len(self._thunks)

    def __repr__(self):
        return "LazyLog(" + ", ".join(f"{k}=<{'ready' if k in self._values else 'lazy'}>" for k in self._thunks) + ")"


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
        self.eval_autograd = config.get("eval_autograd", False)
        self.lazy_log = config.get("lazy_log", True)
        self.graph_forward = config.get("graph_forward", True)

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
        self._fused_fp = None
        # a parent module's load_state_dict recurses through _load_from_state_dict and never calls this module's
        # load_state_dict: the post hook fires in both cases
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_fused())

    # -- fused-engine plumbing ------------------------------------------------------------------
    def invalidate_fused(self):
        """drop the folded-weight engine (and its captured graphs): the next eval forward re-folds"""
        self._fused = None
        self._fused_fp = None
        self.__dict__.pop("_fp_tensors", None)

    def train(self, mode=True):
        self.invalidate_fused()     # parameters / BN statistics may change: re-fold on next eval forward
        return super().train(mode)

    def _apply(self, fn, *a, **kw):       # .to() / .cuda() / .float(): the arena lives on the old device
        self.invalidate_fused()
        return super()._apply(fn, *a, **kw)

    def _fingerprint(self):
        """in-place updates through torch while in eval mode (an optimiser step, `p.add_`, `p.copy_`) bump the
        tensors' version counters; rebinding `p.data` moves the storage.  Writes that bypass torch altogether
        (`p.data.add_`, a raw kernel) need invalidate_fused()."""
        plist = self.__dict__.get("_fp_tensors")
        if plist is None:               # collected once per engine build (invalidate_fused() drops it)
            plist = list(self.parameters()) + list(self.buffers())
            self.__dict__["_fp_tensors"] = plist
        return (tuple(map(operator.attrgetter("_version"), plist)), plist[0].data_ptr(), plist[-1].data_ptr())

    def fused_engine(self):
        fp = self._fingerprint()
        if self._fused is None or fp != self._fused_fp:
            from ..fused import FusedPWCLONet
            self._fused = FusedPWCLONet(self)
            self._fused_fp = fp
        return self._fused

    # -- forward ----------------------------------------------------------------------------------
    def pyramid_geometry(self, xyz_f1, xyz_f2):
        """The coordinates-only, non-differentiable part of the siamese pyramid for both frames stacked (2B clouds): per
        level (FPS indices, centres, neighbour indices).  It depends on the input clouds alone -- not on the weights -- so a
        training loop can compute it for batch i+1 while batch i is still in its backward pass and hand it to
        forward(..., geoms=...) (training.PWCLONetTrainer does, on a second stream)."""
        cur = torch.cat((xyz_f1.permute(0, 2, 1), xyz_f2.permute(0, 2, 1)), dim=0).contiguous().detach()
        geoms = []
        for psa in (self.psa_1, self.psa_2, self.psa_3, self.psa_4):
            g = psa.geometry(cur)
            geoms.append(g)
            cur = g[1]
        return geoms

    def forward(self, xyz_f1, points_f1, xyz_f2, points_f2, bn_decay=None, geoms=None):
        no_graph_needed = (not torch.is_grad_enabled()) or (
            not self.eval_autograd and not xyz_f1.requires_grad and not xyz_f2.requires_grad)
        fused_ok = (self.use_fused and not self.training and no_graph_needed
                    and points_f1 is None and points_f2 is None and xyz_f1.is_cuda)
        if fused_ok:
            with torch.no_grad():
                eng = self.fused_engine()
                if self.graph_forward:
                    pose, mask1, xyz1_l1 = eng.forward_graphed(xyz_f1, xyz_f2, slot=self.__dict__.get("fused_slot", 0))
                else:
                    pose, mask1, xyz1_l1 = eng.forward(xyz_f1, xyz_f2)
        else:
            from ..pytorch_utils import deferred_bn_counters
            with deferred_bn_counters():
                pose, mask1, xyz1_l1 = self._forward_composed(xyz_f1, points_f1, xyz_f2, points_f2, geoms)
        thunks = {
            "embedding_mask": lambda: torch.linalg.norm(
                F.softmax(mask1.detach().cpu(), dim=2).permute(0, 2, 1), dim=-1, ord=2),
            "point_cloud": lambda: xyz1_l1.detach().cpu(),
        }
        log = LazyLog(thunks)
        if not self.lazy_log:
            log = {k: log[k] for k in thunks}
        return pose, log

    def _forward_composed(self, xyz_f1, points_f1, xyz_f2, points_f2, geoms=None):
        xyz_f1_t = xyz_f1.permute(0, 2, 1).contiguous()
        xyz_f2_t = xyz_f2.permute(0, 2, 1).contiguous()
        x1, f1, x2, f2 = [xyz_f1_t], [points_f1], [xyz_f2_t], [points_f2]
        # FPS / gather / kNN depend on coordinates only and are not differentiable: one batch of 2B clouds per level
        # for both frames (per cloud the same indices as the reference's per-frame calls; with 8 pairs per GPU the
        # 16384 -> 2048 sampling is a 2047-round chain on 8 CTAs, twice, otherwise).  The layers keep their per-frame
        # calls: train-mode BatchNorm statistics must be those of one frame's batch, as in the reference.
        Bp = xyz_f1_t.shape[0]
        # (the FPS prefix shortcut -- geometry(tie_in=..., return_tie=True) -- is not used: measured, tracking ties costs more
        # on level 1 than the skipped levels give back as soon as one cloud of the batch has a tie; see fused.py)
        if geoms is None:
            geoms = self.pyramid_geometry(xyz_f1, xyz_f2)
        # The two frames' pyramids are independent until the first cost volume: in training on the GPU the second frame's runs
        # on the branch stream (pytorch_utils.branch_stream), forward and backward.  Both use the same SharedMLPs, so the
        # branch leaves their BatchNorm running statistics alone and its updates are applied after the join -- frame 1
        # first, frame 2 second, as in the reference's sequential calls.
        from ..pytorch_utils import branch_stream, deferred_running_stats, on_branch
        stream = branch_stream(xyz_f1.device) if self.training else None
        for psa, g in zip((self.psa_1, self.psa_2, self.psa_3, self.psa_4), geoms):
            a, b = psa(x1[-1], f1[-1], geom=tuple(t[:Bp] for t in g))
            x1.append(a), f1.append(b)
        branch = on_branch(stream)
        with branch, (deferred_running_stats() if stream is not None else contextlib.nullcontext()) as later:
            for psa, g in zip((self.psa_1, self.psa_2, self.psa_3, self.psa_4), geoms):
                a, b = psa(x2[-1], f2[-1], geom=tuple(t[Bp:] for t in g))
                x2.append(a), f2.append(b)
        branch.join(*(x2[1:] + f2[1:]))
        if later is not None:
            later.apply()
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
