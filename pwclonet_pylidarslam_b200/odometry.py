"""PWCLO-Net as an odometry algorithm with its pose post-processing on the GPU (N4).

`PWCLONetOdometry` has the interface of the reference's `OdometryAlgorithm` (slam/odometry/odometry.py:21-81:
init / process_next_frame / do_process_next_frame / get_relative_poses / get_elapsed, keys `odometry_pc`
and `odometry_pose`) and the structure of `PoseNetOdometry` (slam/odometry/posenet_odometry.py:46-122): keep
the previous cloud, run the network on (current, previous) -- the frame order the KITTI dataset feeds the
network, kitti_odometry_dataset.py:330-343, 462-463 --, turn the finest-level row of pose_params into a
4x4 relative pose exactly as the evaluation does (train.py:875-886: quat2mat, then np.linalg.inv) and
append it.  `get_absolute_poses` chains them as KITTI360_TRANSFORMATIONS.convert_to_absolute does
(slam/common/kitti360_utils.py:406-432).  The forward is the fused sm_100a engine; pose matrices never
leave the GPU until they are read.  No CPU path.
"""
import ctypes
import os
import time

import numpy as np
import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def pose_params_to_matrices(pose_params, invert=True):
    """pose_params float32 CUDA [B,4,7] (row 0 = finest level is used) or [B,7] -> float64 [B,4,4]
    (train.py:875-886; `invert` = the np.linalg.inv the evaluation applies)."""
    if not (isinstance(pose_params, torch.Tensor) and pose_params.is_cuda and pose_params.dtype == torch.float32):
        raise RuntimeError("pose_params must be a float32 CUDA tensor (there is no CPU path)")
    pose_params = pose_params.contiguous()
    if pose_params.dim() == 3:
        B, stride = pose_params.shape[0], pose_params.shape[1] * pose_params.shape[2]
    elif pose_params.dim() == 2 and pose_params.shape[1] == 7:
        B, stride = pose_params.shape[0], 7
    else:
        raise RuntimeError("pose_params must be [B,4,7] or [B,7]")
    out = torch.empty((B, 4, 4), dtype=torch.float64, device=pose_params.device)
    with torch.cuda.device(pose_params.device):
        _lib.check(_lib.lib().pwclo_pose_to_matrix(_p(pose_params), B, stride, 1 if invert else 0, _p(out), _lib.stream_ptr()),
                   "pose_to_matrix")
    return out


def convert_to_absolute(relative_poses, first_transformation=None, dict_semantics=False):
    """float64 CUDA [F,4,4] relative poses -> [F,4,4] absolute poses (kitti360_utils.py:406-432).

    The reference treats `first_transformation` differently per input type; the input here is an array, so the
    default follows its ndarray branch (:412-420):  abs_f = inv(rel_f @ ... @ rel_0 @ first).
    dict_semantics=True gives the dict branch (:422-427):  abs_f = inv(rel_f @ inv(abs_{f-1})), abs_{-1} = first,
    i.e. inv(rel_f @ ... @ rel_0 @ inv(first)).  Without a first transformation the two are the same."""
    if not (isinstance(relative_poses, torch.Tensor) and relative_poses.is_cuda and relative_poses.dtype == torch.float64):
        raise RuntimeError("relative_poses must be a float64 CUDA tensor (there is no CPU path)")
    rel = relative_poses.contiguous()
    F = rel.shape[0]
    first = None
    if first_transformation is not None:
        first = torch.as_tensor(first_transformation, dtype=torch.float64).reshape(4, 4)
        if not dict_semantics:          # the kernel's scan is seeded with abs_{-1}: inv(first) turns it into the array branch
            first = torch.linalg.inv(first.cpu())
        first = first.to(rel.device).contiguous()
    out = torch.empty_like(rel)
    with torch.cuda.device(rel.device):
        _lib.check(_lib.lib().pwclo_accumulate_poses(_p(rel), F, _p(first), _p(out), _lib.stream_ptr()), "accumulate_poses")
    return out


class PWCLONetOdometry:
    """Deep LiDAR odometry on PWCLO-Net.  config keys (all optional): train_dir + checkpoint_file (reference
    checkpoint with a `prediction_module` entry, trainer.py:882-907), num_points (8192), device."""

    def __init__(self, config=None, pose=None, device=None, prediction_module=None, **kwargs):
        self.config = dict(config or {})
        self.elapsed = []
        self.device = torch.device(device or self.config.get("device", "cuda:0"))
        if self.device.type != "cuda":
            raise RuntimeError("PWCLONetOdometry needs a CUDA device (there is no CPU path)")
        self.pose = pose
        self.num_points = int(self.config.get("num_points", 8192))
        if prediction_module is None:
            from .training import _PWCLONetPredictionModule
            prediction_module = _PWCLONetPredictionModule({"device": str(self.device), "num_points": self.num_points})
        self.prediction_module = prediction_module.to(self.device).eval()
        self.checkpoint_path = None
        if self.config.get("train_dir"):
            self.checkpoint_path = os.path.join(self.config["train_dir"], self.config.get("checkpoint_file", "checkpoint.ckp"))
        self.previous_cloud = None
        self._iter = 0
        self.relative_poses = []

    @staticmethod
    def pointcloud_key():
        return "odometry_pc"

    @staticmethod
    def relative_pose_key():
        return "odometry_pose"

    def init(self):
        self.elapsed = []
        self.relative_poses = []
        self.previous_cloud = None
        self._iter = 0
        if self.checkpoint_path:
            sd = torch.load(self.checkpoint_path, map_location=self.device, weights_only=False)
            self.prediction_module.load_state_dict(sd["prediction_module"])
            self.prediction_module.eval()

    def process_next_frame(self, data_dict):
        beginning = time.time()
        self.do_process_next_frame(data_dict)
        self.elapsed.append(time.time() - beginning)

    def _cloud(self, data_dict):
        pc = data_dict[self.pointcloud_key()]
        pc = torch.as_tensor(pc)
        if pc.dim() != 2 or pc.shape[1] < 3 or pc.shape[0] < self.num_points:
            raise RuntimeError(f"`{self.pointcloud_key()}` must be [N>={self.num_points}, >=3]")
        return pc[:self.num_points, :3].to(self.device, dtype=torch.float32, non_blocking=True).contiguous().unsqueeze(0)

    def do_process_next_frame(self, data_dict):
        cloud = self._cloud(data_dict)
        if self._iter == 0:
            rel = torch.eye(4, dtype=torch.float64, device=self.device).unsqueeze(0)
        else:
            with torch.no_grad():
                pose_params, _ = self.prediction_module([cloud, self.previous_cloud])
            rel = pose_params_to_matrices(pose_params, invert=True)
        self.previous_cloud = cloud
        self.relative_poses.append(rel)
        data_dict[self.relative_pose_key()] = rel[0]          # stays on the device until the caller reads it
        self._iter += 1

    def process_pairs(self, clouds):
        """batched form for offline sequences: clouds float32 [F,N,3] on the device -> F relative poses in one
        forward of F-1 frame pairs (pairs are independent: they shard across GPUs, see sharding.py)"""
        F = clouds.shape[0]
        rel = [torch.eye(4, dtype=torch.float64, device=self.device).unsqueeze(0)]
        if F > 1:
            with torch.no_grad():
                pose_params, _ = self.prediction_module([clouds[1:], clouds[:-1]])
            rel.append(pose_params_to_matrices(pose_params, invert=True))
        self.relative_poses = [torch.cat(rel)]
        self.previous_cloud = clouds[-1:].contiguous()
        self._iter = F
        return self.relative_poses[0]

    def get_relative_poses(self):
        return torch.cat(self.relative_poses, dim=0).cpu().numpy()

    def get_absolute_poses(self, first_transformation=None):
        return convert_to_absolute(torch.cat(self.relative_poses, dim=0), first_transformation).cpu().numpy()

    def get_elapsed(self):
        return sum(self.elapsed)


def save_poses(poses, file_path, frames=None):
    """KITTI-style pose text file as the reference writes it (KITTI360_IO.save_poses,
    slam/common/kitti360_utils.py:267-304, used by train.py:935-952): one line per frame,
    `frame r00 r01 r02 tx r10 ... tz`, blank-separated, floats in their shortest round-trip form (what
    pandas.DataFrame.to_csv emits).  poses: [F,4,4] / [F,3,4] / [F,12] array or tensor (a CUDA tensor is read back
    once), or a dict frame -> 4x4; frames default to 0..F-1."""
    if poses is None:
        raise RuntimeError("[save_poses]: poses is None")
    if isinstance(poses, dict):
        keys = sorted(poses.keys())
        frames = [int(k) for k in keys]
        arr = np.stack([np.asarray(poses[k], dtype=np.float64)[:3, :] for k in keys]).reshape(len(keys), 12)
    else:
        if torch.is_tensor(poses):
            poses = poses.detach().cpu().numpy()
        arr = np.asarray(poses, dtype=np.float64)
        if arr.ndim == 3 and arr.shape[1:] in ((4, 4), (3, 4)):
            arr = arr[:, :3, :].reshape(arr.shape[0], 12)
        elif not (arr.ndim == 2 and arr.shape[1] == 12):
            raise RuntimeError("[save_poses]: poses should be an array of size (-1, 4, 4), (-1, 3, 4) or (-1, 12)")
        frames = list(range(arr.shape[0])) if frames is None else [int(f) for f in frames]
    with open(file_path, "w") as f:
        for fr, row in zip(frames, arr):
            f.write(" ".join([str(fr)] + [repr(float(v)) for v in row]) + "\n")
