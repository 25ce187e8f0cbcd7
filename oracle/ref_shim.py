"""oracle/ref_shim.py -- TEST INFRASTRUCTURE ONLY; works only where /root/reference is mounted.

Imports the UNMODIFIED reference PWCLO-Net (slam/models/PWCLONet/pwclo_net.py:32-207) on CPU by
injecting the import stubs listed in SURVEY.md section 9.7 (omegaconf, slam.common.utils,
slam.common.pose) and a CPU `pointnet2_ops._ext` stand-in backed by oracle/pointnet2_cpu.c.
Used by oracle/make_golden.py to generate tests/golden/*.npz and by tests to validate the
travelling restatement (oracle/pwclo_port.py).  The reference sources are read in place, never
copied.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("PWCLO_REFERENCE_ROOT", "/root/reference")
P2_LIB = os.path.join(REF_ROOT, "slam/models/Pointnet2_PyTorch/pointnet2_ops_lib")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "slam/models/PWCLONet"))


class DictConfig(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


def _install_stubs(ext_module):
    os.environ["PYLIDAR_SLAM_PWCLONET_ABS_PATH"] = REF_ROOT
    for p in (P2_LIB, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = DictConfig
    sys.modules.setdefault("omegaconf", oc)
    cu = types.ModuleType("slam.common.utils")

    def assert_debug(cond, msg=""):
        assert cond, msg

    cu.assert_debug = assert_debug
    sys.modules["slam.common.utils"] = cu
    cp = types.ModuleType("slam.common.pose")

    class Pose:
        def __init__(self, pose_type="quaternions"):
            self.pose_type = pose_type

        def num_rot_params(self):
            return 4 if self.pose_type == "quaternions" else 3

    cp.Pose = Pose
    sys.modules["slam.common.pose"] = cp
    # the package __init__ imports `pointnet2_ops.*`; the extension is looked up as pointnet2_ops._ext
    import importlib
    pkg = importlib.import_module("pointnet2_ops") if "pointnet2_ops" in sys.modules else None
    sys.modules["pointnet2_ops._ext"] = ext_module
    if pkg is not None:
        pkg._ext = ext_module


def load_reference(ext_module=None, device="cpu"):
    """Returns (PWCLONet instance in eval mode, module namespace dict)."""
    import torch
    if not available():
        raise RuntimeError("reference tree not mounted")
    if ext_module is None:
        from oracle.cpu_ops import torch_ext
        ext_module = torch_ext
    _install_stubs(ext_module)
    from slam.models.PWCLONet.pwclo_net import PWCLONet  # noqa: E402
    import slam.models.Pointnet2_PyTorch.pointnet2_ops_lib.pointnet2_ops.pointnet2_utils as p2u
    p2u._ext = ext_module
    torch.manual_seed(0)
    net = PWCLONet(DictConfig(num_input_channels=3, sequence_len=2, device=device, scalar_last=False))
    return net.eval()


def reference_modules():
    """The reference python modules (after load_reference) for per-layer checks."""
    import slam.models.Pointnet2_PyTorch.pointnet2_ops_lib.pointnet2_ops.pointnet2_utils as p2u
    import slam.models.Pointnet2_PyTorch.pointnet2_ops_lib.pointnet2_ops.pytorch_utils as ptu
    import slam.models.Pointnet2_PyTorch.pointnet2_ops_lib.pointnet2_ops.pointnet2_modules as p2m
    import slam.models.PWCLONet.PWCLO_utils as pwu
    return {"pointnet2_utils": p2u, "pytorch_utils": ptu, "pointnet2_modules": p2m, "PWCLO_utils": pwu}
