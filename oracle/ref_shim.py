"""oracle/ref_shim.py -- TEST INFRASTRUCTURE ONLY; needs /root/reference or its staged python files (baseline/_ref).

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

def _find_root():
    """/root/reference where it is mounted (this container); on the GPU box the unmodified hot-path python files staged
    by oracle/stage_reference.py into the git-ignored baseline/_ref/"""
    if os.environ.get("PWCLO_REFERENCE_ROOT"):
        return os.environ["PWCLO_REFERENCE_ROOT"]
    if os.path.isdir("/root/reference/slam/models/PWCLONet"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REF_ROOT = _find_root()
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


def load_reference_loss(with_exp_weights=True, init_weights=(0.0, -2.5), loss_weights=(1.0, 1.0)):
    """The UNMODIFIED `_PWCLONetLossModule` (slam/training/loss_modules.py:329-544) on CPU.  Its module
    imports hydra / pyquaternion / omegaconf.OmegaConf and slam.common.{geometry,optimization,projection},
    none of which the loss touches: they are stubbed, the class body runs as written."""
    import dataclasses
    if not available():
        raise RuntimeError("reference tree not mounted")
    _install_stubs(types.ModuleType("pointnet2_ops._ext_stub"))
    sys.modules["omegaconf"].OmegaConf = type("OmegaConf", (), {"create": staticmethod(lambda x: x)})
    hy, hc, hcore, hcs = (types.ModuleType(n) for n in ("hydra", "hydra.conf", "hydra.core", "hydra.core.config_store"))
    hc.dataclass, hc.field, hc.MISSING = dataclasses.dataclass, dataclasses.field, "???"

    class ConfigStore:
        @staticmethod
        def instance():
            return ConfigStore()

        def store(self, **kw):
            pass

    hcs.ConfigStore = ConfigStore
    for n, m in (("hydra", hy), ("hydra.conf", hc), ("hydra.core", hcore), ("hydra.core.config_store", hcs)):
        sys.modules.setdefault(n, m)
    pq = types.ModuleType("pyquaternion")
    pq.Quaternion = object
    sys.modules.setdefault("pyquaternion", pq)
    sys.modules["slam.common.utils"].check_tensor = lambda *a, **k: None
    for name, attrs in (("slam.common.geometry", ("compute_normal_map", "projection_map_to_points")),
                        ("slam.common.optimization", ("_LS_SCHEME", "_WLSScheme", "PointToPlaneCost")),
                        ("slam.common.projection", ("Projector",))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, object)
            sys.modules[name] = m
    import importlib
    lm = importlib.import_module("slam.training.loss_modules")
    cfg = DictConfig(mode="supervised", loss_degrees=False, loss_weights=list(loss_weights),
                     with_exp_weights=with_exp_weights, init_weights=list(init_weights), loss_option="l2_norm",
                     nb_levels=4, device="cpu", scalar_last=False)
    return lm._PWCLONetLossModule(cfg, sys.modules["slam.common.pose"].Pose("quaternions"))
