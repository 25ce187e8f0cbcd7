"""The UNMODIFIED reference model on the B200 (SURVEY 8b row B0, 8c "oracle of record on the GPU").

The reference's hot-path python files are staged (unmodified, git-ignored) by oracle/stage_reference.py into
baseline/_ref/ and its own CUDA extension is compiled for sm_100a by oracle/build_ref_ext.py into oracle/_ref/; both
travel to the GPU box.  Checked here:
  (a) reference PWCLONet + reference extension  vs  this repository's fused forward: sampled coordinates bit-exact,
      pose within 1e-4 m / 1e-5 rad, mask feature within 1e-4 relative (TF32 disabled on both sides);
  (b) the same reference model on top of `_ext.register_as_pointnet2_ops_ext()` (our nine-function drop-in) gives the
      identical pose; with `pt_utils.knn_point` also replaced by ours the pose stays within tolerance (neighbour sets
      differ only inside exact distance ties)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import _common as C

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def setup(cuda, ref_ext):
    from oracle import ref_shim
    from pwclonet_pylidarslam_b200 import synthetic as syn
    from pwclonet_pylidarslam_b200.pwclonet import PWCLONet
    if ref_ext is None or not ref_shim.available():
        pytest.skip("reference extension / staged reference python not available")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = ref_shim.load_reference(ext_module=ref_ext, device="cuda:0").to(cuda).eval()
    mods = ref_shim.reference_modules()
    shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    w = C.weights_for(shapes, 3)
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    ref.load_state_dict(sd)
    ours = PWCLONet({"device": "cuda:0"})
    ours.load_state_dict(sd)                 # the reference's state dict loads unchanged
    ours = ours.to(cuda).eval()
    x1, x2, _ = syn.make_batch(920, 2, 8192)
    a, b = torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda)
    with torch.no_grad():
        pose_ref, log_ref = ref(a, None, b, None)
    return dict(ref=ref, ours=ours, mods=mods, ref_ext=ref_ext, a=a, b=b, pose_ref=pose_ref.cpu().numpy(), log_ref=log_ref)


def test_reference_model_on_gpu_vs_fused_forward(setup):
    with torch.no_grad():
        pose, log = setup["ours"](setup["a"], None, setup["b"], None)
    te, re_ = C.pose_errors(pose.cpu().numpy(), setup["pose_ref"])
    print(f"fused vs unmodified reference on the B200: translation {te:.3e} m, rotation {re_:.3e} rad")
    assert te <= C.TOL_TRANSLATION_M and re_ <= C.TOL_ROTATION_RAD
    lr = setup["log_ref"]
    np.testing.assert_array_equal(log["point_cloud"].numpy(), lr["point_cloud"].numpy())     # FPS + gather: bit-exact
    assert C.rel_err(log["embedding_mask"].numpy(), lr["embedding_mask"].numpy()) <= C.TOL_FEATURE_REL


def test_reference_model_on_our_extension(setup):
    """row B0: swap the extension under the unmodified reference python for this repository's drop-in"""
    from pwclonet_pylidarslam_b200 import _ext
    from pwclonet_pylidarslam_b200.pytorch_utils import knn_point
    p2u, ptu = setup["mods"]["pointnet2_utils"], setup["mods"]["pytorch_utils"]
    mod = _ext.register_as_pointnet2_ops_ext()
    assert sys.modules["pointnet2_ops._ext"] is mod
    old_ext, old_knn = p2u._ext, ptu.knn_point
    try:
        p2u._ext = mod
        with torch.no_grad():
            pose_b, log_b = setup["ref"](setup["a"], None, setup["b"], None)
        np.testing.assert_array_equal(log_b["point_cloud"].numpy(), setup["log_ref"]["point_cloud"].numpy())
        np.testing.assert_allclose(pose_b.cpu().numpy(), setup["pose_ref"], rtol=0, atol=1e-6)   # same ops, same indices
        ptu.knn_point = knn_point
        with torch.no_grad():
            pose_c, _ = setup["ref"](setup["a"], None, setup["b"], None)
        te, re_ = C.pose_errors(pose_c.cpu().numpy(), setup["pose_ref"])
        assert te <= C.TOL_TRANSLATION_M and re_ <= C.TOL_ROTATION_RAD
    finally:
        p2u._ext, ptu.knn_point = old_ext, old_knn
        sys.modules["pointnet2_ops._ext"] = setup["ref_ext"]


def test_reference_imports_our_extension_when_registered_first(setup):
    """the drop-in path a reference user takes: register, then import the reference unchanged -- its
    `import pointnet2_ops._ext` (P2/pointnet2_utils.py:7-8) resolves to our module; fresh interpreter"""
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from pwclonet_pylidarslam_b200 import _ext\n"
        "mod = _ext.register_as_pointnet2_ops_ext()\n"
        "from oracle import ref_shim\n"
        "net = ref_shim.load_reference(ext_module=mod, device='cuda:0').to('cuda:0')\n"
        "p2u = ref_shim.reference_modules()['pointnet2_utils']\n"
        "assert p2u._ext is mod\n"
        "from pwclonet_pylidarslam_b200 import synthetic as syn\n"
        "x1, x2, _ = syn.make_batch(5, 1, 8192)\n"
        "with torch.no_grad():\n"
        "    pose, _ = net(torch.from_numpy(x1).cuda(), None, torch.from_numpy(x2).cuda(), None)\n"
        "assert pose.shape == (1, 4, 7) and bool(torch.isfinite(pose).all())\n"
        "print('DROPIN_OK')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
