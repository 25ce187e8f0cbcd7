"""oracle/stage_reference.py -- TEST INFRASTRUCTURE ONLY.

Stages the reference's hot-path PYTHON files (slam/models/PWCLONet/*.py and
slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops/*.py, unmodified) into the git-ignored
`baseline/_ref/` so that they travel to the GPU box with gpurun, where /root/reference does not exist.  There
tests/test_reference_gpu.py and tools/bench_reference_gpu.py run the UNMODIFIED reference model on the B200 -- on its
own CUDA extension (oracle/_ref, built by oracle/build_ref_ext.py) and on this repository's drop-in
(`_ext.register_as_pointnet2_ops_ext()`).  Nothing staged is tracked by git or imported by the
product package.  Called by __graft_entry__.build() where /root/reference is mounted.
"""
import os
import shutil

SRC = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
P2 = "slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops"
FILES = ["slam/__init__.py", "slam/models/__init__.py", "slam/common/__init__.py",
         "slam/models/PWCLONet/pwclo_net.py", "slam/models/PWCLONet/costvolume.py", "slam/models/PWCLONet/flowpredictor.py",
         "slam/models/PWCLONet/pose_calculator.py", "slam/models/PWCLONet/pose_warp_refinement.py",
         "slam/models/PWCLONet/PWCLO_utils.py",
         P2 + "/__init__.py", P2 + "/_version.py", P2 + "/pointnet2_modules.py", P2 + "/pointnet2_utils.py",
         P2 + "/pytorch_utils.py"]


def staged():
    return os.path.isfile(os.path.join(DST, "slam/models/PWCLONet/pwclo_net.py"))


def stage():
    """copy the files (only where the reference tree is mounted); returns the staging root or None"""
    if not os.path.isdir(os.path.join(SRC, "slam/models/PWCLONet")):
        return DST if staged() else None
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    return DST


if __name__ == "__main__":
    print(stage())
