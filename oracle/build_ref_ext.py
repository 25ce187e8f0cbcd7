"""Build recipe for the REFERENCE's own pointnet2 CUDA extension (test infrastructure only).

Compiles the unmodified sources where they lie under /root/reference
(slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops/_ext-src) for sm_100a and
writes ONLY build outputs into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
Nothing from the reference is copied into the repository.

The resulting module `oracle/_ref/pwclo_ref_ext.so` is the index oracle-of-record on the GPU
(the reference kernels themselves, recompiled for Blackwell).  It is only ever loaded by tests/
and by bench.py's reference legs -- never by the product package.
"""
import os
import sys
import glob

REF_ROOT = os.environ.get("PWCLO_REFERENCE_ROOT", "/root/reference")
EXT_SRC = os.path.join(REF_ROOT, "slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops/_ext-src")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "pwclo_ref_ext"


def built_path():
    p = os.path.join(OUT, NAME + ".so")
    return p if os.path.exists(p) else None


def build(verbose=False):
    if not os.path.isdir(EXT_SRC):
        return built_path()
    if built_path():
        return built_path()
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    srcs = sorted(glob.glob(os.path.join(EXT_SRC, "src", "*.cpp")) + glob.glob(os.path.join(EXT_SRC, "src", "*.cu")))
    load(NAME, sources=srcs, extra_include_paths=[os.path.join(EXT_SRC, "include")],
         extra_cflags=["-O3"], extra_cuda_cflags=["-O3", "-lineinfo"], with_cuda=True,
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return built_path()


def load_module():
    """Import the built reference extension as a python module (needs torch + CUDA libs)."""
    p = built_path()
    if p is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
