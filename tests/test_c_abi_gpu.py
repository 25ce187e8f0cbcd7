"""The drop-in boundary is a C ABI: a plain C program (tests/c_abi_driver.c: CUDA runtime + include/pwclo_b200.h,
no Python, no torch) runs FPS -> gather -> kNN -> group -> ball query through libpwclo_b200.so and compares with the
C oracle bit for bit."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from oracle import cpu_ops
    from pwclonet_pylidarslam_b200 import _lib
    oracle_so = cpu_ops.build()
    exe = str(tmp_path / "c_abi_driver")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-O1", "-std=c11", os.path.join(ROOT, "tests", "c_abi_driver.c"), "-o", exe,
                           f"-I{cuda}/include", f"-L{libdir}", "-lpwclo_b200", f"-L{os.path.dirname(oracle_so)}",
                           "-loracle_pwclo", f"-L{cuda}/lib64", "-lcudart", "-lm",
                           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{os.path.dirname(oracle_so)}", f"-Wl,-rpath,{cuda}/lib64"])
    return exe


def test_c_driver_compiles_and_links(tmp_path):
    """CPU: the header is valid C11 and every symbol the driver uses resolves against the two libraries"""
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_driver_matches_oracle(cuda, tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("bit-exact") == 5 and "OK" in r.stdout
