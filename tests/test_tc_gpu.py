"""tcgen05 / TMEM building blocks: the 3xTF32 tensor-core contraction against fp64."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(32, 64), (64, 128), (160, 128), (192, 64)])
def test_tc_gemm_3xtf32(cuda, K, N):
    from pwclonet_pylidarslam_b200 import _lib, tc_pack
    rng = np.random.default_rng(K + N)
    A = rng.standard_normal((128, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    want = A.astype(np.float64) @ W.astype(np.float64).T
    dA = torch.from_numpy(A).to(cuda)
    dW = torch.from_numpy(tc_pack.pack_tc(W)).to(cuda)
    errs = {}
    for mode in (1, 3):
        D = torch.zeros(128, N, device=cuda)
        rc = _lib.lib().pwclo_tc_selftest(ctypes.c_void_p(dA.data_ptr()), ctypes.c_void_p(dW.data_ptr()), K, N, mode,
                                          ctypes.c_void_p(D.data_ptr()), _lib.stream_ptr())
        _lib.check(rc, "tc_selftest")
        torch.cuda.synchronize()
        errs[mode] = float(np.abs(D.cpu().numpy() - want).max() / np.abs(want).max())
    print(f"K={K} N={N}: rel err tf32 {errs[1]:.2e}, 3xtf32 {errs[3]:.2e}")
    assert errs[1] < 5e-3          # plain TF32: ~1e-3
    assert errs[3] < 5e-6          # error-compensated: fp32 class (2^-21 per product)


@pytest.mark.parametrize("K,N", [(32, 64), (64, 128), (160, 128), (192, 64)])
def test_tc_gemm_hybrid(cuda, K, N):
    """tf32 main product + two bf16 correction products (8 MMAs per 32 inputs instead of 12)."""
    from pwclonet_pylidarslam_b200 import _lib, tc_pack
    rng = np.random.default_rng(K + N + 1)
    A = rng.standard_normal((128, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    want = A.astype(np.float64) @ W.astype(np.float64).T
    dA = torch.from_numpy(A).to(cuda)
    dW = torch.from_numpy(tc_pack.pack_tc2(W)).to(cuda)
    D = torch.zeros(128, N, device=cuda)
    rc = _lib.lib().pwclo_tc_selftest(ctypes.c_void_p(dA.data_ptr()), ctypes.c_void_p(dW.data_ptr()), K, N, 5,
                                      ctypes.c_void_p(D.data_ptr()), _lib.stream_ptr())
    _lib.check(rc, "tc_selftest")
    torch.cuda.synchronize()
    err = float(np.abs(D.cpu().numpy() - want).max() / np.abs(want).max())
    print(f"K={K} N={N}: rel err hybrid {err:.2e}")
    assert err < 5e-6
