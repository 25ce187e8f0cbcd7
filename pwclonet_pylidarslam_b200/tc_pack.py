"""Host-side packing of folded weights for the tcgen05 layer kernels (pack_tc2: the hybrid tf32 + bf16
scheme the kernels use; pack_tc: the plain hi/lo 3xTF32 layout kept for the self-test).

A weight matrix W [N, K] (N outputs, K inputs, K-major like a torch conv weight) is split into
tf32-exact hi / lo parts and stored chunk-wise in the canonical K-major no-swizzle shared-memory
layout of the UMMA descriptor: [K/32 chunks][hi, lo][N/8][k/4 (8)][n%8 (8)][k%4 (4)] fp32, so that a
chunk (32 input channels, hi+lo) is ONE contiguous TMA bulk copy and one MMA B operand per 8 inputs.
"""
import numpy as np


def split_tf32(w):
    w = np.ascontiguousarray(w, np.float32)
    hi = (w.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (w - hi).astype(np.float32)
    return hi, lo


def pack_tc(W):
    """W [N, K] float -> flat fp32 array, K zero-padded to a multiple of 32, N must be a multiple of 8."""
    W = np.asarray(W, np.float64)
    N, K = W.shape
    assert N % 8 == 0
    K32 = (K + 31) // 32 * 32
    Wp = np.zeros((N, K32), np.float32)
    Wp[:, :K] = W.astype(np.float32)
    hi, lo = split_tf32(Wp)
    out = np.empty((K32 // 32, 2, N // 8, 8, 8, 4), np.float32)
    for part, m in enumerate((hi, lo)):
        # m[n, k] -> [chunk, n/8, k/4 within chunk, n%8, k%4]
        v = m.reshape(N // 8, 8, K32 // 32, 8, 4)          # [n/8, n%8, chunk, kc, k%4]
        out[:, part] = v.transpose(2, 0, 3, 1, 4)
    return out.reshape(-1)


def tf32_rna(w):
    """round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32.f32"""
    u = np.ascontiguousarray(w, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def bf16_bits(w):
    """round-to-nearest-even bfloat16 bit patterns (uint16) of an fp32 array"""
    u = np.ascontiguousarray(w, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def pack_tc2(W):
    """Hybrid packing (tc_mma.cuh v2): per 32-input chunk
         [tf32(W)        fp32  [N/8][k/4 (8)][n%8][k%4]]   N*128 B
         [bf16(W)        bf16  [N/8][k/8 (4)][n%8][k%8]]   N*64 B
         [bf16(W-tf32(W)) bf16 same layout]                N*64 B
    returned as a flat fp32 view (N*64 floats per chunk)."""
    W = np.asarray(W, np.float64)
    N, K = W.shape
    assert N % 8 == 0
    K32 = (K + 31) // 32 * 32
    Wp = np.zeros((N, K32), np.float32)
    Wp[:, :K] = W.astype(np.float32)
    hi = tf32_rna(Wp)
    lo = (Wp - hi).astype(np.float32)
    nch = K32 // 32
    out = np.empty((nch, N * 256), np.uint8)
    t = hi.reshape(N // 8, 8, nch, 8, 4).transpose(2, 0, 3, 1, 4)            # [chunk, n/8, k/4, n%8, k%4]
    out[:, : N * 128] = np.ascontiguousarray(t).reshape(nch, -1).view(np.uint8)
    for j, m in enumerate((bf16_bits(Wp), bf16_bits(lo))):
        v = m.reshape(N // 8, 8, nch, 4, 8).transpose(2, 0, 3, 1, 4)         # [chunk, n/8, k/8, n%8, k%8]
        out[:, N * 128 + j * N * 64: N * 128 + (j + 1) * N * 64] = np.ascontiguousarray(v).reshape(nch, -1).view(np.uint8)
    return out.reshape(-1).view(np.float32)
