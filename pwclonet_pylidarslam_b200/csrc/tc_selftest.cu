// Self-test of the tcgen05 3xTF32 building blocks: D[128 x N] = A[128 x K] * W[N x K]^T.
// A is split into hi/lo planes in TMEM, W (hi/lo, host-packed canonical K-major chunks) sits in
// shared memory, the accumulator is read back from TMEM.  tests/test_tc_gpu.py compares with fp64.
#include "tc_mma.cuh"

namespace pwclo {

// wpk layout: [K/32 chunks][2 (hi, lo)][N/8][8 (k/4)][8 (n%8)][4 (k%4)]
__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ wpk,
                                                          int K, int N, int mode, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* wsm = reinterpret_cast<float*>(smem_raw);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbarrier_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  const int COL_HI = 0, COL_LO = 192, COL_D = 384;
  // A row -> hi/lo planes
  for (int c = 0; c < K; c += 16) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = A[(size_t)tid * K + c + i];
    tmem_st16_split(lane_base + COL_HI + c, lane_base + COL_LO + c, v);
  }
  tmem_wait_st();
  const int wfloats = (K / 32) * 2 * N * 32;
  for (int i = tid; i < wfloats; i += 128) wsm[i] = wpk[i];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = tc_idesc_tf32(128, N);
    uint32_t acc = 0;
    for (int ch = 0; ch < K / 32; ++ch) {
      const float* bh = wsm + (size_t)ch * 2 * N * 32;
      const float* bl = bh + N * 32;
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t a_hi = tb + COL_HI + ch * 32 + ks * 8, a_lo = tb + COL_LO + ch * 32 + ks * 8;
        const uint64_t d_bh = tc_smem_desc(bh + ks * 64), d_bl = tc_smem_desc(bl + ks * 64);
        tc_mma_ts(tb + COL_D, a_hi, d_bh, idesc, acc);
        acc = 1;
        if (mode >= 3) {
          tc_mma_ts(tb + COL_D, a_lo, d_bh, idesc, 1);
          tc_mma_ts(tb + COL_D, a_hi, d_bl, idesc, 1);
        }
      }
    }
    tc_commit(&bar);
  }
  mbarrier_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(lane_base + COL_D + c, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


// Hybrid scheme (tc_mma.cuh, "v2"): wpk2 layout per 32-input chunk = [tf32(W): N*32 fp32][bf16(W): N*32 bf16]
// [bf16(W - tf32(W)): N*32 bf16] (N*256 bytes), see tc_pack.pack_tc2.
__global__ void __launch_bounds__(128) tc_selftest2_kernel(const float* __restrict__ A, const float* __restrict__ wpk,
                                                           int K, int N, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* wsm = reinterpret_cast<float*>(smem_raw);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbarrier_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  const int COL_T = 0, COL_HB = 192, COL_LB = 288, COL_D = 384;
  for (int c = 0; c < K; c += 16) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = A[(size_t)tid * K + c + i];
    tmem_st16_hybrid(lane_base + COL_T + c, lane_base + COL_HB + c / 2, lane_base + COL_LB + c / 2, v);
  }
  tmem_wait_st();
  const int wfloats = (K / 32) * N * 64;
  for (int i = tid; i < wfloats; i += 128) wsm[i] = wpk[i];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    const uint32_t it = tc_idesc_tf32(128, N), ib = tc_idesc_bf16(128, N);
    const uint32_t base = smem_addr(wsm);
    for (int ch = 0; ch < K / 32; ++ch) {
      const uint32_t cb = base + (uint32_t)ch * N * 256;
      tc_mma_hybrid_chunk_warp(tb + COL_D, tb + COL_T + ch * 32, tb + COL_HB + ch * 16, tb + COL_LB + ch * 16,
                               tc_desc_at(cb, TC_DESC_TF32), tc_desc_at(cb + N * 128, TC_DESC_BF16),
                               tc_desc_at(cb + N * 192, TC_DESC_BF16), it, ib, ch > 0);
    }
    tc_commit_warp(smem_addr(&bar));
  }
  mbarrier_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(lane_base + COL_D + c, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

}  // namespace pwclo

using namespace pwclo;

// A [128,K] row-major, wpk packed as documented above, D [128,N]; K % 32 == 0, K <= 192, N in {64,128}.
// mode 1: single TF32 product, mode 3: error-compensated 3xTF32 (wpk from tc_pack.pack_tc),
// mode 5: hybrid tf32 + bf16 correction terms (wpk from tc_pack.pack_tc2).
PWCLO_API int pwclo_tc_selftest(const float* A, const float* wpk, int K, int N, int mode, float* D, void* stream) {
  if (!A || !wpk || !D || K <= 0 || K % 32 != 0 || K > 192 || (N != 64 && N != 128)) return PWCLO_EINVAL;
  const size_t smem = (size_t)(K / 32) * 2 * N * 32 * sizeof(float) + 1024;
  if (mode == 5) {
    cudaError_t e2 = cudaFuncSetAttribute(tc_selftest2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) return (int)e2;
    tc_selftest2_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, wpk, K, N, D);
    return launch_status();
  }
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, wpk, K, N, mode, D);
  return launch_status();
}
