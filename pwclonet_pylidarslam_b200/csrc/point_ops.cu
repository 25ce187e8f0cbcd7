// Stand-alone PointNet++ operators for sm_100a: gather / group (+ scatter-add gradients), ball
// query, three_nn, three_interpolate.  They back the reference's `pointnet2_ops._ext` surface
// (EXT/src/bindings.cpp:7-18).  Inside the fused PWCLO-Net forward none of the grouped tensors is
// ever materialised (see layers.cu); these kernels exist for the op-level drop-in API, for
// training through the reference's autograd Functions, and for the op-suite benchmark.
//
// group_points design (HBM-write-bound: the output is K times larger than the input):
//   * a CTA owns (cloud b, a tile of TC channels, a slice of the S*K output positions);
//   * the TC source rows points[b, c0:c0+TC, :] are staged in shared memory with
//     cp.async.bulk (TMA 1-D bulk copy, mbarrier completion) -- random 4-byte gathers then hit
//     shared-memory banks instead of costing one L1 wavefront per distinct 128-B line;
//   * each thread produces 4 consecutive output positions per channel and stores one float4:
//     fully coalesced 16-B stores, index vector read once per tile (int4) and reused by all TC
//     channels.  The reference kernel (EXT/src/group_points_gpu.cu:8-28) runs one CTA per cloud,
//     re-reads idx per channel and writes with a stride of K floats between lanes.
#include <math_constants.h>

#include "common.cuh"

namespace pwclo {

// ---------------------------------------------------------------- TMA bulk copy helpers (1-D)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}
// global -> shared bulk copy; size and both addresses must be multiples of 16 bytes
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- gather
__global__ void gather_points_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int C, int N,
                                     int M, float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int a = idx[(size_t)b * M + j];
  out[((size_t)b * C + c) * M + j] = __ldg(points + ((size_t)b * C + c) * N + a);
}

__global__ void gather_points_grad_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int C,
                                          int N, int M, float* __restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int a = idx[(size_t)b * M + j];
  atomicAdd(grad_points + ((size_t)b * C + c) * N + a, grad_out[((size_t)b * C + c) * M + j]);
}

// ---------------------------------------------------------------- group
constexpr int GROUP_THREADS = 256;

// generic path: any N / alignment.  thread = one output position e in [0, S*K), loops over a
// channel tile; coalesced stores, idx read once per thread.
template <int TC>
__global__ void __launch_bounds__(GROUP_THREADS)
group_points_generic_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int C, int N, int SK,
                            float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * TC;
  const int e = blockIdx.x * GROUP_THREADS + threadIdx.x;
  if (e >= SK) return;
  const int a = idx[(size_t)b * SK + e];
  const float* src = points + ((size_t)b * C + c0) * N + a;
  float* dst = out + ((size_t)b * C + c0) * SK + e;
  float v[TC];
#pragma unroll
  for (int c = 0; c < TC; ++c) v[c] = (c0 + c < C) ? __ldg(src + (size_t)c * N) : 0.f;
#pragma unroll
  for (int c = 0; c < TC; ++c)
    if (c0 + c < C) __stcs(dst + (size_t)c * SK, v[c]);
}

// TMA-staged path: requires N % 4 == 0, SK % 4 == 0 and 16-byte aligned bases.
// grid = (slices of SK, channel tiles, B); dynamic smem = TC * N floats.
template <int TC>
__global__ void __launch_bounds__(GROUP_THREADS)
group_points_tma_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int C, int N, int SK,
                        int slice, float* __restrict__ out) {
  extern __shared__ __align__(16) float rows[];  // [TC][N]
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * TC;
  const int nc = min(TC, C - c0);
  const int e_begin = blockIdx.x * slice;
  const int e_end = min(SK, e_begin + slice);

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)nc * N * sizeof(float);
    mbar_expect_tx(&bar, bytes);
    // rows c0..c0+nc of one cloud are contiguous in [B,C,N]: one bulk copy (<= 1 MiB each)
    const char* src = (const char*)(points + ((size_t)b * C + c0) * N);
    char* dst = (char*)rows;
    uint32_t left = bytes;
    while (left) {
      uint32_t chunk = left > 65536u ? 65536u : left;
      tma_load_1d(dst, src, chunk, &bar);
      dst += chunk; src += chunk; left -= chunk;
    }
  }
  // overlap: fetch this thread's first index vector while the bulk copy is in flight
  const int4* idx4 = reinterpret_cast<const int4*>(idx + (size_t)b * SK);
  int e = e_begin + threadIdx.x * 4;
  int4 a = make_int4(0, 0, 0, 0);
  if (e < e_end) a = __ldg(idx4 + e / 4);
  mbar_wait(&bar, 0);

  float* obase = out + ((size_t)b * C + c0) * SK;
  while (e < e_end) {
    const int en = e + GROUP_THREADS * 4;
    int4 an = make_int4(0, 0, 0, 0);
    if (en < e_end) an = __ldg(idx4 + en / 4);
#pragma unroll
    for (int c = 0; c < TC; ++c) {
      if (c < nc) {
        const float* r = rows + c * N;
        float4 v = make_float4(r[a.x], r[a.y], r[a.z], r[a.w]);
        __stcs(reinterpret_cast<float4*>(obase + (size_t)c * SK + e), v);
      }
    }
    a = an;
    e = en;
  }
}

__global__ void __launch_bounds__(GROUP_THREADS)
group_points_grad_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int C, int N, int SK,
                         float* __restrict__ grad_points) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int e = blockIdx.x * GROUP_THREADS + threadIdx.x;
  if (e >= SK) return;
  const int a = idx[(size_t)b * SK + e];
  atomicAdd(grad_points + ((size_t)b * C + c) * N + a, __ldcs(grad_out + ((size_t)b * C + c) * SK + e));
}

// ---------------------------------------------------------------- ball query
// one warp per query; hits are discovered 32 references at a time and written in ascending index
// order (EXT/src/ball_query_gpu.cu:27-42 semantics: first nsample hits, padding = first hit, zero
// row when there is none).
constexpr int BQ_WARPS = 8;
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz, int n, int m, float radius2,
                  int nsample, int32_t* __restrict__ idx) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * BQ_WARPS + (threadIdx.x >> 5);
  if (j >= m) return;
  xyz += (size_t)b * n * 3;
  const float* q = new_xyz + ((size_t)b * m + j) * 3;
  int32_t* o = idx + ((size_t)b * m + j) * nsample;
  const float qx = q[0], qy = q[1], qz = q[2];
  int cnt = 0, first = 0;
  for (int k0 = 0; k0 < n && cnt < nsample; k0 += 32) {
    const int k = k0 + lane;
    bool hit = false;
    if (k < n) {
      float d2 = dist2_ref_fma(qx - xyz[k * 3 + 0], qy - xyz[k * 3 + 1], qz - xyz[k * 3 + 2]);
      hit = d2 < radius2;
    }
    const unsigned mask = __ballot_sync(PWCLO_FULL_MASK, hit);
    if (mask) {
      if (cnt == 0) first = k0 + __ffs(mask) - 1;
      const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
      if (hit && slot < nsample) o[slot] = k;
      cnt += __popc(mask);
    }
  }
  cnt = min(cnt, nsample);
  for (int l = cnt + lane; l < nsample; l += 32) o[l] = first;  // first == 0 when nothing was hit
}

// ---------------------------------------------------------------- three_nn
// one warp per query: every lane keeps the three best of its strided subset, then the 32 partial
// lists are merged by (distance, index).  Equivalent to the reference's sequential scan with strict
// '<' (EXT/src/interpolate_gpu.cu:27-47): the result is the three smallest (d, k) pairs.
__device__ __forceinline__ bool lt_dk(float d, int k, float d2, int k2) { return d < d2 || (d == d2 && k < k2); }

constexpr int NN3_WARPS = 8;
__global__ void __launch_bounds__(NN3_WARPS * 32)
three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known, int n, int m,
                float* __restrict__ dist2, int32_t* __restrict__ idx) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * NN3_WARPS + (threadIdx.x >> 5);
  if (j >= n) return;
  known += (size_t)b * m * 3;
  const float* u = unknown + ((size_t)b * n + j) * 3;
  const float ux = u[0], uy = u[1], uz = u[2];
  const int BIG = 0x7fffffff;
  float d1 = CUDART_INF_F, d2_ = CUDART_INF_F, d3 = CUDART_INF_F;
  int i1 = BIG, i2 = BIG, i3 = BIG;
  for (int k = lane; k < m; k += 32) {
    float d = dist2_ref_fma(ux - known[k * 3 + 0], uy - known[k * 3 + 1], uz - known[k * 3 + 2]);
    if (d < d1) { d3 = d2_; i3 = i2; d2_ = d1; i2 = i1; d1 = d; i1 = k; }
    else if (d < d2_) { d3 = d2_; i3 = i2; d2_ = d; i2 = k; }
    else if (d < d3) { d3 = d; i3 = k; }
  }
  // three rounds of "global minimum by (d,k), owner pops its head"
  float od[3]; int oi[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float bd = d1; int bi = i1;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      float xd = __shfl_xor_sync(PWCLO_FULL_MASK, bd, off);
      int xi = __shfl_xor_sync(PWCLO_FULL_MASK, bi, off);
      if (lt_dk(xd, xi, bd, bi)) { bd = xd; bi = xi; }
    }
    od[r] = bd; oi[r] = bi;
    if (i1 == bi && bi != BIG) { d1 = d2_; i1 = i2; d2_ = d3; i2 = i3; d3 = CUDART_INF_F; i3 = BIG; }
  }
  if (lane < 3) {
    float d = lane == 0 ? od[0] : (lane == 1 ? od[1] : od[2]);
    int i = lane == 0 ? oi[0] : (lane == 1 ? oi[1] : oi[2]);
    dist2[((size_t)b * n + j) * 3 + lane] = d;                 // +inf where the reference stores (float)1e40
    idx[((size_t)b * n + j) * 3 + lane] = i == BIG ? 0 : i;    // reference index init is 0
  }
}

// ---------------------------------------------------------------- three_interpolate
__global__ void three_interpolate_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx,
                                         const float* __restrict__ weight, int c, int m, int n,
                                         float* __restrict__ out) {
  const int b = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t* ii = idx + ((size_t)b * n + j) * 3;
  const float* w = weight + ((size_t)b * n + j) * 3;
  const float* row = points + ((size_t)b * c + l) * m;
  // nvcc contracts p1*w1 + p2*w2 + p3*w3 of the reference (interpolate_gpu.cu:96-97) into
  // FMUL(p2,w2); FFMA(p1,w1,.); FFMA(p3,w3,.)
  out[((size_t)b * c + l) * n + j] =
      __fmaf_rn(__ldg(row + ii[2]), w[2], __fmaf_rn(__ldg(row + ii[0]), w[0], __fmul_rn(__ldg(row + ii[1]), w[1])));
}

__global__ void three_interpolate_grad_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
                                              const float* __restrict__ weight, int c, int n, int m,
                                              float* __restrict__ grad_points) {
  const int b = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int32_t* ii = idx + ((size_t)b * n + j) * 3;
  const float* w = weight + ((size_t)b * n + j) * 3;
  const float g = grad_out[((size_t)b * c + l) * n + j];
  float* row = grad_points + ((size_t)b * c + l) * m;
  atomicAdd(row + ii[0], __fmul_rn(g, w[0]));
  atomicAdd(row + ii[1], __fmul_rn(g, w[1]));
  atomicAdd(row + ii[2], __fmul_rn(g, w[2]));
}

}  // namespace pwclo

using namespace pwclo;

static inline bool grid_ok(long long y, long long z) { return y <= 65535 && z <= 65535; }

PWCLO_API int pwclo_gather_points(const float* points, const int32_t* idx, int B, int C, int N, int M, float* out,
                                  void* stream) {
  if (!points || !idx || !out || B < 0 || C < 0 || N <= 0 || M < 0) return PWCLO_EINVAL;
  if (B == 0 || C == 0 || M == 0) return PWCLO_OK;
  if (!grid_ok(C, B)) return PWCLO_EUNSUPPORTED;
  dim3 grid(ceil_div(M, 256), C, B);
  gather_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, idx, C, N, M, out);
  return launch_status();
}

PWCLO_API int pwclo_gather_points_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int M,
                                       float* grad_points, void* stream) {
  if (!grad_out || !idx || !grad_points || B < 0 || C < 0 || N <= 0 || M < 0) return PWCLO_EINVAL;
  if (B == 0 || C == 0 || M == 0) return PWCLO_OK;
  if (!grid_ok(C, B)) return PWCLO_EUNSUPPORTED;
  dim3 grid(ceil_div(M, 256), C, B);
  gather_points_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, M, grad_points);
  return launch_status();
}

PWCLO_API int pwclo_group_points(const float* points, const int32_t* idx, int B, int C, int N, int S, int K,
                                 float* out, void* stream) {
  if (!points || !idx || !out || B < 0 || C < 0 || N <= 0 || S < 0 || K < 0) return PWCLO_EINVAL;
  if (B == 0 || C == 0 || S == 0 || K == 0) return PWCLO_OK;
  const long long SKl = (long long)S * K;
  if (SKl > 0x7fffffffLL || !grid_ok(C, B)) return PWCLO_EUNSUPPORTED;
  const int SK = (int)SKl;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = (N % 4 == 0) && (SK % 4 == 0) && (((uintptr_t)points | (uintptr_t)idx | (uintptr_t)out) % 16 == 0);
  // TMA-staged path when the staged rows are re-used enough (output slice >= 2 x N per channel).
  // Channel tile: as many rows as fit in 64 KB of shared memory (3 CTAs per SM stay resident).
  if (aligned && N <= 16384 && SK >= 2 * N) {
    int tc = 1;
    while (tc < 16 && (size_t)(tc * 2) * N * sizeof(float) <= 64 * 1024 && tc < C) tc *= 2;
    const size_t smem = (size_t)tc * N * sizeof(float);
    const int ctile = ceil_div(C, tc);
    // slice: multiple of 1024 outputs, sized so that the grid is ~4 CTAs per SM, never below 2N
    long long want = (long long)kNumSM * 4;
    int slices = (int)((want + (long long)B * ctile - 1) / ((long long)B * ctile));
    if (slices < 1) slices = 1;
    int slice = ceil_div(ceil_div(SK, slices), GROUP_THREADS * 4) * GROUP_THREADS * 4;
    if (slice < 2 * N) slice = ceil_div(2 * N, GROUP_THREADS * 4) * GROUP_THREADS * 4;
    void (*kern)(const float*, const int32_t*, int, int, int, int, float*) =
        tc == 1 ? group_points_tma_kernel<1> : tc == 2 ? group_points_tma_kernel<2>
        : tc == 4 ? group_points_tma_kernel<4> : tc == 8 ? group_points_tma_kernel<8> : group_points_tma_kernel<16>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(ceil_div(SK, slice), ctile, B);
    kern<<<grid, GROUP_THREADS, smem, st>>>(points, idx, C, N, SK, slice, out);
    return launch_status();
  }
  constexpr int TCG = 8;
  dim3 grid(ceil_div(SK, GROUP_THREADS), ceil_div(C, TCG), B);
  group_points_generic_kernel<TCG><<<grid, GROUP_THREADS, 0, st>>>(points, idx, C, N, SK, out);
  return launch_status();
}

PWCLO_API int pwclo_group_points_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int S, int K,
                                      float* grad_points, void* stream) {
  if (!grad_out || !idx || !grad_points || B < 0 || C < 0 || N <= 0 || S < 0 || K < 0) return PWCLO_EINVAL;
  if (B == 0 || C == 0 || S == 0 || K == 0) return PWCLO_OK;
  const long long SKl = (long long)S * K;
  if (SKl > 0x7fffffffLL || !grid_ok(C, B)) return PWCLO_EUNSUPPORTED;
  dim3 grid(ceil_div((int)SKl, GROUP_THREADS), C, B);
  group_points_grad_kernel<<<grid, GROUP_THREADS, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, (int)SKl, grad_points);
  return launch_status();
}

PWCLO_API int pwclo_ball_query(const float* new_xyz, const float* xyz, int B, int n, int m, float radius, int nsample,
                               int32_t* idx, void* stream) {
  if (!new_xyz || !xyz || !idx || B < 0 || n <= 0 || m < 0 || nsample < 0) return PWCLO_EINVAL;
  if (B == 0 || m == 0 || nsample == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  dim3 grid(ceil_div(m, BQ_WARPS), B);
  ball_query_kernel<<<grid, BQ_WARPS * 32, 0, (cudaStream_t)stream>>>(new_xyz, xyz, n, m, radius * radius, nsample, idx);
  return launch_status();
}

PWCLO_API int pwclo_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2,
                             int32_t* idx, void* stream) {
  if (!unknown || !known || !dist2 || !idx || B < 0 || n < 0 || m < 0) return PWCLO_EINVAL;
  if (B == 0 || n == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  dim3 grid(ceil_div(n, NN3_WARPS), B);
  three_nn_kernel<<<grid, NN3_WARPS * 32, 0, (cudaStream_t)stream>>>(unknown, known, n, m, dist2, idx);
  return launch_status();
}

PWCLO_API int pwclo_three_interpolate(const float* points, const int32_t* idx, const float* weight, int B, int c,
                                      int m, int n, float* out, void* stream) {
  if (!points || !idx || !weight || !out || B < 0 || c < 0 || m <= 0 || n < 0) return PWCLO_EINVAL;
  if (B == 0 || c == 0 || n == 0) return PWCLO_OK;
  if (!grid_ok(c, B)) return PWCLO_EUNSUPPORTED;
  dim3 grid(ceil_div(n, 256), c, B);
  three_interpolate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, idx, weight, c, m, n, out);
  return launch_status();
}

PWCLO_API int pwclo_three_interpolate_grad(const float* grad_out, const int32_t* idx, const float* weight, int B,
                                           int c, int n, int m, float* grad_points, void* stream) {
  if (!grad_out || !idx || !weight || !grad_points || B < 0 || c < 0 || m <= 0 || n < 0) return PWCLO_EINVAL;
  if (B == 0 || c == 0 || n == 0) return PWCLO_OK;
  if (!grid_ok(c, B)) return PWCLO_EUNSUPPORTED;
  dim3 grid(ceil_div(n, 256), c, B);
  three_interpolate_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, idx, weight, c, n, m, grad_points);
  return launch_status();
}

PWCLO_API const char* pwclo_version(void) { return "pwclo_b200 0.1 (sm_100a)"; }

PWCLO_API const char* pwclo_error_string(int code) {
  if (code == PWCLO_OK) return "ok";
  if (code == PWCLO_EINVAL) return "invalid argument (null pointer or bad size)";
  if (code == PWCLO_EUNSUPPORTED) return "size not supported by the sm_100a kernels";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown pwclo error";
}
