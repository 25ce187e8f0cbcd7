// Furthest point sampling for sm_100a -- replaces EXT/src/sampling_gpu.cu:69-229 and the orphan
// /sampling_gpu_copy.cu:93-252 of the reference.
//
// Design (B200-first, not a port):
//   * one CTA per cloud; the cloud is block-resident for all m-1 dependent rounds: coordinates and
//     the running minimum distances live in REGISTERS (P points per thread), a second copy of the
//     coordinates sits in shared memory only so that the winner's xyz can be broadcast-read;
//     the reference re-reads 20 B/point from global memory every round and needs a `temp` buffer.
//   * block arg-max = redux.sync (one instruction per warp for the value, one for the tie key),
//     one shared-memory exchange and ONE __syncthreads per round (double-buffered slots); the
//     reference uses a 9-level shared-memory tree with 10 barriers per round.
//   * bit-exact selection: the distance is the same FMUL/FFMA/FFMA chain nvcc emits for the
//     reference, and exact ties are resolved exactly like the reference's tree: the candidate with
//     the smallest (bit-reversed (k mod T), k) wins, T = the reference's block size for this n
//     (EXT/include/cuda_utils.h:13-19).  See DESIGN.md "FPS tie order".
#include "common.cuh"

namespace pwclo {

// key orders candidates exactly as the reference tree does among equal values (smaller wins)
__device__ __forceinline__ unsigned fps_key(unsigned k, int logT) {
  if (logT == 0) return k;
  unsigned low = k & ((1u << logT) - 1u);
  return __brev(low) | (k >> logT);
}
__device__ __forceinline__ unsigned fps_key_to_index(unsigned key, int logT) {
  if (logT == 0) return key;
  unsigned top_mask = ~((1u << (32 - logT)) - 1u);
  return __brev(key & top_mask) | ((key & ~top_mask) << logT);
}

// MODE 0: xyz + mind in registers, MODE 1: xyz in smem, mind in registers, MODE 2: everything in
// global memory (mind in `scratch`), any n.
template <int P, int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float* __restrict__ xyz, int n, int m, int logT, int origin_skip, float* __restrict__ scratch,
           int32_t* __restrict__ idx) {
  extern __shared__ float smem[];
  constexpr int NW = THREADS / 32;
  float* sx = smem;
  float* sy = sx + (MODE == 2 ? 0 : n);
  float* sz = sy + (MODE == 2 ? 0 : n);
  __shared__ unsigned red_val[2][32];
  __shared__ unsigned red_key[2][32];

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)b * n * 3;
  idx += (size_t)b * m;
  float* gmind = (MODE == 2) ? scratch + (size_t)b * n : nullptr;

  float px[MODE == 0 ? P : 1], py[MODE == 0 ? P : 1], pz[MODE == 0 ? P : 1];
  float mind[MODE == 2 ? 1 : P];

  if (MODE != 2) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      int k = tid + j * THREADS;
      float x = 0.f, y = 0.f, z = 0.f;
      bool valid = k < n;
      if (valid) {
        x = xyz[k * 3 + 0]; y = xyz[k * 3 + 1]; z = xyz[k * 3 + 2];
        sx[k] = x; sy[k] = y; sz[k] = z;
        if (origin_skip) {
          float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
          valid = !((double)mag <= 1e-3);
        }
      }
      if (MODE == 0) { px[j] = x; py[j] = y; pz[j] = z; }
      mind[j] = valid ? 1e10f : -1.0f;   // -1 can never beat the running best (strict >, start -1)
    }
  } else {
    for (int k = tid; k < n; k += THREADS) {
      float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
      bool valid = true;
      if (origin_skip) {
        float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        valid = !((double)mag <= 1e-3);
      }
      gmind[k] = valid ? 1e10f : -1.0f;
    }
  }
  if (tid == 0) idx[0] = 0;
  __syncthreads();

  int old = 0;
  for (int r = 1; r < m; ++r) {
    float x1, y1, z1;
    if (MODE == 2) { x1 = xyz[old * 3 + 0]; y1 = xyz[old * 3 + 1]; z1 = xyz[old * 3 + 2]; }
    else { x1 = sx[old]; y1 = sy[old]; z1 = sz[old]; }

    float best = -1.0f;
    int bestk = 0;
    if (MODE == 2) {
      for (int k = tid; k < n; k += THREADS) {
        float d = dist2_ref_fma(xyz[k * 3 + 0] - x1, xyz[k * 3 + 1] - y1, xyz[k * 3 + 2] - z1);
        float d2 = fminf(d, gmind[k]);
        gmind[k] = d2;
        if (d2 > best) { best = d2; bestk = k; }
      }
    } else {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        float x, y, z;
        if (MODE == 0) { x = px[j]; y = py[j]; z = pz[j]; }
        else {
          int k = tid + j * THREADS;
          k = k < n ? k : 0;
          x = sx[k]; y = sy[k]; z = sz[k];
        }
        float d = dist2_ref_fma(x - x1, y - y1, z - z1);
        float d2 = fminf(d, mind[j]);
        mind[j] = d2;
        if (d2 > best) { best = d2; bestk = tid + j * THREADS; }
      }
    }
    // (value, key): value as order-preserving uint (+1 so that "no candidate" = 0)
    unsigned vb = best < 0.f ? 0u : __float_as_uint(best) + 1u;
    unsigned key = fps_key((unsigned)bestk, logT);
    unsigned wv = __reduce_max_sync(PWCLO_FULL_MASK, vb);
    unsigned wk = __reduce_min_sync(PWCLO_FULL_MASK, vb == wv ? key : 0xffffffffu);
    const int buf = r & 1;
    if (lane == 0) { red_val[buf][warp] = wv; red_key[buf][warp] = wk; }
    __syncthreads();
    unsigned v2 = lane < NW ? red_val[buf][lane] : 0u;
    unsigned k2 = lane < NW ? red_key[buf][lane] : 0xffffffffu;
    unsigned bv = __reduce_max_sync(PWCLO_FULL_MASK, v2);
    unsigned bk = __reduce_min_sync(PWCLO_FULL_MASK, v2 == bv ? k2 : 0xffffffffu);
    old = bv == 0u ? 0 : (int)fps_key_to_index(bk, logT);
    if (tid == 0) idx[r] = old;
  }
}

template <int P, int THREADS, int MODE>
static int launch_fps(const float* xyz, int B, int n, int m, int logT, int origin_skip, float* scratch, int32_t* idx,
                      cudaStream_t st) {
  size_t smem = MODE == 2 ? 0 : (size_t)3 * n * sizeof(float);
  auto kern = fps_kernel<P, THREADS, MODE>;
  if (smem > 32 * 1024) {  // static smem (reduction slots) counts against the 48 KB default limit
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<B, THREADS, smem, st>>>(xyz, n, m, logT, origin_skip, scratch, idx);
  return launch_status();
}

}  // namespace pwclo

using namespace pwclo;

// scratch for MODE 2 (n > 16384) is allocated lazily per call size; kept simple on purpose: that
// path is outside every BASELINE configuration and exists so that no size silently fails.
PWCLO_API int pwclo_furthest_point_sampling(const float* xyz, int B, int N, int m, unsigned flags, int32_t* idx,
                                            void* stream) {
  if (!xyz || !idx || B < 0 || N <= 0 || m < 0) return PWCLO_EINVAL;
  if (B == 0 || m == 0) return PWCLO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = (flags & PWCLO_FPS_CAP1024) ? 1024 : 512;
  int T = 1, logT = 0;
  while (T * 2 <= N && T * 2 <= cap) { T *= 2; ++logT; }
  const int skip = (flags & PWCLO_FPS_ORIGIN_SKIP) ? 1 : 0;
  // THREADS must be a multiple of T so that a thread's points share (k mod T): 512 or 1024 (cap 1024)
#define FPS_CASE(P, TH, MODE) return launch_fps<P, TH, MODE>(xyz, B, N, m, logT, skip, nullptr, idx, st)
  if (cap == 1024 || N > 4096) {
    if (N <= 1024) FPS_CASE(1, 1024, 0);
    if (N <= 2048) FPS_CASE(2, 1024, 0);
    if (N <= 4096) FPS_CASE(4, 1024, 0);
    if (N <= 8192) FPS_CASE(8, 1024, 0);
    if (N <= 16384) FPS_CASE(16, 1024, 1);
  } else {
    if (N <= 512) FPS_CASE(1, 512, 0);
    if (N <= 1024) FPS_CASE(2, 512, 0);
    if (N <= 2048) FPS_CASE(4, 512, 0);
    if (N <= 4096) FPS_CASE(8, 512, 0);
  }
#undef FPS_CASE
  // generic path: any N, minimum distances in a global scratch buffer
  float* scratch = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&scratch, (size_t)B * N * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  int rc = launch_fps<1, 1024, 2>(xyz, B, N, m, logT, skip, scratch, idx, st);
  cudaFreeAsync(scratch, st);
  return rc;
}
