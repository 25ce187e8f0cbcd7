// Brute-force k-nearest-neighbour search for sm_100a -- replaces the reference's pure-PyTorch
// `knn_point` (P2/pytorch_utils.py:12-49), which materialises [B,S,N,3] differences (~2 GB of HBM
// traffic per 8192-point cloud) and runs torch.topk.
//
// Design:
//   * the reference set of one cloud is staged once per CTA in shared memory as SoA tiles
//     (<= 8192 points = 96 KB) and re-used by every query the CTA owns; HBM/L2 traffic is the
//     compulsory 12 B/point per CTA instead of 12 B per (query, point) pair;
//   * one warp owns Q queries at a time; per 32-reference chunk each lane evaluates Q distances
//     from one shared-memory read of the reference point (register tiling over queries);
//   * selection (warp-select): the running top-K of a query is a warp-distributed sorted list (lane i
//     holds the i-th best).  A candidate is first tested against a conservative squared-distance
//     bound of the K-th entry (no sqrt); survivors (O(K log(N/K)) per query) are appended to a small
//     shared-memory queue and merged 32 at a time with a warp bitonic sort + merge on 64-bit
//     (distance, index) keys;
//   * the compared quantity is bit-for-bit the reference's: unfused fp32 (q-r)^2 products, the
//     size-3 sum in torch's order (selectable), + 1e-8f, IEEE sqrt; equal distances are ordered by
//     index (torch.topk leaves that order unspecified);
//   * optional fused pose warp of the queries (PW/PWCLO_utils.py:42-63), so the warped cloud of a
//     refinement level never makes a round trip through HBM before its neighbour search.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"

namespace pwclo {

constexpr int KNN_WARPS = 16;       // warps per CTA
constexpr int KNN_MAX_TILE = 8192;  // reference points per shared-memory tile
constexpr int KNN_BUF = 64;         // pending-candidate slots per (warp, query)

typedef unsigned long long u64;
constexpr u64 KNN_INF_KEY = 0x7f800000ffffffffull;

// (distance, index) packed so that integer order == (distance asc, index asc); distances are >= 0
__device__ __forceinline__ u64 knn_key(float v, int i) { return ((u64)__float_as_uint(v) << 32) | (unsigned)i; }

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) {
  unsigned lo = __shfl_xor_sync(PWCLO_FULL_MASK, (unsigned)v, m);
  unsigned hi = __shfl_xor_sync(PWCLO_FULL_MASK, (unsigned)(v >> 32), m);
  return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
  unsigned lo = __shfl_sync(PWCLO_FULL_MASK, (unsigned)v, src);
  unsigned hi = __shfl_sync(PWCLO_FULL_MASK, (unsigned)(v >> 32), src);
  return ((u64)hi << 32) | lo;
}

// Merge up to 32 unsorted candidate keys (one per lane, KNN_INF_KEY = empty) into the sorted
// 32-entry list (lane i = i-th smallest).  Bitonic sort of the candidates (descending, 15
// compare-exchange steps) + bitonic merge (1 + 5 steps): ~170 instructions per 32 candidates instead
// of ~30 per candidate for one-at-a-time insertion.
__device__ __noinline__ u64 knn_merge32(u64 list, u64 cand, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const u64 other = shfl_xor_u64(cand, j);
      const bool desc_block = (lane & k) == 0 || k == 32;   // final pass: whole warp descending
      const bool lower = (lane & j) == 0;
      const bool keep_max = lower == desc_block;
      cand = keep_max ? (cand > other ? cand : other) : (cand < other ? cand : other);
    }
  }
  // list ascending, cand descending: element-wise min is a bitonic sequence holding the 32 smallest
  u64 c = list < cand ? list : cand;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const u64 other = shfl_xor_u64(c, j);
    const bool lower = (lane & j) == 0;
    c = lower ? (c < other ? c : other) : (c > other ? c : other);
  }
  return c;
}

// rd(rd(e1 + e3) + e2) * KNN_LB_SLACK > bound  =>  d2 = rn(rn(A + B) + C) > bound for any association of the
// three squared components with e_i <= its component: the exact sum T >= the rounded-down sum, d2 >= T (1 - 2^-23)
constexpr float KNN_LB_SLACK = 0.99999976f;   // 1 - 2^-22

// Conservative squared-distance bound: d2 > bound  =>  sqrt_rn(d2 + 1e-8f) >= kth  
__device__ __forceinline__ float knn_bound(float kth) { return __fmul_ru(__fmul_ru(kth, kth), 1.0000005f); }

template <int Q, int SUM_ORDER>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S, int K, int tile,
           int q_per_cta, const float* __restrict__ warp_qt, float* __restrict__ warped_out,
           int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  extern __shared__ float smem[];
  float* sx = smem;
  float* sy = sx + tile;
  float* sz = sy + tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* cand_d = sz + tile + (size_t)warp * Q * KNN_BUF;                                        // [Q][KNN_BUF]
  int* cand_i = reinterpret_cast<int*>(sz + tile + (size_t)KNN_WARPS * Q * KNN_BUF) + (size_t)warp * Q * KNN_BUF;

  const int b = blockIdx.y;
  xyz += (size_t)b * N * 3;
  new_xyz += (size_t)b * S * 3;
  const int q_begin = blockIdx.x * q_per_cta;
  const int q_end = min(S, q_begin + q_per_cta);

  PoseQT pose;
  const bool do_warp = warp_qt != nullptr;
  if (do_warp) pose = make_pose(warp_qt + (size_t)b * 7);

  // a cloud that fits one tile is staged exactly once per CTA
  const bool single_tile = N <= tile;
  if (single_tile) {
    for (int i = threadIdx.x; i < N; i += KNN_WARPS * 32) {
      sx[i] = xyz[i * 3 + 0];
      sy[i] = xyz[i * 3 + 1];
      sz[i] = xyz[i * 3 + 2];
    }
    __syncthreads();
  }

  // the CTA's queries are processed in rounds of KNN_WARPS*Q; every round walks all reference tiles.
  for (int q0 = q_begin; q0 < q_end; q0 += KNN_WARPS * Q) {
    float qx[Q], qy[Q], qz[Q], bound[Q];
    u64 list[Q];
    int cnt[Q], qi[Q];
#pragma unroll
    for (int u = 0; u < Q; ++u) {
      qi[u] = q0 + warp * Q + u;
      int qq = min(qi[u], S - 1);
      float x = new_xyz[qq * 3 + 0], y = new_xyz[qq * 3 + 1], z = new_xyz[qq * 3 + 2];
      if (do_warp) {
        warp_point(pose, x, y, z, x, y, z);
        if (warped_out != nullptr && lane == 0 && qi[u] < q_end) {
          float* w = warped_out + ((size_t)b * S + qq) * 3;
          w[0] = x; w[1] = y; w[2] = z;
        }
      }
      qx[u] = x; qy[u] = y; qz[u] = z;
      list[u] = KNN_INF_KEY;
      bound[u] = CUDART_INF_F;
      cnt[u] = 0;
    }

    for (int t0 = 0; t0 < N; t0 += tile) {
      const int tn = min(tile, N - t0);
      if (!single_tile) {
        __syncthreads();  // previous tile / round fully consumed
        for (int i = threadIdx.x; i < tn; i += KNN_WARPS * 32) {
          sx[i] = xyz[(t0 + i) * 3 + 0];
          sy[i] = xyz[(t0 + i) * 3 + 1];
          sz[i] = xyz[(t0 + i) * 3 + 2];
        }
        __syncthreads();
      }

      for (int c0 = 0; c0 < tn; c0 += 32) {
        const int r = c0 + lane;
        const bool rv = r < tn;
        const float rx = rv ? sx[r] : 0.f, ry = rv ? sy[r] : 0.f, rz = rv ? sz[r] : 0.f;
#pragma unroll
        for (int u = 0; u < Q; ++u) {
          const float dx = __fsub_rn(qx[u], rx), dy = __fsub_rn(qy[u], ry), dz = __fsub_rn(qz[u], rz);
          const float xx = __fmul_rn(dx, dx), yy = __fmul_rn(dy, dy), zz = __fmul_rn(dz, dz);
          const float d2 = SUM_ORDER == 0 ? __fadd_rn(__fadd_rn(xx, yy), zz) : __fadd_rn(__fadd_rn(xx, zz), yy);
          const bool pass = rv && d2 <= bound[u];
          const unsigned mask = __ballot_sync(PWCLO_FULL_MASK, pass);
          if (mask) {  // warp-uniform
            float* cd = cand_d + u * KNN_BUF;
            int* ci = cand_i + u * KNN_BUF;
            if (pass) {
              const int slot = cnt[u] + __popc(mask & ((1u << lane) - 1u));
              cd[slot] = d2;
              ci[slot] = t0 + r;
            }
            cnt[u] += __popc(mask);
            if (cnt[u] >= 32) {
              __syncwarp();
              const u64 ck = knn_key(__fsqrt_rn(__fadd_rn(cd[lane], 1e-8f)), ci[lane]);
              const int rest = cnt[u] - 32;
              float md = 0.f; int mi = 0;
              if (lane < rest) { md = cd[32 + lane]; mi = ci[32 + lane]; }
              __syncwarp();
              if (lane < rest) { cd[lane] = md; ci[lane] = mi; }
              cnt[u] = rest;
              list[u] = knn_merge32(list[u], ck, lane);
              const float kth = __uint_as_float((unsigned)(shfl_u64(list[u], K - 1) >> 32));
              bound[u] = knn_bound(kth);
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < Q; ++u) {
      if (cnt[u] > 0) {   // warp-uniform: drain the pending candidates
        __syncwarp();
        const float* cd = cand_d + u * KNN_BUF;
        const int* ci = cand_i + u * KNN_BUF;
        const u64 ck = lane < cnt[u] ? knn_key(__fsqrt_rn(__fadd_rn(cd[lane], 1e-8f)), ci[lane]) : KNN_INF_KEY;
        list[u] = knn_merge32(list[u], ck, lane);
      }
      __syncwarp();
      if (qi[u] < q_end && lane < K) {
        size_t o = ((size_t)b * S + qi[u]) * K + lane;
        idx_out[o] = (int)(unsigned)list[u];
        if (dist_out) dist_out[o] = __uint_as_float((unsigned)(list[u] >> 32));
      }
    }
  }
}

template <int Q>
static int launch_knn(const float* xyz, const float* new_xyz, int B, int N, int S, int K, int sum_order,
                      const float* warp_qt, float* warped_out, int32_t* idx, float* dist, cudaStream_t st) {
  const int tile = min(((N + 31) / 32) * 32, KNN_MAX_TILE);
  const size_t smem = (size_t)3 * tile * sizeof(float) + (size_t)KNN_WARPS * Q * KNN_BUF * 8;
  // queries per CTA: enough CTAs to fill the machine (>= 2 waves of 148 SMs x resident CTAs) while
  // amortising the shared-memory fill of the reference tile over as many queries as possible.
  const int per_round = KNN_WARPS * Q;
  int q_per_cta = per_round;
  const int resident = smem > 110 * 1024 ? 1 : (smem > 72 * 1024 ? 2 : 3);
  while (q_per_cta * 2 <= S && (long long)B * ceil_div(S, q_per_cta * 2) >= 2LL * kNumSM * resident) q_per_cta *= 2;
  dim3 grid(ceil_div(S, q_per_cta), B);
  auto kern = sum_order == PWCLO_KNN_SUM_XY_Z ? knn_kernel<Q, 0> : knn_kernel<Q, 1>;
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<grid, KNN_WARPS * 32, smem, st>>>(xyz, new_xyz, N, S, K, tile, q_per_cta, warp_qt, warped_out, idx, dist);
  return launch_status();
}

}  // namespace pwclo

using namespace pwclo;

PWCLO_API int pwclo_knn(const float* xyz, const float* new_xyz, int B, int N, int S, int K, int sum_order,
                        const float* warp_qt, float* warped_out, int32_t* idx, float* dist, void* stream) {
  if (!xyz || !new_xyz || !idx || B < 0 || N <= 0 || S < 0 || K <= 0) return PWCLO_EINVAL;
  if (K > 32) return PWCLO_EUNSUPPORTED;   // the reference never asks for more than 32 neighbours
  if (K > N) return PWCLO_EINVAL;          // torch.topk raises in the reference
  if (sum_order != PWCLO_KNN_SUM_XY_Z && sum_order != PWCLO_KNN_SUM_XZ_Y) return PWCLO_EINVAL;
  if (B == 0 || S == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  // queries per warp: as many as still leave >= 2 CTAs per SM
  const long long total = (long long)B * S;
  if (total >= 2LL * kNumSM * KNN_WARPS * 4)
    return launch_knn<4>(xyz, new_xyz, B, N, S, K, sum_order, warp_qt, warped_out, idx, dist, st);
  if (total >= 2LL * kNumSM * KNN_WARPS * 2)
    return launch_knn<2>(xyz, new_xyz, B, N, S, K, sum_order, warp_qt, warped_out, idx, dist, st);
  return launch_knn<1>(xyz, new_xyz, B, N, S, K, sum_order, warp_qt, warped_out, idx, dist, st);
}

// =================================================================================================
// Sorted-strip exact kNN (the default path for N <= 8192 when the caller provides a workspace).
//
// Brute force evaluates S*N pairs; for LiDAR clouds the K-th neighbour is ~0.5 m away while the
// cloud spans ~60 m, so almost all of that work is provably useless.  Two kernels:
//   1. knn_presort_kernel: one CTA per cloud sorts the reference points along the axis of largest
//      extent a1 (64-bit (ordered coordinate, index) keys, bitonic sort in shared memory), cuts the
//      sorted sequence into T equal-count strips and sorts every strip along the axis of second
//      largest extent a2.  It writes the strip-major SoA (x, y, z, original index), the axis ids and
//      the a1 boundaries of the strips to the workspace.
//   2. knn_slab_kernel / knn_slab_small_kernel: the sorted cloud is staged in shared memory by ONE
//      TMA bulk copy; a warp (or an 8/16-lane group) visits the strips nearest-first along a1 and,
//      inside a strip, binary-searches its query's a2 position and scans chunks outwards on both
//      sides, nearest side first, until  rd(e1 + e2) > bound, where e1 / e2 are the squared a1
//      distance to the strip and the squared a2 distance to the next unscanned point, and bound is
//      the conservative bound of the current K-th neighbour.  Distances, keys, queueing and merging
//      are exactly those of the brute-force kernel, so the result is bit-identical: the computed
//      d2 = fl(fl(A + B) + C) of a skipped point is >= rd(e1 + e2) > bound (fp32 addition of
//      non-negative terms and subtraction are monotone), hence sqrt(d2+1e-8) > K-th distance.
// =================================================================================================
namespace pwclo {

constexpr int SORT_THREADS = 1024;
constexpr int SLAB_WARPS = 16;

__device__ __forceinline__ unsigned ordered_bits(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int KNN_MAX_STRIPS = 32;
// floats per cloud: x | y | z | id (N4 each) | hdr (axis1, axis2, strips, log2 strip length) | a1 bounds of the strips
// (KNN_MAX_STRIPS + 4) | per-strip min / max of the third axis a3 (KNN_MAX_STRIPS each)
constexpr int KNN_HDR = 4 + (KNN_MAX_STRIPS + 4) + 2 * KNN_MAX_STRIPS;
__host__ __device__ inline size_t knn_ws_stride(int N) { return (size_t)4 * ((N + 3) & ~3) + KNN_HDR; }
// strips for a cloud padded to NP (power of two) points
__host__ __device__ inline int knn_log_strip(int NP, int K, int strips_req) {
  // measured on B200 (tools/run_knn.py): 16 strips for the big clouds, 8 when 32 neighbours are wanted from 2048 points
  int T = strips_req > 0 ? strips_req : (NP >= 4096 || (NP >= 2048 && K <= 16) ? 16 : (NP >= 512 ? 8 : (NP >= 128 ? 4 : 1)));
  if (T > KNN_MAX_STRIPS) T = KNN_MAX_STRIPS;
  int logL = 0;
  while ((1 << logL) < NP) ++logL;
  while (T > 1 && logL > 5) { T >>= 1; --logL; }   // strip length >= 32 points
  return logL;
}

// In-place bitonic sort network on shared-memory keys, stages k = 2 .. kmax (kmax < NP sorts every aligned
// kmax-block on its own, all ascending).  Every warp owns a contiguous run of compare-exchange pairs,
// so the sub-steps whose partner distance stays inside a warp's run need __syncwarp only.
__device__ __forceinline__ void knn_bitonic(u64* keys, int NP, int kmax, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int ppw = max(NP / 2 / (SORT_THREADS / 32), 32);      // pairs per warp
  const int t_begin = warp * ppw, t_end = min(NP / 2, t_begin + ppw);
  for (int k = 2; k <= kmax; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = t_begin + lane; t < t_end; t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j cleared
        const int hi = lo | j;
        const u64 a = keys[lo], c = keys[hi];
        const bool up = k == kmax || (lo & k) == 0;             // last stage: every block ascending
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      const int jn = j > 1 ? (j >> 1) : k;                      // partner distance of the next sub-step
      if (j <= ppw && jn <= ppw) __syncwarp(); else __syncthreads();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SORT_THREADS)
knn_presort_kernel(const float* __restrict__ xyz, int N, int NP, int logL, float* __restrict__ ws,
                   const float* __restrict__ queries, int S, int SP, int* __restrict__ qorder) {
  const int idbits = NP > 8192 ? 14 : 13;                 // width of the point index inside the in-strip sort key
  const unsigned idmask = (1u << idbits) - 1u;
  extern __shared__ __align__(16) unsigned char sort_smem[];
  u64* keys = reinterpret_cast<u64*>(sort_smem);
  __shared__ float red[6][32];
  __shared__ int axis_s, axis2_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)b * N * 3;
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int i = tid; i < N; i += SORT_THREADS)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = xyz[i * 3 + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int off = 16; off; off >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(PWCLO_FULL_MASK, mn[d], off));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(PWCLO_FULL_MASK, mx[d], off));
    }
    if (lane == 0) { red[d][warp] = mn[d]; red[3 + d][warp] = mx[d]; }
  }
  __syncthreads();
  if (tid == 0) {
    float ext[3];
    for (int d = 0; d < 3; ++d) {
      float a = CUDART_INF_F, c = -CUDART_INF_F;
      for (int w = 0; w < SORT_THREADS / 32; ++w) { a = fminf(a, red[d][w]); c = fmaxf(c, red[3 + d][w]); }
      ext[d] = c - a;
    }
    const int a1 = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[2] >= ext[1] ? 2 : 1);
    const int o1 = (a1 + 1) % 3, o2 = (a1 + 2) % 3;
    const int a2 = ext[o1] >= ext[o2] ? o1 : o2;
    // a single strip is a plain slab search: it should run along the longest axis
    const bool one = (1 << logL) >= NP;
    axis_s = one ? a2 : a1;
    axis2_s = one ? a1 : a2;
  }
  __syncthreads();
  const int axis = axis_s, axis2 = axis2_s;
  for (int i = tid; i < NP; i += SORT_THREADS)
    keys[i] = i < N ? (((u64)ordered_bits(xyz[i * 3 + axis]) << 32) | (unsigned)i) : ~0ull;
  __syncthreads();
  knn_bitonic(keys, NP, NP, tid);
  const int N4 = (N + 3) & ~3;
  float* w = ws + (size_t)b * knn_ws_stride(N);
  // ---- strips: equal-count cuts of the a1 order; bounds[t] = a1 of the first point of strip t,
  //      bounds[strips] = a1 of the last point
  const int L = 1 << logL;
  const int strips = (N + L - 1) >> logL;
  if (tid <= strips) {
    const int rnk = min(tid << logL, N - 1);
    w[4 * N4 + 4 + tid] = xyz[(int)(unsigned)keys[rnk] * 3 + axis];
  }
  if (tid == 0) {
    int* h = reinterpret_cast<int*>(w) + 4 * N4;
    h[0] = axis; h[1] = axis2; h[2] = strips; h[3] = logL;
  }
  __syncthreads();   // bounds read before the keys are rewritten
  // ---- every strip sorted along a2: key = (strip, ordered a2, index)
  for (int i = tid; i < NP; i += SORT_THREADS) {
    u64 k2 = ~0ull;
    if (i < N) {
      const unsigned id = (unsigned)keys[i];
      k2 = ((u64)(i >> logL) << (32 + idbits)) | ((u64)ordered_bits(xyz[id * 3 + axis2]) << idbits) | id;
    }
    keys[i] = k2;
  }
  __syncthreads();
  knn_bitonic(keys, NP, L, tid);
  for (int i = tid; i < N4; i += SORT_THREADS) {
    float x = CUDART_INF_F, y = CUDART_INF_F, z = CUDART_INF_F;
    int id = 0;
    if (i < N) {
      id = (int)((unsigned)keys[i] & idmask);
      x = xyz[id * 3 + 0]; y = xyz[id * 3 + 1]; z = xyz[id * 3 + 2];
    }
    w[i] = x; w[N4 + i] = y; w[2 * N4 + i] = z;
    reinterpret_cast<int*>(w)[3 * N4 + i] = id;
  }
  // ---- range of the third axis inside every strip (one warp per strip): a 3-D lower bound for the search
  if (warp < strips) {
    const int axis3 = 3 - axis - axis2;
    float lo3 = CUDART_INF_F, hi3 = -CUDART_INF_F;
    for (int i = (warp << logL) + lane; i < min(N, (warp + 1) << logL); i += 32) {
      const float v = xyz[(int)((unsigned)keys[i] & idmask) * 3 + axis3];
      lo3 = fminf(lo3, v); hi3 = fmaxf(hi3, v);
    }
    for (int off = 16; off; off >>= 1) {
      lo3 = fminf(lo3, __shfl_xor_sync(PWCLO_FULL_MASK, lo3, off));
      hi3 = fmaxf(hi3, __shfl_xor_sync(PWCLO_FULL_MASK, hi3, off));
    }
    if (lane == 0) {
      w[4 * N4 + 4 + KNN_MAX_STRIPS + 4 + warp] = lo3;
      w[4 * N4 + 4 + 2 * KNN_MAX_STRIPS + 4 + warp] = hi3;
    }
  }

  // ---- query visiting order: Morton order (6 bits per axis) of the queries of this cloud, so that the
  // queries a warp processes back to back are neighbours in space (their K-th distance + their mutual
  // distance bounds the next query's K-th distance: a tight pruning bound from the first chunk on)
  if (qorder == nullptr) return;
  __syncthreads();
  queries += (size_t)b * S * 3;
  unsigned* qk = reinterpret_cast<unsigned*>(sort_smem);
  float qmn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, qmx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int i = tid; i < S; i += SORT_THREADS)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = queries[i * 3 + d];
      qmn[d] = fminf(qmn[d], v);
      qmx[d] = fmaxf(qmx[d], v);
    }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int off = 16; off; off >>= 1) {
      qmn[d] = fminf(qmn[d], __shfl_xor_sync(PWCLO_FULL_MASK, qmn[d], off));
      qmx[d] = fmaxf(qmx[d], __shfl_xor_sync(PWCLO_FULL_MASK, qmx[d], off));
    }
    if (lane == 0) { red[d][warp] = qmn[d]; red[3 + d][warp] = qmx[d]; }
  }
  __syncthreads();
  float lo3[3], sc3[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float a = CUDART_INF_F, c = -CUDART_INF_F;
    for (int wv = 0; wv < SORT_THREADS / 32; ++wv) { a = fminf(a, red[d][wv]); c = fmaxf(c, red[3 + d][wv]); }
    lo3[d] = a;
    sc3[d] = c > a ? 63.999f / (c - a) : 0.f;
  }
  for (int i = tid; i < SP; i += SORT_THREADS) {
    unsigned key = 0xffffffffu;
    if (i < S) {
      unsigned code = 0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const unsigned cell = (unsigned)fminf(fmaxf((queries[i * 3 + d] - lo3[d]) * sc3[d], 0.f), 63.f);
#pragma unroll
        for (int bit = 0; bit < 6; ++bit) code |= ((cell >> bit) & 1u) << (3 * bit + d);
      }
      key = (code << 13) | (unsigned)i;
    }
    qk[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= SP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < SP / 2; t += SORT_THREADS) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const unsigned a = qk[lo], c = qk[hi];
        const bool up = (lo & k) == 0;
        if ((a > c) == up) { qk[lo] = c; qk[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < S; i += SORT_THREADS) qorder[(size_t)b * S + i] = (int)(qk[i] & 0x1fffu);
}

__device__ __forceinline__ void slab_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

template <int SUM_ORDER>
__global__ void __launch_bounds__(1024)
knn_slab_kernel(const float* __restrict__ ws, const int* __restrict__ qorder, const float* __restrict__ new_xyz, int N,
                int S, int K, int q_per_cta, const float* __restrict__ warp_qt, float* __restrict__ warped_out, int32_t* __restrict__ idx_out,
                float* __restrict__ dist_out) {
  extern __shared__ __align__(128) unsigned char slab_smem[];
  const int N4 = (N + 3) & ~3;
  float* sx = reinterpret_cast<float*>(slab_smem);
  float* sy = sx + N4;
  float* sz = sy + N4;
  int* sid = reinterpret_cast<int*>(sz + N4);
  int* hdr = sid + N4;                        // [0] = axis
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* cand_d = reinterpret_cast<float*>(hdr + KNN_HDR) + (size_t)warp * KNN_BUF;
  int* cand_i = reinterpret_cast<int*>(reinterpret_cast<float*>(hdr + KNN_HDR) + (size_t)nwarps * KNN_BUF) + (size_t)warp * KNN_BUF;
  __shared__ __align__(8) uint64_t bar;

  const int b = blockIdx.y;
  const float* w = ws + (size_t)b * knn_ws_stride(N);
  if (threadIdx.x == 0) {
    const uint32_t baddr = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(baddr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)(knn_ws_stride(N) * sizeof(float));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(baddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(sx)),
                 "l"(w), "r"(bytes), "r"(baddr)
                 : "memory");
  }
  __syncthreads();   // barrier initialised before anyone polls it
  new_xyz += (size_t)b * S * 3;
  PoseQT pose;
  const bool do_warp = warp_qt != nullptr;
  if (do_warp) pose = make_pose(warp_qt + (size_t)b * 7);
  slab_mbar_wait(&bar, 0);
  const int axis = hdr[0], axis2 = hdr[1], strips = hdr[2], logL = hdr[3];
  const float* sa1b = reinterpret_cast<const float*>(hdr + 4);     // strip bounds along a1 [strips + 1]
  const float* sa3lo = sa1b + KNN_MAX_STRIPS + 4;                   // per-strip range of the third axis
  const float* sa3hi = sa3lo + KNN_MAX_STRIPS;
  const int axis3 = 3 - axis - axis2;
  const float* sa = axis2 == 0 ? sx : (axis2 == 1 ? sy : sz);     // in-strip sort axis a2

  const int flush_at = min(32, max(2 * K, 8));
  // each warp walks a contiguous run of the (Morton-ordered) query list of this CTA
  const int c_begin = blockIdx.x * q_per_cta, c_end = min(S, c_begin + q_per_cta);
  const int per_warp = (c_end - c_begin + nwarps - 1) / nwarps;
  const int w_begin = c_begin + warp * per_warp, w_end = min(c_end, w_begin + per_warp);
  float pqx = 0.f, pqy = 0.f, pqz = 0.f, prev_kth = CUDART_INF_F;   // previous query of this warp
  for (int qpos = w_begin; qpos < w_end; ++qpos) {
    const int q = qorder != nullptr ? qorder[(size_t)b * S + qpos] : qpos;
    float qx = new_xyz[q * 3 + 0], qy = new_xyz[q * 3 + 1], qz = new_xyz[q * 3 + 2];
    if (do_warp) {
      warp_point(pose, qx, qy, qz, qx, qy, qz);
      if (warped_out != nullptr && lane == 0) {
        float* o = warped_out + ((size_t)b * S + q) * 3;
        o[0] = qx; o[1] = qy; o[2] = qz;
      }
    }
    // the K neighbours of the previous query lie within prev_kth + |q - q_prev| of this query, so this
    // query's K-th distance cannot be larger (triangle inequality; rounded up, + 2e-5 relative slack)
    float bound0 = CUDART_INF_F;
    if (prev_kth < CUDART_INF_F) {
      const float ex = qx - pqx, ey = qy - pqy, ez = qz - pqz;
      const float dq = __fsqrt_ru(__fmaf_ru(ez, ez, __fmaf_ru(ey, ey, __fmul_ru(ex, ex))));
      const float rr = __fadd_ru(prev_kth, dq);
      bound0 = __fmul_ru(__fmul_ru(rr, rr), 1.00002f);
    }
    pqx = qx; pqy = qy; pqz = qz;
    const float qa1 = axis == 0 ? qx : (axis == 1 ? qy : qz);
    const float qa = axis2 == 0 ? qx : (axis2 == 1 ? qy : qz);
    const float qa3 = axis3 == 0 ? qx : (axis3 == 1 ? qy : qz);
    // squared a1 distance from the query to strip t (0 inside its closed a1 range); monotone in fp32
    auto strip_d1 = [&](int t) -> float {
      const float d = fmaxf(fmaxf(__fsub_rn(sa1b[t], qa1), __fsub_rn(qa1, sa1b[t + 1])), 0.f);
      return __fmul_rn(d, d);
    };
    // ... plus the squared a3 distance to the strip's a3 range, summed downwards: a lower bound (up to the
    // association order of the three terms, covered by KNN_LB_SLACK) of d2 for every point of strip t
    auto strip_d13 = [&](int t, float e1) -> float {
      const float d = fmaxf(fmaxf(__fsub_rn(sa3lo[t], qa3), __fsub_rn(qa3, sa3hi[t])), 0.f);
      return __fadd_rd(e1, __fmul_rn(d, d));
    };
    // home strip: the last one whose first point is not beyond the query
    int t = __popc(__ballot_sync(PWCLO_FULL_MASK, lane >= 1 && lane < strips && sa1b[lane] <= qa1));
    int tl = t - 1, tr = t + 1;
    float d1 = strip_d1(t);
    float d13 = strip_d13(t, d1);
    u64 list = KNN_INF_KEY;
    float bound = bound0;
    int cnt = 0;
    while (true) {
     if (__fmul_rd(d13, KNN_LB_SLACK) <= bound) {      // otherwise the whole strip is provably too far (a3 range)
      const int s_begin = t << logL, s_end = min(N, s_begin + (1 << logL));
      // first position of the strip with sa[pos] >= qa
      int lo_b = s_begin, hi_b = s_end;
      while (lo_b < hi_b) {
        const int mid = (lo_b + hi_b) >> 1;
        if (sa[mid] < qa) lo_b = mid + 1; else hi_b = mid;
      }
      int left = lo_b - 1, right = lo_b;   // nearest unscanned positions on each side
      while (left >= s_begin || right < s_end) {
        float el = CUDART_INF_F, er = CUDART_INF_F;   // squared a2 distance of the nearest unscanned point
        if (left >= s_begin) { const float d = __fsub_rn(qa, sa[left]); el = __fmul_rn(d, d); }
        if (right < s_end) { const float d = __fsub_rn(qa, sa[right]); er = __fmul_rn(d, d); }
        const bool go_left = el <= er;
        const float e = go_left ? el : er;
        if (!(__fmul_rd(__fadd_rd(e, d13), KNN_LB_SLACK) <= bound)) break;   // provably too far (also ends on inf)
        int pos;
        bool rv;
        if (go_left) { pos = left - lane; left -= 32; rv = pos >= s_begin; }
        else { pos = right + lane; right += 32; rv = pos < s_end; }
        const int pc = rv ? pos : s_begin;
        const float dx = __fsub_rn(qx, sx[pc]), dy = __fsub_rn(qy, sy[pc]), dz = __fsub_rn(qz, sz[pc]);
        const float xx = __fmul_rn(dx, dx), yy = __fmul_rn(dy, dy), zz = __fmul_rn(dz, dz);
        const float d2 = SUM_ORDER == 0 ? __fadd_rn(__fadd_rn(xx, yy), zz) : __fadd_rn(__fadd_rn(xx, zz), yy);
        const bool pass = rv && d2 <= bound;
        const unsigned mask = __ballot_sync(PWCLO_FULL_MASK, pass);
        if (mask) {
          if (pass) {
            const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
            cand_d[slot] = d2;
            cand_i[slot] = sid[pc];
          }
          cnt += __popc(mask);
          if (cnt >= flush_at) {   // small K: merge early so that the pruning bound tightens early
            __syncwarp();
            const u64 ck = lane < cnt ? knn_key(__fsqrt_rn(__fadd_rn(cand_d[lane], 1e-8f)), cand_i[lane]) : KNN_INF_KEY;
            const int rest = max(cnt - 32, 0);
            float md = 0.f; int mi = 0;
            if (lane < rest) { md = cand_d[32 + lane]; mi = cand_i[32 + lane]; }
            __syncwarp();
            if (lane < rest) { cand_d[lane] = md; cand_i[lane] = mi; }
            cnt = rest;
            list = knn_merge32(list, ck, lane);
            bound = fminf(bound, knn_bound(__uint_as_float((unsigned)(shfl_u64(list, K - 1) >> 32))));
          }
        }
      }
     }
      // next strip: the nearer of the two unvisited neighbours along a1
      if (tl < 0 && tr >= strips) break;
      const float dl = tl >= 0 ? strip_d1(tl) : CUDART_INF_F;
      const float dr = tr < strips ? strip_d1(tr) : CUDART_INF_F;
      if (tl >= 0 && (dl <= dr || tr >= strips)) { t = tl--; d1 = dl; } else { t = tr++; d1 = dr; }
      if (!(d1 <= bound)) break;        // every remaining strip is provably too far along a1 alone
      d13 = strip_d13(t, d1);
    }
    if (cnt > 0) {
      __syncwarp();
      const u64 ck = lane < cnt ? knn_key(__fsqrt_rn(__fadd_rn(cand_d[lane], 1e-8f)), cand_i[lane]) : KNN_INF_KEY;
      list = knn_merge32(list, ck, lane);
    }
    __syncwarp();
    prev_kth = __uint_as_float((unsigned)(shfl_u64(list, K - 1) >> 32));
    if (lane < K) {
      const size_t o = ((size_t)b * S + q) * K + lane;
      idx_out[o] = (int)(unsigned)list;
      if (dist_out) dist_out[o] = __uint_as_float((unsigned)(list >> 32));
    }
  }
}


// -------------------------------------------------------------------------------------------------
// Clouds of 8193 .. 16384 points (the 16 384-point level of the training configuration): the sorted cloud does not fit
// shared memory (256 KB), its two halves -- strips [0, hs) and [hs, strips), hs = strips / 2 -- do.  A CTA stages half 0,
// half 1 and half 0 again; a query whose home strip lies in half 0 is searched in pass 0 and finished in pass 1, one whose
// home strip lies in half 1 starts in pass 1 and is finished in pass 2.  Between the passes the running K-list of every
// query of the CTA waits in shared memory.  The second visit starts with the bound of the K-th neighbour found at home,
// so for most queries it ends at the first strip.  Same distances, keys, pruning rule and merge network as
// knn_slab_kernel: bit-identical to the brute-force kernel.
// -------------------------------------------------------------------------------------------------
constexpr int KNN_MAX_N2 = 16384;
constexpr int KNN2_Q_PER_CTA = 256;

template <int SUM_ORDER>
__global__ void __launch_bounds__(1024)
knn_slab2_kernel(const float* __restrict__ ws, const float* __restrict__ new_xyz, int N, int S, int K, int q_per_cta,
                 const float* __restrict__ warp_qt, float* __restrict__ warped_out, int32_t* __restrict__ idx_out,
                 float* __restrict__ dist_out) {
  extern __shared__ __align__(128) unsigned char slab_smem[];
  constexpr int NH = KNN_MAX_N2 / 2;                 // capacity of a half
  const int N4 = (N + 3) & ~3;
  float* sx = reinterpret_cast<float*>(slab_smem);
  float* sy = sx + NH;
  float* sz = sy + NH;
  int* sid = reinterpret_cast<int*>(sz + NH);
  int* hdr = sid + NH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* cand_d = reinterpret_cast<float*>(hdr + KNN_HDR) + (size_t)warp * KNN_BUF;
  int* cand_i = reinterpret_cast<int*>(reinterpret_cast<float*>(hdr + KNN_HDR) + (size_t)nwarps * KNN_BUF) + (size_t)warp * KNN_BUF;
  u64* lists = reinterpret_cast<u64*>(reinterpret_cast<float*>(hdr + KNN_HDR) + (size_t)2 * nwarps * KNN_BUF);   // [q_per_cta][32]
  __shared__ __align__(8) uint64_t bar;

  const int b = blockIdx.y;
  const float* w = ws + (size_t)b * knn_ws_stride(N);
  // strip geometry straight from the record in global memory (needed before anything is staged)
  const int* gh = reinterpret_cast<const int*>(w) + 4 * N4;
  const int axis = gh[0], axis2 = gh[1], strips = gh[2], logL = gh[3];
  const int hs = strips >> 1;                         // first strip of half 1
  const int split = min(hs << logL, N4);              // first sorted position of half 1 (multiple of 32)
  if (split > NH || N4 - split > NH) __trap();         // cannot happen for N <= 16384 with >= 2 strips; never overrun smem
  const uint32_t baddr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(baddr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  new_xyz += (size_t)b * S * 3;
  PoseQT pose;
  const bool do_warp = warp_qt != nullptr;
  if (do_warp) pose = make_pose(warp_qt + (size_t)b * 7);
  const float* sa1b = reinterpret_cast<const float*>(hdr + 4);
  const float* sa3lo = sa1b + KNN_MAX_STRIPS + 4;
  const float* sa3hi = sa3lo + KNN_MAX_STRIPS;
  const int axis3 = 3 - axis - axis2;
  const int flush_at = min(32, max(2 * K, 8));
  const int c_begin = blockIdx.x * q_per_cta, c_end = min(S, c_begin + q_per_cta);
  const int per_warp = (c_end - c_begin + nwarps - 1) / nwarps;
  const int w_begin = c_begin + warp * per_warp, w_end = min(c_end, w_begin + per_warp);

  for (int pass = 0; pass < 3; ++pass) {
    const int half = pass & 1;
    const int base = half ? split : 0;                       // first sorted position held in shared memory
    const int count = half ? N4 - split : split;             // multiple of 4
    __syncthreads();                                         // everybody is done with the previous half
    if (threadIdx.x == 0) {
      const uint32_t arr = (uint32_t)count * 4u;
      const uint32_t bytes = 4u * arr + (pass == 0 ? (uint32_t)(KNN_HDR * sizeof(float)) : 0u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the old half before the async writes
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(baddr), "r"(bytes) : "memory");
      const float* src[4] = {w + base, w + N4 + base, w + 2 * N4 + base, w + 3 * N4 + base};
      float* dst[4] = {sx, sy, sz, reinterpret_cast<float*>(sid)};
#pragma unroll
      for (int a = 0; a < 4; ++a)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(dst[a])),
                     "l"(src[a]), "r"(arr), "r"(baddr)
                     : "memory");
      if (pass == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(hdr)),
                     "l"(w + 4 * N4), "r"((uint32_t)(KNN_HDR * sizeof(float))), "r"(baddr)
                     : "memory");
    }
    slab_mbar_wait(&bar, (uint32_t)(pass & 1));
    const float* ax = sx - base;                             // indexed by the absolute sorted position
    const float* ay = sy - base;
    const float* az = sz - base;
    const int* aid = sid - base;
    const float* sa = axis2 == 0 ? ax : (axis2 == 1 ? ay : az);
    const int lo = half ? hs : 0, hi = half ? strips : hs;   // strips of this half

    for (int qpos = w_begin; qpos < w_end; ++qpos) {
      const int q = qpos;
      float qx = new_xyz[q * 3 + 0], qy = new_xyz[q * 3 + 1], qz = new_xyz[q * 3 + 2];
      if (do_warp) warp_point(pose, qx, qy, qz, qx, qy, qz);
      const float qa1 = axis == 0 ? qx : (axis == 1 ? qy : qz);
      const float qa = axis2 == 0 ? qx : (axis2 == 1 ? qy : qz);
      const float qa3 = axis3 == 0 ? qx : (axis3 == 1 ? qy : qz);
      // home strip over the WHOLE cloud: the last one whose first point is not beyond the query
      const int home = __popc(__ballot_sync(PWCLO_FULL_MASK, lane >= 1 && lane < strips && sa1b[lane] <= qa1));
      const bool home_hi = home >= hs;
      const bool fresh = pass == 0 || (pass == 1 && home_hi);
      if ((pass == 0 && home_hi) || (pass == 2 && !home_hi)) continue;      // warp-uniform
      const bool final_pass = pass == 2 || (pass == 1 && !home_hi);
      if (fresh && do_warp && warped_out != nullptr && lane == 0) {
        float* o = warped_out + ((size_t)b * S + q) * 3;
        o[0] = qx; o[1] = qy; o[2] = qz;
      }
      auto strip_d1 = [&](int t) -> float {
        const float d = fmaxf(fmaxf(__fsub_rn(sa1b[t], qa1), __fsub_rn(qa1, sa1b[t + 1])), 0.f);
        return __fmul_rn(d, d);
      };
      auto strip_d13 = [&](int t, float e1) -> float {
        const float d = fmaxf(fmaxf(__fsub_rn(sa3lo[t], qa3), __fsub_rn(qa3, sa3hi[t])), 0.f);
        return __fadd_rd(e1, __fmul_rn(d, d));
      };
      u64* my_list = lists + (size_t)(qpos - c_begin) * 32;
      u64 list = fresh ? KNN_INF_KEY : my_list[lane];
      float bound = CUDART_INF_F;
      if (!fresh) {
        const unsigned kv = (unsigned)(shfl_u64(list, K - 1) >> 32);
        if (kv != (unsigned)(KNN_INF_KEY >> 32)) bound = knn_bound(__uint_as_float(kv));
      }
      int t = min(max(home, lo), hi - 1);                 // nearest strip of this half
      int tl = t - 1, tr = t + 1;
      float d1 = strip_d1(t);
      float d13 = strip_d13(t, d1);
      int cnt = 0;
      if (d1 <= bound) {
        while (true) {
          if (__fmul_rd(d13, KNN_LB_SLACK) <= bound) {
            const int s_begin = t << logL, s_end = min(N, s_begin + (1 << logL));
            int lo_b = s_begin, hi_b = s_end;
            while (lo_b < hi_b) {
              const int mid = (lo_b + hi_b) >> 1;
              if (sa[mid] < qa) lo_b = mid + 1; else hi_b = mid;
            }
            int left = lo_b - 1, right = lo_b;
            while (left >= s_begin || right < s_end) {
              float el = CUDART_INF_F, er = CUDART_INF_F;
              if (left >= s_begin) { const float d = __fsub_rn(qa, sa[left]); el = __fmul_rn(d, d); }
              if (right < s_end) { const float d = __fsub_rn(qa, sa[right]); er = __fmul_rn(d, d); }
              const bool go_left = el <= er;
              const float e = go_left ? el : er;
              if (!(__fmul_rd(__fadd_rd(e, d13), KNN_LB_SLACK) <= bound)) break;
              int pos;
              bool rv;
              if (go_left) { pos = left - lane; left -= 32; rv = pos >= s_begin; }
              else { pos = right + lane; right += 32; rv = pos < s_end; }
              const int pc = rv ? pos : s_begin;
              const float dx = __fsub_rn(qx, ax[pc]), dy = __fsub_rn(qy, ay[pc]), dz = __fsub_rn(qz, az[pc]);
              const float xx = __fmul_rn(dx, dx), yy = __fmul_rn(dy, dy), zz = __fmul_rn(dz, dz);
              const float d2 = SUM_ORDER == 0 ? __fadd_rn(__fadd_rn(xx, yy), zz) : __fadd_rn(__fadd_rn(xx, zz), yy);
              const bool pass_c = rv && d2 <= bound;
              const unsigned mask = __ballot_sync(PWCLO_FULL_MASK, pass_c);
              if (mask) {
                if (pass_c) {
                  const int slot = cnt + __popc(mask & ((1u << lane) - 1u));
                  cand_d[slot] = d2;
                  cand_i[slot] = aid[pc];
                }
                cnt += __popc(mask);
                if (cnt >= flush_at) {
                  __syncwarp();
                  const u64 ck = lane < cnt ? knn_key(__fsqrt_rn(__fadd_rn(cand_d[lane], 1e-8f)), cand_i[lane]) : KNN_INF_KEY;
                  const int rest = max(cnt - 32, 0);
                  float md = 0.f; int mi = 0;
                  if (lane < rest) { md = cand_d[32 + lane]; mi = cand_i[32 + lane]; }
                  __syncwarp();
                  if (lane < rest) { cand_d[lane] = md; cand_i[lane] = mi; }
                  cnt = rest;
                  list = knn_merge32(list, ck, lane);
                  const unsigned kv = (unsigned)(shfl_u64(list, K - 1) >> 32);
                  if (kv != (unsigned)(KNN_INF_KEY >> 32)) bound = fminf(bound, knn_bound(__uint_as_float(kv)));
                }
              }
            }
          }
          if (tl < lo && tr >= hi) break;
          const float dl = tl >= lo ? strip_d1(tl) : CUDART_INF_F;
          const float dr = tr < hi ? strip_d1(tr) : CUDART_INF_F;
          if (tl >= lo && (dl <= dr || tr >= hi)) { t = tl--; d1 = dl; } else { t = tr++; d1 = dr; }
          if (!(d1 <= bound)) break;
          d13 = strip_d13(t, d1);
        }
      }
      if (cnt > 0) {
        __syncwarp();
        const u64 ck = lane < cnt ? knn_key(__fsqrt_rn(__fadd_rn(cand_d[lane], 1e-8f)), cand_i[lane]) : KNN_INF_KEY;
        list = knn_merge32(list, ck, lane);
      }
      __syncwarp();
      if (final_pass) {
        if (lane < K) {
          const size_t o = ((size_t)b * S + q) * K + lane;
          idx_out[o] = (int)(unsigned)list;
          if (dist_out) dist_out[o] = __uint_as_float((unsigned)(list >> 32));
        }
      } else {
        my_list[lane] = list;
      }
    }
  }
}


// -------------------------------------------------------------------------------------------------
// Small-K variant of the slab search: G = 8 or 16 lanes per query, 32/G queries per warp in lock step.
// The running list, the candidate queue and the bitonic sort/merge networks are G wide, so a merge
// costs ~G/32 of the 32-lane version and four (two) queries share every instruction.  Same keys,
// same pruning rule, same result bits.
// -------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ u64 knn_merge_g(u64 list, u64 cand, int gl) {
#pragma unroll
  for (int k = 2; k <= G; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const u64 other = shfl_xor_u64(cand, j);
      const bool desc_block = (gl & k) == 0 || k == G;
      const bool lower = (gl & j) == 0;
      const bool keep_max = lower == desc_block;
      cand = keep_max ? (cand > other ? cand : other) : (cand < other ? cand : other);
    }
  }
  u64 c = list < cand ? list : cand;
#pragma unroll
  for (int j = G / 2; j > 0; j >>= 1) {
    const u64 other = shfl_xor_u64(c, j);
    const bool lower = (gl & j) == 0;
    c = lower ? (c < other ? c : other) : (c > other ? c : other);
  }
  return c;
}

template <int SUM_ORDER, int G>
__global__ void __launch_bounds__(SLAB_WARPS * 32)
knn_slab_small_kernel(const float* __restrict__ ws, const float* __restrict__ new_xyz, int N, int S, int K, int q_per_cta,
                      const float* __restrict__ warp_qt, float* __restrict__ warped_out, int32_t* __restrict__ idx_out,
                      float* __restrict__ dist_out) {
  constexpr int QPW = 32 / G;
  extern __shared__ __align__(128) unsigned char slab_smem[];
  const int N4 = (N + 3) & ~3;
  float* sx = reinterpret_cast<float*>(slab_smem);
  float* sy = sx + N4;
  float* sz = sy + N4;
  int* sid = reinterpret_cast<int*>(sz + N4);
  int* hdr = sid + N4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / G, gl = lane % G;
  const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (grp * G);
  float* cand_d = reinterpret_cast<float*>(hdr + KNN_HDR) + (size_t)warp * KNN_BUF + grp * 2 * G;
  int* cand_i = reinterpret_cast<int*>(reinterpret_cast<float*>(hdr + KNN_HDR) + (size_t)SLAB_WARPS * KNN_BUF) + (size_t)warp * KNN_BUF + grp * 2 * G;
  __shared__ __align__(8) uint64_t bar;

  const int b = blockIdx.y;
  const float* w = ws + (size_t)b * knn_ws_stride(N);
  if (threadIdx.x == 0) {
    const uint32_t baddr = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(baddr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)(knn_ws_stride(N) * sizeof(float));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(baddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(sx)),
                 "l"(w), "r"(bytes), "r"(baddr)
                 : "memory");
  }
  __syncthreads();
  new_xyz += (size_t)b * S * 3;
  PoseQT pose;
  const bool do_warp = warp_qt != nullptr;
  if (do_warp) pose = make_pose(warp_qt + (size_t)b * 7);
  slab_mbar_wait(&bar, 0);
  const int axis = hdr[0], axis2 = hdr[1], strips = hdr[2], logL = hdr[3];
  const float* sa1b = reinterpret_cast<const float*>(hdr + 4);     // strip bounds along a1 [strips + 1]
  const float* sa3lo = sa1b + KNN_MAX_STRIPS + 4;                   // per-strip range of the third axis
  const float* sa3hi = sa3lo + KNN_MAX_STRIPS;
  const int axis3 = 3 - axis - axis2;
  const float* sa = axis2 == 0 ? sx : (axis2 == 1 ? sy : sz);     // in-strip sort axis a2
  const int nsearch = logL + 1;          // binary-search steps inside a strip: enough for 2^logL + 1 outcomes

  const int flush_at = min(G, max(2 * K, 8));
  const int c_begin = blockIdx.x * q_per_cta, c_end = min(S, c_begin + q_per_cta);
  for (int q0 = c_begin + warp * QPW; q0 < c_end; q0 += SLAB_WARPS * QPW) {
    const int q = q0 + grp;
    const bool has_q = q < c_end;
    const int qq = has_q ? q : c_end - 1;
    float qx = new_xyz[qq * 3 + 0], qy = new_xyz[qq * 3 + 1], qz = new_xyz[qq * 3 + 2];
    if (do_warp) {
      warp_point(pose, qx, qy, qz, qx, qy, qz);
      if (warped_out != nullptr && gl == 0 && has_q) {
        float* o = warped_out + ((size_t)b * S + q) * 3;
        o[0] = qx; o[1] = qy; o[2] = qz;
      }
    }
    const float qa1 = axis == 0 ? qx : (axis == 1 ? qy : qz);
    const float qa = axis2 == 0 ? qx : (axis2 == 1 ? qy : qz);
    const float qa3 = axis3 == 0 ? qx : (axis3 == 1 ? qy : qz);
    auto strip_d1 = [&](int t) -> float {
      const float d = fmaxf(fmaxf(__fsub_rn(sa1b[t], qa1), __fsub_rn(qa1, sa1b[t + 1])), 0.f);
      return __fmul_rn(d, d);
    };
    // ... plus the squared a3 distance to the strip's a3 range, summed downwards: a lower bound (up to the
    // association order of the three terms, covered by KNN_LB_SLACK) of d2 for every point of strip t
    auto strip_d13 = [&](int t, float e1) -> float {
      const float d = fmaxf(fmaxf(__fsub_rn(sa3lo[t], qa3), __fsub_rn(qa3, sa3hi[t])), 0.f);
      return __fadd_rd(e1, __fmul_rn(d, d));
    };
    // per-group state (uniform inside a group): current strip t with a1 distance d1, scan cursors, next strips
    int t = __popc(__ballot_sync(PWCLO_FULL_MASK, gl >= 1 && gl < strips && sa1b[gl] <= qa1) & gmask);   // home strip
    if (G < KNN_MAX_STRIPS)
      t += __popc(__ballot_sync(PWCLO_FULL_MASK, gl + G < strips && sa1b[gl + G] <= qa1) & gmask);
    if (2 * G < KNN_MAX_STRIPS) {
      t += __popc(__ballot_sync(PWCLO_FULL_MASK, gl + 2 * G < strips && sa1b[gl + 2 * G] <= qa1) & gmask);
      t += __popc(__ballot_sync(PWCLO_FULL_MASK, gl + 3 * G < strips && sa1b[gl + 3 * G] <= qa1) & gmask);
    }
    int tl = t - 1, tr = t + 1;
    float d1 = strip_d1(t);
    float d13 = strip_d13(t, d1);
    int s_begin, s_end, left, right;
    bool done = !has_q;
    // enter strip t: binary search of the query's a2 position (no warp collectives inside: may diverge)
    auto enter = [&]() {
      s_begin = t << logL;
      s_end = min(N, s_begin + (1 << logL));
      int lo_b = s_begin, hi_b = s_end;
      while (lo_b < hi_b) {
        const int mid = (lo_b + hi_b) >> 1;
        if (sa[mid] < qa) lo_b = mid + 1; else hi_b = mid;
      }
      left = lo_b - 1; right = lo_b;
    };
    enter();
    u64 list = KNN_INF_KEY;
    float bound = CUDART_INF_F;
    int cnt = 0;
    while (true) {
      float e;
      bool go_left, active;
      // nearest unscanned point of the current strip and whether it can still matter
      auto probe = [&]() {
        float el = CUDART_INF_F, er = CUDART_INF_F;
        if (left >= s_begin) { const float d = __fsub_rn(qa, sa[left]); el = __fmul_rn(d, d); }
        if (right < s_end) { const float d = __fsub_rn(qa, sa[right]); er = __fmul_rn(d, d); }
        go_left = el <= er;
        e = go_left ? el : er;
        active = (left >= s_begin || right < s_end) && __fmul_rd(__fadd_rd(e, d13), KNN_LB_SLACK) <= bound;   // uniform within a group
      };
      probe();
      if (!done && !active) {
        // strip finished: move to the nearer unvisited neighbour strip along a1, or stop
        if (tl < 0 && tr >= strips) {
          done = true;
        } else {
          const float dl = tl >= 0 ? strip_d1(tl) : CUDART_INF_F;
          const float dr = tr < strips ? strip_d1(tr) : CUDART_INF_F;
          if (tl >= 0 && (dl <= dr || tr >= strips)) { t = tl--; d1 = dl; } else { t = tr++; d1 = dr; }
          if (d1 <= bound) {
            d13 = strip_d13(t, d1);
            // a strip that is too far by its a3 range alone is left at once: empty cursors, next round advances
            if (__fmul_rd(d13, KNN_LB_SLACK) <= bound) { enter(); probe(); } else { left = -1; right = 0; s_begin = 0; s_end = 0; active = false; }
          } else {
            done = true;
          }
        }
      }
      active = active && !done;
      if (!__any_sync(PWCLO_FULL_MASK, !done)) break;
      int pos = 0;
      bool rv = false;
      if (active) {
        if (go_left) { pos = left - gl; left -= G; rv = pos >= s_begin; }
        else { pos = right + gl; right += G; rv = pos < s_end; }
      }
      const int pc = rv ? pos : 0;
      const float dx = __fsub_rn(qx, sx[pc]), dy = __fsub_rn(qy, sy[pc]), dz = __fsub_rn(qz, sz[pc]);
      const float xx = __fmul_rn(dx, dx), yy = __fmul_rn(dy, dy), zz = __fmul_rn(dz, dz);
      const float d2 = SUM_ORDER == 0 ? __fadd_rn(__fadd_rn(xx, yy), zz) : __fadd_rn(__fadd_rn(xx, zz), yy);
      const bool pass = rv && d2 <= bound;
      const unsigned mask = __ballot_sync(PWCLO_FULL_MASK, pass);
      if (mask) {   // warp-uniform
        const unsigned gm = mask & gmask;
        if (pass) {
          const int slot = cnt + __popc(gm & ((1u << lane) - 1u));
          cand_d[slot] = d2;
          cand_i[slot] = sid[pc];
        }
        cnt += __popc(gm);
        const bool flush = cnt >= flush_at;
        if (__any_sync(PWCLO_FULL_MASK, flush)) {
          __syncwarp();
          const u64 ck = flush && gl < cnt ? knn_key(__fsqrt_rn(__fadd_rn(cand_d[gl], 1e-8f)), cand_i[gl]) : KNN_INF_KEY;
          const int rest = flush ? max(cnt - G, 0) : 0;
          float md = 0.f; int mi = 0;
          if (gl < rest) { md = cand_d[G + gl]; mi = cand_i[G + gl]; }
          __syncwarp();
          if (gl < rest) { cand_d[gl] = md; cand_i[gl] = mi; }
          if (flush) cnt = rest;
          list = knn_merge_g<G>(list, ck, gl);      // groups that do not flush merge an all-INF batch: a no-op
          // (shuffle executed by every lane: `flush` differs between the groups of a warp)
          const unsigned kv = (unsigned)(__shfl_sync(PWCLO_FULL_MASK, (unsigned)(list >> 32), K - 1, G));
          if (flush) bound = fminf(bound, knn_bound(__uint_as_float(kv)));
        }
      }
    }
    __syncwarp();
    {
      const u64 ck = gl < cnt ? knn_key(__fsqrt_rn(__fadd_rn(cand_d[gl], 1e-8f)), cand_i[gl]) : KNN_INF_KEY;
      list = knn_merge_g<G>(list, ck, gl);
    }
    __syncwarp();
    if (has_q && gl < K) {
      const size_t o = ((size_t)b * S + q) * K + gl;
      idx_out[o] = (int)(unsigned)list;
      if (dist_out) dist_out[o] = __uint_as_float((unsigned)(list >> 32));
    }
  }
}

}  // namespace pwclo

PWCLO_API size_t pwclo_knn_workspace_bytes(int B, int N, int S) {
  if (B <= 0 || N <= 0 || S < 0 || N > KNN_MAX_N2) return 0;
  return (size_t)B * knn_ws_stride(N) * sizeof(float) + (size_t)B * S * sizeof(int);
}

// presort (qorder != nullptr: also the Morton visiting order of the queries) and search, shared by the entry points
static int knn_presort_launch(const float* xyz, int B, int N, int K, float* workspace, const float* queries, int S, int* qorder,
                              cudaStream_t st) {
  int NP = 1, SP = 1;
  while (NP < N) NP <<= 1;
  while (SP < S) SP <<= 1;
  const size_t smem = max((size_t)NP * sizeof(u64), qorder ? (size_t)SP * sizeof(unsigned) : (size_t)0);
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(knn_presort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const char* se = getenv("PWCLO_KNN_STRIPS");
  knn_presort_kernel<<<B, SORT_THREADS, smem, st>>>(xyz, N, NP, knn_log_strip(NP, K, se ? atoi(se) : 0), workspace, queries, S, SP,
                                                    qorder);
  return launch_status();
}

static int knn_search_launch(const float* workspace, const int* qorder, const float* new_xyz, int B, int N, int S, int K,
                             int sum_order, const float* warp_qt, float* warped_out, int32_t* idx, float* dist, cudaStream_t st) {
  if (N > KNN_MAX_TILE) {      // 8193 .. 16384 points: the two halves of the sorted cloud take turns in shared memory
    const int warps = 32, q_per_cta = KNN2_Q_PER_CTA;
    const size_t smem2 = (size_t)4 * (KNN_MAX_N2 / 2) * sizeof(float) + KNN_HDR * sizeof(float) + (size_t)warps * KNN_BUF * 8 +
                         (size_t)q_per_cta * 32 * sizeof(u64) + 128;
    auto k2 = sum_order == PWCLO_KNN_SUM_XY_Z ? knn_slab2_kernel<0> : knn_slab2_kernel<1>;
    cudaError_t e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return (int)e;
    k2<<<dim3(ceil_div(S, q_per_cta), B), warps * 32, smem2, st>>>(workspace, new_xyz, N, S, K, q_per_cta, warp_qt, warped_out, idx,
                                                                    dist);
    return launch_status();
  }
  const size_t smem = knn_ws_stride(N) * sizeof(float) + (size_t)SLAB_WARPS * KNN_BUF * 8 + 128;
  // queries per CTA: amortise the shared-memory fill, keep >= ~3 CTAs per SM in flight overall
  int q_per_cta = SLAB_WARPS;
  while (q_per_cta * 2 <= S && (long long)B * ceil_div(S, q_per_cta * 2) >= 3LL * kNumSM) q_per_cta *= 2;
  dim3 grid(ceil_div(S, q_per_cta), B);
  if (K <= 16 && N >= 16 && !getenv("PWCLO_KNN_NO_SMALL")) {   // G lanes per query, 32/G queries per warp
    void (*ks)(const float*, const float*, int, int, int, int, const float*, float*, int32_t*, float*);
    if (K <= 8) ks = sum_order == PWCLO_KNN_SUM_XY_Z ? knn_slab_small_kernel<0, 8> : knn_slab_small_kernel<1, 8>;
    else ks = sum_order == PWCLO_KNN_SUM_XY_Z ? knn_slab_small_kernel<0, 16> : knn_slab_small_kernel<1, 16>;
    if (smem > 32 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    ks<<<grid, SLAB_WARPS * 32, smem, st>>>(workspace, new_xyz, N, S, K, q_per_cta, warp_qt, warped_out, idx, dist);
    return launch_status();
  }
  auto kern = sum_order == PWCLO_KNN_SUM_XY_Z ? knn_slab_kernel<0> : knn_slab_kernel<1>;
  // a cloud that leaves room for only one CTA per SM gets 32 warps per CTA (the search is latency bound)
  const int warps = smem > 100 * 1024 && !getenv("PWCLO_KNN_16W") ? 32 : SLAB_WARPS;
  const size_t smem_k = knn_ws_stride(N) * sizeof(float) + (size_t)warps * KNN_BUF * 8 + 128;
  if (warps != SLAB_WARPS) {
    q_per_cta = warps;
    while (q_per_cta * 2 <= S && (long long)B * ceil_div(S, q_per_cta * 2) >= 3LL * kNumSM) q_per_cta *= 2;
    grid = dim3(ceil_div(S, q_per_cta), B);
  }
  if (smem_k > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<grid, warps * 32, smem_k, st>>>(workspace, qorder, new_xyz, N, S, K, q_per_cta, warp_qt, warped_out, idx, dist);
  return launch_status();
}

static int knn_check(const void* a, const void* b, const void* c, int B, int N, int S, int K, int sum_order) {
  if (!a || !b || !c || B < 0 || N <= 0 || S < 0 || K <= 0) return PWCLO_EINVAL;
  if (K > 32) return PWCLO_EUNSUPPORTED;
  if (K > N) return PWCLO_EINVAL;
  if (sum_order != PWCLO_KNN_SUM_XY_Z && sum_order != PWCLO_KNN_SUM_XZ_Y) return PWCLO_EINVAL;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  return PWCLO_OK;
}

PWCLO_API int pwclo_knn_sorted(const float* xyz, const float* new_xyz, int B, int N, int S, int K, int sum_order,
                               const float* warp_qt, float* warped_out, int32_t* idx, float* dist, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (int rc = knn_check(xyz, new_xyz, idx, B, N, S, K, sum_order)) return rc;
  if (B == 0 || S == 0) return PWCLO_OK;
  if (N > KNN_MAX_N2 || !workspace || workspace_bytes < pwclo_knn_workspace_bytes(B, N, S) || (uintptr_t)workspace % 16 != 0)
    return pwclo_knn(xyz, new_xyz, B, N, S, K, sum_order, warp_qt, warped_out, idx, dist, stream);
  cudaStream_t st = (cudaStream_t)stream;
  int* qorder = S <= 8192 && N <= KNN_MAX_TILE && getenv("PWCLO_KNN_QORDER") ? reinterpret_cast<int*>((float*)workspace + (size_t)B * knn_ws_stride(N)) : nullptr;
  if (int rc = knn_presort_launch(xyz, B, N, K, (float*)workspace, new_xyz, S, qorder, st)) return rc;
  return knn_search_launch((const float*)workspace, qorder, new_xyz, B, N, S, K, sum_order, warp_qt, warped_out, idx, dist, st);
}

PWCLO_API int pwclo_knn_presort(const float* xyz, int B, int N, int K, void* workspace, size_t workspace_bytes, void* stream) {
  if (!xyz || !workspace || B < 0 || N <= 0 || K <= 0 || K > 32) return PWCLO_EINVAL;
  if (N > KNN_MAX_N2 || B > 65535) return PWCLO_EUNSUPPORTED;
  if (workspace_bytes < pwclo_knn_workspace_bytes(B, N, 0) || (uintptr_t)workspace % 16 != 0) return PWCLO_EINVAL;
  if (B == 0) return PWCLO_OK;
  return knn_presort_launch(xyz, B, N, K, (float*)workspace, nullptr, 0, nullptr, (cudaStream_t)stream);
}

PWCLO_API int pwclo_knn_search(const void* workspace, const float* new_xyz, int B, int N, int S, int K, int sum_order,
                               const float* warp_qt, float* warped_out, int32_t* idx, float* dist, void* stream) {
  if (int rc = knn_check(workspace, new_xyz, idx, B, N, S, K, sum_order)) return rc;
  if (N > KNN_MAX_N2) return PWCLO_EUNSUPPORTED;
  if ((uintptr_t)workspace % 16 != 0) return PWCLO_EINVAL;
  if (B == 0 || S == 0) return PWCLO_OK;
  return knn_search_launch((const float*)workspace, nullptr, new_xyz, B, N, S, K, sum_order, warp_qt, warped_out, idx, dist,
                           (cudaStream_t)stream);
}
