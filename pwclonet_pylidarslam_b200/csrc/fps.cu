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
#include <cooperative_groups.h>

#include <cstdlib>

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
// Prefix property: when the cloud handed to FPS is itself the FPS-ordered output of a run that never saw
// two candidates tie for the arg-max (`tie_in[b] == 0`), the first m samples of this run are 0, 1, ..., m-1 -- the
// running minima are the same floats round for round, so the same points win.  Every kernel below starts with this
// early-out and, when asked (`tie_out`), reports whether its own run saw a tie (conservatively: a set flag may be a
// false alarm, a clear flag is exact).
__device__ __forceinline__ bool fps_prefix_shortcut(const int32_t* __restrict__ tie_in, int32_t* __restrict__ tie_out, int b,
                                                    int m, int32_t* __restrict__ idx, int tid, int threads, bool writer) {
  if (tie_in == nullptr || tie_in[b] != 0) return false;
  if (writer) {
    for (int r = tid; r < m; r += threads) idx[r] = r;
    if (tie_out != nullptr && tid == 0) tie_out[b] = 0;
  }
  return true;
}

template <int P, int THREADS, int MODE, bool TIE, int PICK>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float* __restrict__ xyz, int n, int m, int logT, int origin_skip, float* __restrict__ scratch,
           int32_t* __restrict__ idx, const int32_t* __restrict__ tie_in, int32_t* __restrict__ tie_out) {
  extern __shared__ float smem[];
  constexpr int NW = THREADS / 32;
  float* sx = smem;
  float* sy = sx + (MODE == 2 ? 0 : n);
  float* sz = sy + (MODE == 2 ? 0 : n);
  __shared__ unsigned red_val[2][32];
  __shared__ unsigned red_key[2][32];
  __shared__ unsigned red_tie[2][32];

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)b * n * 3;
  idx += (size_t)b * m;
  if (fps_prefix_shortcut(tie_in, tie_out, b, m, idx, tid, THREADS, true)) return;
  bool tie_acc = MODE == 2;      // the global-memory fallback does not track ties: always "maybe"
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

  if (PICK == 2 && MODE == 0) {          // double pick, see fps_slab_kernel
    __shared__ unsigned red2_val[2][2][32], red2_key[2][2][32];
    int sa = 0, sb = 0;
    bool have_b = false;
    for (int r = 1, it = 0; r < m; ++it) {
      const float xa = sx[sa], ya = sy[sa], za = sz[sa];
      const float xb = sx[sb], yb = sy[sb], zb = sz[sb];
      float best = -1.0f;
      int bj = 0;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        float d2 = fminf(dist2_ref_fma(px[j] - xa, py[j] - ya, pz[j] - za), mind[j]);
        if (have_b) d2 = fminf(d2, dist2_ref_fma(px[j] - xb, py[j] - yb, pz[j] - zb));
        mind[j] = d2;
        if (d2 > best) { best = d2; bj = j; }
      }
      const unsigned vb = best < 0.f ? 0u : __float_as_uint(best) + 1u;
      const unsigned key = fps_key((unsigned)(tid + bj * THREADS), logT);
      const unsigned wv1 = __reduce_max_sync(PWCLO_FULL_MASK, vb);
      const unsigned wk1 = __reduce_min_sync(PWCLO_FULL_MASK, vb == wv1 ? key : 0xffffffffu);
      const bool owner = wv1 != 0u && vb == wv1 && key == wk1;
      float sbest = -1.0f;
      int sj = 0;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const float v = (owner && j == bj) ? -2.0f : mind[j];
        if (v > sbest) { sbest = v; sj = j; }
      }
      const unsigned vs = sbest < 0.f ? 0u : __float_as_uint(sbest) + 1u;
      const unsigned skey = fps_key((unsigned)(tid + sj * THREADS), logT);
      const unsigned wv2 = __reduce_max_sync(PWCLO_FULL_MASK, vs);
      const unsigned wk2 = __reduce_min_sync(PWCLO_FULL_MASK, (vs == wv2 && wv2 != 0u) ? skey : 0xffffffffu);
      const int buf = it & 1;
      if (lane == 0) {
        red2_val[buf][0][warp] = wv1; red2_key[buf][0][warp] = wk1;
        red2_val[buf][1][warp] = wv2; red2_key[buf][1][warp] = wk2;
      }
      __syncthreads();
      const unsigned v1 = lane < NW ? red2_val[buf][0][lane] : 0u, k1 = lane < NW ? red2_key[buf][0][lane] : 0xffffffffu;
      const unsigned v2 = lane < NW ? red2_val[buf][1][lane] : 0u, k2 = lane < NW ? red2_key[buf][1][lane] : 0xffffffffu;
      const unsigned bv1 = __reduce_max_sync(PWCLO_FULL_MASK, v1);
      const unsigned bk1 = __reduce_min_sync(PWCLO_FULL_MASK, v1 == bv1 ? k1 : 0xffffffffu);
      const bool win = bv1 != 0u && v1 == bv1 && k1 == bk1;
      const unsigned cv2 = win ? v2 : v1, ck2 = win ? k2 : k1;
      const unsigned bv2 = __reduce_max_sync(PWCLO_FULL_MASK, cv2);
      const unsigned bk2 = __reduce_min_sync(PWCLO_FULL_MASK, cv2 == bv2 ? ck2 : 0xffffffffu);
      sa = bv1 == 0u ? 0 : (int)fps_key_to_index(bk1, logT);
      if (tid == 0) idx[r] = sa;
      have_b = false;
      sb = sa;
      if (bv1 != 0u && bv2 > 1u && r + 1 < m) {
        const int cand = (int)fps_key_to_index(bk2, logT);
        const float d = dist2_ref_fma(sx[cand] - sx[sa], sy[cand] - sy[sa], sz[cand] - sz[sa]);
        if (!(d < __uint_as_float(bv2 - 1u))) {
          have_b = true;
          sb = cand;
          if (tid == 0) idx[r + 1] = cand;
        }
      }
      r += have_b ? 2 : 1;
    }
    return;
  }

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
    if (TIE && MODE != 2) {
      bool lt = false;
      if (vb == wv && wv != 0u) {
        int c = 0;
#pragma unroll
        for (int j = 0; j < P; ++j) c += (mind[j] == best) ? 1 : 0;
        lt = c > 1;
      }
      const bool wt = __popc(__ballot_sync(PWCLO_FULL_MASK, vb == wv)) > 1 || __any_sync(PWCLO_FULL_MASK, lt);
      if (lane == 0) red_tie[buf][warp] = wt ? 1u : 0u;
    }
    if (lane == 0) { red_val[buf][warp] = wv; red_key[buf][warp] = wk; }
    __syncthreads();
    unsigned v2 = lane < NW ? red_val[buf][lane] : 0u;
    unsigned k2 = lane < NW ? red_key[buf][lane] : 0xffffffffu;
    unsigned bv = __reduce_max_sync(PWCLO_FULL_MASK, v2);
    unsigned bk = __reduce_min_sync(PWCLO_FULL_MASK, v2 == bv ? k2 : 0xffffffffu);
    if (TIE && MODE != 2) {
      const unsigned t2 = lane < NW ? red_tie[buf][lane] : 0u;
      const unsigned eq = __ballot_sync(PWCLO_FULL_MASK, lane < NW && v2 == bv);
      tie_acc |= bv == 0u || __popc(eq) > 1 || __any_sync(PWCLO_FULL_MASK, lane < NW && v2 == bv && t2 != 0u);
    }
    old = bv == 0u ? 0 : (int)fps_key_to_index(bk, logT);
    if (tid == 0) idx[r] = old;
  }
  if (TIE && tid == 0 && tie_out != nullptr) tie_out[b] = tie_acc ? 1 : 0;
}


// -------------------------------------------------------------------------------------------------
// Slab-skipping variant (same result, bit for bit).  The points of a cloud are re-dealt so that warp w
// owns a contiguous slab along the cloud's widest axis.  A warp whose slab is farther from the point
// selected last than its current largest running minimum cannot change any of its minima
// (d >= fl(dx^2) >= fl(dmin^2) > max mind), so it skips the whole update and re-submits its cached
// (value, key).  After the first ~100 rounds only the 2-4 slabs around the new sample do any work.
// Inside a slab points are dealt to (lane, j) in ascending tie-key order, which keeps the per-thread
// "first maximum wins" scan equivalent to the reference's tie order.
// -------------------------------------------------------------------------------------------------
typedef unsigned long long fu64;

__device__ __forceinline__ unsigned fps_ordered_bits(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int NP, int THREADS>
__device__ __forceinline__ void fps_bitonic_sort(fu64* keys, int tid) {
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int t = tid; t < NP / 2; t += THREADS) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const fu64 a = keys[lo], c = keys[hi];
        const bool up = (lo & k) == 0;
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
}

template <int NP, int THREADS>
__device__ __forceinline__ void fps_bitonic_sort32(unsigned* keys, int tid) {
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int t = tid; t < NP / 2; t += THREADS) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const unsigned a = keys[lo], c = keys[hi];
        const bool up = (lo & k) == 0;
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
}

// PICK = 2 ("double pick", exact): a round also finds the SECOND candidate b in the reference's order.  If selecting the
// winner a cannot lower b's running minimum (d(b, a) >= mind[b], the very comparison the next round's update would
// make) and mind[b] > 0, then after a's update every other point still ranks behind b (minima only decrease, a itself
// drops to 0), so b IS the next sample: both indices are written now and the next round applies both updates in one
// pass.  The sampling is a 2047-deep dependent chain bound by the barrier + reduction latency of a round, not by the
// distance updates.  Bit-identical output -- but measured slower on LiDAR clouds (see launch_fps_slab), so it is opt-in.
template <int P, int THREADS, bool TIE, int PICK>
__global__ void __launch_bounds__(THREADS, 1)
fps_slab_kernel(const float* __restrict__ xyz, int n, int m, int logT, int origin_skip, int32_t* __restrict__ idx, int dbg,
                const int32_t* __restrict__ tie_in, int32_t* __restrict__ tie_out) {
  constexpr int NP = P * THREADS;
  constexpr int NW = THREADS / 32;
  extern __shared__ __align__(16) unsigned char fps_smem[];
  unsigned* keys = reinterpret_cast<unsigned*>(fps_smem);         // [NP] (coarse axis coordinate, index) sort keys
  int* sidx = reinterpret_cast<int*>(keys + NP);                  // [NP] sorted slot -> original index
  float* sx = reinterpret_cast<float*>(sidx + NP);                // [n] original order
  float* sy = sx + n;
  float* sz = sy + n;
  __shared__ unsigned red_val[2][32];
  __shared__ unsigned red_key[2][32];
  __shared__ float ext_mn[3][32], ext_mx[3][32];
  __shared__ unsigned red_tie[2][32];
  __shared__ int axis_s;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)b * n * 3;
  idx += (size_t)b * m;
  if (fps_prefix_shortcut(tie_in, tie_out, b, m, idx, tid, THREADS, true)) return;
  bool tie_acc = false;
  unsigned wtie = 0u;                   // cached with (wv, wk): the warp's maximum is attained by more than one point

  // ---- stage the cloud, pick the widest axis
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int k = tid; k < n; k += THREADS) {
    const float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
    sx[k] = x; sy[k] = y; sz[k] = z;
    mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
    mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
    mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int off = 16; off; off >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(PWCLO_FULL_MASK, mn[d], off));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(PWCLO_FULL_MASK, mx[d], off));
    }
    if (lane == 0) { ext_mn[d][warp] = mn[d]; ext_mx[d][warp] = mx[d]; }
  }
  __syncthreads();
  if (tid == 0) {
    float ext[3];
    for (int d = 0; d < 3; ++d) {
      float a = 3.4e38f, c = -3.4e38f;
      for (int w = 0; w < NW; ++w) { a = fminf(a, ext_mn[d][w]); c = fmaxf(c, ext_mx[d][w]); }
      ext[d] = c - a;
    }
    axis_s = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[2] >= ext[1] ? 2 : 1);
    idx[0] = 0;
  }
  __syncthreads();
  const int axis = axis_s;
  const float* sa = axis == 0 ? sx : (axis == 1 ? sy : sz);

  // ---- sort 1 along the axis: key = (top 19 bits of the order-preserving coordinate, 13-bit index); the
  // truncation only coarsens the slab assignment.  sort 2 by (slab, compact tie key): inside a slab the
  // points are ordered by the reference's tie key, so that dealing them lane-major gives every thread
  // its points in ascending tie-key order (the strict '>' scan then keeps the reference's tie order).
  static_assert(NP <= 8192, "13-bit index field");
  for (int i = tid; i < NP; i += THREADS)
    keys[i] = i < n ? ((fps_ordered_bits(sa[i]) & 0xffffe000u) | (unsigned)i) : 0xffffffffu;
  __syncthreads();
  fps_bitonic_sort32<NP, THREADS>(keys, tid);
  {
    unsigned k2[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int i = tid + j * THREADS;
      const unsigned k1 = keys[i];
      const unsigned slab = (unsigned)(i / (32 * P));
      // compact 13-bit tie key: bit-reversed (k mod T) above (k div T)
      const unsigned k = k1 & 0x1fffu;
      const unsigned ck = logT == 0 ? k : (((__brev(k & ((1u << logT) - 1u)) >> (32 - logT)) << (13 - logT)) | (k >> logT));
      k2[j] = k1 == 0xffffffffu ? ((slab << 13) | 0x1fffu) | 0x80000000u : ((slab << 13) | ck);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < P; ++j) keys[tid + j * THREADS] = k2[j];
    __syncthreads();
  }
  fps_bitonic_sort32<NP, THREADS>(keys, tid);

  // ---- deal the points: slot = warp*32P + j*32 + lane
  float px[P], py[P], pz[P], mind[P];
  float slab_lo = 3.4e38f, slab_hi = -3.4e38f;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int slot = warp * 32 * P + j * 32 + lane;
    const unsigned kk = keys[slot];
    const bool exists = (kk & 0x80000000u) == 0u;
    const unsigned ck = kk & 0x1fffu;
    const int k = !exists ? 0
                  : (logT == 0 ? (int)ck
                               : (int)((__brev(ck >> (13 - logT)) >> (32 - logT)) | ((ck & ((1u << (13 - logT)) - 1u)) << logT)));
    sidx[slot] = k;
    const float x = sx[k], y = sy[k], z = sz[k];
    bool valid = exists;
    if (exists) {
      const float c = axis == 0 ? x : (axis == 1 ? y : z);
      if (origin_skip) {
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        valid = !((double)mag <= 1e-3);
      }
      if (valid) { slab_lo = fminf(slab_lo, c); slab_hi = fmaxf(slab_hi, c); }
    }
    px[j] = x; py[j] = y; pz[j] = z;
    mind[j] = valid ? 1e10f : -1.0f;
  }
  for (int off = 16; off; off >>= 1) {
    slab_lo = fminf(slab_lo, __shfl_xor_sync(PWCLO_FULL_MASK, slab_lo, off));
    slab_hi = fmaxf(slab_hi, __shfl_xor_sync(PWCLO_FULL_MASK, slab_hi, off));
  }
  __syncthreads();

  if (PICK == 2) {
    __shared__ unsigned red2_val[2][2][32], red2_key[2][2][32];      // [buffer][first / second of the warp][warp]
    unsigned wv1 = 0u, wk1 = 0xffffffffu, wv2 = 0u, wk2 = 0xffffffffu;   // cached warp top-2 (value, key)
    float wmax = 1e10f;
    int sa = 0, sb = 0;                  // the sample(s) selected by the previous round
    bool have_b = false;
    const int* my_idx = sidx + warp * 32 * P + lane;
    for (int r = 1, it = 0; r < m; ++it) {
      const float xa = sx[sa], ya = sy[sa], za = sz[sa];
      const float xb = sx[sb], yb = sy[sb], zb = sz[sb];
      const float qa = axis == 0 ? xa : (axis == 1 ? ya : za);
      const float qb = axis == 0 ? xb : (axis == 1 ? yb : zb);
      const float dma = fmaxf(fmaxf(__fsub_rn(slab_lo, qa), __fsub_rn(qa, slab_hi)), 0.f);
      const float dmb = fmaxf(fmaxf(__fsub_rn(slab_lo, qb), __fsub_rn(qb, slab_hi)), 0.f);
      if (!(__fmul_rn(dma, dma) > wmax) || (have_b && !(__fmul_rn(dmb, dmb) > wmax))) {   // warp-uniform: the slab may change
        float cv[P];
        int cj[P];
        if (have_b) {
#pragma unroll
          for (int j = 0; j < P; ++j) {
            const float da = dist2_ref_fma(px[j] - xa, py[j] - ya, pz[j] - za);
            const float db = dist2_ref_fma(px[j] - xb, py[j] - yb, pz[j] - zb);
            const float d2 = fminf(fminf(da, mind[j]), db);
            mind[j] = d2; cv[j] = d2; cj[j] = j;
          }
        } else {
#pragma unroll
          for (int j = 0; j < P; ++j) {
            const float d2 = fminf(dist2_ref_fma(px[j] - xa, py[j] - ya, pz[j] - za), mind[j]);
            mind[j] = d2; cv[j] = d2; cj[j] = j;
          }
        }
#pragma unroll
        for (int w = 1; w < P; w <<= 1) {
#pragma unroll
          for (int j = 0; j + w < P; j += 2 * w) {
            const bool take = cv[j + w] > cv[j];
            cv[j] = take ? cv[j + w] : cv[j];
            cj[j] = take ? cj[j + w] : cj[j];
          }
        }
        const float best = cv[0] > -1.0f ? cv[0] : -1.0f;
        const int bj = cv[0] > -1.0f ? cj[0] : 0;
        const unsigned vb = best < 0.f ? 0u : __float_as_uint(best) + 1u;
        wv1 = __reduce_max_sync(PWCLO_FULL_MASK, vb);
        unsigned key = 0xffffffffu;
        if (vb == wv1) key = fps_key((unsigned)my_idx[bj * 32], logT);
        wk1 = __reduce_min_sync(PWCLO_FULL_MASK, key);
        wmax = wv1 == 0u ? -1.0f : __uint_as_float(wv1 - 1u);
        // second of the warp: the owner lane of the first repeats its tournament without that point; the other lanes'
        // candidates are their own first (masking nothing reproduces it)
        const bool owner = wv1 != 0u && vb == wv1 && key == wk1;
#pragma unroll
        for (int j = 0; j < P; ++j) { cv[j] = (owner && j == bj) ? -2.0f : mind[j]; cj[j] = j; }
#pragma unroll
        for (int w = 1; w < P; w <<= 1) {
#pragma unroll
          for (int j = 0; j + w < P; j += 2 * w) {
            const bool take = cv[j + w] > cv[j];
            cv[j] = take ? cv[j + w] : cv[j];
            cj[j] = take ? cj[j + w] : cj[j];
          }
        }
        const float sbest = cv[0] > -1.0f ? cv[0] : -1.0f;
        const int sj = cv[0] > -1.0f ? cj[0] : 0;
        const unsigned vs = sbest < 0.f ? 0u : __float_as_uint(sbest) + 1u;
        wv2 = __reduce_max_sync(PWCLO_FULL_MASK, vs);
        unsigned key2 = 0xffffffffu;
        if (vs == wv2 && wv2 != 0u) key2 = fps_key((unsigned)my_idx[sj * 32], logT);
        wk2 = __reduce_min_sync(PWCLO_FULL_MASK, key2);
      }
      const int buf = it & 1;
      if (lane == 0) {
        red2_val[buf][0][warp] = wv1; red2_key[buf][0][warp] = wk1;
        red2_val[buf][1][warp] = wv2; red2_key[buf][1][warp] = wk2;
      }
      __syncthreads();
      const unsigned v1 = lane < NW ? red2_val[buf][0][lane] : 0u, k1 = lane < NW ? red2_key[buf][0][lane] : 0xffffffffu;
      const unsigned v2 = lane < NW ? red2_val[buf][1][lane] : 0u, k2 = lane < NW ? red2_key[buf][1][lane] : 0xffffffffu;
      const unsigned bv1 = __reduce_max_sync(PWCLO_FULL_MASK, v1);
      const unsigned bk1 = __reduce_min_sync(PWCLO_FULL_MASK, v1 == bv1 ? k1 : 0xffffffffu);
      const bool win = bv1 != 0u && v1 == bv1 && k1 == bk1;           // the warp that holds the first: its second competes
      const unsigned cv2 = win ? v2 : v1, ck2 = win ? k2 : k1;
      const unsigned bv2 = __reduce_max_sync(PWCLO_FULL_MASK, cv2);
      const unsigned bk2 = __reduce_min_sync(PWCLO_FULL_MASK, cv2 == bv2 ? ck2 : 0xffffffffu);
      sa = bv1 == 0u ? 0 : (int)fps_key_to_index(bk1, logT);
      if (tid == 0) idx[r] = sa;
      have_b = false;
      sb = sa;
      if (bv1 != 0u && bv2 > 1u && r + 1 < m) {            // a second candidate with a running minimum > 0
        const int cand = (int)fps_key_to_index(bk2, logT);
        const float d = dist2_ref_fma(sx[cand] - sx[sa], sy[cand] - sy[sa], sz[cand] - sz[sa]);   // candidate minus last
        if (!(d < __uint_as_float(bv2 - 1u))) {            // fminf(d, mind[cand]) would leave mind[cand] unchanged
          have_b = true;
          sb = cand;
          if (tid == 0) idx[r + 1] = cand;
        }
      }
      r += have_b ? 2 : 1;
    }
    return;
  }
  unsigned wv = 0u, wk = 0xffffffffu;   // cached warp (value, key)
  float wmax = 1e10f;                   // cached largest running minimum of the warp (-1: no valid point)
  int old = 0;
  for (int r = 1; r < m; ++r) {
    const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
    const float qa = axis == 0 ? x1 : (axis == 1 ? y1 : z1);
    const float dmin = fmaxf(fmaxf(__fsub_rn(slab_lo, qa), __fsub_rn(qa, slab_hi)), 0.f);
    if (!(__fmul_rn(dmin, dmin) > wmax) && !(dbg && r > 1)) {     // warp-uniform: this slab may change
      const int* my_idx = sidx + warp * 32 * P + lane;
      // all P updates are independent; the arg-max is a log-depth tournament (earlier j wins ties, which
      // is the reference's order because j ascends with the tie key) instead of a P-long select chain
      float cv[P];
      int cj[P];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const float d = dist2_ref_fma(px[j] - x1, py[j] - y1, pz[j] - z1);
        const float d2 = fminf(d, mind[j]);
        mind[j] = d2;
        cv[j] = d2;
        cj[j] = j;
      }
#pragma unroll
      for (int w = 1; w < P; w <<= 1) {
#pragma unroll
        for (int j = 0; j + w < P; j += 2 * w) {
          const bool take = cv[j + w] > cv[j];
          cv[j] = take ? cv[j + w] : cv[j];
          cj[j] = take ? cj[j + w] : cj[j];
        }
      }
      const float best = cv[0] > -1.0f ? cv[0] : -1.0f;
      const int bj = cv[0] > -1.0f ? cj[0] : 0;
      const unsigned vb = best < 0.f ? 0u : __float_as_uint(best) + 1u;
      wv = __reduce_max_sync(PWCLO_FULL_MASK, vb);
      unsigned key = 0xffffffffu;
      if (vb == wv) key = fps_key((unsigned)my_idx[bj * 32], logT);
      wk = __reduce_min_sync(PWCLO_FULL_MASK, key);
      wmax = wv == 0u ? -1.0f : __uint_as_float(wv - 1u);
      if (TIE) {
        bool lt = false;
        if (vb == wv && wv != 0u) {     // the lane(s) holding the warp maximum: is it attained twice inside the thread?
          int c = 0;
#pragma unroll
          for (int j = 0; j < P; ++j) c += (mind[j] == best) ? 1 : 0;
          lt = c > 1;
        }
        wtie = (__popc(__ballot_sync(PWCLO_FULL_MASK, vb == wv)) > 1 || __any_sync(PWCLO_FULL_MASK, lt)) ? 1u : 0u;
      }
    }
    const int buf = r & 1;
    if (lane == 0) {
      red_val[buf][warp] = wv; red_key[buf][warp] = wk;
      if (TIE) red_tie[buf][warp] = wtie;
    }
    __syncthreads();
    const unsigned v2 = lane < NW ? red_val[buf][lane] : 0u;
    const unsigned k2 = lane < NW ? red_key[buf][lane] : 0xffffffffu;
    const unsigned bv = __reduce_max_sync(PWCLO_FULL_MASK, v2);
    const unsigned bk = __reduce_min_sync(PWCLO_FULL_MASK, v2 == bv ? k2 : 0xffffffffu);
    if (TIE) {
      const unsigned t2 = lane < NW ? red_tie[buf][lane] : 0u;
      const unsigned eq = __ballot_sync(PWCLO_FULL_MASK, lane < NW && v2 == bv);
      tie_acc |= bv == 0u || __popc(eq) > 1 || __any_sync(PWCLO_FULL_MASK, lane < NW && v2 == bv && t2 != 0u);
    }
    old = bv == 0u ? 0 : (int)fps_key_to_index(bk, logT);
    if (tid == 0) idx[r] = old;
  }
  if (TIE && tid == 0 && tie_out != nullptr) tie_out[b] = tie_acc ? 1 : 0;
}

struct FpsTie { const int32_t* in; int32_t* out; };

template <int P, int THREADS>
static int launch_fps_slab(const float* xyz, int B, int n, int m, int logT, int origin_skip, int32_t* idx, FpsTie tie,
                           cudaStream_t st) {
  const size_t smem = (size_t)P * THREADS * (sizeof(unsigned) + sizeof(int)) + (size_t)3 * n * sizeof(float);
  const char* pe = getenv("PWCLO_FPS_PICK");
  // Measured on the B200 (profiles/round2_fps_double_pick.json): the double pick is exact but SLOWER (8192 -> 2048: 1.15 ms
  // against 0.89 ms) -- the two best candidates of a round usually sit in the same gap of the sample set, so the second
  // one is rarely independent of the first, and the top-2 bookkeeping lengthens every round.  Opt-in: PWCLO_FPS_PICK=2.
  const bool dual = pe && pe[0] == '2';
  auto kern = tie.out ? fps_slab_kernel<P, THREADS, true, 1>
                      : (dual ? fps_slab_kernel<P, THREADS, false, 2> : fps_slab_kernel<P, THREADS, false, 1>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<B, THREADS, smem, st>>>(xyz, n, m, logT, origin_skip, idx, getenv("PWCLO_FPS_DBG_SKIPALL") ? 1 : 0, tie.in, tie.out);
  return launch_status();
}

// -------------------------------------------------------------------------------------------------
// Cluster variant for FEW clouds (B * 8 <= SM count: training at 8 pairs per GPU, single-pair inference).  One CTA
// per cloud leaves most of the machine idle and makes every one of the m-1 dependent rounds pay for all n points
// on one SM (n = 16384: coordinates do not even fit the register file, 1.7 us per round).  Here a thread-block
// cluster of 8 CTAs owns a cloud: point k lives in the registers of global thread k mod 2048 (so a thread's points
// share k mod T and ascend in k, which keeps the reference's tie order exactly as in the kernels above), every CTA
// finds its local arg-max, writes (value, key, xyz of that point) into a slot of ALL eight CTAs through distributed
// shared memory, one cluster barrier, and every CTA picks the winner from its own eight slots.  Per round: n/8 point
// updates per SM, one __syncthreads, one cluster barrier; slots are double-buffered.  Same result, bit for bit.
// -------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kFpsCluster = 8;
constexpr int kFpsClThreads = 256;

struct FpsCand { unsigned val, key; float x, y, z; unsigned tie; };

template <int P, bool TIE>
__global__ void __launch_bounds__(kFpsClThreads, 1)
fps_cluster_kernel(const float* __restrict__ xyz, int n, int m, int logT, int origin_skip, int32_t* __restrict__ idx,
                   const int32_t* __restrict__ tie_in, int32_t* __restrict__ tie_out) {
  constexpr int THREADS = kFpsClThreads, CL = kFpsCluster, TT = THREADS * CL, NW = THREADS / 32;
  __shared__ float sx[P * THREADS], sy[P * THREADS], sz[P * THREADS];     // this CTA's points, slot = j*THREADS + tid
  __shared__ unsigned red_val[2][NW], red_key[2][NW], red_tie[2][NW];
  __shared__ FpsCand cand[2][CL];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gt = rank * THREADS + tid;                 // global thread of the cloud: owns points gt, gt + TT, ...
  xyz += (size_t)b * n * 3;
  idx += (size_t)b * m;
  // the flag is per cloud, so all CTAs of the cluster take the same branch (no cluster barrier is left half-entered)
  if (fps_prefix_shortcut(tie_in, tie_out, b, m, idx, tid, THREADS, rank == 0)) return;
  bool tie_acc = false;

  float px[P], py[P], pz[P], mind[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int k = gt + j * TT;
    float x = 0.f, y = 0.f, z = 0.f;
    bool valid = k < n;
    if (valid) {
      x = xyz[k * 3 + 0]; y = xyz[k * 3 + 1]; z = xyz[k * 3 + 2];
      if (origin_skip) {
        const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
        valid = !((double)mag <= 1e-3);
      }
    }
    px[j] = x; py[j] = y; pz[j] = z;
    sx[j * THREADS + tid] = x; sy[j * THREADS + tid] = y; sz[j * THREADS + tid] = z;
    mind[j] = valid ? 1e10f : -1.0f;
  }
  if (rank == 0 && tid == 0) idx[0] = 0;
  float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];         // the first sample is point 0
  __syncthreads();

  for (int r = 1; r < m; ++r) {
    float best = -1.0f;
    int bestk = 0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float d = dist2_ref_fma(px[j] - x1, py[j] - y1, pz[j] - z1);
      const float d2 = fminf(d, mind[j]);
      mind[j] = d2;
      if (d2 > best) { best = d2; bestk = gt + j * TT; }
    }
    const unsigned vb = best < 0.f ? 0u : __float_as_uint(best) + 1u;
    const unsigned key = fps_key((unsigned)bestk, logT);
    const unsigned wv = __reduce_max_sync(PWCLO_FULL_MASK, vb);
    const unsigned wk = __reduce_min_sync(PWCLO_FULL_MASK, vb == wv ? key : 0xffffffffu);
    const int buf = r & 1;
    if (TIE) {
      bool lt = false;
      if (vb == wv && wv != 0u) {
        int c = 0;
#pragma unroll
        for (int j = 0; j < P; ++j) c += (mind[j] == best) ? 1 : 0;
        lt = c > 1;
      }
      const bool wt = __popc(__ballot_sync(PWCLO_FULL_MASK, vb == wv)) > 1 || __any_sync(PWCLO_FULL_MASK, lt);
      if (lane == 0) red_tie[buf][warp] = wt ? 1u : 0u;
    }
    if (lane == 0) { red_val[buf][warp] = wv; red_key[buf][warp] = wk; }
    __syncthreads();
    if (warp == 0) {
      const unsigned v2 = lane < NW ? red_val[buf][lane] : 0u;
      const unsigned k2 = lane < NW ? red_key[buf][lane] : 0xffffffffu;
      const unsigned bv = __reduce_max_sync(PWCLO_FULL_MASK, v2);
      const unsigned bk = __reduce_min_sync(PWCLO_FULL_MASK, v2 == bv ? k2 : 0xffffffffu);
      unsigned ctie = 0u;
      if (TIE) {
        const unsigned t2 = lane < NW ? red_tie[buf][lane] : 0u;
        const unsigned eq = __ballot_sync(PWCLO_FULL_MASK, lane < NW && v2 == bv);
        ctie = (__popc(eq) > 1 || __any_sync(PWCLO_FULL_MASK, lane < NW && v2 == bv && t2 != 0u)) ? 1u : 0u;
      }
      if (lane < CL) {
        FpsCand c;
        c.val = bv; c.key = bk; c.x = 0.f; c.y = 0.f; c.z = 0.f; c.tie = ctie;
        if (bv != 0u) {
          const int k = (int)fps_key_to_index(bk, logT);           // a point of this CTA
          const int slot = (k / TT) * THREADS + (k % TT - rank * THREADS);
          c.x = sx[slot]; c.y = sy[slot]; c.z = sz[slot];
        }
        *cluster.map_shared_rank(&cand[buf][rank], lane) = c;      // my candidate into CTA `lane`'s slot table
      }
    }
    cluster.sync();
    unsigned bv = 0u, bk = 0xffffffffu;
    int who = 0;
#pragma unroll
    for (int c = 0; c < CL; ++c) {
      const unsigned v = cand[buf][c].val, k = cand[buf][c].key;
      if (v > bv || (v == bv && k < bk)) { bv = v; bk = k; who = c; }
    }
    if (TIE) {
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < CL; ++c) cnt += (cand[buf][c].val == bv) ? 1 : 0;
      tie_acc |= bv == 0u || cnt > 1 || cand[buf][who].tie != 0u;
    }
    int old = 0;
    if (bv != 0u) {
      old = (int)fps_key_to_index(bk, logT);
      x1 = cand[buf][who].x; y1 = cand[buf][who].y; z1 = cand[buf][who].z;
    } else {
      x1 = xyz[0]; y1 = xyz[1]; z1 = xyz[2];                        // no valid candidate: the reference re-selects index 0
    }
    if (rank == 0 && tid == 0) idx[r] = old;
  }
  if (TIE && rank == 0 && tid == 0 && tie_out != nullptr) tie_out[b] = tie_acc ? 1 : 0;
  cluster.sync();        // nobody exits while a neighbour may still write into its slot table
}

template <int P>
static int launch_fps_cluster(const float* xyz, int B, int n, int m, int logT, int origin_skip, int32_t* idx, FpsTie tie,
                              cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kFpsCluster, B, 1);
  cfg.blockDim = dim3(kFpsClThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kFpsCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int32_t* tin = tie.in;
  int32_t* tout = tie.out;
  cudaError_t e = tie.out ? cudaLaunchKernelEx(&cfg, fps_cluster_kernel<P, true>, xyz, n, m, logT, origin_skip, idx, tin, tout)
                          : cudaLaunchKernelEx(&cfg, fps_cluster_kernel<P, false>, xyz, n, m, logT, origin_skip, idx, tin, tout);
  return e == cudaSuccess ? launch_status() : (int)e;
}

template <int P, int THREADS, int MODE>
static int launch_fps(const float* xyz, int B, int n, int m, int logT, int origin_skip, float* scratch, int32_t* idx,
                      FpsTie tie, cudaStream_t st) {
  size_t smem = MODE == 2 ? 0 : (size_t)3 * n * sizeof(float);
  const char* pe = getenv("PWCLO_FPS_PICK");
  const bool dual = MODE == 0 && pe && pe[0] == '2';
  auto kern = tie.out ? fps_kernel<P, THREADS, MODE, true, 1>
                      : (dual ? fps_kernel<P, THREADS, MODE, false, 2> : fps_kernel<P, THREADS, MODE, false, 1>);
  if (smem > 32 * 1024) {  // static smem (reduction slots) counts against the 48 KB default limit
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<B, THREADS, smem, st>>>(xyz, n, m, logT, origin_skip, scratch, idx, tie.in, tie.out);
  return launch_status();
}

}  // namespace pwclo

using namespace pwclo;

// scratch for MODE 2 (n > 16384) is allocated lazily per call size; kept simple on purpose: that
// path is outside every BASELINE configuration and exists so that no size silently fails.
PWCLO_API int pwclo_furthest_point_sampling(const float* xyz, int B, int N, int m, unsigned flags, int32_t* idx,
                                            void* stream) {
  return pwclo_furthest_point_sampling_prefix(xyz, B, N, m, flags, idx, nullptr, nullptr, stream);
}

PWCLO_API int pwclo_furthest_point_sampling_prefix(const float* xyz, int B, int N, int m, unsigned flags, int32_t* idx,
                                                   const int32_t* tie_in, int32_t* tie_out, void* stream) {
  if (!xyz || !idx || B < 0 || N <= 0 || m < 0) return PWCLO_EINVAL;
  if (tie_in && m > N) return PWCLO_EINVAL;        // a prefix of length m needs m input points
  if (B == 0 || m == 0) return PWCLO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const FpsTie tie = {tie_in, tie_out};
  const int cap = (flags & PWCLO_FPS_CAP1024) ? 1024 : 512;
  int T = 1, logT = 0;
  while (T * 2 <= N && T * 2 <= cap) { T *= 2; ++logT; }
  const int skip = (flags & PWCLO_FPS_ORIGIN_SKIP) ? 1 : 0;
  // few clouds: a cluster of 8 CTAs per cloud (see fps_cluster_kernel); PWCLO_FPS_CLUSTER=0 disables
  {
    const char* ce = getenv("PWCLO_FPS_CLUSTER");
    const bool allow = !(ce && ce[0] == '0');
    // Measured on the B200 (tools/bench_fps_cluster.py): a round costs 0.8-1.1 us whatever n is -- the cluster barrier
    // and the DSMEM exchange, not the point updates -- so the cluster wins only where one CTA per cloud is slow:
    // n > 8192 (16 clouds x 16384 -> 2048: 2.23 ms against 3.57 ms); at n <= 8192 the register-resident slab kernel
    // needs 0.43 us per round and stays the choice (PWCLO_FPS_CLUSTER=2 forces the cluster there, for tests).
    const bool force = ce && ce[0] == '2';
    if (allow && m >= 64 && N >= 2048 && N <= 16384 && (N > 8192 || force) && B * kFpsCluster <= kNumSM) {
      if (N <= 4096) return launch_fps_cluster<2>(xyz, B, N, m, logT, skip, idx, tie, st);
      if (N <= 8192) return launch_fps_cluster<4>(xyz, B, N, m, logT, skip, idx, tie, st);
      return launch_fps_cluster<8>(xyz, B, N, m, logT, skip, idx, tie, st);
    }
  }
  // THREADS must be a multiple of T so that a thread's points share (k mod T): 512 or 1024 (cap 1024)
  // slab-skipping kernel: worth its two in-kernel sorts once there are enough rounds and points
  const char* slab_min = getenv("PWCLO_FPS_SLAB_MIN_N");
  if (m >= 256 && N >= (slab_min ? atoi(slab_min) : 2049) && !getenv("PWCLO_FPS_NO_SLAB")) {
    const bool wide = getenv("PWCLO_FPS_SLAB8") == nullptr;     // 512 threads x 16 points: fewer warps per barrier (7 % faster)
    if (cap == 1024) {
      if (N <= 4096) return launch_fps_slab<4, 1024>(xyz, B, N, m, logT, skip, idx, tie, st);
      if (N <= 8192) return launch_fps_slab<8, 1024>(xyz, B, N, m, logT, skip, idx, tie, st);
    } else {
      if (N <= 2048) return launch_fps_slab<4, 512>(xyz, B, N, m, logT, skip, idx, tie, st);
      if (N <= 4096) return launch_fps_slab<8, 512>(xyz, B, N, m, logT, skip, idx, tie, st);
      if (N <= 8192) return wide ? launch_fps_slab<16, 512>(xyz, B, N, m, logT, skip, idx, tie, st)
                                 : launch_fps_slab<8, 1024>(xyz, B, N, m, logT, skip, idx, tie, st);
    }
  }
#define FPS_CASE(P, TH, MODE) return launch_fps<P, TH, MODE>(xyz, B, N, m, logT, skip, nullptr, idx, tie, st)
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
  int rc = launch_fps<1, 1024, 2>(xyz, B, N, m, logT, skip, scratch, idx, tie, st);
  cudaFreeAsync(scratch, st);
  return rc;
}
