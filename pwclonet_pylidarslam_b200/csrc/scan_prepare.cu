// scan_prepare.cu -- KITTI scan -> network input on the GPU.
//
// Reference (host numpy, per frame): slam/dataset/kitti_odometry_dataset.py:375-397 reads the .bin
// scan float32[n,4], multiplies by the 4x4 velo->cam calibration `Tr` in float64, then filter_pcd
// (:149-172) drops the ground (y > 1.1) and everything outside |x|,|z| < 30 m and draws `npoints`
// survivors without replacement in random order (np.random.choice); with fewer survivors than npoints
// it keeps them all (ascending index) and pads with draws with replacement; with none it draws from
// the whole scan.  The result is cast to float32 (optionally after the training augmentation
// transform, :404-446).
//
// Here: two launches over the whole batch of scans, every scan point read from HBM once (16 B/point) plus a
// 2-byte key prefix per point written and read back (L2-resident).  The random subset is defined by a
// counter-based generator so that it is reproducible and checkable: point i of scan s gets the 31-bit key
// philox4x32-10(counter = (32*(i/128) + i%32, s, 0, 0), key = seed)[(i/32)%4] >> 1 (one Philox call per lane
// of the warp that streams a 128-point block), and the sample is the npoints survivors with the smallest
// (key, index), emitted in that order -- a uniformly random subset in uniformly random order, the
// distribution np.random.choice(replace=False) draws from.
//   scan_keys_kernel    grid (8, scans): the 128-point blocks of a scan are dealt round-robin to 8 CTAs; each
//                       streams its blocks with coalesced float4 loads: fp64 transform, crop, key; it writes the
//                       15-bit key prefix of every point (or `rejected`) and its 1024-bin histogram of prefixes.
//                       Pure streaming, no inter-CTA dependency: this is the HBM-bound part.
//   scan_select_kernel  one CTA per scan, everything in shared memory: sums the 8 histograms and finds the coarse
//                       bin holding the npoints-th smallest key; streams the prefixes once, re-bins the
//                       ~npoints + M/1024 candidates below it over their own key range (~npoints/1024 keys per
//                       bin) and compacts them; counting-sorts them by fine bin with their full keys (recomputed
//                       densely); the rank of a key inside its bin (~8 comparisons) is its output position;
//                       gathers the chosen raw points (L2 hits), transforms and writes them.
// A first version did all of this in one launch with an 8-CTA thread-block cluster per scan (histograms reduced
// through distributed shared memory): bit-identical results, but its nine cluster barriers and DSMEM latency
// chains cost ~50 us per cluster with only 33 clusters resident -- 0.26 ms per 128 scans against 0.0x ms here
// (profiles/r1l_scan_ncu_summary.txt).
#include "common.cuh"

namespace pwclo {

constexpr int kPrepThreads = 512;
constexpr int kPrepSplit = 8;            // CTAs per scan in the streaming kernel
constexpr int kPrepBinBits = 10;
constexpr int kPrepBins = 1 << kPrepBinBits;   // top bits of the 31-bit key
constexpr int kPrepBoundaryCap = 512;    // keys inside the boundary bin (expected survivors / kPrepBins)
constexpr int kPrepBlock = 128;          // points per warp step: 4 coalesced float4 loads + one Philox call per lane
constexpr unsigned short kPrepRejected = 0xffffu;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ unsigned pick4(const uint4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }

// row-major 3x4 affine map in float64 in the order np.matmul(Tr, [x y z 1]^T) accumulates its 4-term dot
// products: one multiply, then fused multiply-adds in storage order (the same expression is checked bit
// for bit against np.matmul in tests/test_scan_cpu.py)
__device__ __forceinline__ void affine3x4(const double* __restrict__ T, double x, double y, double z, double& X, double& Y,
                                          double& Z) {
  X = __dadd_rn(__fma_rn(T[2], z, __fma_rn(T[1], y, __dmul_rn(T[0], x))), T[3]);
  Y = __dadd_rn(__fma_rn(T[6], z, __fma_rn(T[5], y, __dmul_rn(T[4], x))), T[7]);
  Z = __dadd_rn(__fma_rn(T[10], z, __fma_rn(T[9], y, __dmul_rn(T[8], x))), T[11]);
}

// filter_pcd of the two datasets the reference trains PWCLO-Net on:
//   KITTI odometry (kitti_odometry_dataset.py:151-159, camera-style frame): not (y > 1.1) and |x| < 30 and |z| < 30
//   KITTI-360 (kitti_360_dataset_2.py:113-123, velodyne frame, no Tr):       not (z < -(1.73-0.3)) and |x| < near and |y| < near
struct CropSpec {
  int ground_axis, ground_sign;   // reject when sign * (P[axis] - thr) > 0
  int near_a, near_b;             // keep when |P[a]| < near and |P[b]| < near (strict, both signs, as the reference)
  double ground_thr, near_thr;
};

__device__ __forceinline__ bool keep_point(const CropSpec& c, double X, double Y, double Z) {
  const double g = c.ground_axis == 0 ? X : (c.ground_axis == 1 ? Y : Z);
  const double a = c.near_a == 0 ? X : (c.near_a == 1 ? Y : Z);
  const double b = c.near_b == 0 ? X : (c.near_b == 1 ? Y : Z);
  const bool ground = c.ground_sign > 0 ? (g > c.ground_thr) : (g < c.ground_thr);
  return !ground && (a < c.near_thr) && (a > -c.near_thr) && (b < c.near_thr) && (b > -c.near_thr);
}

struct PrepCtl {
  int bstar;       // boundary bin (kPrepBins when there are fewer keys than wanted)
  int below;       // keys in bins < bstar
  int bpop;        // keys inside the boundary bin
  int total;       // all keys of the histogram
  int ncand;       // compacted candidates
  int pad[3];
};

// 31-bit selection key of point i of a scan: lane (i & 31) of the warp that streams the 128-point block
// i >> 7 draws one Philox counter for its four points i, i+32, i+64, i+96 of that block
__device__ __forceinline__ unsigned point_key(int i, unsigned scan, uint2 pkey) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)(((i >> 7) << 5) | (i & 31)), scan, 0u, 0u), pkey);
  return pick4(r, (i >> 5) & 3) >> 1;
}

// exclusive scan of one value per thread over the CTA; returns the exclusive prefix, *total = CTA sum
template <int T>
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* warp_tot, unsigned* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(PWCLO_FULL_MASK, incl, o);
    if (lane >= o) incl += u;
  }
  __syncthreads();                                           // previous users of warp_tot are done
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned before = 0, sum = 0;
#pragma unroll
  for (int w = 0; w < T / 32; ++w) {
    const unsigned c = warp_tot[w];
    before += w < warp ? c : 0u;
    sum += c;
  }
  *total = sum;
  return before + incl - v;
}

// exclusive-scan the 1024-bin histogram (kPrepBins / T bins per thread), write the bin starts and locate the bin
// in which the `want`-th smallest key lies
template <int T>
__device__ void scan_locate(const unsigned* hist, unsigned* starts, unsigned* warp_tot, int want_i, PrepCtl* ctl) {
  constexpr int PER = kPrepBins / T;
  static_assert(PER == 1 || PER == 2, "one or two bins per thread");
  const int tid = threadIdx.x;
  const unsigned c0 = hist[PER * tid], c1 = PER == 2 ? hist[PER * tid + 1] : 0u;
  if (tid == 0) ctl->bstar = kPrepBins;
  unsigned total;
  const unsigned mine = c0 + c1;
  const unsigned excl = block_exclusive_scan<T>(mine, warp_tot, &total);   // (its barriers order the bstar reset)
  starts[PER * tid] = excl;
  if (PER == 2) starts[PER * tid + 1] = excl + c0;
  if (tid == T - 1) { starts[kPrepBins] = total; ctl->total = (int)total; }
  const unsigned want = (unsigned)want_i;
  if (excl < want && excl + mine >= want) {                  // exactly one thread when there are >= want keys
    if (excl + c0 >= want) { ctl->bstar = PER * tid; ctl->below = (int)excl; ctl->bpop = (int)c0; }
    else { ctl->bstar = PER * tid + 1; ctl->below = (int)(excl + c0); ctl->bpop = (int)c1; }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// launch 1: stream every scan point once
__global__ void __launch_bounds__(kPrepThreads, 2)
scan_keys_kernel(const float4* __restrict__ raw, const long long* __restrict__ offsets, const double* __restrict__ Tr,
                 int tr_per_scan, CropSpec crop, unsigned seed_lo, unsigned seed_hi, unsigned short* __restrict__ pref,
                 unsigned* __restrict__ ghist) {
  __shared__ unsigned hist[kPrepBins];
  __shared__ double sT[12];
  const int rank = blockIdx.x, scan = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kPrepThreads / 32;
  const long long base = offsets[scan];
  const int n = (int)(offsets[scan + 1] - base);
  const float4* pts = raw + base;
  unsigned short* pf = pref + base;
  const int nblk = (n + kPrepBlock - 1) / kPrepBlock;
  const uint2 pkey = make_uint2(seed_lo, seed_hi);
  for (int b = tid; b < kPrepBins; b += kPrepThreads) hist[b] = 0;
  if (tid < 12) sT[tid] = Tr[(tr_per_scan ? (size_t)scan * 12 : 0) + tid];
  __syncthreads();
  // a warp takes 128 consecutive points per step: 4 x 512 B coalesced loads, one Philox call per lane.  Blocks are
  // dealt round-robin (a scan is stored beam by beam and whole beams are cropped away: contiguous eighths would
  // leave some CTAs with all the survivors -- and all the histogram atomics)
  for (int g = rank + warp * kPrepSplit; g < nblk; g += kWarps * kPrepSplit) {
    const int b0 = g << 7;
    float4 p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = b0 + 32 * k + lane;
      if (i < n) p[k] = __ldcs(pts + i);
    }
    const uint4 r = philox4x32_10(make_uint4((unsigned)((g << 5) | lane), (unsigned)scan, 0u, 0u), pkey);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = b0 + 32 * k + lane;
      if (i < n) {
        unsigned short v = kPrepRejected;
        double X, Y, Z;
        affine3x4(sT, (double)p[k].x, (double)p[k].y, (double)p[k].z, X, Y, Z);
        if (keep_point(crop, X, Y, Z)) {
          const unsigned top15 = pick4(r, k) >> 17;              // (key31 = word >> 1) >> 16
          v = (unsigned short)top15;
          atomicAdd(&hist[top15 >> (15 - kPrepBinBits)], 1u);
        }
        pf[i] = v;
      }
    }
  }
  __syncthreads();
  unsigned* gh = ghist + ((size_t)scan * kPrepSplit + rank) * kPrepBins;
  for (int b = tid; b < kPrepBins; b += kPrepThreads) gh[b] = hist[b];
}

// list entry: key31 << 32 | fine bin << 22 | point index (< 2^22); the bin is a monotonic function of the key,
// so comparing whole entries orders by (key, index)
constexpr int kPrepIdxBits = 22;
constexpr unsigned kPrepIdxMask = (1u << kPrepIdxBits) - 1u;
constexpr int kPrepCandSlack = 2048;     // candidates beyond npoints (expected M/1024)
constexpr int kSelThreads = 1024;        // select kernel: one CTA per scan, latency bound -> as many threads as fit

// ---------------------------------------------------------------------------------------------------------------
// launch 2: one CTA per scan selects, orders, gathers
__global__ void __launch_bounds__(kSelThreads, 1)
scan_select_kernel(const float4* __restrict__ raw, const long long* __restrict__ offsets, const double* __restrict__ Tr,
                   int tr_per_scan, const double* __restrict__ post, unsigned seed_lo, unsigned seed_hi, int npoints,
                   const unsigned short* __restrict__ pref, const unsigned* __restrict__ ghist, float* __restrict__ out,
                   int* __restrict__ sel_idx, int* __restrict__ survivors) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* hist = reinterpret_cast<unsigned*>(smem_raw);                    // [kPrepBins]
  unsigned* starts = hist + kPrepBins;                                       // [kPrepBins + 2]
  unsigned long long* list = reinterpret_cast<unsigned long long*>(starts + kPrepBins + 2);   // [npoints + kPrepBoundaryCap]
  unsigned* cand = reinterpret_cast<unsigned*>(list + npoints + kPrepBoundaryCap);            // [npoints + kPrepCandSlack]
  __shared__ PrepCtl ctl;
  __shared__ unsigned warp_tot[kSelThreads / 32];
  __shared__ double sT[12], sP[12];
  const int list_cap = npoints + kPrepBoundaryCap, cand_cap = npoints + kPrepCandSlack;
  const int scan = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const long long base = offsets[scan];
  const int n = (int)(offsets[scan + 1] - base);
  const float4* pts = raw + base;
  const unsigned short* pf = pref + base;
  const uint2 pkey = make_uint2(seed_lo, seed_hi);
  const bool too_big = n > (int)kPrepIdxMask;

  {  // coarse histogram of the scan = sum of the streaming CTAs' histograms
    const unsigned* gh = ghist + (size_t)scan * kPrepSplit * kPrepBins;
    for (int b = tid; b < kPrepBins; b += kSelThreads) {
      unsigned c = 0;
#pragma unroll
      for (int r = 0; r < kPrepSplit; ++r) c += gh[r * kPrepBins + b];
      hist[b] = c;
    }
  }
  if (tid < 12) {
    sT[tid] = Tr[(tr_per_scan ? (size_t)scan * 12 : 0) + tid];
    sP[tid] = post ? post[(size_t)scan * 12 + tid] : 0.0;
  }
  if (tid == 0) ctl.ncand = 0;
  __syncthreads();
  scan_locate<kSelThreads>(hist, starts, warp_tot, npoints, &ctl);
  const int cstar = ctl.bstar;
  const int M = ctl.total;
  const bool short_mode = cstar == kPrepBins || too_big;   // fewer survivors than npoints: keep them all, in index order
  __syncthreads();
  int have;

  if (!short_mode) {
    // Only keys whose 15-bit prefix is below R = (cstar+1) << 5 can be chosen: about npoints + M/1024 candidates.
    // Re-bin exactly those over [0, R) so that a bin holds ~npoints/1024 keys whatever the survivor count is, and
    // compact them (one shared-memory atomic per warp step).
    const unsigned R = (unsigned)(cstar + 1) << (15 - kPrepBinBits);
    for (int b = tid; b < kPrepBins; b += kSelThreads) hist[b] = 0;
    __syncthreads();
    // 8 prefixes per 16-byte load (aligned on the whole buffer: the first / last load of a scan may straddle its
    // neighbours and is masked), 4 loads in flight per thread
    const long long a0 = base & ~7ll;
    const uint4* pv = reinterpret_cast<const uint4*>(pref + a0);
    const int nvec = (int)((base + n - a0 + 7) >> 3);
    const int shift = (int)(base - a0);
    for (int v0 = 0; v0 < nvec; v0 += 4 * kSelThreads) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * kSelThreads + tid;
        q[u] = v < nvec ? __ldg(pv + v) : make_uint4(~0u, ~0u, ~0u, ~0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i_first = ((v0 + u * kSelThreads + tid) << 3) - shift;          // scan index of the first of 8
        const unsigned w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
        unsigned cmask = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const unsigned t15 = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
          const int i = i_first + e;
          if (t15 < R && i >= 0 && i < n) cmask |= 1u << e;
        }
        const unsigned c = __popc(cmask);
        // exclusive prefix of c over the warp, one shared-memory atomic per warp
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(PWCLO_FULL_MASK, incl, o);
          if (lane >= o) incl += t;
        }
        const unsigned wtotal = __shfl_sync(PWCLO_FULL_MASK, incl, 31);
        if (wtotal) {
          unsigned wbase = 0;
          if (lane == 0) wbase = (unsigned)atomicAdd(&ctl.ncand, (int)wtotal);
          wbase = __shfl_sync(PWCLO_FULL_MASK, wbase, 0);
          unsigned slot = wbase + incl - c;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (cmask & (1u << e)) {
              const unsigned t15 = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
              const unsigned fb = (t15 << kPrepBinBits) / R;
              atomicAdd(&hist[fb], 1u);
              if (slot < (unsigned)cand_cap) cand[slot] = (fb << kPrepIdxBits) | (unsigned)(i_first + e);
              ++slot;
            }
          }
        }
      }
    }
    __syncthreads();
    scan_locate<kSelThreads>(hist, starts, warp_tot, npoints, &ctl);
    const int bstar = ctl.bstar, below = ctl.below, bpop = ctl.bpop;
    const int ncand = min(ctl.ncand, cand_cap);
    const bool overflow = ctl.ncand > cand_cap || bpop > kPrepBoundaryCap;
    // counting sort by fine bin, with the full keys (dense: every thread has work)
    for (int c = tid; c < ncand; c += kSelThreads) {
      const unsigned e = cand[c];
      const unsigned bin = e >> kPrepIdxBits;
      if (bin > (unsigned)bstar) continue;
      const unsigned slot = starts[bin] + (atomicSub(&hist[bin], 1u) - 1u);
      if (slot < (unsigned)list_cap)
        list[slot] = ((unsigned long long)point_key((int)(e & kPrepIdxMask), (unsigned)scan, pkey) << 32) | e;
    }
    __syncthreads();
    const int filled = min(list_cap, below + bpop);
    if (tid == 0 && survivors) survivors[scan] = overflow ? -1 : M;
    // rank inside the bin = final position; gather, transform, write -- 4 entries per thread step so that the four
    // gathers are in flight together
    for (int j0 = 0; j0 < filled; j0 += 4 * kSelThreads) {
      int pos[4], idx[4];
      float4 p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * kSelThreads + tid;
        pos[u] = -1;
        idx[u] = 0;
        if (j < filled) {
          const unsigned long long e = list[j];
          const unsigned bin = ((unsigned)e >> kPrepIdxBits) & (kPrepBins - 1);
          const int s0 = (int)starts[bin], e0 = min((int)starts[bin + 1], filled);
          int ps = s0;
          for (int i = s0; i < e0; ++i) ps += list[i] < e ? 1 : 0;
          if (ps < npoints) { pos[u] = ps; idx[u] = (int)((unsigned)e & kPrepIdxMask); }   // the boundary bin holds a few too many
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (pos[u] >= 0) p[u] = __ldg(pts + idx[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (pos[u] < 0) continue;
        double X, Y, Z;
        affine3x4(sT, (double)p[u].x, (double)p[u].y, (double)p[u].z, X, Y, Z);
        if (post) { double A, B, C; affine3x4(sP, X, Y, Z, A, B, C); X = A; Y = B; Z = C; }
        float* o = out + ((size_t)scan * npoints + pos[u]) * 3;
        o[0] = (float)X; o[1] = (float)Y; o[2] = (float)Z;
        if (sel_idx) sel_idx[(size_t)scan * npoints + pos[u]] = idx[u];
      }
    }
    return;
  }

  // ---- fewer survivors than npoints (rare): all of them in ascending index, then draws with replacement
  {
    unsigned running = 0;
    for (int i0 = 0; i0 < n && !too_big; i0 += kSelThreads) {
      const int i = i0 + tid;
      const bool keep = i < n && pf[i] != kPrepRejected;
      unsigned total;
      const unsigned excl = block_exclusive_scan<kSelThreads>(keep ? 1u : 0u, warp_tot, &total);
      if (keep && running + excl < (unsigned)list_cap) list[running + excl] = (unsigned long long)(unsigned)i;
      running += total;
    }
    __syncthreads();
    have = too_big ? 0 : min(M, list_cap);
    if (tid == 0 && survivors) survivors[scan] = too_big ? -1 : M;
  }
  for (int j = tid; j < npoints; j += kSelThreads) {
    int idx;
    if (j < have) {
      idx = (int)(unsigned)list[j];
    } else {
      // draws with replacement (from the survivors, or from the whole scan when there are none)
      const int d = j - have;
      const uint4 r = philox4x32_10(make_uint4((unsigned)(d >> 2), (unsigned)scan, 1u, 0u), pkey);
      const unsigned u = pick4(r, d & 3);
      idx = have > 0 ? (int)(unsigned)list[__umulhi(u, (unsigned)have)] : (n > 0 ? (int)__umulhi(u, (unsigned)n) : -1);
    }
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (idx >= 0) {
      const float4 p = __ldg(pts + idx);
      double X, Y, Z;
      affine3x4(sT, (double)p.x, (double)p.y, (double)p.z, X, Y, Z);
      if (post) { double A, B, C; affine3x4(sP, X, Y, Z, A, B, C); X = A; Y = B; Z = C; }
      ox = (float)X; oy = (float)Y; oz = (float)Z;
    }
    float* o = out + ((size_t)scan * npoints + j) * 3;
    o[0] = ox; o[1] = oy; o[2] = oz;
    if (sel_idx) sel_idx[(size_t)scan * npoints + j] = idx;
  }
}

}  // namespace pwclo

// workspace: key prefixes (2 B per point, rounded to 256 B) + 8 x 1024 histogram counters per scan
PWCLO_API size_t pwclo_prepare_scans_workspace_bytes(long long total_points, int nscan) {
  if (total_points < 0 || nscan < 0) return 0;
  const size_t pref = (((size_t)total_points * 2 + 255) / 256) * 256;
  return pref + (size_t)nscan * pwclo::kPrepSplit * pwclo::kPrepBins * 4;
}

PWCLO_API int pwclo_prepare_scans(const float* raw, const long long* offsets, int nscan, long long total_points,
                                  const double* Tr, int tr_per_scan, const double* post, unsigned long long seed, int npoints,
                                  float* out, int32_t* sel_idx, int32_t* survivors, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  // KITTI odometry crop (kitti_odometry_dataset.py:151-159)
  return pwclo_prepare_scans_crop(raw, offsets, nscan, total_points, Tr, tr_per_scan, post, seed, npoints, 1, +1, 1.1, 0, 2, 30.0,
                                  out, sel_idx, survivors, workspace, workspace_bytes, stream);
}

PWCLO_API int pwclo_prepare_scans_crop(const float* raw, const long long* offsets, int nscan, long long total_points,
                                       const double* Tr, int tr_per_scan, const double* post, unsigned long long seed,
                                       int npoints, int ground_axis, int ground_sign, double ground_thr, int near_axis_a,
                                       int near_axis_b, double near_thr, float* out, int32_t* sel_idx, int32_t* survivors,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pwclo;
  if (ground_axis < 0 || ground_axis > 2 || near_axis_a < 0 || near_axis_a > 2 || near_axis_b < 0 || near_axis_b > 2 ||
      ground_sign == 0)
    return PWCLO_EINVAL;
  CropSpec crop;
  crop.ground_axis = ground_axis; crop.ground_sign = ground_sign; crop.near_a = near_axis_a; crop.near_b = near_axis_b;
  crop.ground_thr = ground_thr; crop.near_thr = near_thr;
  if (!raw || !offsets || !Tr || !out || !workspace || nscan < 0 || npoints <= 0 || total_points < 0) return PWCLO_EINVAL;
  if (((uintptr_t)raw & 15) != 0 || ((uintptr_t)workspace & 15) != 0) return PWCLO_EINVAL;
  if (workspace_bytes < pwclo_prepare_scans_workspace_bytes(total_points, nscan)) return PWCLO_EINVAL;
  if (nscan == 0) return PWCLO_OK;
  const size_t smem = (size_t)(kPrepBins * 2 + 2) * 4 + (size_t)(npoints + kPrepBoundaryCap) * 8 +
                      (size_t)(npoints + kPrepCandSlack) * 4;
  if (smem > 226 * 1024) return PWCLO_EUNSUPPORTED;
  {   // the attribute is per device: set it on every call (as every other launcher here does)
    cudaError_t e = cudaFuncSetAttribute(scan_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  unsigned short* pref = reinterpret_cast<unsigned short*>(workspace);
  unsigned* ghist = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(workspace) +
                                                (((size_t)total_points * 2 + 255) / 256) * 256);
  const unsigned slo = (unsigned)(seed & 0xffffffffull), shi = (unsigned)(seed >> 32);
  cudaStream_t st = (cudaStream_t)stream;
  scan_keys_kernel<<<dim3(kPrepSplit, nscan, 1), kPrepThreads, 0, st>>>(reinterpret_cast<const float4*>(raw), offsets, Tr,
                                                                       tr_per_scan ? 1 : 0, crop, slo, shi, pref, ghist);
  int rc = launch_status();
  if (rc) return rc;
  scan_select_kernel<<<nscan, kSelThreads, smem, st>>>(reinterpret_cast<const float4*>(raw), offsets, Tr, tr_per_scan ? 1 : 0,
                                                        post, slo, shi, npoints, pref, ghist, out, sel_idx, survivors);
  return launch_status();
}
