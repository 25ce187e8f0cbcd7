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
// Here: ONE kernel, one thread-block cluster of 8 CTAs per scan, every scan point read from HBM exactly
// once (16 B/point, the roofline of this path).  The random subset is defined by a counter-based
// generator so that it is reproducible and checkable: point i of scan s gets the 31-bit key
// philox4x32-10(counter = (i/4, s, 0, 0), key = seed)[i%4] >> 1, and the sample is the npoints survivors
// with the smallest (key, index), emitted in that order -- a uniformly random subset in uniformly
// random order, the distribution np.random.choice(replace=False) draws from.
//   phase 1  each CTA streams its eighth of the scan: float4 loads, fp64 transform, mask, key; keys stay
//            in shared memory, an 11-bit key histogram is built with shared-memory atomics
//   phase 2  CTA 0 sums the 8 histograms through distributed shared memory and finds the bin b* in which
//            the npoints-th smallest key lies
//   phase 3  every CTA appends its keys below b* to CTA 0's list and those inside b* to a small boundary
//            list (DSMEM atomics + stores)
//   phase 4  CTA 0 sorts the boundary list, completes the list to exactly npoints, bitonic-sorts it by
//            (key, index), gathers the chosen raw points, transforms and writes them
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pwclo {

constexpr int kPrepThreads = 1024;
constexpr int kPrepCluster = 8;
constexpr int kPrepBins = 2048;          // top 11 bits of the 31-bit key
constexpr int kPrepBoundaryCap = 1024;   // keys inside the boundary bin (expected survivors/2048)
constexpr unsigned kPrepRejected = 0xffffffffu;

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

__device__ __forceinline__ bool keep_point(double X, double Y, double Z) {
  // filter_pcd: not (y > 1.1) and -30 < x < 30 and -30 < z < 30   (kitti_odometry_dataset.py:151-159)
  return !(Y > 1.1) && (X < 30.0) && (X > -30.0) && (Z < 30.0) && (Z > -30.0);
}

struct PrepCtl {
  int count;       // entries in the main list
  int bcount;      // entries in the boundary list
  int bstar;       // boundary bin (kPrepBins when there are fewer survivors than npoints)
  int below;       // survivors in bins < bstar
  int survivors;   // M
  int pad[3];
};

// bitonic sort of `len` (power of two) 64-bit keys in shared memory, ascending, all threads of the CTA
__device__ void bitonic_sort_u64(unsigned long long* a, int len) {
  for (int k = 2; k <= len; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (len >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // lower index of the pair
        const int p = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long x = a[i], y = a[p];
        if ((x > y) == up) { a[i] = y; a[p] = x; }
      }
      __syncthreads();
    }
  }
}

__global__ void __cluster_dims__(kPrepCluster, 1, 1) __launch_bounds__(kPrepThreads, 1)
scan_prepare_kernel(const float4* __restrict__ raw, const long long* __restrict__ offsets, const double* __restrict__ Tr,
                    int tr_per_scan, const double* __restrict__ post, unsigned seed_lo, unsigned seed_hi, int npoints,
                    int chunk_cap, int list_cap, float* __restrict__ out, int* __restrict__ sel_idx,
                    int* __restrict__ survivors) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* hist = reinterpret_cast<unsigned*>(smem_raw);                                    // [kPrepBins]
  unsigned long long* blist = reinterpret_cast<unsigned long long*>(hist + kPrepBins);       // [kPrepBoundaryCap]
  unsigned long long* list = blist + kPrepBoundaryCap;                                       // [list_cap]
  unsigned* keys = reinterpret_cast<unsigned*>(list + list_cap);                             // [chunk_cap]
  __shared__ PrepCtl ctl;
  __shared__ unsigned warp_tot[kPrepThreads / 32];
  __shared__ double sT[12], sP[12];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int scan = blockIdx.y;
  const int tid = threadIdx.x;
  const long long base = offsets[scan];
  const int n = (int)(offsets[scan + 1] - base);
  const float4* pts = raw + base;
  // eighths rounded up to whole groups of 4 points (one Philox call serves 4 points)
  int chunk = ((n + 4 * kPrepCluster - 1) / (4 * kPrepCluster)) * 4;
  const bool too_big = chunk > chunk_cap;            // host sized the keys region from max_points
  if (too_big) chunk = 0;
  const int lo = rank * chunk, hi = min(n, lo + chunk);       // lo stays a multiple of 4 (may lie beyond n: empty slice)
  const uint2 pkey = make_uint2(seed_lo, seed_hi);

  for (int b = tid; b < kPrepBins; b += kPrepThreads) hist[b] = 0;
  if (tid < 12) {
    sT[tid] = Tr[(tr_per_scan ? (size_t)scan * 12 : 0) + tid];
    sP[tid] = post ? post[(size_t)scan * 12 + tid] : 0.0;
  }
  if (tid == 0) { ctl.count = 0; ctl.bcount = 0; ctl.bstar = kPrepBins; ctl.below = 0; ctl.survivors = 0; }
  __syncthreads();

  // ---- phase 1: stream the slice once
  for (int g = (lo >> 2) + tid; (g << 2) < hi; g += kPrepThreads) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)g, (unsigned)scan, 0u, 0u), pkey);
    const int i0 = g << 2;
    float4 p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j < hi) p[j] = __ldcs(pts + i0 + j);          // streaming: read once
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < hi) {
        double X, Y, Z;
        affine3x4(sT, (double)p[j].x, (double)p[j].y, (double)p[j].z, X, Y, Z);
        const unsigned k31 = pick4(r, j) >> 1;
        const bool keep = keep_point(X, Y, Z);
        keys[i0 + j - lo] = keep ? k31 : kPrepRejected;
        if (keep) atomicAdd(&hist[k31 >> 20], 1u);
      }
    }
  }
  cluster.sync();

  // ---- phase 2: CTA 0 reduces the histograms over the cluster and locates the boundary bin
  if (rank == 0) {
    unsigned c0 = 0, c1 = 0;
    for (int r = 0; r < kPrepCluster; ++r) {
      const unsigned* h = cluster.map_shared_rank(hist, r);
      c0 += h[2 * tid];
      c1 += h[2 * tid + 1];
    }
    // exclusive scan of (c0 + c1) over the 1024 threads
    const unsigned mine = c0 + c1;
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(PWCLO_FULL_MASK, incl, o);
      if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      unsigned w = warp_tot[tid], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(PWCLO_FULL_MASK, wi, o);
        if (tid >= o) wi += v;
      }
      warp_tot[tid] = wi - w;                                  // exclusive warp offsets
      if (tid == 31) ctl.survivors = (int)wi;
    }
    __syncthreads();
    const unsigned excl = warp_tot[tid >> 5] + incl - mine;
    const unsigned want = (unsigned)npoints;
    if (excl < want && excl + mine >= want) {                  // exactly one thread when M >= npoints
      if (excl + c0 >= want) { ctl.bstar = 2 * tid; ctl.below = (int)excl; }
      else { ctl.bstar = 2 * tid + 1; ctl.below = (int)(excl + c0); }
    }
    __syncthreads();
  }
  cluster.sync();
  PrepCtl* ctl0 = cluster.map_shared_rank(&ctl, 0);
  const int bstar = ctl0->bstar;
  const bool short_mode = bstar == kPrepBins;                  // fewer survivors than npoints: keep them all, by index

  // ---- phase 3: append the selected keys to CTA 0's lists
  {
    unsigned long long* list0 = cluster.map_shared_rank(list, 0);
    unsigned long long* blist0 = cluster.map_shared_rank(blist, 0);
    for (int j = tid; j < hi - lo; j += kPrepThreads) {
      const unsigned k = keys[j];
      if (k == kPrepRejected) continue;
      const int bin = (int)(k >> 20);
      const unsigned idx = (unsigned)(lo + j);
      if (short_mode) {
        const int pos = atomicAdd(&ctl0->count, 1);
        if (pos < list_cap) list0[pos] = (unsigned long long)idx;
      } else if (bin < bstar) {
        const int pos = atomicAdd(&ctl0->count, 1);
        if (pos < list_cap) list0[pos] = ((unsigned long long)k << 32) | idx;
      } else if (bin == bstar) {
        const int pos = atomicAdd(&ctl0->bcount, 1);
        if (pos < kPrepBoundaryCap) blist0[pos] = ((unsigned long long)k << 32) | idx;
      }
    }
  }
  cluster.sync();
  if (rank != 0) return;          // nothing reads this CTA's shared memory any more

  // ---- phase 4 (CTA 0): complete, sort, gather
  const int M = ctl.survivors;
  int have = min(ctl.count, list_cap);
  bool overflow = too_big || ctl.count > list_cap;
  if (!short_mode) {
    const int bc = ctl.bcount;
    if (bc > kPrepBoundaryCap) overflow = true;
    const int bn = min(bc, kPrepBoundaryCap);
    for (int j = bn + tid; j < kPrepBoundaryCap; j += kPrepThreads) blist[j] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(blist, kPrepBoundaryCap);
    const int need = min(npoints - have, bn);
    for (int j = tid; j < need; j += kPrepThreads) list[have + j] = blist[j];
    have += max(need, 0);
  }
  int p2 = 1;
  while (p2 < max(have, 2)) p2 <<= 1;
  for (int j = have + tid; j < p2; j += kPrepThreads) list[j] = ~0ull;
  __syncthreads();
  bitonic_sort_u64(list, p2);

  if (tid == 0 && survivors) survivors[scan] = overflow ? -1 : M;
  for (int j = tid; j < npoints; j += kPrepThreads) {
    int idx;
    if (j < have) {
      idx = (int)(unsigned)(list[j] & 0xffffffffull);
    } else {
      // pad with draws with replacement (from the survivors, or from the whole scan when there are none)
      const int d = j - have;
      const uint4 r = philox4x32_10(make_uint4((unsigned)(d >> 2), (unsigned)scan, 1u, 0u), pkey);
      const unsigned u = pick4(r, d & 3);
      idx = have > 0 ? (int)(unsigned)(list[__umulhi(u, (unsigned)have)] & 0xffffffffull) : (n > 0 ? (int)__umulhi(u, (unsigned)n) : -1);
    }
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (idx >= 0) {
      const float4 p = pts[idx];
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

PWCLO_API int pwclo_prepare_scans(const float* raw, const long long* offsets, int nscan, int max_points, const double* Tr,
                                  int tr_per_scan, const double* post, unsigned long long seed, int npoints, float* out,
                                  int32_t* sel_idx, int32_t* survivors, void* stream) {
  using namespace pwclo;
  if (!raw || !offsets || !Tr || !out || nscan < 0 || npoints <= 0 || max_points <= 0) return PWCLO_EINVAL;
  if (((uintptr_t)raw & 15) != 0) return PWCLO_EINVAL;
  if (nscan == 0) return PWCLO_OK;
  int chunk_cap = ((max_points + 4 * kPrepCluster - 1) / (4 * kPrepCluster)) * 4;
  int list_cap = 2;
  while (list_cap < npoints) list_cap <<= 1;
  const size_t smem = (size_t)kPrepBins * 4 + (size_t)kPrepBoundaryCap * 8 + (size_t)list_cap * 8 + (size_t)chunk_cap * 4;
  if (smem > 227 * 1024 - 1024) return PWCLO_EUNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(scan_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dim3 grid(kPrepCluster, nscan, 1);
  scan_prepare_kernel<<<grid, kPrepThreads, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(raw), offsets, Tr, tr_per_scan ? 1 : 0, post, (unsigned)(seed & 0xffffffffull),
      (unsigned)(seed >> 32), npoints, chunk_cap, list_cap, out, sel_idx, survivors);
  return launch_status();
}
