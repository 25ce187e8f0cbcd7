// Fused PWCLO-Net layer kernels (inference, BatchNorm folded) for sm_100a.
//
//   pwclo_set_conv      PointnetSAModulePWCLONet / PointnetFPModulePWCLONet.mlp  (P2/pointnet2_modules.py:179-245, :480-506)
//   pwclo_pointwise_mlp FlowPredictor / set-upconv post_mlp                      (PW/flowpredictor.py:53-84, pointnet2_modules.py:508-515)
//   pwclo_cost_volume_1 CostVolume, first (point-to-patch) aggregation           (PW/costvolume.py:75-144)
//   pwclo_cost_volume_2 CostVolume, second (patch-to-patch) aggregation          (PW/costvolume.py:150-188)
//   pwclo_pose_head     softmax-over-points mask, PoseCalculator, pose composition (PW/pose_calculator.py:47-87,
//                       PW/pose_warp_refinement.py:120-148, PW/pwclo_net.py:172-205)
//
// Data layout in HBM: coordinates [B,N,3], features POINT-MAJOR [B,N,C] (one neighbour = one
// contiguous 64..512-byte row, fetched with float4 loads), neighbour indices [B,S,K] int32.
// One CTA owns R = 64 or 128 rows = (R/K points) x (K neighbours); grouped tensors exist only in
// shared memory.  See mlp_core.cuh for the GEMM/weight-streaming machinery.
#include <math_constants.h>

#include <cstdlib>

#include "mlp_core.cuh"

namespace pwclo {

struct Smem {
  // carve-out helper: 128-byte aligned slices of the dynamic shared memory window
  unsigned char* p;
  __device__ explicit Smem(unsigned char* base) : p(base) {}
  template <typename T>
  __device__ T* take(size_t n) {
    T* r = reinterpret_cast<T*>(p);
    p += (n * sizeof(T) + 127) & ~(size_t)127;
    return r;
  }
};
static inline size_t al128(size_t bytes) { return (bytes + 127) & ~(size_t)127; }

// ------------------------------------------------------------------------------------------------
// set conv / set upconv:  X = [feat(C) | xyz_nbr - xyz_ctr (3)]  -> 2 or 3 layers -> max over K
// ------------------------------------------------------------------------------------------------
struct SAArgs {
  const float* xyz;      // [B,N,3] reference set
  const float* feats;    // [B,N,C] or nullptr: the neighbour's absolute xyz is the feature (C = 3)
  const float* new_xyz;  // [B,S,3] centres
  const int32_t* idx;    // [B,S,K]
  float* out;            // [B,S,cout_last]
  int N, S, K, C, nlayers, ldA, ldB;
  Layer l[3];
};

template <int R>
__global__ void __launch_bounds__(LT) set_conv_kernel(const SAArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem sm(smem_raw);
  float* bufA = sm.take<float>((size_t)R * a.ldA);
  float* bufB = sm.take<float>((size_t)R * a.ldB);
  float* wbuf = sm.take<float>(2 * KC * CBMAX);
  int* idx_s = sm.take<int>(R);
  float* ctr = sm.take<float>(3 * R);
  uint64_t* bars = sm.take<uint64_t>(2);

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int P = R / a.K;
  const int p0 = blockIdx.x * P;
  WeightPipe pipe;
  pipe.init(wbuf, bars);
  __syncthreads();
  pipe.prime(a.l[0]);

  for (int r = tid; r < R; r += LT) {
    const int p = r / a.K, k = r - p * a.K;
    const bool valid = p < P && p0 + p < a.S;
    idx_s[r] = valid ? a.idx[((size_t)b * a.S + p0 + p) * a.K + k] : 0;
  }
  for (int i = tid; i < P * 3; i += LT) {
    const int p = min(p0 + i / 3, a.S - 1);
    ctr[i] = a.new_xyz[((size_t)b * a.S + p) * 3 + i % 3];
  }
  __syncthreads();
  const int C = a.C, k4 = a.l[0].k4;
  for (int r = tid; r < R; r += LT) {
    const int p = min(r / a.K, P - 1);
    const float* q = a.xyz + ((size_t)b * a.N + idx_s[r]) * 3;
    float* x = bufA + (size_t)r * a.ldA;
    const float qx = q[0], qy = q[1], qz = q[2];
    if (a.feats == nullptr) { x[0] = qx; x[1] = qy; x[2] = qz; }
    x[C + 0] = __fsub_rn(qx, ctr[p * 3 + 0]);
    x[C + 1] = __fsub_rn(qy, ctr[p * 3 + 1]);
    x[C + 2] = __fsub_rn(qz, ctr[p * 3 + 2]);
    for (int c = C + 3; c < k4; ++c) x[c] = 0.f;
  }
  if (a.feats != nullptr) {
    const int C4 = C >> 2;
    for (int e = tid; e < R * C4; e += LT) {
      const int r = e / C4, c4 = e - r * C4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.feats + ((size_t)b * a.N + idx_s[r]) * C) + c4);
      *reinterpret_cast<float4*>(bufA + (size_t)r * a.ldA + 4 * c4) = v;
    }
  }
  __syncthreads();

  gemm_layer<R, true>(bufA, a.ldA, bufB, a.ldB, 0, a.l[0], &a.l[1], pipe);
  const float* res;
  int ldr;
  if (a.nlayers == 2) {
    gemm_layer<R, true>(bufB, a.ldB, bufA, a.ldA, 0, a.l[1], nullptr, pipe);
    res = bufA; ldr = a.ldA;
  } else {
    gemm_layer<R, true>(bufB, a.ldB, bufA, a.ldA, 0, a.l[1], &a.l[2], pipe);
    gemm_layer<R, true>(bufA, a.ldA, bufB, a.ldB, 0, a.l[2], nullptr, pipe);
    res = bufB; ldr = a.ldB;
  }
  const int co = a.l[a.nlayers - 1].cout;
  for (int e = tid; e < P * co; e += LT) {
    const int p = e / co, c = e - p * co;
    if (p0 + p >= a.S) continue;
    const float* y = res + (size_t)(p * a.K) * ldr + c;
    float m = y[0];
    for (int k = 1; k < a.K; ++k) m = fmaxf(m, y[(size_t)k * ldr]);
    a.out[((size_t)b * a.S + p0 + p) * co + c] = m;
  }
}


// ------------------------------------------------------------------------------------------------
// narrow set conv (pyramid levels 1-3: widths <= 64).  One lane per (point, neighbour) row: the row's
// inputs and all intermediate activations live in registers, weights are broadcast-read from shared
// memory, and the max over the K neighbours is a log-step shuffle transpose-reduce.  No activation
// ever touches shared memory; the GEMM-tile kernel above wastes most of its tile on these widths.
// ------------------------------------------------------------------------------------------------
template <int R, int CIN4, int COUT>
__device__ __forceinline__ void small_layer(const float (&x)[R][CIN4], float (&y)[R][COUT], const float* __restrict__ w,
                                            const float* __restrict__ b) {
  // accumulators live as packed pairs (channels n, n+1) for the whole k loop; a broadcast LDS.128 of four weights is two
  // ready-made operand pairs, the activation is duplicated once per (row, k) and feeds COUT / 2 FFMA2
  f32x2_t acc[R][COUT / 2];
#pragma unroll
  for (int n = 0; n < COUT; n += 4) {
    const float4 bv = *reinterpret_cast<const float4*>(b + n);
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r][n / 2] = f2_pack(bv.x, bv.y); acc[r][n / 2 + 1] = f2_pack(bv.z, bv.w); }
  }
#pragma unroll
  for (int k = 0; k < CIN4; ++k) {
    f32x2_t xx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) xx[r] = f2_pack(x[r][k], x[r][k]);
#pragma unroll
    for (int n = 0; n < COUT; n += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(w + k * COUT + n);   // one broadcast LDS.128 feeds 2 R FFMA2
      const f32x2_t w01 = f2_pack(wv.x, wv.y), w23 = f2_pack(wv.z, wv.w);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][n / 2] = f2_fma(xx[r], w01, acc[r][n / 2]);
        acc[r][n / 2 + 1] = f2_fma(xx[r], w23, acc[r][n / 2 + 1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int n = 0; n < COUT; n += 2) {
      float lo, hi;
      f2_unpack(acc[r][n / 2], lo, hi);
      y[r][n] = fmaxf(lo, 0.f);
      y[r][n + 1] = fmaxf(hi, 0.f);
    }
}

// R rows per lane (the same neighbour slot of R different points): each weight read is shared by R rows,
// which halves the shared-memory traffic that bounds the one-row version at ~1/3 of the FFMA peak.
template <int K, int C, int C1, int C2, int C3, int R>
__global__ void __launch_bounds__(256) set_conv_small_kernel(const SAArgs a, int total_points) {
  constexpr int CIN4 = (C + 3 + 3) & ~3;
  constexpr int PPW = 32 / K;              // points per warp and row slot
  __shared__ __align__(16) float w1[CIN4 * C1], w2[C1 * C2], w3[C2 * C3], b1[C1], b2[C2], b3[C3];
  for (int i = threadIdx.x; i < CIN4 * C1; i += 256) w1[i] = a.l[0].w[i];
  for (int i = threadIdx.x; i < C1 * C2; i += 256) w2[i] = a.l[1].w[i];
  for (int i = threadIdx.x; i < C2 * C3; i += 256) w3[i] = a.l[2].w[i];
  for (int i = threadIdx.x; i < C1; i += 256) b1[i] = a.l[0].b[i];
  for (int i = threadIdx.x; i < C2; i += 256) b2[i] = a.l[1].b[i];
  for (int i = threadIdx.x; i < C3; i += 256) b3[i] = a.l[2].b[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane / K, k = lane % K;
  const int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * 8;
#pragma unroll 1
  for (int g0 = warp_global * PPW * R; g0 < total_points; g0 += nwarps * PPW * R) {
    asm volatile("" ::: "memory");                    // keep the (loop-invariant) weights in shared memory, not registers
    float x[R][CIN4];
    int gr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int g = min(g0 + r * PPW + sub, total_points - 1);    // (cloud, point) flattened
      gr[r] = g;
      const int b = g / a.S;
      const int n = a.idx[(size_t)g * K + k];
      const float* q = a.xyz + ((size_t)b * a.N + n) * 3;
      const float* ctr = a.new_xyz + (size_t)g * 3;
      const float qx = q[0], qy = q[1], qz = q[2];
      if (a.feats == nullptr) {
        x[r][0] = qx; x[r][1] = qy; x[r][2] = qz;
      } else {
        const float4* f = reinterpret_cast<const float4*>(a.feats + ((size_t)b * a.N + n) * C);
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
          const float4 v = __ldg(f + c);
          x[r][4 * c] = v.x; x[r][4 * c + 1] = v.y; x[r][4 * c + 2] = v.z; x[r][4 * c + 3] = v.w;
        }
      }
      x[r][C] = __fsub_rn(qx, ctr[0]); x[r][C + 1] = __fsub_rn(qy, ctr[1]); x[r][C + 2] = __fsub_rn(qz, ctr[2]);
#pragma unroll
      for (int c = C + 3; c < CIN4; ++c) x[r][c] = 0.f;
    }
    float h1[R][C1], h2[R][C2], v[R][C3];
    small_layer<R, CIN4, C1>(x, h1, w1, b1);
    small_layer<R, C1, C2>(h1, h2, w2, b2);
    small_layer<R, C2, C3>(h2, v, w3, b3);
    // max over the K lanes of a point: transpose-reduce, every step halves the live values per lane
    int nlive = C3, chan = 0;
#pragma unroll
    for (int off = K / 2; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
      if (nlive > 1) {
        const int half = nlive / 2;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < C3 / 2; ++i) {
            if (i < half) {
              const float send = upper ? v[r][i] : v[r][i + half];
              const float keep = upper ? v[r][i + half] : v[r][i];
              v[r][i] = fmaxf(keep, __shfl_xor_sync(PWCLO_FULL_MASK, send, off));
            }
          }
        chan += upper ? half : 0;
        nlive = half;
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r][0] = fmaxf(v[r][0], __shfl_xor_sync(PWCLO_FULL_MASK, v[r][0], off));
      }
    }
    // lane now holds `nlive` consecutive channels starting at `chan`
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (g0 + r * PPW + sub < total_points) {
        float* o = a.out + (size_t)gr[r] * C3;
        if (nlive == 1) {
          // when K >= C3 several lanes hold the same channel: let the lowest one write
          const int dup = K / C3 > 1 ? K / C3 : 1;        // lanes per channel
          if ((k % dup) == 0 || C3 >= K) o[chan] = v[r][0];
        } else {
#pragma unroll
          for (int i = 0; i < C3; ++i)
            if (i < nlive) o[chan + i] = v[r][i];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// point-wise MLP on the concatenation of up to three point-major tensors (1 or 2 layers)
// ------------------------------------------------------------------------------------------------
struct PWArgs {
  const float* src[3];
  int c[3];
  int nsrc, rows, nlayers, ldA, ldB;
  float* out;  // [rows, cout_last]
  Layer l[2];
};

template <int R>
__global__ void __launch_bounds__(LT) pointwise_kernel(const PWArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem sm(smem_raw);
  float* bufA = sm.take<float>((size_t)R * a.ldA);
  float* bufB = sm.take<float>((size_t)R * a.ldB);
  float* wbuf = sm.take<float>(2 * KC * CBMAX);
  uint64_t* bars = sm.take<uint64_t>(2);
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * R;
  WeightPipe pipe;
  pipe.init(wbuf, bars);
  __syncthreads();
  pipe.prime(a.l[0]);

  int col = 0;
  for (int s = 0; s < a.nsrc; ++s) {
    const int C4 = a.c[s] >> 2;
    for (int e = tid; e < R * C4; e += LT) {
      const int r = e / C4, c4 = e - r * C4;
      const int gr = min(r0 + r, a.rows - 1);
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.src[s] + (size_t)gr * a.c[s]) + c4);
      *reinterpret_cast<float4*>(bufA + (size_t)r * a.ldA + col + 4 * c4) = v;
    }
    col += a.c[s];
  }
  __syncthreads();
  gemm_layer<R, true>(bufA, a.ldA, bufB, a.ldB, 0, a.l[0], a.nlayers == 2 ? &a.l[1] : nullptr, pipe);
  const float* res = bufB;
  int ldr = a.ldB;
  if (a.nlayers == 2) {
    gemm_layer<R, true>(bufB, a.ldB, bufA, a.ldA, 0, a.l[1], nullptr, pipe);
    res = bufA; ldr = a.ldA;
  }
  const int co = a.l[a.nlayers - 1].cout, co4 = co >> 2;
  for (int e = tid; e < R * co4; e += LT) {
    const int r = e / co4, c4 = e - r * co4;
    if (r0 + r >= a.rows) continue;
    *reinterpret_cast<float4*>(a.out + (size_t)(r0 + r) * co + 4 * c4) =
        *reinterpret_cast<const float4*>(res + (size_t)r * ldr + 4 * c4);
  }
}

// ------------------------------------------------------------------------------------------------
// cost volume, first aggregation.  Rows = (point of frame 1, one of its K neighbours in frame 2).
//   X  = [geo(10)+2 zero | f1(C) | f2 nbr(C)]         geo = (p, q, q-p, |q-p|)
//   h  = mlp1(X) (->128->64->64),  enc = conv(geo) (->64),  a = mlp2([enc | h]) (->128->64)
//   out[p] = sum_k softmax_k(a)[k] * h[k]
// ------------------------------------------------------------------------------------------------
struct CV1Args {
  const float* wxyz;  // [B,S,3] (warped) frame-1 coordinates
  const float* f1;    // [B,S,C]
  const float* xyz2;  // [B,N,3]
  const float* f2;    // [B,N,C]
  const int32_t* idx; // [B,S,K] neighbours in frame 2
  float* out;         // [B,S,64]
  int S, N, K, C, ldX, ldH;
  Layer m1[3], enc, m2[2];
};

__device__ __forceinline__ void write_geo(float* x, float px, float py, float pz, float qx, float qy, float qz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  // torch.sum over the (strided) channel axis accumulates sequentially: (dx2 + dy2) + dz2, then + 1e-20, sqrt
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  x[0] = px; x[1] = py; x[2] = pz; x[3] = qx; x[4] = qy; x[5] = qz; x[6] = dx; x[7] = dy; x[8] = dz;
  x[9] = __fsqrt_rn(__fadd_rn(n2, 1e-20f));
  x[10] = 0.f; x[11] = 0.f;
}

// softmax over the K rows of a point for channel c of `att`, weighted sum of `val`
__device__ __forceinline__ float softmax_pool(const float* att, int ld_att, const float* val, int ld_val, int K) {
  float m = att[0];
  for (int k = 1; k < K; ++k) m = fmaxf(m, att[(size_t)k * ld_att]);
  float z = 0.f, s = 0.f;
  for (int k = 0; k < K; ++k) {
    const float e = expf(att[(size_t)k * ld_att] - m);
    z += e;
    s = fmaf(e, val[(size_t)k * ld_val], s);
  }
  return s / z;
}

template <int R>
__global__ void __launch_bounds__(LT) cost_volume1_kernel(const CV1Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem sm(smem_raw);
  float* bufX = sm.take<float>((size_t)R * a.ldX);
  float* bufH = sm.take<float>((size_t)R * a.ldH);
  float* wbuf = sm.take<float>(2 * KC * CBMAX);
  int* idx_s = sm.take<int>(R);
  uint64_t* bars = sm.take<uint64_t>(2);
  const int tid = threadIdx.x, b = blockIdx.y;
  const int P = R / a.K, p0 = blockIdx.x * P, C = a.C;
  WeightPipe pipe;
  pipe.init(wbuf, bars);
  __syncthreads();
  pipe.prime(a.m1[0]);

  for (int r = tid; r < R; r += LT) {
    const int p = r / a.K, k = r - p * a.K;
    const bool valid = p < P && p0 + p < a.S;
    const int gp = min(p0 + min(p, P - 1), a.S - 1);
    const int n = valid ? a.idx[((size_t)b * a.S + gp) * a.K + k] : 0;
    idx_s[r] = n;
    const float* pp = a.wxyz + ((size_t)b * a.S + gp) * 3;
    const float* qq = a.xyz2 + ((size_t)b * a.N + n) * 3;
    write_geo(bufX + (size_t)r * a.ldX, pp[0], pp[1], pp[2], qq[0], qq[1], qq[2]);
  }
  __syncthreads();
  const int C4 = C >> 2;
  for (int e = tid; e < R * C4; e += LT) {
    const int r = e / C4, c4 = e - r * C4;
    const int gp = min(p0 + min(r / a.K, P - 1), a.S - 1);
    const float4 u = __ldg(reinterpret_cast<const float4*>(a.f1 + ((size_t)b * a.S + gp) * C) + c4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.f2 + ((size_t)b * a.N + idx_s[r]) * C) + c4);
    float* x = bufX + (size_t)r * a.ldX + 12 + 4 * c4;
    *reinterpret_cast<float4*>(x) = u;
    *reinterpret_cast<float4*>(x + C) = v;
  }
  __syncthreads();

  gemm_layer<R, true>(bufX, a.ldX, bufH, a.ldH, 0, a.m1[0], &a.m1[1], pipe);        // h1 -> H[0:128)
  gemm_layer<R, true>(bufH, a.ldH, bufX, a.ldX, 12, a.m1[1], &a.m1[2], pipe);       // h2 -> X[12:76)
  gemm_layer<R, true>(bufX + 12, a.ldX, bufH, a.ldH, 64, a.m1[2], &a.enc, pipe);    // h3 -> H[64:128)
  gemm_layer<R, true>(bufX, a.ldX, bufH, a.ldH, 0, a.enc, &a.m2[0], pipe);          // enc -> H[0:64)
  gemm_layer<R, true>(bufH, a.ldH, bufX, a.ldX, 0, a.m2[0], &a.m2[1], pipe);        // a1 -> X[0:128)
  gemm_layer<R, true>(bufX, a.ldX, bufH, a.ldH, 0, a.m2[1], nullptr, pipe);         // a2 -> H[0:64)

  for (int e = tid; e < P * 64; e += LT) {
    const int p = e >> 6, c = e & 63;
    if (p0 + p >= a.S) continue;
    const float* base = bufH + (size_t)(p * a.K) * a.ldH;
    a.out[((size_t)b * a.S + p0 + p) * 64 + c] = softmax_pool(base + c, a.ldH, base + 64 + c, a.ldH, a.K);
  }
}

// ------------------------------------------------------------------------------------------------
// cost volume, second aggregation.  Rows = (point p, one of its K self-neighbours n in frame 1).
//   X = [enc2(geo2) (64) | f1[p] (C) | e1[n] (64)] -> mlp3 (->128->64) -> softmax over K -> sum_k w * e1[n]
// ------------------------------------------------------------------------------------------------
struct CV2Args {
  const float* wxyz;  // [B,S,3]
  const float* f1;    // [B,S,C]
  const float* e1;    // [B,S,64] output of the first aggregation
  const int32_t* idx; // [B,S,K] self neighbours
  float* out;         // [B,S,64]
  int S, K, C, ldX, ldH;
  Layer enc, m3[2];
};

template <int R>
__global__ void __launch_bounds__(LT) cost_volume2_kernel(const CV2Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem sm(smem_raw);
  float* bufX = sm.take<float>((size_t)R * a.ldX);
  float* bufH = sm.take<float>((size_t)R * a.ldH);
  float* bufG = sm.take<float>((size_t)R * 12);
  float* wbuf = sm.take<float>(2 * KC * CBMAX);
  int* idx_s = sm.take<int>(R);
  uint64_t* bars = sm.take<uint64_t>(2);
  const int tid = threadIdx.x, b = blockIdx.y;
  const int P = R / a.K, p0 = blockIdx.x * P, C = a.C;
  WeightPipe pipe;
  pipe.init(wbuf, bars);
  __syncthreads();
  pipe.prime(a.enc);

  for (int r = tid; r < R; r += LT) {
    const int p = r / a.K, k = r - p * a.K;
    const bool valid = p < P && p0 + p < a.S;
    const int gp = min(p0 + min(p, P - 1), a.S - 1);
    const int n = valid ? a.idx[((size_t)b * a.S + gp) * a.K + k] : 0;
    idx_s[r] = n;
    const float* pp = a.wxyz + ((size_t)b * a.S + gp) * 3;
    const float* qq = a.wxyz + ((size_t)b * a.S + n) * 3;
    write_geo(bufG + (size_t)r * 12, pp[0], pp[1], pp[2], qq[0], qq[1], qq[2]);
  }
  __syncthreads();
  const int C4 = C >> 2;
  for (int e = tid; e < R * C4; e += LT) {
    const int r = e / C4, c4 = e - r * C4;
    const int gp = min(p0 + min(r / a.K, P - 1), a.S - 1);
    *reinterpret_cast<float4*>(bufX + (size_t)r * a.ldX + 64 + 4 * c4) =
        __ldg(reinterpret_cast<const float4*>(a.f1 + ((size_t)b * a.S + gp) * C) + c4);
  }
  for (int e = tid; e < R * 16; e += LT) {
    const int r = e >> 4, c4 = e & 15;
    *reinterpret_cast<float4*>(bufX + (size_t)r * a.ldX + 64 + C + 4 * c4) =
        __ldg(reinterpret_cast<const float4*>(a.e1 + ((size_t)b * a.S + idx_s[r]) * 64) + c4);
  }
  __syncthreads();

  gemm_layer<R, true>(bufG, 12, bufX, a.ldX, 0, a.enc, &a.m3[0], pipe);             // enc2 -> X[0:64)
  gemm_layer<R, true>(bufX, a.ldX, bufH, a.ldH, 0, a.m3[0], &a.m3[1], pipe);        // a1 -> H[0:128)
  gemm_layer<R, true>(bufH, a.ldH, bufX, a.ldX, 0, a.m3[1], nullptr, pipe);         // a2 -> X[0:64)

  for (int e = tid; e < P * 64; e += LT) {
    const int p = e >> 6, c = e & 63;
    if (p0 + p >= a.S) continue;
    const float* base = bufX + (size_t)(p * a.K) * a.ldX;
    a.out[((size_t)b * a.S + p0 + p) * 64 + c] = softmax_pool(base + c, a.ldX, base + 64 + C + c, a.ldX, a.K);
  }
}

// ------------------------------------------------------------------------------------------------
// pose head: W = softmax over points of the mask; s = sum_n emb*W; conv1d 64->256 (+b); q = 256->4,
// normalised; t = 256->3; optional composition with the coarse pose; writes the row of pose_params.
// One CTA per batch element.
// ------------------------------------------------------------------------------------------------
struct PoseArgs {
  const float* emb;    // [B,S,64]
  const float* mask;   // [B,S,64]
  const float* wqt; const float* bqt;   // [256,64], [256]
  const float* wq; const float* bq;     // [4,256], [4]
  const float* wt; const float* bt;     // [3,256], [3]
  const float* coarse; // [B,7] (q,t) of the coarser level or nullptr
  float* qt_out;       // [B,7] refined (q, t), q NOT re-normalised (as the reference propagates it)
  float* pose_row;     // pose_params + level*7, row stride 28 floats: (t, q/|q|)
  int S;
};

constexpr int POSE_PARTS = 16;   // 1024 threads: 16 slices of the point axis x 64 channels
__global__ void __launch_bounds__(POSE_PARTS * 64) pose_head_kernel(const PoseArgs a) {
  __shared__ float pm[POSE_PARTS][64], pz[POSE_PARTS][64], ps[POSE_PARTS][64];
  __shared__ float sv[64], big[256], qd[4], td[3];
  const int tid = threadIdx.x, b = blockIdx.x;
  const int c = tid & 63, part = tid >> 6;
  const float* E = a.emb + (size_t)b * a.S * 64;
  const float* M = a.mask + (size_t)b * a.S * 64;
  float m = -CUDART_INF_F, z = 0.f, s = 0.f;
  for (int n = part; n < a.S; n += POSE_PARTS) {
    const float v = M[(size_t)n * 64 + c];
    const float f = E[(size_t)n * 64 + c];
    if (v > m) { const float sc = expf(m - v); z *= sc; s *= sc; m = v; }
    const float e = expf(v - m);
    z += e;
    s = fmaf(e, f, s);
  }
  pm[part][c] = m; pz[part][c] = z; ps[part][c] = s;
  __syncthreads();
  if (tid < 64) {
    float mm = pm[0][tid];
    for (int i = 1; i < POSE_PARTS; ++i) mm = fmaxf(mm, pm[i][tid]);
    float zz = 0.f, ss = 0.f;
    for (int i = 0; i < POSE_PARTS; ++i) {
      const float sc = pz[i][tid] > 0.f ? expf(pm[i][tid] - mm) : 0.f;
      zz += pz[i][tid] * sc;
      ss += ps[i][tid] * sc;
    }
    sv[tid] = ss / zz;
  }
  __syncthreads();
  if (tid < 256) {
    float acc = a.bqt[tid];
    const float* w = a.wqt + (size_t)tid * 64;
    for (int k = 0; k < 64; ++k) acc = fmaf(w[k], sv[k], acc);
    big[tid] = acc;
  }
  __syncthreads();
  if (tid < 7 * 32) {  // 7 outputs, one warp each
    const int o = tid >> 5, lane = tid & 31;
    const float* w = o < 4 ? a.wq + (size_t)o * 256 : a.wt + (size_t)(o - 4) * 256;
    float acc = 0.f;
    for (int k = lane; k < 256; k += 32) acc = fmaf(w[k], big[k], acc);
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(PWCLO_FULL_MASK, acc, off);
    if (lane == 0) {
      if (o < 4) qd[o] = acc + a.bq[o];
      else td[o - 4] = acc + a.bt[o - 4];
    }
  }
  __syncthreads();
  if (tid == 0) {
    // q_det / (sqrt(sum q^2 + 1e-10) + 1e-10)     (PW/pose_calculator.py:69)
    float n = sqrtf(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3] + 1e-10f) + 1e-10f;
    float q0 = qd[0] / n, q1 = qd[1] / n, q2 = qd[2] / n, q3 = qd[3] / n;
    float t0 = td[0], t1 = td[1], t2 = td[2];
    if (a.coarse != nullptr) {
      const float* cq = a.coarse + (size_t)b * 7;
      // t = q_det (x) [0,t_coarse] (x) q_det^-1 + t_det    (PW/pose_warp_refinement.py:148)
      float qt[7] = {q0, q1, q2, q3, t0, t1, t2};
      PoseQT pq = make_pose(qt);
      warp_point(pq, cq[4], cq[5], cq[6], t0, t1, t2);
      // q = q_det (x) q_coarse                              (PW/pose_warp_refinement.py:139)
      const float c0 = cq[0], c1 = cq[1], c2 = cq[2], c3 = cq[3];
      const float r0 = q0 * c0 - q1 * c1 - q2 * c2 - q3 * c3;
      const float r1 = q0 * c1 + q1 * c0 + q2 * c3 - q3 * c2;
      const float r2 = q0 * c2 - q1 * c3 + q2 * c0 + q3 * c1;
      const float r3 = q0 * c3 + q1 * c2 - q2 * c1 + q3 * c0;
      q0 = r0; q1 = r1; q2 = r2; q3 = r3;
    }
    float* o = a.qt_out + (size_t)b * 7;
    o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3; o[4] = t0; o[5] = t1; o[6] = t2;
    // pose_params row: (t, q / (sqrt(sum q^2 + 1e-10) + 1e-10))   (PW/pwclo_net.py:195-205)
    n = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3 + 1e-10f) + 1e-10f;
    float* pr = a.pose_row + (size_t)b * 28;
    pr[0] = t0; pr[1] = t1; pr[2] = t2; pr[3] = q0 / n; pr[4] = q1 / n; pr[5] = q2 / n; pr[6] = q3 / n;
  }
}

// ------------------------------------------------------------------------------------------------
// small layout helpers
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows3_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx, int N, int M,
                                    float* __restrict__ out) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const float* s = xyz + ((size_t)b * N + idx[(size_t)b * M + j]) * 3;
  float* d = out + ((size_t)b * M + j) * 3;
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

// [B,C,N] -> [B,N,C] (C small) and back
__global__ void transpose_cn_kernel(const float* __restrict__ in, int C, int N, float* __restrict__ out, int to_point_major) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C * N) return;
  if (to_point_major) {
    const int n = e / C, c = e - n * C;
    out[((size_t)b * N + n) * C + c] = in[((size_t)b * C + c) * N + n];
  } else {
    const int c = e / N, n = e - c * N;
    out[((size_t)b * C + c) * N + n] = in[((size_t)b * N + n) * C + c];
  }
}

static Layer make_layer(const pwclo_layer_t& l) {
  Layer L;
  L.w = l.w; L.b = l.b; L.k4 = round4(l.cin); L.cout = l.cout;
  return L;
}
static bool layer_ok(const pwclo_layer_t& l) {
  return l.w && l.b && l.cin > 0 && (l.cout == 8 || l.cout == 16 || l.cout == 32 || (l.cout > 0 && l.cout % 64 == 0)) &&
         ((uintptr_t)l.w % 16 == 0) && ((uintptr_t)l.b % 16 == 0);
}

template <typename KernT, typename ArgsT>
static int launch_layer(KernT kern, const ArgsT& args, dim3 grid, size_t smem, cudaStream_t st) {
  if (smem > 227 * 1024) return PWCLO_EUNSUPPORTED;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<grid, LT, smem, st>>>(args);
  return launch_status();
}

constexpr size_t kRingBytes = 2 * KC * CBMAX * sizeof(float) + 128;

}  // namespace pwclo

using namespace pwclo;

PWCLO_API int pwclo_set_conv(const float* xyz, const float* feats, const float* new_xyz, const int32_t* idx, int B,
                             int N, int S, int K, int C, const pwclo_layer_t* layers, int nlayers, float* out,
                             void* stream) {
  if (!xyz || !new_xyz || !idx || !layers || !out || B < 0 || N <= 0 || S <= 0 || K <= 0) return PWCLO_EINVAL;
  if (nlayers < 2 || nlayers > 3 || K > 64) return PWCLO_EUNSUPPORTED;
  if (feats == nullptr) C = 3;
  if (feats != nullptr && (C % 4 != 0 || (uintptr_t)feats % 16 != 0)) return PWCLO_EUNSUPPORTED;
  if (B == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  SAArgs a;
  a.xyz = xyz; a.feats = feats; a.new_xyz = new_xyz; a.idx = idx; a.out = out;
  a.N = N; a.S = S; a.K = K; a.C = C; a.nlayers = nlayers;
  int widthA = 0, widthB = 0;
  for (int i = 0; i < nlayers; ++i) {
    if (!layer_ok(layers[i])) return PWCLO_EINVAL;
    a.l[i] = make_layer(layers[i]);
  }
  if (layers[0].cin != C + 3) return PWCLO_EINVAL;
  for (int i = 1; i < nlayers; ++i)
    if (layers[i].cin != layers[i - 1].cout) return PWCLO_EINVAL;
  widthA = max(a.l[0].k4, a.l[1].cout);
  widthB = max(a.l[0].cout, nlayers == 3 ? a.l[2].cout : 0);
  a.ldA = ld_for(widthA);
  a.ldB = ld_for(widthB);
  bool narrow = false;
  for (int i = 0; i < nlayers; ++i) narrow |= a.l[i].cout == 8;
  auto smem_for = [&](int R) {
    return al128((size_t)R * a.ldA * 4) + al128((size_t)R * a.ldB * 4) + kRingBytes + al128(R * 4) + al128(3 * R * 4) + 128;
  };
  cudaStream_t st = (cudaStream_t)stream;
  if (nlayers == 3 && !getenv("PWCLO_NO_SMALL_SA")) {   // register-resident kernel for the narrow pyramid levels
    const int c1 = a.l[0].cout, c2 = a.l[1].cout, c3 = a.l[2].cout;
    const int total = B * S;
#define SMALL_CASE(KK, CC, A1, A2, A3, RR)                                                               \
  if (K == KK && C == CC && c1 == A1 && c2 == A2 && c3 == A3) {                                          \
    const int ppw = 32 / KK * RR;                                                                        \
    const int blocks = min(ceil_div(total, 8 * ppw), kNumSM * 8);                                        \
    set_conv_small_kernel<KK, CC, A1, A2, A3, RR><<<blocks, 256, 0, st>>>(a, total);                     \
    return launch_status();                                                                              \
  }
    const bool one_row = getenv("PWCLO_SA_ONE_ROW") != nullptr;
    if (!one_row) {
      SMALL_CASE(32, 3, 8, 8, 16, 2)
      SMALL_CASE(32, 16, 16, 16, 32, 2)
    }
    SMALL_CASE(32, 3, 8, 8, 16, 1)
    SMALL_CASE(32, 16, 16, 16, 32, 1)
    SMALL_CASE(16, 32, 32, 32, 64, 1)
#undef SMALL_CASE
  }
  if (narrow || (smem_for(128) <= 100 * 1024 && K <= 128)) {
    dim3 grid(ceil_div(S, 128 / K), B);
    return launch_layer(set_conv_kernel<128>, a, grid, smem_for(128), st);
  }
  dim3 grid(ceil_div(S, 64 / K), B);
  return launch_layer(set_conv_kernel<64>, a, grid, smem_for(64), st);
}

PWCLO_API int pwclo_pointwise_mlp(const float* const* src, const int* channels, int nsrc, int rows,
                                  const pwclo_layer_t* layers, int nlayers, float* out, void* stream) {
  if (!src || !channels || !layers || !out || nsrc < 1 || nsrc > 3 || rows < 0) return PWCLO_EINVAL;
  if (nlayers < 1 || nlayers > 2) return PWCLO_EUNSUPPORTED;
  if (rows == 0) return PWCLO_OK;
  PWArgs a;
  int cin = 0;
  for (int s = 0; s < 3; ++s) { a.src[s] = nullptr; a.c[s] = 0; }
  for (int s = 0; s < nsrc; ++s) {
    if (!src[s] || channels[s] % 4 != 0 || (uintptr_t)src[s] % 16 != 0) return PWCLO_EUNSUPPORTED;
    a.src[s] = src[s]; a.c[s] = channels[s]; cin += channels[s];
  }
  a.nsrc = nsrc; a.rows = rows; a.nlayers = nlayers; a.out = out;
  for (int i = 0; i < nlayers; ++i) {
    if (!layer_ok(layers[i]) || layers[i].cout < 16) return PWCLO_EINVAL;
    a.l[i] = make_layer(layers[i]);
  }
  if (layers[0].cin != cin || (nlayers == 2 && layers[1].cin != layers[0].cout)) return PWCLO_EINVAL;
  a.ldA = ld_for(max(a.l[0].k4, nlayers == 2 ? a.l[1].cout : 0));
  a.ldB = ld_for(a.l[0].cout);
  const size_t smem = al128((size_t)64 * a.ldA * 4) + al128((size_t)64 * a.ldB * 4) + kRingBytes + 128;
  return launch_layer(pointwise_kernel<64>, a, dim3(ceil_div(rows, 64)), smem, (cudaStream_t)stream);
}

PWCLO_API int pwclo_cost_volume_1(const float* wxyz, const float* f1, const float* xyz2, const float* f2,
                                  const int32_t* idx, int B, int S, int N, int K, int C, const pwclo_layer_t* mlp1,
                                  const pwclo_layer_t* enc, const pwclo_layer_t* mlp2, float* out, void* stream) {
  if (!wxyz || !f1 || !xyz2 || !f2 || !idx || !mlp1 || !enc || !mlp2 || !out || B < 0 || S <= 0 || N <= 0 || K <= 0)
    return PWCLO_EINVAL;
  if (C % 4 != 0 || K > 64 || ((uintptr_t)f1 | (uintptr_t)f2) % 16 != 0) return PWCLO_EUNSUPPORTED;
  if (B == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  CV1Args a;
  a.wxyz = wxyz; a.f1 = f1; a.xyz2 = xyz2; a.f2 = f2; a.idx = idx; a.out = out;
  a.S = S; a.N = N; a.K = K; a.C = C;
  for (int i = 0; i < 3; ++i) { if (!layer_ok(mlp1[i])) return PWCLO_EINVAL; a.m1[i] = make_layer(mlp1[i]); }
  for (int i = 0; i < 2; ++i) { if (!layer_ok(mlp2[i])) return PWCLO_EINVAL; a.m2[i] = make_layer(mlp2[i]); }
  if (!layer_ok(*enc)) return PWCLO_EINVAL;
  a.enc = make_layer(*enc);
  // the kernel's buffer choreography is specific to the reference's widths (PW/costvolume.py:40-58)
  if (mlp1[0].cin != 2 * C + 10 || mlp1[0].cout != 128 || mlp1[1].cout != 64 || mlp1[2].cout != 64 || enc->cin != 10 ||
      enc->cout != 64 || mlp2[0].cin != 128 || mlp2[0].cout != 128 || mlp2[1].cout != 64)
    return PWCLO_EUNSUPPORTED;
  a.m1[0].k4 = 2 * C + 12;   // geo is stored as 12 columns (10 + 2 zero) in front of the features
  a.enc.k4 = 12;
  a.ldX = ld_for(max(2 * C + 12, 128));
  a.ldH = ld_for(128);
  auto smem_for = [&](int R) {
    return al128((size_t)R * a.ldX * 4) + al128((size_t)R * a.ldH * 4) + kRingBytes + al128(R * 4) + 128;
  };
  cudaStream_t st = (cudaStream_t)stream;
  if (getenv("PWCLO_CV_R128")) {
    return launch_layer(cost_volume1_kernel<128>, a, dim3(ceil_div(S, 128 / K), B), smem_for(128), st);
  }
  return launch_layer(cost_volume1_kernel<64>, a, dim3(ceil_div(S, 64 / K), B), smem_for(64), st);
}

PWCLO_API int pwclo_cost_volume_2(const float* wxyz, const float* f1, const float* e1, const int32_t* idx, int B,
                                  int S, int K, int C, const pwclo_layer_t* enc, const pwclo_layer_t* mlp3, float* out,
                                  void* stream) {
  if (!wxyz || !f1 || !e1 || !idx || !enc || !mlp3 || !out || B < 0 || S <= 0 || K <= 0) return PWCLO_EINVAL;
  if (C % 4 != 0 || K > 64 || ((uintptr_t)f1 | (uintptr_t)e1) % 16 != 0) return PWCLO_EUNSUPPORTED;
  if (B == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  CV2Args a;
  a.wxyz = wxyz; a.f1 = f1; a.e1 = e1; a.idx = idx; a.out = out;
  a.S = S; a.K = K; a.C = C;
  if (!layer_ok(*enc) || !layer_ok(mlp3[0]) || !layer_ok(mlp3[1])) return PWCLO_EINVAL;
  if (enc->cin != 10 || enc->cout != 64 || mlp3[0].cin != 128 + C || mlp3[0].cout != 128 || mlp3[1].cin != 128 ||
      mlp3[1].cout != 64)
    return PWCLO_EUNSUPPORTED;
  a.enc = make_layer(*enc);
  a.enc.k4 = 12;
  a.m3[0] = make_layer(mlp3[0]);
  a.m3[1] = make_layer(mlp3[1]);
  a.ldX = ld_for(128 + C);
  a.ldH = ld_for(128);
  auto smem_for = [&](int R) {
    return al128((size_t)R * a.ldX * 4) + al128((size_t)R * a.ldH * 4) + al128((size_t)R * 12 * 4) + kRingBytes +
           al128(R * 4) + 128;
  };
  cudaStream_t st = (cudaStream_t)stream;
  if (getenv("PWCLO_CV_R128")) {
    return launch_layer(cost_volume2_kernel<128>, a, dim3(ceil_div(S, 128 / K), B), smem_for(128), st);
  }
  return launch_layer(cost_volume2_kernel<64>, a, dim3(ceil_div(S, 64 / K), B), smem_for(64), st);
}

PWCLO_API int pwclo_pose_head(const float* emb, const float* mask, int B, int S, const float* wqt, const float* bqt,
                              const float* wq, const float* bq, const float* wt, const float* bt, const float* coarse_qt,
                              float* qt_out, float* pose_params, int level, void* stream) {
  if (!emb || !mask || !wqt || !bqt || !wq || !bq || !wt || !bt || !qt_out || !pose_params || B < 0 || S <= 0 ||
      level < 0 || level > 3)
    return PWCLO_EINVAL;
  if (B == 0) return PWCLO_OK;
  PoseArgs a;
  a.emb = emb; a.mask = mask; a.wqt = wqt; a.bqt = bqt; a.wq = wq; a.bq = bq; a.wt = wt; a.bt = bt;
  a.coarse = coarse_qt; a.qt_out = qt_out; a.pose_row = pose_params + level * 7; a.S = S;
  pose_head_kernel<<<B, POSE_PARTS * 64, 0, (cudaStream_t)stream>>>(a);
  return launch_status();
}

PWCLO_API int pwclo_gather_rows3(const float* xyz, const int32_t* idx, int B, int N, int M, float* out, void* stream) {
  if (!xyz || !idx || !out || B < 0 || N <= 0 || M < 0) return PWCLO_EINVAL;
  if (B == 0 || M == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  gather_rows3_kernel<<<dim3(ceil_div(M, 256), B), 256, 0, (cudaStream_t)stream>>>(xyz, idx, N, M, out);
  return launch_status();
}

PWCLO_API int pwclo_transpose(const float* in, int B, int C, int N, int to_point_major, float* out, void* stream) {
  if (!in || !out || B < 0 || C <= 0 || N <= 0) return PWCLO_EINVAL;
  if (B == 0) return PWCLO_OK;
  if (B > 65535) return PWCLO_EUNSUPPORTED;
  transpose_cn_kernel<<<dim3(ceil_div(C * N, 256), B), 256, 0, (cudaStream_t)stream>>>(in, C, N, out, to_point_major);
  return launch_status();
}
