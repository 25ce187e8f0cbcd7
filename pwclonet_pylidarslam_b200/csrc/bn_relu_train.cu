// bn_relu_train.cu -- train-mode BatchNorm + ReLU of the shared-MLP layers, forward and backward.
//
// Reference: every layer of a SharedMLP is nn.Conv2d(1x1, bias=False) -> nn.BatchNorm2d(eps=1e-5) -> ReLU
// (P2/pytorch_utils.py:52-167); in training BatchNorm uses the statistics of the batch over (B, S, K) and
// updates the running statistics with `momentum` (unbiased variance), which the reference leaves to cuDNN plus a
// separate ReLU kernel.  On the [B,C,S,K] tensors of this network cuDNN's spatial backward
// (bn_bw_1C11_kernel_new) runs at ~1.4 TB/s and was 26 % of the GPU time of a training step
// (profiles/r1q_train_step_launches_summary.txt).  Here: two bandwidth-shaped launches per direction, ReLU fused.
//   forward   stats  : per channel sum and sum of squares, double accumulators, grid (C, split)
//             apply  : mean / biased variance / invstd from the partials, running statistics (one CTA per
//                      channel), y = max(0, (x - mean) * invstd * gamma + beta)          12 B per element in total
//   backward  reduce : g = dy * [y > 0] (the mask is recomputed from x),  sum g,  sum g * xhat
//             apply  : dx = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma, dbeta   20 B per element
// x, y, dy, dx are [B, C, HW] contiguous (NCHW with HW = S*K, or NCL).
#include "common.cuh"

namespace pwclo {

constexpr int kBnThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(PWCLO_FULL_MASK, v, o);
  return v;
}

// block sum of two doubles; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b, double (*red)[kBnThreads / 32]) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; b = 0.0;
    for (int w = 0; w < kBnThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
  }
}

// the elements of channel c handled by CTA (c, s): vec4 indices [v_lo, v_hi) of the B*HW/4 vectors of the channel
// (HW % 4 == 0) or element indices when VEC == 1
template <int VEC>
__device__ __forceinline__ void cta_range(long long per_channel, int split, long long& lo, long long& hi) {
  const long long units = per_channel / VEC;
  const long long per = (units + split - 1) / split;
  lo = (long long)blockIdx.y * per;
  hi = lo + per < units ? lo + per : units;
}

template <int VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const float* __restrict__ x, int B, int C, int HW, int split, double* __restrict__ partial) {
  __shared__ double red[2][kBnThreads / 32];
  const int c = blockIdx.x;
  long long lo, hi;
  cta_range<VEC>((long long)B * HW, split, lo, hi);
  const int row = HW / VEC;
  double s = 0.0, q = 0.0;
  for (long long u = lo + threadIdx.x; u < hi; u += kBnThreads) {
    const long long b = u / row, r = u - b * row;
    const float* p = x + ((size_t)b * C + c) * HW + (size_t)r * VEC;
    if (VEC == 4) {
      const float4 v = *reinterpret_cast<const float4*>(p);
      s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      const float v = *p;
      s += v;
      q += (double)v * v;
    }
  }
  block_sum2(s, q, red);
  if (threadIdx.x == 0) {
    partial[((size_t)c * split + blockIdx.y) * 2 + 0] = s;
    partial[((size_t)c * split + blockIdx.y) * 2 + 1] = q;
  }
}

template <int VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_relu_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int B, int C,
                     int HW, int split, const double* __restrict__ partial, float eps, float momentum,
                     float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ y,
                     float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ float sh[2];
  const int c = blockIdx.x;
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < split; ++i) { s += partial[((size_t)c * split + i) * 2]; q += partial[((size_t)c * split + i) * 2 + 1]; }
    const double n = (double)B * HW;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    sh[0] = (float)mean;
    sh[1] = invstd;
    if (blockIdx.y == 0) {
      save_mean[c] = (float)mean;
      save_invstd[c] = invstd;
      if (running_mean) {        // nn.BatchNorm: running = (1 - momentum) * running + momentum * batch (unbiased variance)
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
  __syncthreads();
  const float mean = sh[0];
  const float scale = sh[1] * gamma[c], shift = beta[c];
  long long lo, hi;
  cta_range<VEC>((long long)B * HW, split, lo, hi);
  const int row = HW / VEC;
  for (long long u = lo + threadIdx.x; u < hi; u += kBnThreads) {
    const long long b = u / row, r = u - b * row;
    const size_t off = ((size_t)b * C + c) * HW + (size_t)r * VEC;
    if (VEC == 4) {
      float4 v = *reinterpret_cast<const float4*>(x + off);
      v.x = fmaxf(fmaf(v.x - mean, scale, shift), 0.f);
      v.y = fmaxf(fmaf(v.y - mean, scale, shift), 0.f);
      v.z = fmaxf(fmaf(v.z - mean, scale, shift), 0.f);
      v.w = fmaxf(fmaf(v.w - mean, scale, shift), 0.f);
      *reinterpret_cast<float4*>(y + off) = v;
    } else {
      y[off] = fmaxf(fmaf(x[off] - mean, scale, shift), 0.f);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                     int B, int C, int HW, int split, double* __restrict__ partial) {
  __shared__ double red[2][kBnThreads / 32];
  const int c = blockIdx.x;
  const float mean = save_mean[c], invstd = save_invstd[c];
  const float scale = invstd * gamma[c], shift = beta[c];
  long long lo, hi;
  cta_range<VEC>((long long)B * HW, split, lo, hi);
  const int row = HW / VEC;
  double sg = 0.0, sgx = 0.0;
  for (long long u = lo + threadIdx.x; u < hi; u += kBnThreads) {
    const long long b = u / row, r = u - b * row;
    const size_t off = ((size_t)b * C + c) * HW + (size_t)r * VEC;
    float xv[4], gv[4];
    if (VEC == 4) {
      const float4 a = *reinterpret_cast<const float4*>(x + off), g4 = *reinterpret_cast<const float4*>(dy + off);
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
      gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
    } else {
      xv[0] = x[off]; gv[0] = dy[off];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float d = xv[k] - mean;
      const float g = fmaf(d, scale, shift) > 0.f ? gv[k] : 0.f;      // same expression as the forward: same mask
      sg += g;
      sgx += (double)g * (double)(d * invstd);
    }
  }
  block_sum2(sg, sgx, red);
  if (threadIdx.x == 0) {
    partial[((size_t)c * split + blockIdx.y) * 2 + 0] = sg;
    partial[((size_t)c * split + blockIdx.y) * 2 + 1] = sgx;
  }
}

template <int VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                    int B, int C, int HW, int split, const double* __restrict__ partial, float* __restrict__ dx,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float sh[2];
  const int c = blockIdx.x;
  if (threadIdx.x == 0) {
    double sg = 0.0, sgx = 0.0;
    for (int i = 0; i < split; ++i) { sg += partial[((size_t)c * split + i) * 2]; sgx += partial[((size_t)c * split + i) * 2 + 1]; }
    const double n = (double)B * HW;
    sh[0] = (float)(sg / n);
    sh[1] = (float)(sgx / n);
    if (blockIdx.y == 0) {
      dgamma[c] = (float)sgx;
      dbeta[c] = (float)sg;
    }
  }
  __syncthreads();
  const float mg = sh[0], mgx = sh[1];
  const float mean = save_mean[c], invstd = save_invstd[c];
  const float scale = invstd * gamma[c], shift = beta[c];
  long long lo, hi;
  cta_range<VEC>((long long)B * HW, split, lo, hi);
  const int row = HW / VEC;
  for (long long u = lo + threadIdx.x; u < hi; u += kBnThreads) {
    const long long b = u / row, r = u - b * row;
    const size_t off = ((size_t)b * C + c) * HW + (size_t)r * VEC;
    float xv[4], gv[4], ov[4];
    if (VEC == 4) {
      const float4 a = *reinterpret_cast<const float4*>(x + off), g4 = *reinterpret_cast<const float4*>(dy + off);
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
      gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
    } else {
      xv[0] = x[off]; gv[0] = dy[off];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float d = xv[k] - mean;
      const float g = fmaf(d, scale, shift) > 0.f ? gv[k] : 0.f;
      ov[k] = scale * (g - mg - (d * invstd) * mgx);
    }
    if (VEC == 4) *reinterpret_cast<float4*>(dx + off) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    else dx[off] = ov[0];
  }
}

inline int bn_split(int B, int C, int HW) {
  // enough CTAs to fill the machine (4 per SM), at least ~2k elements each
  const long long per_channel = (long long)B * HW;
  long long want = (4LL * kNumSM + C - 1) / C;
  const long long most = (per_channel + 2047) / 2048;
  if (want > most) want = most;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

}  // namespace pwclo

PWCLO_API size_t pwclo_bn_relu_workspace_bytes(int B, int C, int HW) {
  if (B <= 0 || C <= 0 || HW <= 0) return 0;
  return (size_t)C * pwclo::bn_split(B, C, HW) * 2 * sizeof(double);
}

PWCLO_API int pwclo_bn_relu_train_fwd(const float* x, const float* gamma, const float* beta, int B, int C, int HW, float eps,
                                      float momentum, float* running_mean, float* running_var, float* y, float* save_mean,
                                      float* save_invstd, void* workspace, void* stream) {
  using namespace pwclo;
  if (!x || !gamma || !beta || !y || !save_mean || !save_invstd || !workspace || B <= 0 || C <= 0 || HW <= 0) return PWCLO_EINVAL;
  if ((running_mean == nullptr) != (running_var == nullptr)) return PWCLO_EINVAL;
  const int split = bn_split(B, C, HW);
  double* part = reinterpret_cast<double*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = HW % 4 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  dim3 grid(C, split);
  if (vec) {
    bn_stats_kernel<4><<<grid, kBnThreads, 0, st>>>(x, B, C, HW, split, part);
    bn_apply_relu_kernel<4><<<grid, kBnThreads, 0, st>>>(x, gamma, beta, B, C, HW, split, part, eps, momentum, running_mean,
                                                         running_var, y, save_mean, save_invstd);
  } else {
    bn_stats_kernel<1><<<grid, kBnThreads, 0, st>>>(x, B, C, HW, split, part);
    bn_apply_relu_kernel<1><<<grid, kBnThreads, 0, st>>>(x, gamma, beta, B, C, HW, split, part, eps, momentum, running_mean,
                                                         running_var, y, save_mean, save_invstd);
  }
  return launch_status();
}

PWCLO_API int pwclo_bn_relu_train_bwd(const float* x, const float* dy, const float* gamma, const float* beta,
                                      const float* save_mean, const float* save_invstd, int B, int C, int HW, float* dx,
                                      float* dgamma, float* dbeta, void* workspace, void* stream) {
  using namespace pwclo;
  if (!x || !dy || !gamma || !beta || !save_mean || !save_invstd || !dx || !dgamma || !dbeta || !workspace || B <= 0 || C <= 0 ||
      HW <= 0)
    return PWCLO_EINVAL;
  const int split = bn_split(B, C, HW);
  double* part = reinterpret_cast<double*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = HW % 4 == 0 && (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0;
  dim3 grid(C, split);
  if (vec) {
    bn_bwd_reduce_kernel<4><<<grid, kBnThreads, 0, st>>>(x, dy, gamma, beta, save_mean, save_invstd, B, C, HW, split, part);
    bn_bwd_apply_kernel<4><<<grid, kBnThreads, 0, st>>>(x, dy, gamma, beta, save_mean, save_invstd, B, C, HW, split, part, dx,
                                                        dgamma, dbeta);
  } else {
    bn_bwd_reduce_kernel<1><<<grid, kBnThreads, 0, st>>>(x, dy, gamma, beta, save_mean, save_invstd, B, C, HW, split, part);
    bn_bwd_apply_kernel<1><<<grid, kBnThreads, 0, st>>>(x, dy, gamma, beta, save_mean, save_invstd, B, C, HW, split, part, dx,
                                                        dgamma, dbeta);
  }
  return launch_status();
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of a 1x1 convolution with very few channels (the first set-conv layers: 6 -> 8 -> 8 on
// [B, C, S*K] with S*K*B ~ 5e5 positions): dW[o][c] = sum over positions of dy[b,o,p] * x[b,c,p].  As a GEMM this is
// M x N = 8 x 6 with K = 524 288: cuBLAS picks gemmSN_TN and needs 0.67 ms per call
// (profiles/r1w_train_step_launches_fusedbn_summary.txt).  Here every thread keeps the CO x CI products in registers
// over a strided slice of the positions (coalesced rows), a CTA reduces them, a second launch adds the CTA partials in
// a fixed order (deterministic).
namespace pwclo {

constexpr int kWgThreads = 256;

template <int CI, int CO>
__global__ void __launch_bounds__(kWgThreads)
conv1x1_wgrad_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, long long HW,
                             float* __restrict__ partial) {
  __shared__ float red[kWgThreads / 32][CI * CO];
  float acc[CO][CI];
#pragma unroll
  for (int o = 0; o < CO; ++o)
#pragma unroll
    for (int c = 0; c < CI; ++c) acc[o][c] = 0.f;
  const long long total = (long long)B * HW;
  for (long long p = (long long)blockIdx.x * kWgThreads + threadIdx.x; p < total; p += (long long)gridDim.x * kWgThreads) {
    const long long b = p / HW, r = p - b * HW;
    float xv[CI], gv[CO];
#pragma unroll
    for (int c = 0; c < CI; ++c) xv[c] = x[((size_t)b * CI + c) * HW + r];
#pragma unroll
    for (int o = 0; o < CO; ++o) gv[o] = dy[((size_t)b * CO + o) * HW + r];
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int c = 0; c < CI; ++c) acc[o][c] = fmaf(gv[o], xv[c], acc[o][c]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 0; o < CO; ++o)
#pragma unroll
    for (int c = 0; c < CI; ++c) {
      float v = acc[o][c];
#pragma unroll
      for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(PWCLO_FULL_MASK, v, s);
      if (lane == 0) red[warp][o * CI + c] = v;
    }
  __syncthreads();
  if (threadIdx.x < CI * CO) {
    float v = 0.f;
    for (int w = 0; w < kWgThreads / 32; ++w) v += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * CI * CO + threadIdx.x] = v;
  }
}

__global__ void conv1x1_wgrad_final_kernel(const float* __restrict__ partial, int nparts, int n, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = 0.0;
  for (int p = 0; p < nparts; ++p) v += (double)partial[(size_t)p * n + i];
  dw[i] = (float)v;
}

inline int wgrad_parts(int B, long long HW) {
  long long want = ((long long)B * HW + 4 * kWgThreads - 1) / (4 * kWgThreads);
  if (want > 4 * kNumSM) want = 4 * kNumSM;
  if (want < 1) want = 1;
  return (int)want;
}

template <int CI, int CO>
static int launch_wgrad(const float* x, const float* dy, int B, long long HW, float* dw, float* ws, cudaStream_t st) {
  const int parts = wgrad_parts(B, HW);
  conv1x1_wgrad_partial_kernel<CI, CO><<<parts, kWgThreads, 0, st>>>(x, dy, B, HW, ws);
  conv1x1_wgrad_final_kernel<<<1, 128, 0, st>>>(ws, parts, CI * CO, dw);
  return launch_status();
}

}  // namespace pwclo

PWCLO_API size_t pwclo_conv1x1_wgrad_workspace_bytes(int B, int CI, int CO, long long HW) {
  if (B <= 0 || CI <= 0 || CO <= 0 || HW <= 0) return 0;
  return (size_t)pwclo::wgrad_parts(B, HW) * CI * CO * sizeof(float);
}

// supported (CI, CO): (6,8), (8,8), (8,16), (3,8) -- returns PWCLO_EUNSUPPORTED otherwise (the caller keeps torch's GEMM)
PWCLO_API int pwclo_conv1x1_wgrad(const float* x, const float* dy, int B, int CI, int CO, long long HW, float* dw,
                                  void* workspace, void* stream) {
  using namespace pwclo;
  if (!x || !dy || !dw || !workspace || B <= 0 || HW <= 0) return PWCLO_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = reinterpret_cast<float*>(workspace);
  if (CI == 6 && CO == 8) return launch_wgrad<6, 8>(x, dy, B, HW, dw, ws, st);
  if (CI == 8 && CO == 8) return launch_wgrad<8, 8>(x, dy, B, HW, dw, ws, st);
  if (CI == 8 && CO == 16) return launch_wgrad<8, 16>(x, dy, B, HW, dw, ws, st);
  if (CI == 3 && CO == 8) return launch_wgrad<3, 8>(x, dy, B, HW, dw, ws, st);
  return PWCLO_EUNSUPPORTED;
}

// y[b,o,p] = sum_c W[o][c] * x[b,c,p] for the same skinny shapes (forward), or with `transpose_w` the input gradient
// dx[b,c,p] = sum_o W[o][c] * dy[b,o,p] (CI / CO then name the channels of the tensor read / written).  One pass,
// float4 per channel row: the op is pure streaming ((CI + CO) * 4 B per position) where the GEMM libraries spend
// 0.2 ms per call on a 64x64 tensor-op tile that is 90 % padding.
namespace pwclo {

template <int CI, int CO>
__global__ void __launch_bounds__(256)
conv1x1_small_kernel(const float* __restrict__ x, const float* __restrict__ w, int transpose_w, int B, long long HW4,
                     float* __restrict__ y) {
  __shared__ float sw[CO * CI];
  for (int i = threadIdx.x; i < CO * CI; i += 256) {
    const int o = i / CI, c = i - o * CI;
    sw[i] = transpose_w ? w[c * CO + o] : w[o * CI + c];      // transposed: w is stored [CI_out_of_conv = CI here][CO]
  }
  __syncthreads();
  const long long total = (long long)B * HW4;
  for (long long u = (long long)blockIdx.x * 256 + threadIdx.x; u < total; u += (long long)gridDim.x * 256) {
    const long long b = u / HW4, r = u - b * HW4;
    float4 xv[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) xv[c] = reinterpret_cast<const float4*>(x + ((size_t)b * CI + c) * HW4 * 4)[r];
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const float ww = sw[o * CI + c];
        a.x = fmaf(ww, xv[c].x, a.x); a.y = fmaf(ww, xv[c].y, a.y); a.z = fmaf(ww, xv[c].z, a.z); a.w = fmaf(ww, xv[c].w, a.w);
      }
      reinterpret_cast<float4*>(y + ((size_t)b * CO + o) * HW4 * 4)[r] = a;
    }
  }
}

template <int CI, int CO>
static int launch_small(const float* x, const float* w, int tr, int B, long long HW, float* y, cudaStream_t st) {
  const long long units = (long long)B * (HW / 4);
  long long blocks = (units + 255) / 256;
  if (blocks > 8 * kNumSM) blocks = 8 * kNumSM;
  conv1x1_small_kernel<CI, CO><<<(int)blocks, 256, 0, st>>>(x, w, tr, B, HW / 4, y);
  return launch_status();
}

}  // namespace pwclo

// x [B,CI,HW] -> y [B,CO,HW]; w is the conv weight: [CO][CI] (transpose_w = 0) or, for the input gradient, the weight
// [CI][CO] of the forward conv read transposed (transpose_w = 1).  HW % 4 == 0, 16-byte aligned tensors.
PWCLO_API int pwclo_conv1x1_small(const float* x, const float* w, int transpose_w, int B, int CI, int CO, long long HW, float* y,
                                  void* stream) {
  using namespace pwclo;
  if (!x || !w || !y || B <= 0 || HW <= 0) return PWCLO_EINVAL;
  if (HW % 4 != 0 || (((uintptr_t)x | (uintptr_t)y) & 15) != 0) return PWCLO_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int tr = transpose_w ? 1 : 0;
  if (CI == 6 && CO == 8) return launch_small<6, 8>(x, w, tr, B, HW, y, st);
  if (CI == 8 && CO == 8) return launch_small<8, 8>(x, w, tr, B, HW, y, st);
  if (CI == 8 && CO == 16) return launch_small<8, 16>(x, w, tr, B, HW, y, st);
  if (CI == 16 && CO == 8) return launch_small<16, 8>(x, w, tr, B, HW, y, st);
  if (CI == 3 && CO == 8) return launch_small<3, 8>(x, w, tr, B, HW, y, st);
  return PWCLO_EUNSUPPORTED;
}
