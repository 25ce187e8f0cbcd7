// train_ops.cu -- training-side rows: the PWCLO-Net loss with its gradient in ONE launch, and the
// Adam update of the flat parameter arena in ONE launch.
//
// Reference: _PWCLONetLossModule.forward (slam/training/loss_modules.py:424-544) with
// ExponentialWeights.forward (:171-196).  The reference builds the value out of ~120 element-wise
// torch launches on [B,4,7] tensors and autograd replays as many for the gradient; here one CTA
// evaluates the four per-level rotation / translation terms, the weighted total, d loss / d pred and
// d loss / d s.  The problem is tiny (B x 28 floats): the kernel is launch-latency bound by design.
#include <math.h>

#include "common.cuh"

namespace pwclo {

constexpr int kLossThreads = 256;
constexpr int kLossTerms = 8;   // rot_l, trans_l for the 4 levels

// out[16]: 0 loss | 1..4 loss_l1..l4 | 5..8 loss_rot_l1..l4 | 9..12 loss_trans_l1..l4 | 13,14 s (or weights) | 15 B
__global__ void __launch_bounds__(kLossThreads)
pose_loss_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ s, int B,
                 int with_exp, float* __restrict__ out, float* __restrict__ grad_pred, float* __restrict__ grad_s) {
  __shared__ float red[kLossTerms][kLossThreads / 32];
  const float s_t = s[0], s_q = s[1];
  // d loss / d (level term): level weights 0.2,0.4,0.8,1.6 for rows 0..3 (row 0 = finest level,
  // loss_modules.py:531), times exp(-s) (ExponentialWeights) or the fixed weight (:519-522)
  const float wt = with_exp ? expf(-s_t) : s_t;
  const float wq = with_exp ? expf(-s_q) : s_q;
  const float invB = 1.f / (float)B, inv3B = 1.f / (3.f * (float)B);

  float acc[kLossTerms];
#pragma unroll
  for (int i = 0; i < kLossTerms; ++i) acc[i] = 0.f;

  for (int item = threadIdx.x; item < B * 4; item += kLossThreads) {
    const int b = item >> 2, l = item & 3;
    const float* p = pred + (size_t)item * 7;
    const float* g = gt + (size_t)b * 7;
    const float lw = 0.2f * (float)(1 << l);
    // translation: mean over (b, c) of sqrt((t-g)^2 + 1e-10)            (:382-384)
    float tsum = 0.f, gtr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = p[c] - g[c];
      const float r = sqrtf(d * d + 1e-10f);
      tsum += r;
      gtr[c] = lw * wt * inv3B * d / r;
    }
    // rotation: q/(sqrt(sum q^2 + 1e-10) + 1e-10), then mean over b of sqrt(sum (qn-g)^2 + 1e-10)   (:370-373, :388-391)
    const float q0 = p[3], q1 = p[4], q2 = p[5], q3 = p[6];
    const float r = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3 + 1e-10f);
    const float n = r + 1e-10f;
    const float e0 = q0 / n - g[3], e1 = q1 / n - g[4], e2 = q2 / n - g[5], e3 = q3 / n - g[6];
    const float d = sqrtf(e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3 + 1e-10f);
    if (grad_pred) {
      float* gp = grad_pred + (size_t)item * 7;
      gp[0] = gtr[0]; gp[1] = gtr[1]; gp[2] = gtr[2];
      // d d / d qn_i = e_i / d ; d qn_i / d q_j = delta_ij / n - q_i q_j / (n^2 r)
      const float k = lw * wq * invB / d;
      const float h0 = k * e0, h1 = k * e1, h2 = k * e2, h3 = k * e3;
      const float hq = (h0 * q0 + h1 * q1 + h2 * q2 + h3 * q3) / (n * n * r);
      gp[3] = h0 / n - q0 * hq;
      gp[4] = h1 / n - q1 * hq;
      gp[5] = h2 / n - q2 * hq;
      gp[6] = h3 / n - q3 * hq;
    }
    // branch-free scatter into the per-level accumulators
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] += (l == j) ? d : 0.f;
      acc[4 + j] += (l == j) ? tsum : 0.f;
    }
  }
  // block reduction in a fixed order: shuffle tree inside a warp, then warp 0 adds the 8 partials
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kLossTerms; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(PWCLO_FULL_MASK, v, o);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float rot[4], tr[4], lvl[4], total = 0.f, ds_t = 0.f, ds_q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < kLossThreads / 32; ++w) { a += red[i][w]; b += red[4 + i][w]; }
      rot[i] = a * invB;
      tr[i] = b * inv3B;
      const float lw = 0.2f * (float)(1 << i);
      lvl[i] = with_exp ? (tr[i] * wt + s_t) + (rot[i] * wq + s_q) : tr[i] * wt + rot[i] * wq;
      ds_t += lw * (1.f - tr[i] * wt);
      ds_q += lw * (1.f - rot[i] * wq);
    }
    // 1.6*L4 + 0.8*L3 + 0.4*L2 + 0.2*L1, added in the reference's order (:531)
    total = 1.6f * lvl[3] + 0.8f * lvl[2] + 0.4f * lvl[1] + 0.2f * lvl[0];
    out[0] = total;
#pragma unroll
    for (int i = 0; i < 4; ++i) { out[1 + i] = lvl[i]; out[5 + i] = rot[i]; out[9 + i] = tr[i]; }
    out[13] = s_t; out[14] = s_q; out[15] = (float)B;
    if (grad_s) {
      if (with_exp) { grad_s[0] = ds_t; grad_s[1] = ds_q; }
      else {
        // fixed weights: d loss / d weight = sum_l lw * term_l (not a trained quantity in the reference)
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 4; ++i) { a += 0.2f * (float)(1 << i) * tr[i]; b += 0.2f * (float)(1 << i) * rot[i]; }
        grad_s[0] = a; grad_s[1] = b;
      }
    }
  }
}

// torch.optim.Adam (the optimiser the reference builds, slam/training/trainer.py:309-323; betas (0.9, 0.999),
// L2 weight decay added to the gradient) over ONE flat fp32 arena holding every trainable parameter:
// 16 B read + 12 B written per element, float4-vectorised, grid-stride.  Operation order follows
// torch's single-tensor implementation so results agree to fp32 rounding:
//   g += wd*p ; m = lerp(m, g, 1-b1) ; v = b2*v + (1-b2)*g*g ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_over_bc1, float b1, float b2,
                                         float eps, float wd, float sqrt_bc2, float gscale) {
  g *= gscale;
  g = fmaf(wd, p, g);
  m = fmaf(1.f - b1, g - m, m);
  v = fmaf(1.f - b2, g * g, b2 * v);
  const float denom = sqrtf(v) / sqrt_bc2 + eps;
  p = fmaf(-lr_over_bc1, m / denom, p);
}

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 size_t n, float lr_over_bc1, float b1, float b2, float eps, float wd, float sqrt_bc2, float gscale) {
  const size_t n4 = n >> 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n not a multiple of 4)
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_one(p[i], g[i], m[i], v[i], lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
}

// Pose warp of a cloud, p' = q (x) [0,p] (x) q^-1 + t with q^-1 = conj(q) / (|q|^2 + 1e-10) (PW/PWCLO_utils.py:31-63), as ONE
// launch forward and ONE backward for training: the reference's torch expression is ~60 element-wise launches forward
// and ~150 in autograd's replay, three times per step.  xyz / out are channel-major [B,3,N] (the layout the composed
// path holds), q [B,4] scalar first, t [B,3].  Forward values are op-for-op those of the expression (warp_point).
// Backward, with s = |q|^2 + 1e-10, q = (w, v), u = (w^2 - v.v) p + 2 (v.p) v + 2 w (v x p) = s (p' - t), g = dL/dp':
//   dL/dt = sum g          dL/dp = [(w^2 - v.v) g + 2 (v.g) v - 2 w (v x g)] / s
//   dL/dw = sum [2 w (p.g) + 2 (v x p).g] / s - (u.g) 2 w / s^2
//   dL/dv = sum [-2 (p.g) v + 2 (v.g) p + 2 (v.p) g + 2 w (p x g)] / s - (u.g) 2 v / s^2
constexpr int kWarpThreads = 256;

__global__ void __launch_bounds__(kWarpThreads)
warp_fwd_kernel(const float* __restrict__ xyz, const float* __restrict__ q, const float* __restrict__ t, int N,
                float* __restrict__ out) {
  const int b = blockIdx.y;
  __shared__ float qt[7];
  if (threadIdx.x < 4) qt[threadIdx.x] = q[b * 4 + threadIdx.x];
  else if (threadIdx.x < 7) qt[threadIdx.x] = t[b * 3 + threadIdx.x - 4];
  __syncthreads();
  const PoseQT P = make_pose(qt);
  const float* x = xyz + (size_t)b * 3 * N;
  float* o = out + (size_t)b * 3 * N;
  for (int n = blockIdx.x * kWarpThreads + threadIdx.x; n < N; n += gridDim.x * kWarpThreads) {
    float ox, oy, oz;
    warp_point(P, x[n], x[N + n], x[2 * N + n], ox, oy, oz);
    o[n] = ox; o[N + n] = oy; o[2 * N + n] = oz;
  }
}

__global__ void __launch_bounds__(kWarpThreads)
warp_bwd_kernel(const float* __restrict__ xyz, const float* __restrict__ q, const float* __restrict__ gout, int N,
                float* __restrict__ dxyz, float* __restrict__ dq, float* __restrict__ dt) {
  const int b = blockIdx.x;
  __shared__ double red[7][kWarpThreads / 32];
  const double w = q[b * 4 + 0], vx = q[b * 4 + 1], vy = q[b * 4 + 2], vz = q[b * 4 + 3];
  const double s = w * w + vx * vx + vy * vy + vz * vz + 1e-10;
  const double ww = w * w - (vx * vx + vy * vy + vz * vz);
  const float* x = xyz + (size_t)b * 3 * N;
  const float* g = gout + (size_t)b * 3 * N;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};     // dw, dvx, dvy, dvz, dtx, dty, dtz
  for (int n = threadIdx.x; n < N; n += kWarpThreads) {
    const double px = x[n], py = x[N + n], pz = x[2 * N + n];
    const double gx = g[n], gy = g[N + n], gz = g[2 * N + n];
    const double vp = vx * px + vy * py + vz * pz, vg = vx * gx + vy * gy + vz * gz, pg = px * gx + py * gy + pz * gz;
    // v x p, p x g, v x g
    const double cx = vy * pz - vz * py, cy = vz * px - vx * pz, cz = vx * py - vy * px;
    const double hx = py * gz - pz * gy, hy = pz * gx - px * gz, hz = px * gy - py * gx;
    const double ux = ww * px + 2.0 * vp * vx + 2.0 * w * cx, uy = ww * py + 2.0 * vp * vy + 2.0 * w * cy,
                 uz = ww * pz + 2.0 * vp * vz + 2.0 * w * cz;
    const double ug = (ux * gx + uy * gy + uz * gz) * 2.0 / (s * s);
    acc[0] += (2.0 * w * pg + 2.0 * (cx * gx + cy * gy + cz * gz)) / s - ug * w;
    acc[1] += (-2.0 * pg * vx + 2.0 * vg * px + 2.0 * vp * gx + 2.0 * w * hx) / s - ug * vx;
    acc[2] += (-2.0 * pg * vy + 2.0 * vg * py + 2.0 * vp * gy + 2.0 * w * hy) / s - ug * vy;
    acc[3] += (-2.0 * pg * vz + 2.0 * vg * pz + 2.0 * vp * gz + 2.0 * w * hz) / s - ug * vz;
    acc[4] += gx; acc[5] += gy; acc[6] += gz;
    if (dxyz) {
      const double ex = vy * gz - vz * gy, ey = vz * gx - vx * gz, ez = vx * gy - vy * gx;      // v x g
      float* d = dxyz + (size_t)b * 3 * N;
      d[n] = (float)((ww * gx + 2.0 * vg * vx - 2.0 * w * ex) / s);
      d[N + n] = (float)((ww * gy + 2.0 * vg * vy - 2.0 * w * ey) / s);
      d[2 * N + n] = (float)((ww * gz + 2.0 * vg * vz - 2.0 * w * ez) / s);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(PWCLO_FULL_MASK, v, o);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0.0;
    for (int k = 0; k < kWarpThreads / 32; ++k) v += red[threadIdx.x][k];
    if (threadIdx.x < 4) dq[b * 4 + threadIdx.x] = (float)v;
    else dt[b * 3 + threadIdx.x - 4] = (float)v;
  }
}

// Device-resident step counter and learning rate, so that a whole training step (forward, backward,
// all-reduce, Adam) can be captured once in a CUDA graph and replayed: nothing step-dependent is baked into
// kernel arguments.  state_i[0] = step count (incremented here), state_f = {lr (written by the host between
// replays), lr / (1 - beta1^t), sqrt(1 - beta2^t)} computed in double exactly as the host path does.
__global__ void adam_tick_kernel(int* __restrict__ state_i, float* __restrict__ state_f, float beta1, float beta2) {
  const int t = state_i[0] + 1;
  state_i[0] = t;
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  state_f[1] = (float)((double)state_f[0] / bc1);
  state_f[2] = (float)sqrt(bc2);
}

__global__ void __launch_bounds__(256)
adam_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     size_t n, const float* __restrict__ state_f, float b1, float b2, float eps, float wd, float gscale) {
  const float lr_over_bc1 = state_f[1], sqrt_bc2 = state_f[2];
  const size_t n4 = n >> 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_one(p[i], g[i], m[i], v[i], lr_over_bc1, b1, b2, eps, wd, sqrt_bc2, gscale);
}

}  // namespace pwclo

// -------------------------------------------------------------------------------------------------
// Max over the neighbour axis for training (F.max_pool2d(x, [1, K]) of P2/pointnet2_modules.py:239-243, :499-506 and
// its backward): x [rows, K] contiguous -> y [rows], arg [rows] (first maximum wins, NaN propagates: ATen's rule), and
// dx [rows, K] = dy at arg, 0 elsewhere.  One thread per row, K consecutive floats per thread (a warp covers one contiguous
// 128 * K byte region), every dx element written exactly once: no atomics, no zero-fill pass.
// -------------------------------------------------------------------------------------------------
namespace pwclo {
__global__ void maxpool_lastdim_fwd_kernel(const float* __restrict__ x, long long rows, int K, float* __restrict__ y,
                                           unsigned char* __restrict__ arg) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* xr = x + r * K;
  float best = xr[0];
  int bi = 0;
  if ((K & 3) == 0) {                     // 16-byte loads (rows start on 16-byte boundaries when K % 4 == 0)
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    for (int k4 = 0; k4 < (K >> 2); ++k4) {
      const float4 q = x4[k4];
      const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (v[i] > best || (v[i] != v[i] && best == best)) { best = v[i]; bi = 4 * k4 + i; }
    }
  } else {
    for (int k = 1; k < K; ++k) {
      const float v = xr[k];
      if (v > best || (v != v && best == best)) { best = v; bi = k; }
    }
  }
  y[r] = best;
  arg[r] = (unsigned char)bi;
}
__global__ void maxpool_lastdim_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ arg, long long rows,
                                           int K, float* __restrict__ dx) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float g = dy[r];
  const int a = arg[r];
  float* o = dx + r * K;
  if ((K & 3) == 0) {
    float4* o4 = reinterpret_cast<float4*>(o);
    for (int k4 = 0; k4 < (K >> 2); ++k4) {
      const int k = 4 * k4;
      o4[k4] = make_float4(k == a ? g : 0.f, k + 1 == a ? g : 0.f, k + 2 == a ? g : 0.f, k + 3 == a ? g : 0.f);
    }
  } else {
    for (int k = 0; k < K; ++k) o[k] = k == a ? g : 0.f;
  }
}
}  // namespace pwclo

PWCLO_API int pwclo_maxpool_lastdim_fwd(const float* x, long long rows, int K, float* y, unsigned char* arg, void* stream) {
  if (!x || !y || !arg || rows < 0 || K <= 0 || K > 255) return PWCLO_EINVAL;
  if (rows == 0) return PWCLO_OK;
  const long long blocks = (rows + 255) / 256;
  if (blocks > 2147483647LL) return PWCLO_EUNSUPPORTED;
  pwclo::maxpool_lastdim_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, K, y, arg);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_maxpool_lastdim_bwd(const float* dy, const unsigned char* arg, long long rows, int K, float* dx,
                                        void* stream) {
  if (!dy || !arg || !dx || rows < 0 || K <= 0 || K > 255) return PWCLO_EINVAL;
  if (rows == 0) return PWCLO_OK;
  const long long blocks = (rows + 255) / 256;
  if (blocks > 2147483647LL) return PWCLO_EUNSUPPORTED;
  pwclo::maxpool_lastdim_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, arg, rows, K, dx);
  return pwclo::launch_status();
}

// -------------------------------------------------------------------------------------------------
// Geometry channels of the cost volume for training (PW/costvolume.py:94-105, :159-169): for a centre p [B,3,S] and its
// grouped neighbours g [B,3,S,K]:  out [B,10,S,K] = (p, g, g - p, sqrt(|g - p|^2 + 1e-20)), and the gradient of both
// inputs.  The reference composes tile / sub / square / sum / add / sqrt / cat (7 launches forward, ~12 backward).
// -------------------------------------------------------------------------------------------------
namespace pwclo {
__global__ void cost_geometry_fwd_kernel(const float* __restrict__ p, const float* __restrict__ g, int S, int K,
                                         long long total, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, s, k)
  if (i >= total) return;
  const long long SK = (long long)S * K;
  const long long b = i / SK, sk = i - b * SK;
  const int s = (int)(sk / K);
  const float* pb = p + b * 3 * S + s;
  const float* gb = g + b * 3 * SK + sk;
  const float px = pb[0], py = pb[S], pz = pb[2 * (long long)S];
  const float gx = gb[0], gy = gb[SK], gz = gb[2 * SK];
  const float dx = __fsub_rn(gx, px), dy = __fsub_rn(gy, py), dz = __fsub_rn(gz, pz);
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  float* o = out + b * 10 * SK + sk;
  o[0] = px; o[SK] = py; o[2 * SK] = pz;
  o[3 * SK] = gx; o[4 * SK] = gy; o[5 * SK] = gz;
  o[6 * SK] = dx; o[7 * SK] = dy; o[8 * SK] = dz;
  o[9 * SK] = __fsqrt_rn(__fadd_rn(n2, 1e-20f));
}
// one thread per (b, s): its K neighbours; grad_p is their sum (no atomics), grad_g written once
__global__ void cost_geometry_bwd_kernel(const float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ go,
                                         int S, int K, long long points, float* __restrict__ gp, float* __restrict__ gg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, s)
  if (i >= points) return;
  const long long SK = (long long)S * K;
  const long long b = i / S;
  const int s = (int)(i - b * S);
  const float* pb = p + b * 3 * S + s;
  const float px = pb[0], py = pb[S], pz = pb[2 * (long long)S];
  const float* gb = g + b * 3 * SK + (long long)s * K;
  const float* ob = go + b * 10 * SK + (long long)s * K;
  float* ggb = gg ? gg + b * 3 * SK + (long long)s * K : nullptr;
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int k = 0; k < K; ++k) {
    const float dx = __fsub_rn(gb[k], px), dy = __fsub_rn(gb[SK + k], py), dz = __fsub_rn(gb[2 * SK + k], pz);
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float e = __fsqrt_rn(__fadd_rn(n2, 1e-20f));
    const float w = ob[9 * SK + k] / e;                       // d sqrt(n2 + eps) / d d = d / e
    const float tx = ob[6 * SK + k] + w * dx, ty = ob[7 * SK + k] + w * dy, tz = ob[8 * SK + k] + w * dz;   // via d = g - p
    if (ggb) {
      ggb[k] = ob[3 * SK + k] + tx;
      ggb[SK + k] = ob[4 * SK + k] + ty;
      ggb[2 * SK + k] = ob[5 * SK + k] + tz;
    }
    ax += ob[k] - tx;
    ay += ob[SK + k] - ty;
    az += ob[2 * SK + k] - tz;
  }
  if (gp) {
    float* gpb = gp + b * 3 * S + s;
    gpb[0] = ax; gpb[S] = ay; gpb[2 * (long long)S] = az;
  }
}
}  // namespace pwclo

PWCLO_API int pwclo_cost_geometry_fwd(const float* center, const float* grouped, int B, int S, int K, float* out, void* stream) {
  if (!center || !grouped || !out || B < 0 || S <= 0 || K <= 0) return PWCLO_EINVAL;
  const long long total = (long long)B * S * K;
  if (total == 0) return PWCLO_OK;
  const long long blocks = (total + 255) / 256;
  if (blocks > 2147483647LL) return PWCLO_EUNSUPPORTED;
  pwclo::cost_geometry_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(center, grouped, S, K, total, out);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_cost_geometry_bwd(const float* center, const float* grouped, const float* grad_out, int B, int S, int K,
                                      float* grad_center, float* grad_grouped, void* stream) {
  if (!center || !grouped || !grad_out || B < 0 || S <= 0 || K <= 0) return PWCLO_EINVAL;
  const long long points = (long long)B * S;
  if (points == 0 || (!grad_center && !grad_grouped)) return PWCLO_OK;
  const long long blocks = (points + 127) / 128;
  if (blocks > 2147483647LL) return PWCLO_EUNSUPPORTED;
  pwclo::cost_geometry_bwd_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(center, grouped, grad_out, S, K, points,
                                                                                     grad_center, grad_grouped);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_warp_fwd(const float* xyz, const float* q, const float* t, int B, int N, float* out, void* stream) {
  if (!xyz || !q || !t || !out || B < 0 || N < 0) return PWCLO_EINVAL;
  if (B == 0 || N == 0) return PWCLO_OK;
  int gx = pwclo::ceil_div(N, pwclo::kWarpThreads);
  if (gx > 64) gx = 64;
  pwclo::warp_fwd_kernel<<<dim3(gx, B, 1), pwclo::kWarpThreads, 0, (cudaStream_t)stream>>>(xyz, q, t, N, out);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_warp_bwd(const float* xyz, const float* q, const float* grad_out, int B, int N, float* grad_xyz,
                             float* grad_q, float* grad_t, void* stream) {
  if (!xyz || !q || !grad_out || !grad_q || !grad_t || B < 0 || N < 0) return PWCLO_EINVAL;
  if (B == 0) return PWCLO_OK;
  pwclo::warp_bwd_kernel<<<B, pwclo::kWarpThreads, 0, (cudaStream_t)stream>>>(xyz, q, grad_out, N, grad_xyz, grad_q, grad_t);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                                  int32_t* state_i, float* state_f, float beta1, float beta2, float eps, float weight_decay,
                                  float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !state_i || !state_f) return PWCLO_EINVAL;
  if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) return PWCLO_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  pwclo::adam_tick_kernel<<<1, 1, 0, st>>>(state_i, state_f, beta1, beta2);
  if (n == 0) return pwclo::launch_status();
  const size_t n4 = (n + 3) / 4;
  int blocks = (int)((n4 + 255) / 256);
  const int cap = pwclo::kNumSM * 8;
  if (blocks > cap) blocks = cap;
  pwclo::adam_step_dev_kernel<<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, state_f, beta1, beta2, eps,
                                                      weight_decay, grad_scale);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int step,
                              float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                              void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) return PWCLO_EINVAL;
  if (n == 0) return PWCLO_OK;
  if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) return PWCLO_EINVAL;
  // bias corrections in double on the host, as torch does with python floats
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float lr_over_bc1 = (float)((double)lr / bc1), sqrt_bc2 = (float)sqrt(bc2);
  const size_t n4 = (n + 3) / 4;
  int blocks = (int)((n4 + 255) / 256);
  const int cap = pwclo::kNumSM * 8;
  if (blocks > cap) blocks = cap;
  pwclo::adam_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr_over_bc1, beta1,
                                                                    beta2, eps, weight_decay, sqrt_bc2, grad_scale);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_pose_loss(const float* pred, const float* gt, const float* s, int B, int with_exp_weights, float* out,
                              float* grad_pred, float* grad_s, void* stream) {
  if (!pred || !gt || !s || !out || B <= 0) return PWCLO_EINVAL;
  pwclo::pose_loss_kernel<<<1, pwclo::kLossThreads, 0, (cudaStream_t)stream>>>(pred, gt, s, B, with_exp_weights ? 1 : 0, out,
                                                                               grad_pred, grad_s);
  return pwclo::launch_status();
}

// -------------------------------------------------------------------------------------------------
// Attentive pooling of the cost volume in training (PW/costvolume.py:139-145, :181-188):
//   out[r] = sum_k softmax_k(w[r, :])[k] * x[r, k]        w, x [rows, K] (rows = B*C*S, K neighbours, contiguous)
// and its backward.  The reference composes softmax / mul / sum (3 launches forward, ~6 backward, four passes over the
// [B,C,S,K] tensors); here one thread owns a row (K <= 32 values in registers), one pass forward, one pass backward with
// the softmax recomputed from w (nothing but the inputs is saved).
//   d out / d x_k = p_k                      d out / d w_k = p_k (x_k - out)
// -------------------------------------------------------------------------------------------------
namespace pwclo {
template <int K>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[K]) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int i = 0; i < K / 4; ++i) {
      const float4 q = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int i = 0; i < K / 2; ++i) {
      const float2 q = reinterpret_cast<const float2*>(p)[i];
      v[2 * i] = q.x; v[2 * i + 1] = q.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = p[i];
  }
}
template <int K>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[K]) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int i = 0; i < K / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int i = 0; i < K / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) p[i] = v[i];
  }
}
// softmax of a row in place (w -> p), torch's formula: exp(w - max) / sum
template <int K>
__device__ __forceinline__ void softmax_row(float (&w)[K]) {
  float m = w[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m = fmaxf(m, w[k]);
  float z = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { w[k] = expf(w[k] - m); z += w[k]; }
#pragma unroll
  for (int k = 0; k < K; ++k) w[k] = w[k] / z;
}
template <int K>
__global__ void __launch_bounds__(128) softmax_pool_fwd_kernel(const float* __restrict__ w, const float* __restrict__ x,
                                                                long long rows, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float p[K], xv[K];
  load_row<K>(w + r * K, p);
  load_row<K>(x + r * K, xv);
  softmax_row<K>(p);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) acc = __fadd_rn(acc, __fmul_rn(p[k], xv[k]));      // mul then sum, as the reference composes it
  out[r] = acc;
}
template <int K>
__global__ void __launch_bounds__(128) softmax_pool_bwd_kernel(const float* __restrict__ w, const float* __restrict__ x,
                                                                const float* __restrict__ g, long long rows,
                                                                float* __restrict__ dw, float* __restrict__ dx) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float p[K], xv[K];
  load_row<K>(w + r * K, p);
  load_row<K>(x + r * K, xv);
  softmax_row<K>(p);
  float out = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) out = __fadd_rn(out, __fmul_rn(p[k], xv[k]));
  const float gr = g[r];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float gp = gr * p[k];
    xv[k] = gp * (xv[k] - out);      // d w_k
    p[k] = gp;                       // d x_k
  }
  store_row<K>(dw + r * K, xv);
  store_row<K>(dx + r * K, p);
}
template <int K>
static int launch_softmax_pool(const float* w, const float* x, const float* g, long long rows, float* out, float* dw, float* dx,
                               cudaStream_t st) {
  const long long blocks = (rows + 127) / 128;
  if (blocks > 2147483647LL) return PWCLO_EUNSUPPORTED;
  if (g == nullptr) softmax_pool_fwd_kernel<K><<<(unsigned)blocks, 128, 0, st>>>(w, x, rows, out);
  else softmax_pool_bwd_kernel<K><<<(unsigned)blocks, 128, 0, st>>>(w, x, g, rows, dw, dx);
  return launch_status();
}
static int dispatch_softmax_pool(const float* w, const float* x, const float* g, long long rows, int K, float* out, float* dw,
                                 float* dx, cudaStream_t st) {
  switch (K) {
    case 4: return launch_softmax_pool<4>(w, x, g, rows, out, dw, dx, st);
    case 6: return launch_softmax_pool<6>(w, x, g, rows, out, dw, dx, st);
    case 8: return launch_softmax_pool<8>(w, x, g, rows, out, dw, dx, st);
    case 16: return launch_softmax_pool<16>(w, x, g, rows, out, dw, dx, st);
    case 32: return launch_softmax_pool<32>(w, x, g, rows, out, dw, dx, st);
    default: return PWCLO_EUNSUPPORTED;
  }
}
}  // namespace pwclo

// K in {4, 6, 8, 16, 32} (the neighbour counts of PWCLO-Net's cost volumes); PWCLO_EUNSUPPORTED otherwise (the caller keeps
// the composed expression).  Tensors 16-byte aligned.
PWCLO_API int pwclo_softmax_pool_fwd(const float* w, const float* x, long long rows, int K, float* out, void* stream) {
  if (!w || !x || !out || rows < 0 || K <= 0) return PWCLO_EINVAL;
  if (rows == 0) return PWCLO_OK;
  if ((((uintptr_t)w | (uintptr_t)x) & 15) != 0) return PWCLO_EINVAL;
  return pwclo::dispatch_softmax_pool(w, x, nullptr, rows, K, out, nullptr, nullptr, (cudaStream_t)stream);
}

PWCLO_API int pwclo_softmax_pool_bwd(const float* w, const float* x, const float* grad_out, long long rows, int K, float* grad_w,
                                     float* grad_x, void* stream) {
  if (!w || !x || !grad_out || !grad_w || !grad_x || rows < 0 || K <= 0) return PWCLO_EINVAL;
  if (rows == 0) return PWCLO_OK;
  if ((((uintptr_t)w | (uintptr_t)x | (uintptr_t)grad_w | (uintptr_t)grad_x) & 15) != 0) return PWCLO_EINVAL;
  return pwclo::dispatch_softmax_pool(w, x, grad_out, rows, K, nullptr, grad_w, grad_x, (cudaStream_t)stream);
}
