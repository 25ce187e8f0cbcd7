// pose_post.cu -- pose post-processing on the GPU.
//
// Reference (host numpy, one frame at a time): train.py:875-886 turns pose_params[i,0,:] = (t, q) into a
// 4x4 matrix with quat2mat (train.py:762-796, fp32 scalar arithmetic under this image's NumPy 2) and
// inverts it (np.linalg.inv, float64); KITTI360_TRANSFORMATIONS.convert_to_absolute
// (slam/common/kitti360_utils.py:406-432, dict branch) chains them: abs_f = inv(rel_f @ inv(abs_{f-1})).
//   pwclo_pose_to_matrix     one thread per frame: fp32 quat2mat op for op, float64 affine inverse
//   pwclo_accumulate_poses   abs_f = abs_{f-1} * inv(rel_f) for a whole sequence as a parallel prefix product
//                            (warp-shuffle scan of 3x4 float64 affine maps, one CTA)
#include "common.cuh"

namespace pwclo {

struct Aff {          // rows 0..2 of a 4x4 with last row (0,0,0,1)
  double m[12];
};

__device__ __forceinline__ Aff aff_mul(const Aff& a, const Aff& b) {       // a * b
  Aff c;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double s = a.m[r * 4 + 0] * b.m[0 * 4 + k];
      s = fma(a.m[r * 4 + 1], b.m[1 * 4 + k], s);
      s = fma(a.m[r * 4 + 2], b.m[2 * 4 + k], s);
      if (k == 3) s += a.m[r * 4 + 3];
      c.m[r * 4 + k] = s;
    }
  }
  return c;
}

// exact-arithmetic formula of the inverse of [R t; 0 1] for a general (not necessarily orthonormal) R:
// adjugate / determinant, then -R^-1 t
__device__ __forceinline__ Aff aff_inv(const Aff& a) {
  const double a00 = a.m[0], a01 = a.m[1], a02 = a.m[2], a10 = a.m[4], a11 = a.m[5], a12 = a.m[6], a20 = a.m[8],
               a21 = a.m[9], a22 = a.m[10];
  const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  const double id = 1.0 / det;
  Aff r;
  r.m[0] = c00 * id; r.m[1] = (a02 * a21 - a01 * a22) * id; r.m[2] = (a01 * a12 - a02 * a11) * id;
  r.m[4] = c01 * id; r.m[5] = (a00 * a22 - a02 * a20) * id; r.m[6] = (a02 * a10 - a00 * a12) * id;
  r.m[8] = c02 * id; r.m[9] = (a01 * a20 - a00 * a21) * id; r.m[10] = (a00 * a11 - a01 * a10) * id;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    r.m[i * 4 + 3] = -(r.m[i * 4 + 0] * a.m[3] + r.m[i * 4 + 1] * a.m[7] + r.m[i * 4 + 2] * a.m[11]);
  return r;
}

__device__ __forceinline__ Aff aff_identity() {
  Aff a;
#pragma unroll
  for (int i = 0; i < 12; ++i) a.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
  return a;
}

__device__ __forceinline__ void aff_store(double* o, const Aff& a) {
#pragma unroll
  for (int i = 0; i < 12; ++i) o[i] = a.m[i];
  o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
}

__device__ __forceinline__ Aff aff_load(const double* p) {
  Aff a;
#pragma unroll
  for (int i = 0; i < 12; ++i) a.m[i] = p[i];
  return a;
}

__global__ void pose_to_matrix_kernel(const float* __restrict__ pose, int B, int stride, int invert,
                                      double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* p = pose + (size_t)b * stride;
  const float w = p[3], x = p[4], y = p[5], z = p[6];
  // quat2mat (train.py:781-796), every operation a separately rounded fp32 op
  const float Nq = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w, w), __fmul_rn(x, x)), __fmul_rn(y, y)), __fmul_rn(z, z));
  Aff a = aff_identity();
  if (!(Nq < 1e-8f)) {
    const float s = __fdiv_rn(2.0f, Nq);
    const float X = __fmul_rn(x, s), Y = __fmul_rn(y, s), Z = __fmul_rn(z, s);
    const float wX = __fmul_rn(w, X), wY = __fmul_rn(w, Y), wZ = __fmul_rn(w, Z);
    const float xX = __fmul_rn(x, X), xY = __fmul_rn(x, Y), xZ = __fmul_rn(x, Z);
    const float yY = __fmul_rn(y, Y), yZ = __fmul_rn(y, Z), zZ = __fmul_rn(z, Z);
    a.m[0] = (double)__fsub_rn(1.0f, __fadd_rn(yY, zZ)); a.m[1] = (double)__fsub_rn(xY, wZ); a.m[2] = (double)__fadd_rn(xZ, wY);
    a.m[4] = (double)__fadd_rn(xY, wZ); a.m[5] = (double)__fsub_rn(1.0f, __fadd_rn(xX, zZ)); a.m[6] = (double)__fsub_rn(yZ, wX);
    a.m[8] = (double)__fsub_rn(xZ, wY); a.m[9] = (double)__fadd_rn(yZ, wX); a.m[10] = (double)__fsub_rn(1.0f, __fadd_rn(xX, yY));
  }
  a.m[3] = (double)p[0]; a.m[7] = (double)p[1]; a.m[11] = (double)p[2];
  if (invert) a = aff_inv(a);
  aff_store(out + (size_t)b * 16, a);
}

__device__ __forceinline__ Aff aff_shfl_up(const Aff& a, int delta) {
  Aff r;
#pragma unroll
  for (int i = 0; i < 12; ++i) r.m[i] = __shfl_up_sync(PWCLO_FULL_MASK, a.m[i], delta);
  return r;
}

constexpr int kAccThreads = 1024;

__global__ void __launch_bounds__(kAccThreads)
accumulate_poses_kernel(const double* __restrict__ rel, int F, const double* __restrict__ first, double* __restrict__ out) {
  __shared__ Aff warp_tot[kAccThreads / 32];
  __shared__ Aff carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = first ? aff_load(first) : aff_identity();
  __syncthreads();
  for (int f0 = 0; f0 < F; f0 += kAccThreads) {
    const int f = f0 + tid;
    Aff m = f < F ? aff_inv(aff_load(rel + (size_t)f * 16)) : aff_identity();
    // inclusive scan of the ordered product inside the warp: m_t <- m_{t-o} * m_t
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const Aff up = aff_shfl_up(m, o);
      if (lane >= o) m = aff_mul(up, m);
    }
    if (lane == 31) warp_tot[warp] = m;
    __syncthreads();
    if (warp == 0) {
      Aff w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const Aff up = aff_shfl_up(w, o);
        if (lane >= o) w = aff_mul(up, w);
      }
      warp_tot[lane] = w;           // inclusive prefix over warps
    }
    __syncthreads();
    Aff pre = carry_s;
    if (warp > 0) pre = aff_mul(pre, warp_tot[warp - 1]);
    m = aff_mul(pre, m);
    if (f < F) aff_store(out + (size_t)f * 16, m);
    __syncthreads();
    if (tid == kAccThreads - 1) carry_s = m;     // identity padding keeps the last real prefix
    __syncthreads();
  }
}

}  // namespace pwclo

PWCLO_API int pwclo_pose_to_matrix(const float* pose_params, int B, int row_stride, int invert, double* out, void* stream) {
  if (!pose_params || !out || B < 0 || row_stride < 7) return PWCLO_EINVAL;
  if (B == 0) return PWCLO_OK;
  pwclo::pose_to_matrix_kernel<<<pwclo::ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(pose_params, B, row_stride,
                                                                                         invert ? 1 : 0, out);
  return pwclo::launch_status();
}

PWCLO_API int pwclo_accumulate_poses(const double* rel, int F, const double* first, double* out, void* stream) {
  if (!rel || !out || F < 0) return PWCLO_EINVAL;
  if (F == 0) return PWCLO_OK;
  pwclo::accumulate_poses_kernel<<<1, pwclo::kAccThreads, 0, (cudaStream_t)stream>>>(rel, F, first, out);
  return pwclo::launch_status();
}
