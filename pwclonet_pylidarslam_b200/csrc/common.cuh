// Shared device/host helpers for libpwclo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pwclo_b200.h"

#define PWCLO_API extern "C" __attribute__((visibility("default")))

#define PWCLO_FULL_MASK 0xffffffffu

namespace pwclo {

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PWCLO_OK : (int)e;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Squared distance as the reference's CUDA kernels compute it after nvcc's contraction of
// dx*dx + dy*dy + dz*dz: FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)  (checked in the SASS of the
// sm_100a build of EXT/src/sampling_gpu.cu:103-104, ball_query_gpu.cu:31-32, interpolate_gpu.cu:31).
__device__ __forceinline__ float dist2_ref_fma(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// sm_100a packed fp32 arithmetic (SASS FADD2 / FMUL2 / FFMA2): two independent round-to-nearest operations per
// instruction -- bit for bit the results of the scalar instructions, half the issue slots.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_sub(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// dist2_ref_fma for two candidates at once: fma(dz,dz, fma(dx,dx, dy*dy)) per half, the reference's contraction order
__device__ __forceinline__ f32x2_t dist2_ref_fma_x2(f32x2_t dx, f32x2_t dy, f32x2_t dz) {
  return f2_fma(dz, dz, f2_fma(dx, dx, f2_mul(dy, dy)));
}

// Pose warp of one point, op-for-op the reference's torch expression tree
// (PW/PWCLO_utils.py:31-63 with mul_q_point :100-132 and mul_point_q :66-97): every product and
// sum is a separately rounded fp32 operation, left to right.
struct PoseQT {
  float q0, q1, q2, q3;      // rotation, scalar first
  float i0, i1, i2, i3;      // conj(q) / (|q|^2 + 1e-10)
  float t0, t1, t2;
};

__device__ __forceinline__ PoseQT make_pose(const float* __restrict__ qt) {
  PoseQT p;
  p.q0 = qt[0]; p.q1 = qt[1]; p.q2 = qt[2]; p.q3 = qt[3];
  p.t0 = qt[4]; p.t1 = qt[5]; p.t2 = qt[6];
  float n2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.q0, p.q0), __fmul_rn(p.q1, p.q1)), __fmul_rn(p.q2, p.q2)),
                       __fmul_rn(p.q3, p.q3));
  n2 = __fadd_rn(n2, 1e-10f);
  p.i0 = __fdiv_rn(p.q0, n2);
  p.i1 = __fdiv_rn(-p.q1, n2);
  p.i2 = __fdiv_rn(-p.q2, n2);
  p.i3 = __fdiv_rn(-p.q3, n2);
  return p;
}

__device__ __forceinline__ void warp_point(const PoseQT& p, float x, float y, float z, float& ox, float& oy,
                                           float& oz) {
  // r = q (x) [0, x, y, z]
  const float p0 = 0.f;
  float r0 = __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(p.q0, p0), __fmul_rn(p.q1, x)), __fmul_rn(p.q2, y)), __fmul_rn(p.q3, z));
  float r1 = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.q0, x), __fmul_rn(p.q1, p0)), __fmul_rn(p.q2, z)), __fmul_rn(p.q3, y));
  float r2 = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(p.q0, y), __fmul_rn(p.q1, z)), __fmul_rn(p.q2, p0)), __fmul_rn(p.q3, x));
  float r3 = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(p.q0, z), __fmul_rn(p.q1, y)), __fmul_rn(p.q2, x)), __fmul_rn(p.q3, p0));
  // s = r (x) q^-1, vector part
  float s1 = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, p.i1), __fmul_rn(r1, p.i0)), __fmul_rn(r2, p.i3)), __fmul_rn(r3, p.i2));
  float s2 = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(r0, p.i2), __fmul_rn(r1, p.i3)), __fmul_rn(r2, p.i0)), __fmul_rn(r3, p.i1));
  float s3 = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(r0, p.i3), __fmul_rn(r1, p.i2)), __fmul_rn(r2, p.i1)), __fmul_rn(r3, p.i0));
  ox = __fadd_rn(s1, p.t0);
  oy = __fadd_rn(s2, p.t1);
  oz = __fadd_rn(s3, p.t2);
}

}  // namespace pwclo
