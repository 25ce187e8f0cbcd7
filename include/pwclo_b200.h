/*
 * pwclo_b200.h -- C ABI of libpwclo_b200.so, the B200-native (sm_100a) implementation of the
 * PWCLO-Net point-cloud hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in
 *     `_host`; tensors are dense, row-major, fp32 values / int32 indices;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is
 *     enqueued asynchronously, nothing synchronises, nothing allocates (callers own outputs and
 *     workspaces);
 *   - return value: 0 on success, a positive cudaError_t if a launch failed, or a negative
 *     PWCLO_E* code for argument errors.  Nothing ever calls exit() (the reference's
 *     CUDA_CHECK_ERRORS macro does: _ext-src/include/cuda_utils.h:30-39).
 *
 * Reference paths below are relative to /root/reference;
 *   EXT = slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops/_ext-src
 *   P2  = slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops
 *   PW  = slam/models/PWCLONet
 */
#ifndef PWCLO_B200_H
#define PWCLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PWCLO_OK 0
#define PWCLO_EINVAL (-1)      /* bad size / null pointer */
#define PWCLO_EUNSUPPORTED (-2) /* size outside what the kernels were built for */

/* flags of pwclo_furthest_point_sampling */
#define PWCLO_FPS_ORIGIN_SKIP 1u   /* skip candidates with |p|^2 <= 1e-3 (EXT/src/sampling_gpu.cu:100-101) */
#define PWCLO_FPS_CAP1024 2u       /* tie order of the 1024-thread variant (/sampling_gpu_copy.cu:143-148) */

/* summation order of the kNN squared distance (DESIGN.md "kNN formulation") */
#define PWCLO_KNN_SUM_XY_Z 0  /* (dx2+dy2)+dz2 : torch CPU reduction order */
#define PWCLO_KNN_SUM_XZ_Y 1  /* (dx2+dz2)+dy2 : torch CUDA reduction order */

const char *pwclo_version(void);
const char *pwclo_error_string(int code);

/* ---- B0: the nine functions of pointnet2_ops._ext (EXT/src/bindings.cpp:7-18) ------------- */

/* furthest_point_sampling(points[B,N,3], nsamples) -> idx[B,m]   EXT/src/sampling.cpp:66-87,
 * kernel EXT/src/sampling_gpu.cu:69-173.  Bit-exact incl. the reference's tie order.  No temp
 * buffer: the running minimum distances live in registers. */
int pwclo_furthest_point_sampling(const float *xyz, int B, int N, int m, unsigned flags,
                                  int32_t *idx, void *stream);

/* gather_points(points[B,C,N], idx[B,M]) -> out[B,C,M]           EXT/src/sampling.cpp:15-39 */
int pwclo_gather_points(const float *points, const int32_t *idx, int B, int C, int N, int M,
                        float *out, void *stream);
/* gather_points_grad(grad_out[B,C,M], idx[B,M], N) -> grad_points[B,C,N] (+=; caller zero-fills)
 *                                                                EXT/src/sampling.cpp:41-65 */
int pwclo_gather_points_grad(const float *grad_out, const int32_t *idx, int B, int C, int N, int M,
                             float *grad_points, void *stream);

/* group_points(points[B,C,N], idx[B,S,K]) -> out[B,C,S,K]        EXT/src/group_points.cpp:12-36 */
int pwclo_group_points(const float *points, const int32_t *idx, int B, int C, int N, int S, int K,
                       float *out, void *stream);
/* group_points_grad(grad_out[B,C,S,K], idx, N) -> grad_points[B,C,N] (+=; caller zero-fills)
 *                                                                EXT/src/group_points.cpp:38-62 */
int pwclo_group_points_grad(const float *grad_out, const int32_t *idx, int B, int C, int N, int S,
                            int K, float *grad_points, void *stream);

/* ball_query(new_xyz[B,m,3], xyz[B,n,3], radius, nsample) -> idx[B,m,nsample]
 *                                                                EXT/src/ball_query.cpp:9-32 */
int pwclo_ball_query(const float *new_xyz, const float *xyz, int B, int n, int m, float radius,
                     int nsample, int32_t *idx, void *stream);

/* three_nn(unknown[B,n,3], known[B,m,3]) -> dist2[B,n,3], idx[B,n,3]   EXT/src/interpolate.cpp:17-42 */
int pwclo_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                   int32_t *idx, void *stream);
/* three_interpolate(points[B,c,m], idx[B,n,3], weight[B,n,3]) -> out[B,c,n]  interpolate.cpp:44-72 */
int pwclo_three_interpolate(const float *points, const int32_t *idx, const float *weight, int B,
                            int c, int m, int n, float *out, void *stream);
/* three_interpolate_grad(grad_out[B,c,n], idx, weight, m) -> grad_points[B,c,m] (+=)  :74-99 */
int pwclo_three_interpolate_grad(const float *grad_out, const int32_t *idx, const float *weight,
                                 int B, int c, int n, int m, float *grad_points, void *stream);

/* ---- knn_point (pure PyTorch in the reference: P2/pytorch_utils.py:12-49) ----------------- */
/* idx[B,S,K] = indices of the K smallest sqrt(sum((new_xyz-xyz)^2)+1e-8), ascending, ties broken
 * by lower index.  dist (optional, may be NULL) receives the sorted distances.
 * warp_qt (optional, may be NULL): [B,7] = (q0..q3 scalar-first, t0..t2); when given, the query
 * points are first transformed by q (x) [0,p] (x) q^-1 + t exactly as PW/PWCLO_utils.py:42-63 and
 * (if warped_out != NULL) the transformed queries are written to warped_out[B,S,3]. */
int pwclo_knn(const float *xyz, const float *new_xyz, int B, int N, int S, int K, int sum_order,
              const float *warp_qt, float *warped_out, int32_t *idx, float *dist, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PWCLO_B200_H */
