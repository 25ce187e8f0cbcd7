/*
 * pwclo_b200.h -- C ABI of libpwclo_b200.so, the B200-native (sm_100a) implementation of the
 * PWCLO-Net point-cloud hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in
 *     `_host`; tensors are dense, row-major, fp32 values / int32 indices;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is
 *     enqueued asynchronously, nothing synchronises, nothing allocates (callers own outputs and
 *     workspaces; the one exception is FPS with N > 16384, which takes a stream-ordered temporary);
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

#include <stddef.h>
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
/* Same sampling with the prefix property made usable (PW/pwclo_net.py:130-138,164 chains FPS on FPS
 * output: P2/pointnet2_modules.py:200-206 at every level): FPS of a cloud that is itself the
 * FPS-ordered output of a tie-free run returns 0..m-1.
 *   tie_out (optional, int32[B]): set to 1 for the clouds whose run saw two candidates tie for an
 *     arg-max (or ran out of valid candidates), else 0.  Conservative: 1 may be a false alarm.
 *   tie_in  (optional, int32[B]): the tie_out of the run that PRODUCED xyz (xyz[b] = the points it
 *     selected, in selection order, same flags).  Clouds with tie_in[b] == 0 get idx = 0..m-1
 *     without running the m-1 dependent rounds; the others are sampled for real.  Bit-identical to
 *     pwclo_furthest_point_sampling either way.  Requires m <= N. */
int pwclo_furthest_point_sampling_prefix(const float *xyz, int B, int N, int m, unsigned flags,
                                         int32_t *idx, const int32_t *tie_in, int32_t *tie_out,
                                         void *stream);

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

/* Same contract and bit-identical result as pwclo_knn, but exact search with spatial pruning: the
 * reference points of every cloud are first sorted into `workspace` (pwclo_knn_workspace_bytes(B,N,S)
 * bytes, 16-byte aligned, caller-owned scratch) -- along their widest axis into equal-count strips,
 * every strip along the second-widest axis, plus the strips' ranges along the third -- and each query
 * then visits strips / scans outwards from its own position only as far as its current K-th distance
 * allows.  Clouds of 8193 .. 16384 points are searched half by half (the halves of the sorted cloud take turns in shared
 * memory).  Falls back to the brute-force kernel when N > 16384 or the workspace is missing / too small.
 * The workspace is [B] records of pwclo_knn_workspace_bytes(1,N,0) bytes, so the record range of a
 * sub-batch can be passed on its own.
 * pwclo_knn_presort + pwclo_knn_search split the two phases so that several searches against the same
 * reference clouds (PW/pose_warp_refinement.py:101-148 searches xyz2 of a level twice, xyz1 once more)
 * share one sort; K of the presort only tunes the strip count. */
size_t pwclo_knn_workspace_bytes(int B, int N, int S);
int pwclo_knn_sorted(const float *xyz, const float *new_xyz, int B, int N, int S, int K,
                     int sum_order, const float *warp_qt, float *warped_out, int32_t *idx,
                     float *dist, void *workspace, size_t workspace_bytes, void *stream);
int pwclo_knn_presort(const float *xyz, int B, int N, int K, void *workspace,
                      size_t workspace_bytes, void *stream);
int pwclo_knn_search(const void *workspace, const float *new_xyz, int B, int N, int S, int K,
                     int sum_order, const float *warp_qt, float *warped_out, int32_t *idx,
                     float *dist, void *stream);

/* ---- B2: fused inference layers (BatchNorm folded by the host side) ------------------------------
 * Feature tensors are POINT-MAJOR [B,N,C] fp32 (the reference keeps [B,C,N]; pwclo_transpose
 * converts), coordinates [B,N,3], neighbour indices [B,S,K] int32.
 *
 * pwclo_layer_t: one folded 1x1 convolution  y = relu(W x + b).
 *   w: [cout/CB][k4][CB] fp32, CB = min(cout,64), k4 = (cin+3)&~3, rows k >= cin zero; i.e. column
 *      blocks of 64 outputs, each stored k-major (input channel major).  16-byte aligned.
 *   b: [cout] fp32, 16-byte aligned.  cout must be 8, 16, 32 or a multiple of 64.
 * The order of the input channels (k) is the kernel's internal concat order, stated per function.
 */
typedef struct {
  const float *w;
  const float *b;
  int cin;
  int cout;
} pwclo_layer_t;

/* set conv (PointnetSAModulePWCLONet.forward, P2/pointnet2_modules.py:208-243) and the grouped half
 * of set upconv (PointnetFPModulePWCLONet.forward, :480-506):
 *   out[b,s,:] = max_k MLP([feats[b,idx[b,s,k],:] (C) | xyz[b,idx[b,s,k]] - new_xyz[b,s] (3)])
 * feats == NULL: the neighbour's absolute xyz takes the place of the features (C = 3, level 1).
 * layers[0].cin = C + 3 with input order (features, xyz_diff); 2 or 3 layers. */
int pwclo_set_conv(const float *xyz, const float *feats, const float *new_xyz, const int32_t *idx,
                   int B, int N, int S, int K, int C, const pwclo_layer_t *layers, int nlayers,
                   float *out, void *stream);

/* shared MLP on the channel concatenation of up to 3 point-major tensors with `rows` rows
 * (FlowPredictor.forward, PW/flowpredictor.py:53-84; post_mlp of set upconv): 1 or 2 layers. */
int pwclo_pointwise_mlp(const float *const *src, const int *channels, int nsrc, int rows,
                        const pwclo_layer_t *layers, int nlayers, float *out, void *stream);

/* CostVolume.forward first aggregation (PW/costvolume.py:75-144): out[B,S,64].
 * mlp1[0] input order: (geo(10), 0, 0, f1(C), f2_grouped(C)) = 2C+12 rows; enc input (geo(10),0,0);
 * geo = (p, q, q-p, |q-p|); mlp2[0] input (enc(64), mlp1_out(64)). */
int pwclo_cost_volume_1(const float *wxyz, const float *f1, const float *xyz2, const float *f2,
                        const int32_t *idx, int B, int S, int N, int K, int C,
                        const pwclo_layer_t *mlp1 /*3*/, const pwclo_layer_t *enc,
                        const pwclo_layer_t *mlp2 /*2*/, float *out, void *stream);

/* CostVolume.forward second aggregation (PW/costvolume.py:150-188): out[B,S,64].
 * mlp3[0] input order: (enc2(64), f1(C), e1_grouped(64)). */
int pwclo_cost_volume_2(const float *wxyz, const float *f1, const float *e1, const int32_t *idx,
                        int B, int S, int K, int C, const pwclo_layer_t *enc,
                        const pwclo_layer_t *mlp3 /*2*/, float *out, void *stream);

/* softmax over points of `mask`, PoseCalculator (PW/pose_calculator.py:47-87), composition with the
 * coarse pose (PW/pose_warp_refinement.py:139,148; coarse_qt == NULL at the coarsest level) and the
 * row `level` (0 = finest) of pose_params[B,4,7] = (t, q/|q|) (PW/pwclo_net.py:195-205).
 * wqt [256,64], wq [4,256], wt [3,256] row-major (plain Conv1d weights), qt_out [B,7] = (q, t). */
int pwclo_pose_head(const float *emb, const float *mask, int B, int S, const float *wqt,
                    const float *bqt, const float *wq, const float *bq, const float *wt,
                    const float *bt, const float *coarse_qt, float *qt_out, float *pose_params,
                    int level, void *stream);

/* Tensor-core versions (tcgen05.mma, activations resident in TMEM) of the four layer kernels above;
 * same arguments and results (fp32-class accuracy: D += tf32(A)*tf32(W) + bf16(A-tf32(A))*bf16(W) +
 * bf16(A)*bf16(W-tf32(W)), kind::tf32 + 2 x kind::f16 MMAs, relative error ~1e-6).
 * pwclo_layer_t.w is packed for the tensor cores, per chunk of 32 input channels (cout*256 bytes):
 *   [tf32_rna(W)          fp32, element (n,k) at float ((n/8)*8 + k/4)*32 + (n%8)*4 + k%4 ]  cout*128 B
 *   [bf16_rn(W)           bf16, element (n,k) at half  ((n/8)*4 + k/8)*64 + (n%8)*8 + k%8 ]  cout*64 B
 *   [bf16_rn(W - tf32(W)) bf16, same layout]                                                cout*64 B
 * (canonical K-major no-swizzle UMMA layouts; reference packer: pwclonet_pylidarslam_b200/tc_pack.py
 * pack_tc2).  cout must be 64 or 128.  Input-channel orders: set_conv (features(C), xyz_diff(3), 0 x13);
 * cost_volume_1 mlp1[0] (f1(C), f2(C), geo(10), 0 x6), enc (geo(10), 0 x6); cost_volume_2 as above. */
int pwclo_set_conv_tc(const float *xyz, const float *feats, const float *new_xyz, const int32_t *idx,
                      int B, int N, int S, int K, int C, const pwclo_layer_t *layers, int nlayers,
                      float *out, void *stream);
int pwclo_pointwise_mlp_tc(const float *const *src, const int *channels, int nsrc, int rows,
                           const pwclo_layer_t *layers, int nlayers, float *out, void *stream);
int pwclo_cost_volume_1_tc(const float *wxyz, const float *f1, const float *xyz2, const float *f2,
                           const int32_t *idx, int B, int S, int N, int K, int C,
                           const pwclo_layer_t *mlp1, const pwclo_layer_t *enc,
                           const pwclo_layer_t *mlp2, float *out, void *stream);
int pwclo_cost_volume_2_tc(const float *wxyz, const float *f1, const float *e1, const int32_t *idx,
                           int B, int S, int K, int C, const pwclo_layer_t *enc,
                           const pwclo_layer_t *mlp3, float *out, void *stream);

/* Diagnostic: D[128,N] = A[128,K] * W[N,K]^T on the tcgen05 tensor cores (A split hi/lo in TMEM, W packed
 * hi/lo in the canonical K-major chunks, see csrc/tc_mma.cuh).  mode 1 = plain TF32, 3 = 3xTF32 (wpk from
 * tc_pack.pack_tc), 5 = tf32 + bf16 correction terms (wpk from tc_pack.pack_tc2, the layout above). */
int pwclo_tc_selftest(const float *A, const float *wpk, int K, int N, int mode, float *D, void *stream);

/* out[b,j,:] = xyz[b,idx[b,j],:]  (gather_operation on [B,N,3], P2/pointnet2_modules.py:200-206) */
int pwclo_gather_rows3(const float *xyz, const int32_t *idx, int B, int N, int M, float *out,
                       void *stream);
/* [B,C,N] -> [B,N,C] (to_point_major != 0) or back */
int pwclo_transpose(const float *in, int B, int C, int N, int to_point_major, float *out,
                    void *stream);

/* ---- training-side rows ------------------------------------------------ */

/* _PWCLONetLossModule.forward (slam/training/loss_modules.py:424-544, ExponentialWeights :171-196)
 * and its gradient in one launch.  pred[B,4,7] = (t, q) rows finest level first, gt[B,7] = (t, q),
 * s[2] = (s_trans, s_rot) of ExponentialWeights (with_exp_weights != 0) or the fixed (trans, rot)
 * weights.  out[16] = loss | loss_l1..l4 | loss_rot_l1..l4 | loss_trans_l1..l4 | s[0], s[1] | B.
 * grad_pred[B,4,7] and grad_s[2] (either may be NULL) receive d loss / d pred and d loss / d s. */
int pwclo_pose_loss(const float *pred, const float *gt, const float *s, int B, int with_exp_weights,
                    float *out, float *grad_pred, float *grad_s, void *stream);

/* torch.optim.Adam step (the reference's optimiser, slam/training/trainer.py:309-323) over one flat,
 * 16-byte aligned fp32 arena of n elements: param / exp_avg / exp_avg_sq updated in place from grad.
 * grad is multiplied by grad_scale first (1/world_size after the gradient all-reduce), then L2 weight
 * decay is added; `step` counts from 1.  One launch, 28 bytes of HBM traffic per element. */
int pwclo_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, size_t n,
                    int step, float lr, float beta1, float beta2, float eps, float weight_decay,
                    float grad_scale, void *stream);
/* The same update with the step counter and the learning rate in device memory, so that a training step can
 * be captured once in a CUDA graph and replayed: state_i[0] = completed steps (incremented by the call),
 * state_f[0] = learning rate (the host may rewrite it between replays), state_f[1..2] = scratch. */
int pwclo_adam_step_dev(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, size_t n,
                        int32_t *state_i, float *state_f, float beta1, float beta2, float eps,
                        float weight_decay, float grad_scale, void *stream);
/* Pose warp for training (PW/PWCLO_utils.py:31-63): out[B,3,N] = q (x) [0, xyz] (x) q^-1 + t, xyz channel-major [B,3,N],
 * q [B,4] scalar first, t [B,3]; and its gradient: grad_q [B,4], grad_t [B,3] (overwritten), grad_xyz [B,3,N] (NULL =
 * not needed).  One launch each. */
int pwclo_warp_fwd(const float *xyz, const float *q, const float *t, int B, int N, float *out, void *stream);
int pwclo_warp_bwd(const float *xyz, const float *q, const float *grad_out, int B, int N, float *grad_xyz,
                   float *grad_q, float *grad_t, void *stream);

/* Geometry channels of the cost volume in training (PW/costvolume.py:94-105, :159-169): center [B,3,S], grouped
 * [B,3,S,K] -> out [B,10,S,K] = (p, g, g - p, sqrt(|g - p|^2 + 1e-20)); backward: grad_center [B,3,S] and grad_grouped
 * [B,3,S,K] (either may be NULL; both overwritten, no atomics).  One launch each. */
int pwclo_cost_geometry_fwd(const float *center, const float *grouped, int B, int S, int K, float *out, void *stream);
int pwclo_cost_geometry_bwd(const float *center, const float *grouped, const float *grad_out, int B, int S, int K,
                            float *grad_center, float *grad_grouped, void *stream);

/* Max over the neighbour axis in training (F.max_pool2d(x, [1, K]), P2/pointnet2_modules.py:239-243, :499-506) and its
 * backward: x [rows, K] -> y [rows], arg uint8 [rows] (first maximum wins, NaN propagates); dx [rows, K] = dy at arg, 0
 * elsewhere (every element written: no zero-fill needed).  K <= 255. */
int pwclo_maxpool_lastdim_fwd(const float *x, long long rows, int K, float *y, unsigned char *arg, void *stream);
int pwclo_maxpool_lastdim_bwd(const float *dy, const unsigned char *arg, long long rows, int K, float *dx, void *stream);

/* Attentive pooling of the cost volume in training (PW/costvolume.py:139-145 and :181-188: F.softmax(w, dim=3), the
 * product with the grouped features, torch.sum over the neighbour axis) and its backward: w, x [rows, K] contiguous
 * (rows = B*C*S), out [rows] = sum_k softmax(w[r,:])[k] * x[r,k]; grad_w, grad_x [rows, K] from grad_out [rows] (the
 * softmax is recomputed from w).  K in {4, 6, 8, 16, 32}, else PWCLO_EUNSUPPORTED.  16-byte aligned tensors. */
int pwclo_softmax_pool_fwd(const float *w, const float *x, long long rows, int K, float *out, void *stream);
int pwclo_softmax_pool_bwd(const float *w, const float *x, const float *grad_out, long long rows, int K, float *grad_w,
                           float *grad_x, void *stream);

/* Train-mode BatchNorm + ReLU of a shared-MLP layer (P2/pytorch_utils.py:86-167: nn.BatchNorm2d(eps) over
 * (B, S, K) followed by ReLU), forward and backward, two launches each.  x, y, dy, dx: [B, C, HW] contiguous.
 * Forward: batch mean / biased variance per channel (double accumulation), running statistics updated as
 * nn.BatchNorm does (momentum, unbiased variance; pass NULL for both to skip), save_mean / save_invstd [C] for
 * the backward.  Backward: dx, dgamma [C], dbeta [C] (overwritten) from x, dy (the ReLU mask is recomputed from x).
 * workspace: pwclo_bn_relu_workspace_bytes(B, C, HW) bytes, 8-byte aligned. */
size_t pwclo_bn_relu_workspace_bytes(int B, int C, int HW);
int pwclo_bn_relu_train_fwd(const float *x, const float *gamma, const float *beta, int B, int C, int HW,
                            float eps, float momentum, float *running_mean, float *running_var,
                            float *y, float *save_mean, float *save_invstd, void *workspace, void *stream);
int pwclo_bn_relu_train_bwd(const float *x, const float *dy, const float *gamma, const float *beta,
                            const float *save_mean, const float *save_invstd, int B, int C, int HW,
                            float *dx, float *dgamma, float *dbeta, void *workspace, void *stream);
/* 1x1 convolutions with very few channels (the first set-conv layers, 6 -> 8 -> 8 -> 16 on ~5e5 positions), for training:
 * streaming forward / input gradient and a two-launch deterministic weight gradient dW[CO][CI] = sum dy * x.  Supported
 * (CI, CO): forward / input gradient (6,8) (8,8) (8,16) (16,8) (3,8), weight gradient (6,8) (8,8) (8,16) (3,8);
 * anything else returns PWCLO_EUNSUPPORTED and the caller keeps the library GEMM. */
int pwclo_conv1x1_small(const float *x, const float *w, int transpose_w, int B, int CI, int CO, long long HW,
                        float *y, void *stream);
size_t pwclo_conv1x1_wgrad_workspace_bytes(int B, int CI, int CO, long long HW);
int pwclo_conv1x1_wgrad(const float *x, const float *dy, int B, int CI, int CO, long long HW, float *dw,
                        void *workspace, void *stream);

/* ---- input pipeline ----------------------------------------------------------- */

/* KITTI scans -> network input, replacing the per-frame numpy code of
 * slam/dataset/kitti_odometry_dataset.py:375-397 (Tr transform in float64) and filter_pcd :149-172
 * (drop y > 1.1 and |x|,|z| >= 30, keep `npoints` random survivors without replacement, pad with
 * replacement when there are fewer), plus the optional augmentation transform of :404-446.
 *   raw          packed scans, float32 [total_points,4] (x, y, z, reflectance), 16-byte aligned
 *   offsets      int64 [nscan+1], point offset of every scan in `raw` (device memory); scans < 2^22 points
 *   Tr           float64 [12] (or [nscan,12] when tr_per_scan != 0): rows 0..2 of the 4x4 calibration
 *   post         float64 [nscan,12] applied after the selection, before the float32 cast; NULL = none
 *   seed         the sample is the npoints survivors with the smallest (philox4x32-10 key, index)
 *   out          float32 [nscan,npoints,3];  sel_idx int32 [nscan,npoints] source rows (may be NULL);
 *   survivors    int32 [nscan] points that passed the filter, -1 on an internal capacity overflow (may be NULL)
 *   workspace    pwclo_prepare_scans_workspace_bytes(total_points, nscan) bytes of device memory, 16-byte aligned
 * Two launches; every scan point is read from HBM once (16 B per point). */
size_t pwclo_prepare_scans_workspace_bytes(long long total_points, int nscan);
int pwclo_prepare_scans(const float *raw, const long long *offsets, int nscan, long long total_points,
                        const double *Tr, int tr_per_scan, const double *post,
                        unsigned long long seed, int npoints, float *out, int32_t *sel_idx,
                        int32_t *survivors, void *workspace, size_t workspace_bytes, void *stream);
/* The same with the crop spelled out: a point is dropped when ground_sign * (P[ground_axis] - ground_thr) > 0 or
 * when |P[near_axis_a]| or |P[near_axis_b]| is not below near_thr.  pwclo_prepare_scans = (1, +1, 1.1, 0, 2, 30.0);
 * KITTI-360 (slam/dataset/kitti_360_dataset_2.py:113-123: velodyne frame, Tr = identity, comparisons in float32) =
 * (2, -1, (double)(float)-1.43, 0, 1, (double)(float)near_treshold). */
int pwclo_prepare_scans_crop(const float *raw, const long long *offsets, int nscan, long long total_points,
                             const double *Tr, int tr_per_scan, const double *post,
                             unsigned long long seed, int npoints, int ground_axis, int ground_sign,
                             double ground_thr, int near_axis_a, int near_axis_b, double near_thr,
                             float *out, int32_t *sel_idx, int32_t *survivors, void *workspace,
                             size_t workspace_bytes, void *stream);

/* ---- pose post-processing ------------------------------------------------ */

/* pose_params rows (t[3], q[4] scalar first; `row_stride` floats apart, 28 for the finest-level row of
 * [B,4,7]) -> float64 4x4 matrices [B,16]: quat2mat of train.py:762-796 (fp32, op for op) with t in the
 * last column; invert != 0 returns the inverse as train.py:885 does (np.linalg.inv). */
int pwclo_pose_to_matrix(const float *pose_params, int B, int row_stride, int invert, double *out,
                         void *stream);
/* KITTI360_TRANSFORMATIONS.convert_to_absolute, dict branch (slam/common/kitti360_utils.py:424-427):
 * abs_f = inv(rel_f @ inv(abs_{f-1})) = abs_{f-1} * inv(rel_f), abs_{-1} = first (NULL = identity);
 * rel, out float64 [F,16].  One CTA, parallel prefix product. */
int pwclo_accumulate_poses(const double *rel, int F, const double *first, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PWCLO_B200_H */
