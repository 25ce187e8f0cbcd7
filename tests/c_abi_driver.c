/* tests/c_abi_driver.c -- a plain C caller of libpwclo_b200.so (no Python, no torch): allocates device
 * buffers with the CUDA runtime, calls the C ABI of include/pwclo_b200.h on a CUDA stream and compares the
 * results with the C oracle (oracle/pointnet2_cpu.c, linked as liboracle_pwclo.so) bit for bit.
 * Built and run by tests/test_c_abi_gpu.py.  Exit code 0 = every check passed. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pwclo_b200.h"

void oracle_fps(const float *xyz, int B, int n, int m, int origin_skip, int thread_cap, int32_t *idx);
void oracle_knn(const float *xyz, const float *new_xyz, int B, int N, int S, int K, int sum_order, int32_t *idx, float *dist_out);
void oracle_group_points(const float *points, const int32_t *idx, int B, int C, int N, int S, int K, float *out);
void oracle_ball_query(const float *new_xyz, const float *xyz, int B, int n, int m, float radius, int nsample, int32_t *idx);

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } \
  } while (0)
#define PW(x)                                                                          \
  do {                                                                                 \
    int rc_ = (x);                                                                     \
    if (rc_ != 0) { fprintf(stderr, "%s: %s (%d)\n", #x, pwclo_error_string(rc_), rc_); return 3; } \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float frand(void) {   /* xorshift64*, uniform in [-20, 20) */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (float)((rng_state * 0x2545F4914F6CDD1Dull) >> 40) / 16777216.0f * 40.0f - 20.0f;
}

int main(void) {
  const int B = 3, N = 4096, M = 512, K = 16, C = 8, NS = 8;
  printf("%s\n", pwclo_version());
  float *xyz = (float *)malloc(sizeof(float) * B * N * 3), *feat = (float *)malloc(sizeof(float) * B * C * N);
  for (int i = 0; i < B * N * 3; ++i) xyz[i] = frand();
  for (int i = 0; i < B * C * N; ++i) feat[i] = frand();
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  float *d_xyz, *d_new, *d_feat, *d_grp;
  int32_t *d_fps, *d_knn, *d_ball;
  CK(cudaMalloc((void **)&d_xyz, sizeof(float) * B * N * 3));
  CK(cudaMalloc((void **)&d_new, sizeof(float) * B * M * 3));
  CK(cudaMalloc((void **)&d_feat, sizeof(float) * B * C * N));
  CK(cudaMalloc((void **)&d_grp, sizeof(float) * B * C * M * K));
  CK(cudaMalloc((void **)&d_fps, sizeof(int32_t) * B * M));
  CK(cudaMalloc((void **)&d_knn, sizeof(int32_t) * B * M * K));
  CK(cudaMalloc((void **)&d_ball, sizeof(int32_t) * B * M * NS));
  CK(cudaMemcpyAsync(d_xyz, xyz, sizeof(float) * B * N * 3, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_feat, feat, sizeof(float) * B * C * N, cudaMemcpyHostToDevice, st));

  /* FPS -> gather the sampled coordinates -> kNN around them -> group features -> ball query */
  PW(pwclo_furthest_point_sampling(d_xyz, B, N, M, PWCLO_FPS_ORIGIN_SKIP, d_fps, st));
  PW(pwclo_gather_rows3(d_xyz, d_fps, B, N, M, d_new, st));
  PW(pwclo_knn(d_xyz, d_new, B, N, M, K, PWCLO_KNN_SUM_XY_Z, NULL, NULL, d_knn, NULL, st));
  PW(pwclo_group_points(d_feat, d_knn, B, C, N, M, K, d_grp, st));
  PW(pwclo_ball_query(d_new, d_xyz, B, N, M, 2.5f, NS, d_ball, st));

  int32_t *fps = (int32_t *)malloc(sizeof(int32_t) * B * M), *knn = (int32_t *)malloc(sizeof(int32_t) * B * M * K);
  int32_t *ball = (int32_t *)malloc(sizeof(int32_t) * B * M * NS);
  float *grp = (float *)malloc(sizeof(float) * B * C * M * K), *newx = (float *)malloc(sizeof(float) * B * M * 3);
  CK(cudaMemcpyAsync(fps, d_fps, sizeof(int32_t) * B * M, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(knn, d_knn, sizeof(int32_t) * B * M * K, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ball, d_ball, sizeof(int32_t) * B * M * NS, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(grp, d_grp, sizeof(float) * B * C * M * K, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(newx, d_new, sizeof(float) * B * M * 3, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));

  /* the same chain on the CPU oracle */
  int32_t *o_fps = (int32_t *)malloc(sizeof(int32_t) * B * M), *o_knn = (int32_t *)malloc(sizeof(int32_t) * B * M * K);
  int32_t *o_ball = (int32_t *)malloc(sizeof(int32_t) * B * M * NS);
  float *o_grp = (float *)malloc(sizeof(float) * B * C * M * K), *o_new = (float *)malloc(sizeof(float) * B * M * 3);
  oracle_fps(xyz, B, N, M, 1, 512, o_fps);
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < M; ++j) memcpy(o_new + ((size_t)b * M + j) * 3, xyz + ((size_t)b * N + o_fps[b * M + j]) * 3, 12);
  oracle_knn(xyz, o_new, B, N, M, K, 0, o_knn, NULL);
  oracle_group_points(feat, o_knn, B, C, N, M, K, o_grp);
  oracle_ball_query(o_new, xyz, B, N, M, 2.5f, NS, o_ball);

  int bad = 0;
  bad += memcmp(fps, o_fps, sizeof(int32_t) * B * M) != 0;
  printf("fps   %s\n", memcmp(fps, o_fps, sizeof(int32_t) * B * M) ? "MISMATCH" : "bit-exact");
  bad += memcmp(newx, o_new, sizeof(float) * B * M * 3) != 0;
  printf("gather %s\n", memcmp(newx, o_new, sizeof(float) * B * M * 3) ? "MISMATCH" : "bit-exact");
  bad += memcmp(knn, o_knn, sizeof(int32_t) * B * M * K) != 0;
  printf("knn   %s\n", memcmp(knn, o_knn, sizeof(int32_t) * B * M * K) ? "MISMATCH" : "bit-exact");
  bad += memcmp(grp, o_grp, sizeof(float) * B * C * M * K) != 0;
  printf("group %s\n", memcmp(grp, o_grp, sizeof(float) * B * C * M * K) ? "MISMATCH" : "bit-exact");
  bad += memcmp(ball, o_ball, sizeof(int32_t) * B * M * NS) != 0;
  printf("ball  %s\n", memcmp(ball, o_ball, sizeof(int32_t) * B * M * NS) ? "MISMATCH" : "bit-exact");

  /* error behaviour: argument errors are negative codes, nothing calls exit() */
  if (pwclo_furthest_point_sampling(NULL, B, N, M, 0, d_fps, st) != PWCLO_EINVAL) { printf("EINVAL expected\n"); ++bad; }
  printf(bad ? "FAILED\n" : "OK\n");
  return bad ? 1 : 0;
}
