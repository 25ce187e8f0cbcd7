/*
 * oracle/pointnet2_cpu.c -- TEST INFRASTRUCTURE ONLY (never imported by the product package).
 *
 * Plain-C, single-threaded CPU restatement of the algorithms on the PWCLO-Net point-cloud hot
 * path, written from the reference's behaviour (not its text).  Each function cites the
 * reference file:line it follows.  Paths are relative to /root/reference; EXT =
 * slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops/_ext-src/src and
 * P2 = slam/models/Pointnet2_PyTorch/pointnet2_ops_lib/pointnet2_ops.
 *
 * Floating point: compile with -O2 -ffp-contract=off.  Where nvcc contracts the reference's
 * CUDA expression a*a + b*b + c*c into FMUL,FFMA,FFMA we write the fmaf() chain explicitly;
 * where the reference runs separate torch kernels (kNN) we use unfused fp32 operations.
 *
 * Parity status: pinned against (i) the unmodified reference Python model executed in the build
 * container through oracle/ref_shim.py (golden vectors in tests/golden, generator script
 * oracle/make_golden.py) and (ii) on the GPU box, the reference's own CUDA kernels recompiled for
 * sm_100a (oracle/_ref, recipe oracle/build_ref_ext.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* EXT/../include/cuda_utils.h:13-19  opt_n_threads(): largest power of two <= work, clamp [1,cap] */
static int ref_block_threads(int work, int cap) {
  int p = 1;
  while (p * 2 <= work && p * 2 <= cap) p *= 2;
  return p < 1 ? 1 : p;
}

EXPORT int oracle_ref_block_threads(int work, int cap) { return ref_block_threads(work, cap); }

/* squared distance exactly as nvcc compiles  dx*dx + dy*dy + dz*dz : the SECOND product is the
 * plain FMUL, the first and third are fused (SASS of the sm_100a build of EXT/sampling_gpu.cu:
 * FMUL dy*dy ; FFMA dx*dx+. ; FFMA dz*dz+.  -- checked with cuobjdump). */
static inline float dist2_fma(float dx, float dy, float dz) {
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/*
 * Furthest point sampling.  EXT/sampling_gpu.cu:69-173 (kernel), EXT/sampling.cpp:66-87 (temp=1e10,
 * idx zero-initialised).  The reference runs T = opt_n_threads(n) threads; thread t scans
 * k = t, t+T, ... with a strict '>' (first maximum wins, start (-1, 0)), then a shared-memory
 * tree halves the active range; on equal values the lower slot wins (sampling_gpu.cu:60-66).
 * We simulate exactly that: per-slot scan followed by the same tree.
 *   origin_skip != 0 : candidates with x^2+y^2+z^2 <= 1e-3 (float value vs double literal) are
 *                      neither updated nor selectable (sampling_gpu.cu:100-101).
 *   origin_skip == 0, thread_cap = 1024 : the orphan variant /sampling_gpu_copy.cu:129-131,143-148.
 */
EXPORT void oracle_fps(const float *xyz, int B, int n, int m, int origin_skip, int thread_cap, int32_t *idx) {
  int T = ref_block_threads(n, thread_cap);
  float *temp = (float *)malloc(sizeof(float) * (size_t)n);
  float *dv = (float *)malloc(sizeof(float) * (size_t)T);
  int *di = (int *)malloc(sizeof(int) * (size_t)T);
  for (int b = 0; b < B; ++b) {
    const float *p = xyz + (size_t)b * n * 3;
    int32_t *out = idx + (size_t)b * m;
    for (int j = 0; j < m; ++j) out[j] = 0;
    if (m <= 0) continue;
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    int old = 0;
    for (int j = 1; j < m; ++j) {
      float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int t = 0; t < T; ++t) {
        float best = -1.0f;
        int besti = 0;
        for (int k = t; k < n; k += T) {
          float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          if (origin_skip) {
            float mag = dist2_fma(x2, y2, z2);
            if ((double)mag <= 1e-3) continue;
          }
          float d = dist2_fma(x2 - x1, y2 - y1, z2 - z1);
          float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          if (d2 > best) { best = d2; besti = k; }
        }
        dv[t] = best;
        di[t] = besti;
      }
      for (int h = T / 2; h >= 1; h /= 2) {
        for (int t = 0; t < h; ++t) {
          float v1 = dv[t], v2 = dv[t + h];
          if (v2 > v1) { dv[t] = v2; di[t] = di[t + h]; }
        }
      }
      old = di[0];
      out[j] = old;
    }
  }
  free(temp); free(dv); free(di);
}

/* EXT/sampling_gpu.cu:8-20  out[b,c,j] = points[b,c,idx[b,j]] */
EXPORT void oracle_gather_points(const float *points, const int32_t *idx, int B, int C, int N, int M, float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < M; ++j)
        out[((size_t)b * C + c) * M + j] = points[((size_t)b * C + c) * N + idx[(size_t)b * M + j]];
}

/* EXT/sampling_gpu.cu:34-47  scatter-add (the reference uses float atomics: order undefined;
 * the oracle accumulates in double and rounds once, tests use a tolerance). */
EXPORT void oracle_gather_points_grad(const float *grad_out, const int32_t *idx, int B, int C, int N, int M, float *grad_points) {
  double *acc = (double *)calloc((size_t)N, sizeof(double));
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      memset(acc, 0, sizeof(double) * (size_t)N);
      for (int j = 0; j < M; ++j) acc[idx[(size_t)b * M + j]] += (double)grad_out[((size_t)b * C + c) * M + j];
      for (int k = 0; k < N; ++k) grad_points[((size_t)b * C + c) * N + k] = (float)acc[k];
    }
  free(acc);
}

/* EXT/group_points_gpu.cu:8-28  out[b,c,j,k] = points[b,c,idx[b,j,k]] */
EXPORT void oracle_group_points(const float *points, const int32_t *idx, int B, int C, int N, int S, int K, float *out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float *row = points + ((size_t)b * C + c) * N;
      for (size_t e = 0; e < (size_t)S * K; ++e)
        out[((size_t)b * C + c) * S * K + e] = row[idx[(size_t)b * S * K + e]];
    }
}

/* EXT/group_points_gpu.cu:43-64  scatter-add of grad_out[b,c,j,k] into grad_points[b,c,idx] */
EXPORT void oracle_group_points_grad(const float *grad_out, const int32_t *idx, int B, int C, int N, int S, int K, float *grad_points) {
  double *acc = (double *)calloc((size_t)N, sizeof(double));
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      memset(acc, 0, sizeof(double) * (size_t)N);
      for (size_t e = 0; e < (size_t)S * K; ++e)
        acc[idx[(size_t)b * S * K + e]] += (double)grad_out[((size_t)b * C + c) * S * K + e];
      for (int k = 0; k < N; ++k) grad_points[((size_t)b * C + c) * N + k] = (float)acc[k];
    }
  free(acc);
}

/* EXT/ball_query_gpu.cu:9-44  first nsample hits (ascending index) with d2 < r*r; all slots are
 * pre-filled with the first hit; rows without hit stay zero (ball_query.cpp:20-22 zero-inits). */
EXPORT void oracle_ball_query(const float *new_xyz, const float *xyz, int B, int n, int m, float radius, int nsample, int32_t *idx) {
  float r2 = radius * radius;
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < m; ++j) {
      const float *q = new_xyz + ((size_t)b * m + j) * 3;
      int32_t *o = idx + ((size_t)b * m + j) * nsample;
      for (int l = 0; l < nsample; ++l) o[l] = 0;
      int cnt = 0;
      for (int k = 0; k < n && cnt < nsample; ++k) {
        const float *p = xyz + ((size_t)b * n + k) * 3;
        float d2 = dist2_fma(q[0] - p[0], q[1] - p[1], q[2] - p[2]);
        if (d2 < r2) {
          if (cnt == 0) for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt++] = k;
        }
      }
    }
}

/* EXT/interpolate_gpu.cu:9-59  three nearest of `known` per `unknown`; fp32 d compared with the
 * double running bests (init 1e40), strict '<' so the lowest index wins ties; dist2 stored as float. */
EXPORT void oracle_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int32_t *idx) {
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < n; ++j) {
      const float *u = unknown + ((size_t)b * n + j) * 3;
      double b1 = 1e40, b2 = 1e40, b3 = 1e40;
      int i1 = 0, i2 = 0, i3 = 0;
      for (int k = 0; k < m; ++k) {
        const float *p = known + ((size_t)b * m + k) * 3;
        float d = dist2_fma(u[0] - p[0], u[1] - p[1], u[2] - p[2]);
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
        else if (d < b3) { b3 = d; i3 = k; }
      }
      float *od = dist2 + ((size_t)b * n + j) * 3;
      int32_t *oi = idx + ((size_t)b * n + j) * 3;
      od[0] = (float)b1; od[1] = (float)b2; od[2] = (float)b3;
      oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

/* EXT/interpolate_gpu.cu:72-101  out[b,c,j] = p[i1]*w1 + p[i2]*w2 + p[i3]*w3, contracted by nvcc to
 * FMUL,FFMA,FFMA = fma(p3,w3, fma(p1,w1, p2*w2))  (second product is the plain FMUL). */
EXPORT void oracle_three_interpolate(const float *points, const int32_t *idx, const float *weight, int B, int c, int m, int n, float *out) {
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        const int32_t *ii = idx + ((size_t)b * n + j) * 3;
        const float *w = weight + ((size_t)b * n + j) * 3;
        const float *row = points + ((size_t)b * c + l) * m;
        out[((size_t)b * c + l) * n + j] = fmaf(row[ii[2]], w[2], fmaf(row[ii[0]], w[0], row[ii[1]] * w[1]));
      }
}

/* EXT/interpolate_gpu.cu:116-143  three atomic adds of grad_out*w per element (order undefined in
 * the reference; double accumulation here, tolerance in tests). */
EXPORT void oracle_three_interpolate_grad(const float *grad_out, const int32_t *idx, const float *weight, int B, int c, int n, int m, float *grad_points) {
  double *acc = (double *)calloc((size_t)m, sizeof(double));
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < c; ++l) {
      memset(acc, 0, sizeof(double) * (size_t)m);
      for (int j = 0; j < n; ++j) {
        const int32_t *ii = idx + ((size_t)b * n + j) * 3;
        const float *w = weight + ((size_t)b * n + j) * 3;
        float g = grad_out[((size_t)b * c + l) * n + j];
        acc[ii[0]] += (double)(g * w[0]);
        acc[ii[1]] += (double)(g * w[1]);
        acc[ii[2]] += (double)(g * w[2]);
      }
      for (int k = 0; k < m; ++k) grad_points[((size_t)b * c + l) * m + k] = (float)acc[k];
    }
  free(acc);
}

/*
 * kNN.  P2/pytorch_utils.py:12-49: dist = sqrt(sum((q-r)^2, -1) + 1e-8) as separate fp32 torch
 * kernels (no contraction), then topk(k, largest=False, sorted).  The summation order of the
 * size-3 reduction is implementation defined in torch: sequential (x2+y2)+z2 on CPU (sum_order 0),
 * (x2+z2)+y2 in torch's CUDA reduce kernel (sum_order 1, see DESIGN.md "kNN formulation").
 * topk's order among equal distances is unspecified; the oracle (and the CUDA kernel) use
 * (distance ascending, index ascending).  dist_out may be NULL.
 */
static inline float knn_dist(const float *q, const float *r, int sum_order) {
  float dx = q[0] - r[0], dy = q[1] - r[1], dz = q[2] - r[2];
  float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  float s = sum_order == 0 ? (xx + yy) + zz : (xx + zz) + yy;
  return sqrtf(s + 1e-8f);
}

EXPORT void oracle_knn(const float *xyz, const float *new_xyz, int B, int N, int S, int K, int sum_order, int32_t *idx, float *dist_out) {
  float *bv = (float *)malloc(sizeof(float) * (size_t)K);
  int *bi = (int *)malloc(sizeof(int) * (size_t)K);
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < S; ++j) {
      const float *q = new_xyz + ((size_t)b * S + j) * 3;
      int cnt = 0;
      for (int k = 0; k < N; ++k) {
        float d = knn_dist(q, xyz + ((size_t)b * N + k) * 3, sum_order);
        if (cnt == K && !(d < bv[K - 1])) continue;
        int pos = cnt < K ? cnt : K - 1;
        while (pos > 0 && bv[pos - 1] > d) { bv[pos] = bv[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bv[pos] = d; bi[pos] = k;
        if (cnt < K) ++cnt;
      }
      for (int l = 0; l < K; ++l) {
        idx[((size_t)b * S + j) * K + l] = l < cnt ? bi[l] : 0;
        if (dist_out) dist_out[((size_t)b * S + j) * K + l] = l < cnt ? bv[l] : INFINITY;
      }
    }
  free(bv); free(bi);
}

/* Full distance row (for tie-group analysis in tests): out[b,j,k] = knn_dist */
EXPORT void oracle_knn_distances(const float *xyz, const float *new_xyz, int B, int N, int S, int sum_order, float *out) {
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < S; ++j)
      for (int k = 0; k < N; ++k)
        out[((size_t)b * S + j) * N + k] = knn_dist(new_xyz + ((size_t)b * S + j) * 3, xyz + ((size_t)b * N + k) * 3, sum_order);
}
