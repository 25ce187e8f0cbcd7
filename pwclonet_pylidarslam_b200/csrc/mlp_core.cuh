// Fused shared-MLP machinery for the PWCLO-Net layer kernels (sm_100a).
//
// Every PWCLO-Net layer is  gather(idx) -> concat -> 2..5 x [1x1 conv + BN + ReLU] -> max / softmax
// pool over the neighbour axis.  With BatchNorm folded (inference) each conv is  y = relu(W x + b)
// on a [rows x channels] tile where rows = (points of the CTA) x (neighbours).  The tile never
// leaves the SM: activations ping-pong between two shared-memory buffers, weights are streamed
// through a double-buffered shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx) issued by one elected thread, prefetching across layer boundaries.
//
// Arithmetic is fp32 FFMA with sequential-k accumulation (features must match the reference to 1e-4
// relative through ~10 chained layers and two softmaxes; see DESIGN.md "why not bf16/tf32 yet").
//
// Shared-memory layout of an activation tile: row-major [R][ld], ld = 4*odd so that the float4
// loads/stores of 8 consecutive rows fall into 8 distinct 16-byte bank groups (conflict-free).
// Thread mapping of one GEMM (256 threads = 8 warps, WN warps along the output columns, WM = 8/WN
// along the rows): lane l of row-warp wm owns rows  wm*32*TM + l + 32*i (i < TM)  and the TN = CB/WN
// consecutive columns of its column-warp; x is read as float4 over 4 consecutive k, weights as
// float4 over 4 consecutive columns (warp-wide broadcast).
#pragma once
#include "common.cuh"

namespace pwclo {

constexpr int LT = 256;     // threads per CTA of every layer kernel
constexpr int KC = 32;      // weight rows (k) per streamed chunk
constexpr int CBMAX = 64;   // output columns per column block

__host__ __device__ constexpr int round4(int c) { return (c + 3) & ~3; }
// smallest multiple of 4 that is >= c and whose quotient by 4 is odd
__host__ __device__ constexpr int ld_for(int c) { return (((c + 3) / 4) | 1) * 4; }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarrier_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarrier_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  }
}
// polling variant with back-off: for the many waiters of a long phase, so that they do not steal
// issue slots from the single MMA-issuing / TMA-issuing warps that share their scheduler
__device__ __forceinline__ void mbarrier_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// One folded conv layer.  w = [cout/CB][k4][CB] with CB = min(cout, 64): column blocks of 64
// outputs, each k-major (rows >= the true input width are zero), so that every streamed chunk
// (32 k-rows of one column block) is one contiguous TMA bulk copy.  b = [cout].
struct Layer {
  const float* w;
  const float* b;
  int k4;     // round4(input channels)
  int cout;   // multiple of 8
};

// Double-buffered weight ring shared by all GEMMs of a kernel.
struct WeightPipe {
  float* buf;       // smem [2][KC*CBMAX]
  uint64_t* bar;    // smem [2]
  uint32_t stage;   // global stage counter (buffer = stage & 1, parity = (stage >> 1) & 1)
  bool primed;      // the next GEMM's stage 0 has already been issued

  __device__ __forceinline__ void init(float* b, uint64_t* m) {
    buf = b; bar = m; stage = 0; primed = false;
    if (threadIdx.x == 0) {
      mbarrier_init(&bar[0], 1);
      mbarrier_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  // thread 0 only: stream rows [k0,k0+kc) of column block `blk` (cb columns) into ring slot `s & 1`
  __device__ __forceinline__ void issue(const Layer& L, int blk, int k0, int kc, int cb, uint32_t s) {
    float* dst = buf + (s & 1) * (KC * CBMAX);
    uint64_t* b = &bar[s & 1];
    const uint32_t bytes = (uint32_t)(kc * cb * sizeof(float));
    mbarrier_expect_tx(b, bytes);
    tma_bulk_g2s(dst, L.w + ((size_t)blk * L.k4 + k0) * cb, bytes, b);
  }
  // issue stage 0 of a layer ahead of time (e.g. before the gather phase)
  __device__ __forceinline__ void prime(const Layer& L) {
    if (threadIdx.x == 0) {
      const int cb = L.cout < CBMAX ? L.cout : CBMAX;
      issue(L, 0, 0, L.k4 < KC ? L.k4 : KC, cb, stage);
    }
    primed = true;
  }
};

// Y[:, ycol0 + (0..cout)) = act(X[:, 0..k4) * W + b)   for the R rows of the tile.
// `next`: layer whose first weight chunk is prefetched while the last chunk of this one is consumed.
template <int R, int CB, bool RELU>
__device__ __forceinline__ void gemm_block(const float* __restrict__ X, int ldx, float* __restrict__ Y, int ldy,
                                           int ycol0, const Layer& L, const Layer* next, WeightPipe& pipe) {
  constexpr int WN = CB / 4 < 8 ? CB / 4 : 8;
  constexpr int WM = 8 / WN;
  constexpr int TM = R / (32 * WM);
  constexpr int TN = CB / WN;
  static_assert(TM >= 1 && TN % 4 == 0, "bad GEMM tiling");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wn = warp % WN, wm = warp / WN;
  const int row0 = wm * 32 * TM + lane;
  const int nblk = L.cout / CB;
  const int nchunk = (L.k4 + KC - 1) / KC;
  const int nstage = nblk * nchunk;

  if (!pipe.primed && threadIdx.x == 0) pipe.issue(L, 0, 0, L.k4 < KC ? L.k4 : KC, CB, pipe.stage);
  pipe.primed = false;

  float acc[TM][TN];
  for (int s = 0; s < nstage; ++s) {
    const int blk = s / nchunk, ch = s - blk * nchunk;
    const int k0 = ch * KC;
    const int kc = min(KC, L.k4 - k0);
    // prefetch the following stage (own next chunk, or the next layer's first chunk)
    if (threadIdx.x == 0) {
      if (s + 1 < nstage) {
        const int blk1 = (s + 1) / nchunk, ch1 = (s + 1) - blk1 * nchunk;
        pipe.issue(L, blk1, ch1 * KC, min(KC, L.k4 - ch1 * KC), CB, pipe.stage + 1);
      } else if (next != nullptr) {
        const int cb1 = next->cout < CBMAX ? next->cout : CBMAX;
        pipe.issue(*next, 0, 0, next->k4 < KC ? next->k4 : KC, cb1, pipe.stage + 1);
      }
    }
    if (ch == 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    }
    mbarrier_wait(&pipe.bar[pipe.stage & 1], (pipe.stage >> 1) & 1);
    const float* wb = pipe.buf + (pipe.stage & 1) * (KC * CBMAX) + wn * TN;
    const float* xb = X + (size_t)row0 * ldx + k0;
#pragma unroll 2
    for (int kk = 0; kk < kc; kk += 4) {
      float4 xv[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) xv[i] = *reinterpret_cast<const float4*>(xb + (size_t)(32 * i) * ldx + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 wv[TN / 4];
#pragma unroll
        for (int j = 0; j < TN / 4; ++j) wv[j] = *reinterpret_cast<const float4*>(wb + (kk + q) * CB + 4 * j);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float x = q == 0 ? xv[i].x : (q == 1 ? xv[i].y : (q == 2 ? xv[i].z : xv[i].w));
#pragma unroll
          for (int j = 0; j < TN / 4; ++j) {
            acc[i][4 * j + 0] = fmaf(x, wv[j].x, acc[i][4 * j + 0]);
            acc[i][4 * j + 1] = fmaf(x, wv[j].y, acc[i][4 * j + 1]);
            acc[i][4 * j + 2] = fmaf(x, wv[j].z, acc[i][4 * j + 2]);
            acc[i][4 * j + 3] = fmaf(x, wv[j].w, acc[i][4 * j + 3]);
          }
        }
      }
    }
    if (ch == nchunk - 1) {  // epilogue of this column block: bias, activation, store
      const int col = blk * CB + wn * TN;
#pragma unroll
      for (int j = 0; j < TN / 4; ++j) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(L.b + col + 4 * j));
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          float4 o;
          o.x = acc[i][4 * j + 0] + bv.x; o.y = acc[i][4 * j + 1] + bv.y;
          o.z = acc[i][4 * j + 2] + bv.z; o.w = acc[i][4 * j + 3] + bv.w;
          if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          *reinterpret_cast<float4*>(Y + (size_t)(row0 + 32 * i) * ldy + ycol0 + col + 4 * j) = o;
        }
      }
    }
    ++pipe.stage;
    __syncthreads();  // ring slot free for the stage after next; Y visible after the last stage
  }
  if (next != nullptr) pipe.primed = true;
}

// Dispatch on the layer width (cout is 8, 16, 32 or a multiple of 64).
template <int R, bool RELU>
__device__ __forceinline__ void gemm_layer(const float* X, int ldx, float* Y, int ldy, int ycol0, const Layer& L,
                                           const Layer* next, WeightPipe& pipe) {
  if (L.cout >= 64) gemm_block<R, 64, RELU>(X, ldx, Y, ldy, ycol0, L, next, pipe);
  else if (L.cout == 32) gemm_block<R, 32, RELU>(X, ldx, Y, ldy, ycol0, L, next, pipe);
  else if (L.cout == 16) gemm_block<R, 16, RELU>(X, ldx, Y, ldy, ycol0, L, next, pipe);
  else if constexpr (R >= 128) gemm_block<R, 8, RELU>(X, ldx, Y, ldy, ycol0, L, next, pipe);
}

}  // namespace pwclo
