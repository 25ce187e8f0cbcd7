// Tensor-core (tcgen05 / TMEM) versions of the fused PWCLO-Net layer kernels.
//
// One CTA = one 128-row tile (rows = points x neighbours), persistent over tiles.  Warp roles:
//   warps 0-15 "row" warps: warp w serves TMEM lane quadrant w % 4 (tile rows 32*(w%4) .. +31, one row
//              per lane) and column slice w / 4 (every 4th group of 16 columns), so four warps share
//              a quadrant and hide each other's TMEM / global latencies.  They gather the layer
//              input straight into TMEM (a tf32 plane + packed bf16 value / residual planes), run every epilogue
//              (TMEM accumulator -> +bias -> ReLU -> operand planes of the next layer and/or an fp32
//              copy in shared memory) and the final pooling over the neighbour axis;
//   warp 16    TMA producer: streams the packed weight chunks (32 input channels each) of
//              every layer through a 3-slot shared-memory ring with cp.async.bulk + mbarriers;
//   warp 17    MMA issuer: one thread issues tcgen05.mma with the A operand in TMEM and the B operand
//              in shared memory: D += tf32(A)*tf32(W) [kind::tf32, K = 8] + bf16(A - tf32(A))*bf16(W)
//              + bf16(A)*bf16(W - tf32(W)) [kind::f16, K = 16]: fp32-class accuracy (~1e-6) from
//              8 MMAs per 32 inputs; accumulator in TMEM; tcgen05.commit releases ring slots and
//              signals the epilogue.
// Activations therefore never touch shared or global memory between the layers of a chain.
//
// TMEM column map (512 columns x 128 lanes): tf32 plane [0,192), bf16 value plane [192,288) and bf16
// residual plane [288,384) (two inputs per column), accumulator D [384,512).
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>

#include "tc_mma.cuh"

namespace pwclo {

constexpr int TC_ROWS = 128;
constexpr int TC_SLICES = 4;                      // row warps per TMEM lane quadrant (column slices)
constexpr int TC_ROW_THREADS = 128 * TC_SLICES;   // 16 row warps: warp w -> quadrant w % 4, slice w / 4
constexpr int TC_THREADS = TC_ROW_THREADS + 64;   // + TMA producer warp + MMA issuer warp
constexpr int TC_SLOTS = 3;
constexpr int TC_SLOT_FLOATS = 2 * 128 * 32;   // hi + lo chunk of a 128-wide layer (32 KB)
constexpr int TC_MAX_LAYERS = 6;
constexpr int TC_MAX_SIG = 16;                 // "operand columns ready" signals per tile
constexpr int TC_MAX_CHUNKS = 6;               // 32-input chunks per layer (K <= 192)

struct TcLayer {
  const float* w;   // packed per 32-input chunk [tf32(W) | bf16(W) | bf16(W - tf32(W))], n*256 bytes (tc_pack.pack_tc2)
  const float* b;   // [n]
  int seg_col[2];   // A-operand column segments in the planes (multiples of 8)
  int seg_n[2];
  int n;            // 64 or 128 outputs
  int out_col;      // >= 0: operand planes of the next layer start here; -1: not written
  int out_smem;     // 0: none, 1: fp32 copy to S0, 2: fp32 copy to S1
  // pipeline schedule (filled by tc_schedule on the host)
  int dsel;                               // accumulator region (0 / 1)
  int sig_base;                           // index of the signal of this layer's first 32 output columns
  unsigned char wait_sig[TC_MAX_CHUNKS];  // signals that must have completed before input chunk c is consumed
};

enum { TC_SA = 0, TC_PW = 1, TC_CV1 = 2, TC_CV2 = 3 };

struct TcArgs {
  // geometry / gather sources (meaning depends on the mode)
  const float* xyz_ref;   // SA: reference xyz [B,N,3]   CV1: xyz2 [B,N,3]     CV2: warped xyz [B,S,3]
  const float* xyz_ctr;   // SA: centres [B,S,3]         CV1/CV2: warped xyz [B,S,3]
  const float* f_ref;     // SA: feats [B,N,C]           CV1: f2 [B,N,C]       CV2: e1 [B,S,64]
  const float* f_ctr;     //                             CV1/CV2: f1 [B,S,C]
  const float* src[3];    // PW: concatenated sources
  int c_src[3];
  int nsrc;
  const int32_t* idx;     // [B,S,K]
  float* out;
  int B, N, S, K, C;      // PW: S = total rows, K = 1
  int ldS0, ldS1;
  int nlayers;
  int col_hb, col_lb, col_d[2];   // TMEM columns: packed bf16 value / residual planes, accumulator regions (tf32 plane at 0)
  int debug;              // bit 0: skip weight streaming (timing experiments only; results are garbage)
  long long* dbg;         // optional per-phase clock64 stamps of CTA 0, tile 1 (profiling aid)
  TcLayer l[TC_MAX_LAYERS];
};

// 32-bit shared-window addresses are computed ONCE per kernel: converting a generic pointer inside a
// loop costs an S2UR SR_CgaCtaId (hundreds of cycles) per use.
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void rows_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_ROW_THREADS) : "memory"); }

// write 16 consecutive fp32 values of this thread's row into the operand planes at input column `col`
#define plane_store16(lane_base, col, v) \
  tmem_st16_hybrid((lane_base) + (col), (lane_base) + a.col_hb + ((col) >> 1), (lane_base) + a.col_lb + ((col) >> 1), v)

__device__ __forceinline__ void load16(const float* __restrict__ p, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

// geometry group of the cost volumes: (p, q, q - p, |q - p|, 0 x6).  The raw coordinates are loaded by
// geo_load (no arithmetic on the loaded values: the loads stay in flight) and expanded by geo_finish.
__device__ __forceinline__ void geo_load(float (&g)[16], const float* __restrict__ pp, const float* __restrict__ qq) {
  g[0] = __ldg(pp); g[1] = __ldg(pp + 1); g[2] = __ldg(pp + 2); g[3] = __ldg(qq); g[4] = __ldg(qq + 1); g[5] = __ldg(qq + 2);
}
__device__ __forceinline__ void geo_finish(float (&g)[16]) {
  const float dx = __fsub_rn(g[3], g[0]), dy = __fsub_rn(g[4], g[1]), dz = __fsub_rn(g[5], g[2]);
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  g[6] = dx; g[7] = dy; g[8] = dz;
  g[9] = __fsqrt_rn(__fadd_rn(n2, 1e-20f));
#pragma unroll
  for (int i = 10; i < 16; ++i) g[i] = 0.f;
}

// pooling over the neighbour axis (max for set conv, softmax-weighted sum for the cost volumes) of a
// finished tile whose last activations sit in S0 (and S1), executed by all row threads: one
// (point, 4-channel group) item per thread and step, 128-bit shared-memory reads
template <int MODE>
__device__ __forceinline__ void tc_pool(const TcArgs& a, const float* S0, const float* S1, int tid, int tile, int b,
                                        int p0, int P) {
  const int co = a.l[a.nlayers - 1].n;
  const int co4 = co >> 2;
  if (MODE == TC_PW) {
    for (int e = tid; e < TC_ROWS * co4; e += TC_ROW_THREADS) {
      const int rr = e / co4, c4 = e - rr * co4;
      const int row = tile * TC_ROWS + rr;
      if (row < a.S)
        *reinterpret_cast<float4*>(a.out + (size_t)row * co + 4 * c4) =
            *reinterpret_cast<const float4*>(S0 + (size_t)rr * a.ldS0 + 4 * c4);
    }
  } else if (MODE == TC_SA) {
    const int st = a.ldS0 >> 2;   // row stride in float4
    for (int e = tid; e < P * co4; e += TC_ROW_THREADS) {
      const int p = e / co4, c4 = e - p * co4;
      if (p0 + p >= a.S) continue;
      const float4* y = reinterpret_cast<const float4*>(S0 + (size_t)(p * a.K) * a.ldS0) + c4;
      float4 m = y[0];
#pragma unroll 4
      for (int k = 1; k < a.K; ++k) {
        const float4 t = y[k * st];
        m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y); m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
      }
      *reinterpret_cast<float4*>(a.out + ((size_t)b * a.S + p0 + p) * co + 4 * c4) = m;
    }
  } else {
    const int st0 = a.ldS0 >> 2, st1 = a.ldS1 >> 2;
    for (int e = tid; e < P * 16; e += TC_ROW_THREADS) {
      const int p = e >> 4, c4 = e & 15;
      if (p0 + p >= a.S) continue;
      const float4* att = reinterpret_cast<const float4*>(S0 + (size_t)(p * a.K) * a.ldS0) + c4;
      const float4* val = reinterpret_cast<const float4*>(S1 + (size_t)(p * a.K) * a.ldS1) + c4;
      float4 m = att[0];
#pragma unroll 4
      for (int k = 1; k < a.K; ++k) {
        const float4 t = att[k * st0];
        m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y); m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
      }
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f), acc = z;
#pragma unroll 2
      for (int k = 0; k < a.K; ++k) {
        const float4 t = att[k * st0], v = val[k * st1];
        const float ex = __expf(t.x - m.x), ey = __expf(t.y - m.y);   // 2 ulp: far inside the 1e-4 feature budget
        const float ez = __expf(t.z - m.z), ew = __expf(t.w - m.w);
        z.x += ex; z.y += ey; z.z += ez; z.w += ew;
        acc.x = fmaf(ex, v.x, acc.x); acc.y = fmaf(ey, v.y, acc.y); acc.z = fmaf(ez, v.z, acc.z); acc.w = fmaf(ew, v.w, acc.w);
      }
      *reinterpret_cast<float4*>(a.out + ((size_t)b * a.S + p0 + p) * 64 + 4 * c4) =
          make_float4(__fdividef(acc.x, z.x), __fdividef(acc.y, z.y), __fdividef(acc.z, z.z), __fdividef(acc.w, z.w));
    }
  }
}

// Gather: the layer-0 operand columns of a tile row come in groups of 16; group `grp` of the row is
// fetched into registers (all groups of a thread before any TMEM store, so their global-memory
// latencies overlap).  Returns the operand-plane column of the group.
//   SA : [feat(C) | xyz_nbr - xyz_ctr (3) + zeros]
//   CV1: [f1(C) | f2 nbr(C)] at 0, geo(10) + 6 zeros at 128
//   CV2: geo(10) + 6 zeros at 0 (consumed by the first layer), f1(C) at 64, e1 nbr(64) at 64 + C
//   PW : the concatenated sources
template <int MODE>
__device__ __forceinline__ int tc_groups(const TcArgs& a) {
  if (MODE == TC_SA) return (a.C >> 4) + 1;
  if (MODE == TC_CV1) return (a.C >> 3) + 1;
  if (MODE == TC_CV2) return 1 + (a.C >> 4) + 4;
  int c = 0;
  for (int s = 0; s < a.nsrc; ++s) c += a.c_src[s];
  return c >> 4;
}
template <int MODE>
__device__ __forceinline__ int tc_fetch(const TcArgs& a, int grp, int b, int gp, int n, float (&v)[16]) {
  if (MODE == TC_PW) {      // gp = row
    int c = 16 * grp, s = 0;
    while (c >= a.c_src[s]) { c -= a.c_src[s]; ++s; }
    load16(a.src[s] + (size_t)gp * a.c_src[s] + c, v);
    return 16 * grp;
  } else if (MODE == TC_SA) {
    const int c = 16 * grp;
    if (c < a.C) {
      load16(a.f_ref + ((size_t)b * a.N + n) * a.C + c, v);
    } else {
      geo_load(v, a.xyz_ctr + ((size_t)b * a.S + gp) * 3, a.xyz_ref + ((size_t)b * a.N + n) * 3);   // finished by tc_finish
    }
    return c;
  } else if (MODE == TC_CV1) {
    const int c = 16 * grp;
    if (c < 2 * a.C) {
      load16(c < a.C ? a.f_ctr + ((size_t)b * a.S + gp) * a.C + c : a.f_ref + ((size_t)b * a.N + n) * a.C + (c - a.C), v);
      return c;
    }
    geo_load(v, a.xyz_ctr + ((size_t)b * a.S + gp) * 3, a.xyz_ref + ((size_t)b * a.N + n) * 3);
    return 128;
  } else {   // TC_CV2
    if (grp == 0) {
      geo_load(v, a.xyz_ctr + ((size_t)b * a.S + gp) * 3, a.xyz_ref + ((size_t)b * a.S + n) * 3);
      return 0;
    }
    const int c = 16 * (grp - 1);
    if (c < a.C) load16(a.f_ctr + ((size_t)b * a.S + gp) * a.C + c, v);
    else load16(a.f_ref + ((size_t)b * a.S + n) * 64 + (c - a.C), v);
    return 64 + c;
  }
}

// arithmetic on the raw coordinates of the geometry / xyz group (run right before the TMEM store)
template <int MODE>
__device__ __forceinline__ void tc_finish(const TcArgs& a, int col, float (&v)[16]) {
  if (MODE == TC_SA && col == a.C) {
    const float dx = __fsub_rn(v[3], v[0]), dy = __fsub_rn(v[4], v[1]), dz = __fsub_rn(v[5], v[2]);   // neighbour - centre
    v[0] = dx; v[1] = dy; v[2] = dz;
#pragma unroll
    for (int i = 3; i < 16; ++i) v[i] = 0.f;
  }
  if ((MODE == TC_CV1 && col == 128) || (MODE == TC_CV2 && col == 0)) geo_finish(v);
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_mlp_kernel(const TcArgs a, int ntiles, int tiles_per_cloud) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);                       // [TC_SLOTS][TC_SLOT_FLOATS]
  float* S0 = ring + TC_SLOTS * TC_SLOT_FLOATS;                            // [128][ldS0]
  float* S1 = S0 + (size_t)TC_ROWS * a.ldS0;                               // [128][ldS1]
  __shared__ __align__(8) uint64_t full_bar[TC_SLOTS], empty_bar[TC_SLOTS], d_ready[2], sig_bar[TC_MAX_SIG];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == TC_ROW_THREADS / 32 + 1) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < TC_SLOTS; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&empty_bar[s], 1); }
    mbarrier_init(&d_ready[0], 1);
    mbarrier_init(&d_ready[1], 1);
    for (int i = 0; i < TC_MAX_SIG; ++i) mbarrier_init(&sig_bar[i], TC_ROW_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t full_a = smem_addr(&full_bar[0]), empty_a = smem_addr(&empty_bar[0]);   // + 8 * slot
  const uint32_t dready_a = smem_addr(&d_ready[0]), sig_a = smem_addr(&sig_bar[0]);       // + 8 * region / signal
  const uint32_t ring_addr = smem_addr(ring);

  if (warp == TC_ROW_THREADS / 32) {
    // ============================== TMA producer ==============================
    if (lane == 0 && !(a.debug & 1)) {
      uint32_t use = 0;   // global chunk counter
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int L = 0; L < a.nlayers; ++L) {
          const TcLayer& ly = a.l[L];
          const int ktot = ly.seg_n[0] + ly.seg_n[1];
          const int nchunk = (ktot + 31) / 32;
          const uint32_t bytes = (uint32_t)(2 * ly.n * 32 * sizeof(float));
          for (int ch = 0; ch < nchunk; ++ch, ++use) {
            const int s = use % TC_SLOTS;
            const uint32_t n = use / TC_SLOTS;          // how many times this slot was filled before
            if (n > 0) mbar_wait_a(empty_a + 8 * s, (n - 1) & 1);
            mbar_expect_tx_a(full_a + 8 * s, bytes);
            tma_bulk_g2s_a(ring_addr + s * (uint32_t)(TC_SLOT_FLOATS * sizeof(float)), ly.w + (size_t)ch * 2 * ly.n * 32, bytes,
                           full_a + 8 * s);
          }
        }
      }
    }
  } else if (warp == TC_ROW_THREADS / 32 + 1) {
    // ============================== MMA issuer ==============================
    // The whole warp walks the (warp-uniform) schedule so that descriptors and TMEM addresses stay in
    // uniform registers; only the tcgen05 instructions themselves are issued by the elected lane.
    // A layer's chunk c is issued as soon as the signals it depends on (wait_sig) have completed, i.e.
    // while the row warps are still producing the later input columns from the previous accumulator.
    const bool leader = lane == 0;
    uint32_t use = 0, tile_par = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tile_par ^= 1) {
      int seen = 0;   // signals of this tile already waited for
      for (int L = 0; L < a.nlayers; ++L) {
        const TcLayer& ly = a.l[L];
        const int n = ly.n;
        const int col0 = ly.seg_col[0], col1 = ly.seg_col[1];
        const int h0 = ly.seg_n[0] >> 4;                      // 16-input halves in segment 0
        const int halves = h0 + (ly.seg_n[1] >> 4);
        const uint32_t idesc_t = tc_idesc_tf32(128, n), idesc_b = tc_idesc_bf16(128, n);
        const uint32_t hb_off = (uint32_t)n * 128u, lb_off = (uint32_t)n * 192u;   // bf16(W), bf16(W - tf32(W)) blocks
        const uint32_t d_tmem = tb + (uint32_t)a.col_d[ly.dsel];
        const bool mstamp = a.dbg != nullptr && blockIdx.x == 0 && tile == (int)gridDim.x && leader;
        uint32_t acc = 0;
        for (int hh = 0, c = 0; hh < halves; hh += 2, ++c, ++use) {
          const int need = ly.wait_sig[c];
          while (seen < need) {
            mbar_wait_a(sig_a + 8 * seen, tile_par);
            ++seen;
          }
          if (mstamp && c == 0) a.dbg[32 + 2 * L] = clock64();
          const uint32_t s = use % TC_SLOTS;
          if (!(a.debug & 1)) mbar_wait_a(full_a + 8 * s, (use / TC_SLOTS) & 1);
          tc_fence_after();
          const uint32_t slot = ring_addr + s * (uint32_t)(TC_SLOT_FLOATS * sizeof(float));
          const uint64_t dt = tc_desc_at(slot, TC_DESC_TF32);
          const uint64_t dhb = tc_desc_at(slot + hb_off, TC_DESC_BF16), dlb = tc_desc_at(slot + lb_off, TC_DESC_BF16);
          // full chunk inside one column segment (the common case): 8 MMAs from one asm block
          const bool seg0 = hh + 2 <= h0, seg1 = hh >= h0;
          if (a.debug & 2) {
            // timing experiment: no MMAs at all
          } else if (hh + 2 <= halves && (seg0 || seg1)) {
            const uint32_t col = (uint32_t)(seg0 ? col0 + 16 * hh : col1 + 16 * (hh - h0));
            tc_mma_hybrid_chunk_warp(d_tmem, tb + col, tb + a.col_hb + (col >> 1), tb + a.col_lb + (col >> 1), dt, dhb, dlb,
                                     idesc_t, idesc_b, acc);
            acc = 1;
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (hh + j < halves) {
                const int h = hh + j;
                const uint32_t col = (uint32_t)(h < h0 ? col0 + 16 * h : col1 + 16 * (h - h0));
                tc_mma_hybrid_half_warp(d_tmem, tb + col, tb + a.col_hb + (col >> 1), tb + a.col_lb + (col >> 1),
                                        dt + (uint64_t)(32 * j), dhb + (uint64_t)(16 * j), dlb + (uint64_t)(16 * j), idesc_t,
                                        idesc_b, acc);
                acc = 1;
              }
            }
          }
          tc_commit_warp(empty_a + 8 * s);                    // slot reusable once these MMAs have read it
        }
        tc_commit_warp(dready_a + 8 * ly.dsel);               // accumulator complete
        if (mstamp) a.dbg[33 + 2 * L] = clock64();
      }
    }
  } else {
    // ============================== row warps ==============================
    const int quad = warp & 3, slice = warp >> 2;
    const int r = quad * 32 + lane;                // tile row == TMEM lane
    const uint32_t lane_base = tb + ((uint32_t)(quad * 32) << 16);
    uint32_t d_cnt0 = 0, d_cnt1 = 0;
    bool have_prev = false;
    int prev_tile = 0, prev_b = 0, prev_p0 = 0, prev_P = TC_ROWS;
    const int ngroups = tc_groups<MODE>(a);
    const int P_tile = MODE == TC_PW ? TC_ROWS : TC_ROWS / a.K;          // points per tile
    const int p_row = MODE == TC_PW ? 0 : r / a.K, k_row = MODE == TC_PW ? 0 : r - p_row * a.K;
    // Software pipeline of the gather: the global loads of tile t+1 (fv / fcol) are issued before the
    // last epilogue of tile t and stored into TMEM after it; the neighbour index of tile t+2 (n_next) is
    // requested at the same time, so no global-memory latency sits between two tiles.
    auto row_index = [&](int t) -> int {          // neighbour index of this thread's row in tile t
      if (MODE == TC_PW || t >= ntiles) return 0;
      const int tb_ = t / tiles_per_cloud;
      const int tp0 = (t - tb_ * tiles_per_cloud) * P_tile;
      const bool tvalid = p_row < P_tile && tp0 + p_row < a.S;
      const int tgp = min(tp0 + min(p_row, P_tile - 1), a.S - 1);
      return tvalid ? a.idx[((size_t)tb_ * a.S + tgp) * a.K + k_row] : 0;
    };
    float fv[3][16];
    int fcol[3];
    auto fetch_tile = [&](int t, int n) {
      int tb_ = 0, tgp;
      if (MODE == TC_PW) {
        tgp = min(t * TC_ROWS + r, a.S - 1);
      } else {
        tb_ = t / tiles_per_cloud;
        const int tp0 = (t - tb_ * tiles_per_cloud) * P_tile;
        tgp = min(tp0 + min(p_row, P_tile - 1), a.S - 1);
      }
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int grp = slice + TC_SLICES * g;
        fcol[g] = grp < ngroups ? tc_fetch<MODE>(a, grp, tb_, tgp, n, fv[g]) : -1;
      }
    };
    int n_next = 0;
    if ((int)blockIdx.x < ntiles) {
      fetch_tile(blockIdx.x, row_index(blockIdx.x));
      n_next = row_index(blockIdx.x + gridDim.x);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // ---------------- the gathered layer-0 input goes into the planes ----------------
      // (16-column groups are dealt round-robin to the 4 slices of a quadrant: group slice + 4 g)
      int b = 0, p0 = 0, P = TC_ROWS;
      if (a.dbg != nullptr && blockIdx.x == 0 && tile == (int)gridDim.x && tid == 0) a.dbg[22] = clock64();
      if (MODE != TC_PW) {
        b = tile / tiles_per_cloud;
        P = P_tile;
        p0 = (tile - b * tiles_per_cloud) * P;
      }
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        if (fcol[g] >= 0) {
          tc_finish<MODE>(a, fcol[g], fv[g]);
          plane_store16(lane_base, fcol[g], fv[g]);
          if (MODE == TC_CV2 && fcol[g] >= 64 + a.C) {      // fp32 copy of the e1 neighbour features for the pooling
            float* s1 = S1 + (size_t)r * a.ldS1 + (fcol[g] - 64 - a.C);
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(s1 + i) = make_float4(fv[g][i], fv[g][i + 1], fv[g][i + 2], fv[g][i + 3]);
          }
        }
      }
      const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && tile == (int)gridDim.x && tid == 0;
      if (stamp) a.dbg[0] = clock64();
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive_a(sig_a);            // signal 0: layer-0 operands complete
      if (stamp) a.dbg[1] = clock64();
      if (have_prev) {      // previous tile's pooling, overlapped with this tile's first GEMM
        tc_pool<MODE>(a, S0, S1, tid, prev_tile, prev_b, prev_p0, prev_P);
        have_prev = false;
        rows_sync();        // S0 / S1 free before this tile's epilogues write them
      }
      if (stamp) a.dbg[23] = clock64();

      // ---------------- layer epilogues ----------------
      for (int L = 0; L < a.nlayers; ++L) {
        const TcLayer& ly = a.l[L];
        if (L == a.nlayers - 1 && tile + (int)gridDim.x < ntiles) {
          fetch_tile(tile + gridDim.x, n_next);
          n_next = row_index(tile + 2 * gridDim.x);
        }
        if (ly.dsel == 0) { mbar_wait_a(dready_a, d_cnt0 & 1); ++d_cnt0; }
        else { mbar_wait_a(dready_a + 8, d_cnt1 & 1); ++d_cnt1; }
        tc_fence_after();
        if (stamp) a.dbg[2 + 2 * L] = clock64();
        const uint32_t d_lane = lane_base + (uint32_t)a.col_d[ly.dsel];
        float* sdst = ly.out_smem == 1 ? S0 + (size_t)r * a.ldS0 : (ly.out_smem == 2 ? S1 + (size_t)r * a.ldS1 : nullptr);
        const bool planes = ly.out_col >= 0;
        const int nch = ly.n >> 5;
        for (int j = 0; j < nch; ++j) {
          // all 16 row warps work on every 32-column chunk (8 columns per slice), so the MMA warp can start
          // the next layer on chunk 0 after 1/nch of the epilogue
          const int c = 32 * j + 8 * slice;
          float v[8];
          tmem_ld8(d_lane + c, v);
#pragma unroll
          for (int i = 0; i < 8; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(ly.b + c + i));
            v[i] = fmaxf(v[i] + bv.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + bv.y, 0.f);
            v[i + 2] = fmaxf(v[i + 2] + bv.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + bv.w, 0.f);
          }
          if (planes) {
            const int oc = ly.out_col + c;
            tmem_st8_hybrid(lane_base + oc, lane_base + a.col_hb + (oc >> 1), lane_base + a.col_lb + (oc >> 1), v);
          }
          if (sdst != nullptr) {
            *reinterpret_cast<float4*>(sdst + c) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(sdst + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (planes) {
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_a(sig_a + 8 * (ly.sig_base + j));
          }
        }
        if (stamp) a.dbg[3 + 2 * L] = clock64();
      }
      tc_fence_before();
      rows_sync();   // S0 / S1 complete
      if (stamp) a.dbg[20] = clock64();

      // pooling of this tile is deferred: it runs after the NEXT tile's gather has been handed to the MMA
      // warp, so it overlaps with that tile's first GEMM (cost volume 2 pools first: its gather writes S1)
      if (MODE == TC_CV2) {
        tc_pool<MODE>(a, S0, S1, tid, tile, b, p0, P);
        rows_sync();
      } else {
        have_prev = true; prev_tile = tile; prev_b = b; prev_p0 = p0; prev_P = P;
      }
      if (stamp) a.dbg[21] = clock64();
    }
    if (have_prev) tc_pool<MODE>(a, S0, S1, tid, prev_tile, prev_b, prev_p0, prev_P);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_ROW_THREADS / 32 + 1) tmem_dealloc(tb, 512);
}

static bool tc_layer_set(TcLayer& L, const pwclo_layer_t& src, int col0, int n0, int col1, int n1, int out_col, int out_smem) {
  if (!src.w || !src.b || (src.cout != 64 && src.cout != 128) || (uintptr_t)src.w % 16 != 0 || (uintptr_t)src.b % 16 != 0)
    return false;
  if (n0 % 16 != 0 || n1 % 16 != 0 || col0 % 16 != 0 || col1 % 16 != 0 || n0 <= 0) return false;
  L.w = src.w; L.b = src.b; L.n = src.cout;
  L.seg_col[0] = col0; L.seg_n[0] = n0; L.seg_col[1] = col1; L.seg_n[1] = n1;
  L.out_col = out_col; L.out_smem = out_smem;
  return true;
}

// Host-side pipeline schedule: TMEM column map, accumulator regions, and for every input chunk of every
// layer the number of row-warp signals that must have completed before the MMA warp may consume it.
// Signal 0 = gather done; then every layer that writes operand planes emits one signal per 32 output
// columns, in column order.  `gather_cols` = operand columns written by the gather.
static bool tc_schedule(TcArgs& a, int gather_cols) {
  int maxcol = gather_cols;
  for (int L = 0; L < a.nlayers; ++L) {
    const TcLayer& ly = a.l[L];
    for (int s = 0; s < 2; ++s)
      if (ly.seg_n[s] > 0) maxcol = max(maxcol, ly.seg_col[s] + ly.seg_n[s]);
    if (ly.out_col >= 0) maxcol = max(maxcol, ly.out_col + ly.n);
    if ((ly.seg_n[0] + ly.seg_n[1] + 31) / 32 > TC_MAX_CHUNKS) return false;
  }
  if (maxcol > 192) return false;
  const int kt = maxcol <= 160 ? 160 : 192;          // plane width; 2*kt + 128 (+ 64) accumulator columns
  a.col_hb = kt; a.col_lb = kt + kt / 2;
  a.col_d[0] = 2 * kt;
  const bool two = kt <= 160 && !getenv("PWCLO_TC_NO_OVERLAP");   // second (64-column) region available
  a.col_d[1] = two ? 2 * kt + 128 : 2 * kt;
  // accumulator regions: alternate wherever possible (128-wide layers only fit region 0)
  int best = -1, best_mask = 0;
  for (int mask = 0; mask < (1 << a.nlayers); ++mask) {
    int score = 0;
    bool ok = true;
    for (int L = 0; L < a.nlayers && ok; ++L) {
      const int d = (mask >> L) & 1;
      if (d == 1 && (!two || a.l[L].n > 64)) ok = false;
      if (L > 0 && d != ((mask >> (L - 1)) & 1)) ++score;
    }
    if (ok && score > best) { best = score; best_mask = mask; }
  }
  int writer[192];                                   // signal count that covers the last write of each column
  for (int c = 0; c < 192; ++c) writer[c] = 1;       // gather (also covers columns nobody wrote: zero weights there)
  int nsig = 1;
  int done_after[TC_MAX_LAYERS + 1];                 // signals emitted once layer L's epilogue has finished
  for (int L = 0; L < a.nlayers; ++L) {
    TcLayer& ly = a.l[L];
    ly.dsel = (best_mask >> L) & 1;
    // the accumulator region must be free: the epilogue of the previous user has finished
    int floor_sig = 1;
    for (int Lp = L - 1; Lp >= 0; --Lp)
      if (a.l[Lp].dsel == ly.dsel || a.col_d[0] == a.col_d[1]) { floor_sig = done_after[Lp]; break; }
    // ... and so has the epilogue of layer L-2 (keeps the accumulator-ready barrier phases from lapping)
    if (L >= 2) floor_sig = max(floor_sig, done_after[L - 2]);
    const int ktot = ly.seg_n[0] + ly.seg_n[1];
    for (int c = 0; c * 32 < ktot; ++c) {
      int need = floor_sig;
      for (int k = c * 32; k < min(ktot, c * 32 + 32); ++k) {
        const int col = k < ly.seg_n[0] ? ly.seg_col[0] + k : ly.seg_col[1] + (k - ly.seg_n[0]);
        need = max(need, writer[col]);
      }
      if (c > 0) need = max(need, (int)ly.wait_sig[c - 1]);
      ly.wait_sig[c] = (unsigned char)need;
    }
    ly.sig_base = nsig;
    if (ly.out_col >= 0) {
      for (int j = 0; j < ly.n / 32; ++j) {
        for (int c = 0; c < 32; ++c) writer[ly.out_col + 32 * j + c] = nsig + 1;
        ++nsig;
      }
    }
    done_after[L] = nsig;
    if (nsig > TC_MAX_SIG) return false;
  }
  // a layer without plane output must be the last one (its epilogue end is not signalled)
  for (int L = 0; L + 1 < a.nlayers; ++L)
    if (a.l[L].out_col < 0) return false;
  return true;
}

template <int MODE>
static int tc_launch(TcArgs& a, int ntiles, int tiles_per_cloud, int gather_cols, cudaStream_t st) {
  if (a.ldS1 == 0) a.ldS1 = 4;
  if (const char* d = getenv("PWCLO_TC_DEBUG")) a.debug = atoi(d);
  if (!tc_schedule(a, gather_cols)) return PWCLO_EUNSUPPORTED;
  static long long* dbg_buf = nullptr;
  if (getenv("PWCLO_TC_STAMPS")) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 64 * sizeof(long long));
    a.dbg = dbg_buf;
  }
  const size_t smem = (size_t)TC_SLOTS * TC_SLOT_FLOATS * 4 + (size_t)TC_ROWS * (a.ldS0 + a.ldS1) * 4 + 1024;
  if (smem > 227 * 1024) return PWCLO_EUNSUPPORTED;
  auto kern = tc_mlp_kernel<MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = min(ntiles, kNumSM);
  kern<<<grid, TC_THREADS, smem, st>>>(a, ntiles, tiles_per_cloud);
  if (a.dbg) {
    long long h[64];
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc stamps mode %d] gather_start %lld arrive +%lld pool_end +%lld |", MODE, h[22] - h[0], h[1] - h[0], h[23] - h[0]);
    for (int L = 0; L < a.nlayers; ++L)
      fprintf(stderr, " L%d: mma_start +%lld issued +%lld dready_seen +%lld epi_end +%lld |", L, h[32 + 2 * L] - h[0],
              h[33 + 2 * L] - h[0], h[2 + 2 * L] - h[0], h[3 + 2 * L] - h[0]);
    fprintf(stderr, " pooled_sync +%lld tile_end +%lld\n", h[20] - h[0], h[21] - h[0]);
  }
  return launch_status();
}

}  // namespace pwclo

using namespace pwclo;

// The *_tc entry points take layers packed by tc_pack.pack_tc (see include/pwclo_b200.h).

PWCLO_API int pwclo_set_conv_tc(const float* xyz, const float* feats, const float* new_xyz, const int32_t* idx, int B,
                                int N, int S, int K, int C, const pwclo_layer_t* layers, int nlayers, float* out,
                                void* stream) {
  if (!xyz || !feats || !new_xyz || !idx || !layers || !out || B <= 0 || N <= 0 || S <= 0 || K <= 0) return PWCLO_EINVAL;
  if (nlayers < 2 || nlayers > 3 || K > 128 || C % 16 != 0 || C + 16 > 192 || (uintptr_t)feats % 16 != 0) return PWCLO_EUNSUPPORTED;
  TcArgs a = {};
  a.xyz_ref = xyz; a.xyz_ctr = new_xyz; a.f_ref = feats; a.idx = idx; a.out = out;
  a.B = B; a.N = N; a.S = S; a.K = K; a.C = C; a.nlayers = nlayers;
  int kin = C + 16;
  for (int i = 0; i < nlayers; ++i) {
    const bool last = i == nlayers - 1;
    if (!tc_layer_set(a.l[i], layers[i], 0, kin, 0, 0, last ? -1 : 0, last ? 1 : 0)) return PWCLO_EUNSUPPORTED;
    kin = layers[i].cout;
  }
  a.ldS0 = ld_for(layers[nlayers - 1].cout);
  a.ldS1 = 0;
  const int P = TC_ROWS / K, tpc = ceil_div(S, P);
  return tc_launch<TC_SA>(a, B * tpc, tpc, C + 16, (cudaStream_t)stream);
}

PWCLO_API int pwclo_pointwise_mlp_tc(const float* const* src, const int* channels, int nsrc, int rows,
                                     const pwclo_layer_t* layers, int nlayers, float* out, void* stream) {
  if (!src || !channels || !layers || !out || nsrc < 1 || nsrc > 3 || rows <= 0) return PWCLO_EINVAL;
  if (nlayers < 1 || nlayers > 2) return PWCLO_EUNSUPPORTED;
  TcArgs a = {};
  int cin = 0;
  for (int s = 0; s < nsrc; ++s) {
    if (!src[s] || channels[s] % 16 != 0 || (uintptr_t)src[s] % 16 != 0) return PWCLO_EUNSUPPORTED;
    a.src[s] = src[s]; a.c_src[s] = channels[s]; cin += channels[s];
  }
  if (cin > 192) return PWCLO_EUNSUPPORTED;
  a.nsrc = nsrc; a.S = rows; a.K = 1; a.out = out; a.nlayers = nlayers;
  int kin = cin;
  for (int i = 0; i < nlayers; ++i) {
    const bool last = i == nlayers - 1;
    if (!tc_layer_set(a.l[i], layers[i], 0, kin, 0, 0, last ? -1 : 0, last ? 1 : 0)) return PWCLO_EUNSUPPORTED;
    kin = layers[i].cout;
  }
  a.ldS0 = ld_for(layers[nlayers - 1].cout);
  const int ntiles = ceil_div(rows, TC_ROWS);
  return tc_launch<TC_PW>(a, ntiles, 1, cin, (cudaStream_t)stream);
}

PWCLO_API int pwclo_cost_volume_1_tc(const float* wxyz, const float* f1, const float* xyz2, const float* f2,
                                     const int32_t* idx, int B, int S, int N, int K, int C, const pwclo_layer_t* mlp1,
                                     const pwclo_layer_t* enc, const pwclo_layer_t* mlp2, float* out, void* stream) {
  if (!wxyz || !f1 || !xyz2 || !f2 || !idx || !mlp1 || !enc || !mlp2 || !out || B <= 0 || S <= 0 || N <= 0 || K <= 0)
    return PWCLO_EINVAL;
  if (C % 16 != 0 || 2 * C > 128 || K > 128 || ((uintptr_t)f1 | (uintptr_t)f2) % 16 != 0) return PWCLO_EUNSUPPORTED;
  if (mlp1[0].cout != 128 || mlp1[1].cout != 64 || mlp1[2].cout != 64 || enc->cout != 64 || mlp2[0].cout != 128 || mlp2[1].cout != 64)
    return PWCLO_EUNSUPPORTED;
  TcArgs a = {};
  a.xyz_ref = xyz2; a.xyz_ctr = wxyz; a.f_ref = f2; a.f_ctr = f1; a.idx = idx; a.out = out;
  a.B = B; a.N = N; a.S = S; a.K = K; a.C = C; a.nlayers = 6;
  bool ok = tc_layer_set(a.l[0], mlp1[0], 0, 2 * C, 128, 16, 0, 0)      // [f1 | f2 | geo] -> h1 at 0
         && tc_layer_set(a.l[1], mlp1[1], 0, 128, 0, 0, 0, 0)           // h2 at 0
         && tc_layer_set(a.l[2], mlp1[2], 0, 64, 0, 0, 64, 2)           // h3 at 64 (+ fp32 copy S1)
         && tc_layer_set(a.l[3], *enc, 128, 16, 0, 0, 0, 0)             // enc at 0
         && tc_layer_set(a.l[4], mlp2[0], 0, 128, 0, 0, 0, 0)           // a1 at 0
         && tc_layer_set(a.l[5], mlp2[1], 0, 128, 0, 0, -1, 1);         // a2 -> S0
  if (!ok) return PWCLO_EUNSUPPORTED;
  a.ldS0 = ld_for(64); a.ldS1 = ld_for(64);
  const int P = TC_ROWS / K, tpc = ceil_div(S, P);
  return tc_launch<TC_CV1>(a, B * tpc, tpc, 144, (cudaStream_t)stream);
}

PWCLO_API int pwclo_cost_volume_2_tc(const float* wxyz, const float* f1, const float* e1, const int32_t* idx, int B,
                                     int S, int K, int C, const pwclo_layer_t* enc, const pwclo_layer_t* mlp3, float* out,
                                     void* stream) {
  if (!wxyz || !f1 || !e1 || !idx || !enc || !mlp3 || !out || B <= 0 || S <= 0 || K <= 0) return PWCLO_EINVAL;
  if (C % 16 != 0 || C > 64 || K > 128 || ((uintptr_t)f1 | (uintptr_t)e1) % 16 != 0) return PWCLO_EUNSUPPORTED;
  if (enc->cout != 64 || mlp3[0].cout != 128 || mlp3[1].cout != 64) return PWCLO_EUNSUPPORTED;
  TcArgs a = {};
  a.xyz_ref = wxyz; a.xyz_ctr = wxyz; a.f_ref = e1; a.f_ctr = f1; a.idx = idx; a.out = out;
  a.B = B; a.N = S; a.S = S; a.K = K; a.C = C; a.nlayers = 3;
  bool ok = tc_layer_set(a.l[0], *enc, 0, 16, 0, 0, 0, 0)               // geo -> enc2 at 0
         && tc_layer_set(a.l[1], mlp3[0], 0, 128 + C, 0, 0, 0, 0)       // [enc2 | f1 | e1 nbr] -> a1 at 0
         && tc_layer_set(a.l[2], mlp3[1], 0, 128, 0, 0, -1, 1);         // a2 -> S0
  if (!ok) return PWCLO_EUNSUPPORTED;
  a.ldS0 = ld_for(64); a.ldS1 = ld_for(64);
  const int P = TC_ROWS / K, tpc = ceil_div(S, P);
  return tc_launch<TC_CV2>(a, B * tpc, tpc, 128 + C, (cudaStream_t)stream);
}
