// Tensor-core (tcgen05 / TMEM) versions of the fused PWCLO-Net layer kernels.
//
// One CTA = one 128-row tile (rows = points x neighbours), persistent over tiles.  Warp roles:
//   warps 0-15 "row" warps: warp w serves TMEM lane quadrant w % 4 (tile rows 32*(w%4) .. +31, one row
//              per lane) and column slice w / 4 (every 4th group of 16 columns), so four warps share
//              a quadrant and hide each other's TMEM / global latencies.  They gather the layer
//              input straight into TMEM (split into tf32-exact hi / lo planes), run every epilogue
//              (TMEM accumulator -> +bias -> ReLU -> hi/lo planes of the next layer and/or an fp32
//              copy in shared memory) and the final pooling over the neighbour axis;
//   warp 16    TMA producer: streams the packed hi/lo weight chunks (32 input channels each) of
//              every layer through a 3-slot shared-memory ring with cp.async.bulk + mbarriers;
//   warp 17    MMA issuer: one thread issues tcgen05.mma.kind::tf32 with the A operand in TMEM and
//              the B operand in shared memory: D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (3xTF32,
//              fp32-class accuracy), accumulator in TMEM; tcgen05.commit releases ring slots and
//              signals the epilogue.
// Activations therefore never touch shared or global memory between the layers of a chain.
//
// TMEM column map (512 columns x 128 lanes): HI plane [0,192), LO plane [192,384), D [384,512).
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
constexpr int COL_HI = 0, COL_LO = 192, COL_D = 384;
constexpr int TC_MAX_LAYERS = 6;

struct TcLayer {
  const float* w;   // packed [nchunk][2][n/8][8][8][4]  (tc_pack.py)
  const float* b;   // [n]
  int seg_col[2];   // A-operand column segments in the planes (multiples of 8)
  int seg_n[2];
  int n;            // 64 or 128 outputs
  int out_col;      // >= 0: hi/lo planes of the next layer start here; -1: not written
  int out_smem;     // 0: none, 1: fp32 copy to S0, 2: fp32 copy to S1
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

// write 16 consecutive fp32 values of this thread's row into the hi/lo planes at column `col`
__device__ __forceinline__ void plane_store16(uint32_t lane_base, int col, const float (&v)[16]) {
  tmem_st16_split(lane_base + COL_HI + col, lane_base + COL_LO + col, v);
}

__device__ __forceinline__ void load16(const float* __restrict__ p, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

__device__ __forceinline__ void geo16(float (&g)[16], const float* __restrict__ pp, const float* __restrict__ qq) {
  const float px = pp[0], py = pp[1], pz = pp[2], qx = qq[0], qy = qq[1], qz = qq[2];
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  g[0] = px; g[1] = py; g[2] = pz; g[3] = qx; g[4] = qy; g[5] = qz; g[6] = dx; g[7] = dy; g[8] = dz;
  g[9] = __fsqrt_rn(__fadd_rn(n2, 1e-20f));
#pragma unroll
  for (int i = 10; i < 16; ++i) g[i] = 0.f;
}

// pooling over the neighbour axis (max for set conv, softmax-weighted sum for the cost volumes) of a
// finished tile whose last activations sit in S0 (and S1), executed by all row threads
template <int MODE>
__device__ __forceinline__ void tc_pool(const TcArgs& a, const float* S0, const float* S1, int tid, int tile, int b,
                                        int p0, int P) {
  const int co = a.l[a.nlayers - 1].n;
  if (MODE == TC_PW) {
    const int co4 = co >> 2;
    for (int e = tid; e < TC_ROWS * co4; e += TC_ROW_THREADS) {
      const int rr = e / co4, c4 = e - rr * co4;
      const int row = tile * TC_ROWS + rr;
      if (row < a.S)
        *reinterpret_cast<float4*>(a.out + (size_t)row * co + 4 * c4) =
            *reinterpret_cast<const float4*>(S0 + (size_t)rr * a.ldS0 + 4 * c4);
    }
  } else if (MODE == TC_SA) {
    for (int e = tid; e < P * co; e += TC_ROW_THREADS) {
      const int p = e / co, c = e - p * co;
      if (p0 + p >= a.S) continue;
      const float* y = S0 + (size_t)(p * a.K) * a.ldS0 + c;
      float m = y[0];
      for (int k = 1; k < a.K; ++k) m = fmaxf(m, y[(size_t)k * a.ldS0]);
      a.out[((size_t)b * a.S + p0 + p) * co + c] = m;
    }
  } else {
    // two (point, channel) items per thread in flight: the per-item chain (LDS -> exp -> fma) is latency bound
    const int items = P * 64;
    for (int e0 = tid; e0 < items; e0 += 2 * TC_ROW_THREADS) {
      const int e1 = e0 + TC_ROW_THREADS;
      const bool has1 = e1 < items;
      const int pA = e0 >> 6, cA = e0 & 63;
      const int pB = has1 ? e1 >> 6 : pA, cB = has1 ? e1 & 63 : cA;
      const float* attA = S0 + (size_t)(pA * a.K) * a.ldS0 + cA;
      const float* valA = S1 + (size_t)(pA * a.K) * a.ldS1 + cA;
      const float* attB = S0 + (size_t)(pB * a.K) * a.ldS0 + cB;
      const float* valB = S1 + (size_t)(pB * a.K) * a.ldS1 + cB;
      float mA = attA[0], mB = attB[0];
#pragma unroll 4
      for (int k = 1; k < a.K; ++k) {
        mA = fmaxf(mA, attA[(size_t)k * a.ldS0]);
        mB = fmaxf(mB, attB[(size_t)k * a.ldS0]);
      }
      float zA = 0.f, sA = 0.f, zB = 0.f, sB = 0.f;
#pragma unroll 4
      for (int k = 0; k < a.K; ++k) {
        const float exA = __expf(attA[(size_t)k * a.ldS0] - mA);   // 2 ulp: far inside the 1e-4 feature budget
        const float exB = __expf(attB[(size_t)k * a.ldS0] - mB);
        zA += exA; zB += exB;
        sA = fmaf(exA, valA[(size_t)k * a.ldS1], sA);
        sB = fmaf(exB, valB[(size_t)k * a.ldS1], sB);
      }
      if (p0 + pA < a.S) a.out[((size_t)b * a.S + p0 + pA) * 64 + cA] = __fdividef(sA, zA);
      if (has1 && p0 + pB < a.S) a.out[((size_t)b * a.S + p0 + pB) * 64 + cB] = __fdividef(sB, zB);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_mlp_kernel(const TcArgs a, int ntiles, int tiles_per_cloud) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);                       // [TC_SLOTS][TC_SLOT_FLOATS]
  float* S0 = ring + TC_SLOTS * TC_SLOT_FLOATS;                            // [128][ldS0]
  float* S1 = S0 + (size_t)TC_ROWS * a.ldS0;                               // [128][ldS1]
  __shared__ __align__(8) uint64_t full_bar[TC_SLOTS], empty_bar[TC_SLOTS], d_ready, a_ready;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == TC_ROW_THREADS / 32 + 1) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < TC_SLOTS; ++s) { mbarrier_init(&full_bar[s], 1); mbarrier_init(&empty_bar[s], 1); }
    mbarrier_init(&d_ready, 1);
    mbarrier_init(&a_ready, TC_ROW_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t full_a = smem_addr(&full_bar[0]), empty_a = smem_addr(&empty_bar[0]);   // + 8 * slot
  const uint32_t dready_a = smem_addr(&d_ready), aready_a = smem_addr(&a_ready);
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
    const bool leader = lane == 0;
    constexpr uint64_t DESC_FIXED = (8ull << 16) | (64ull << 32) | (1ull << 46);   // LBO 128 B, SBO 1024 B, sm100
    uint32_t use = 0, a_cnt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int L = 0; L < a.nlayers; ++L) {
        {
          const int n = a.l[L].n;
          const int col0 = a.l[L].seg_col[0], col1 = a.l[L].seg_col[1];
          const int ks0 = a.l[L].seg_n[0] >> 3;
          const int ksteps = ks0 + (a.l[L].seg_n[1] >> 3);
          const uint32_t idesc = tc_idesc_tf32(128, n);
          const uint32_t lo_off = (uint32_t)n * 128u;           // byte offset of the lo half of a chunk
          mbar_wait_a(aready_a, a_cnt & 1);                     // the A planes of this layer are complete
          ++a_cnt;
          tc_fence_after();
          const bool mstamp = a.dbg != nullptr && blockIdx.x == 0 && tile == (int)gridDim.x && leader;
          if (mstamp) a.dbg[32 + 2 * L] = clock64();
          uint32_t acc = 0;
          for (int k0 = 0; k0 < ksteps; k0 += 4, ++use) {
            const uint32_t s = use % TC_SLOTS;
            if (!(a.debug & 1)) mbar_wait_a(full_a + 8 * s, (use / TC_SLOTS) & 1);
            tc_fence_after();
            const uint32_t slot = ring_addr + s * (uint32_t)(TC_SLOT_FLOATS * sizeof(float));
            const uint64_t dh0 = DESC_FIXED | (uint64_t)((slot & 0x3ffffu) >> 4);
            const uint64_t dl0 = DESC_FIXED | (uint64_t)(((slot + lo_off) & 0x3ffffu) >> 4);
            const int kn = min(4, ksteps - k0);
            // full chunk inside one column segment (the common case): 12 MMAs from one asm block
            const bool seg0 = k0 + 4 <= ks0, seg1 = k0 >= ks0;
            if (kn == 4 && (seg0 || seg1)) {
              const uint32_t col = (uint32_t)(seg0 ? col0 + 8 * k0 : col1 + 8 * (k0 - ks0));
              tc_mma3x4_ts_warp(tb + COL_D, tb + COL_HI + col, tb + COL_LO + col, dh0, dl0, idesc, acc);
              acc = 1;
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j < kn) {
                  const int ks = k0 + j;
                  const uint32_t col = (uint32_t)(ks < ks0 ? col0 + 8 * ks : col1 + 8 * (ks - ks0));
                  const uint64_t dh = dh0 + (uint64_t)(16 * j), dl = dl0 + (uint64_t)(16 * j);   // +256 B per k-step
                  tc_mma3_ts_warp(tb + COL_D, tb + COL_HI + col, tb + COL_LO + col, dh, dl, idesc, acc);
                  acc = 1;
                }
              }
            }
            tc_commit_warp(empty_a + 8 * s);                    // slot reusable once these MMAs have read it
          }
          tc_commit_warp(dready_a);                             // accumulator complete
          if (mstamp) a.dbg[33 + 2 * L] = clock64();
        }
      }
    }
  } else {
    // ============================== row warps ==============================
    const int quad = warp & 3, slice = warp >> 2;
    const int r = quad * 32 + lane;                // tile row == TMEM lane
    const uint32_t lane_base = tb + ((uint32_t)(quad * 32) << 16);
    uint32_t d_cnt = 0;
    bool have_prev = false;
    int prev_tile = 0, prev_b = 0, prev_p0 = 0, prev_P = TC_ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // ---------------- gather the layer-0 input into the planes ----------------
      // 16-column groups are dealt round-robin to the 4 slices of a quadrant (group counter `grp`)
      int b = 0, p0 = 0, P = TC_ROWS, grp = 0;
      if (MODE == TC_PW) {
        const int row = min(tile * TC_ROWS + r, a.S - 1);
        int col = 0;
        for (int s = 0; s < a.nsrc; ++s) {
          const float* src = a.src[s] + (size_t)row * a.c_src[s];
          for (int c = 0; c < a.c_src[s]; c += 16, ++grp) {
            if ((grp & (TC_SLICES - 1)) != slice) continue;
            float v[16];
            load16(src + c, v);
            plane_store16(lane_base, col + c, v);
          }
          col += a.c_src[s];
        }
      } else {
        b = tile / tiles_per_cloud;
        P = TC_ROWS / a.K;
        p0 = (tile - b * tiles_per_cloud) * P;
        const int p = r / a.K, k = r - p * a.K;
        const bool valid = p < P && p0 + p < a.S;
        const int gp = min(p0 + min(p, P - 1), a.S - 1);
        const int n = valid ? a.idx[((size_t)b * a.S + gp) * a.K + k] : 0;
        if (MODE == TC_SA) {
          // planes: [feat(C) | xyz_nbr - xyz_ctr (3) + zeros]
          const float* f = a.f_ref + ((size_t)b * a.N + n) * a.C;
          for (int c = 0; c < a.C; c += 16, ++grp) {
            if ((grp & (TC_SLICES - 1)) != slice) continue;
            float v[16];
            load16(f + c, v);
            plane_store16(lane_base, c, v);
          }
          if ((grp & (TC_SLICES - 1)) == slice) {
            const float* q = a.xyz_ref + ((size_t)b * a.N + n) * 3;
            const float* ctr = a.xyz_ctr + ((size_t)b * a.S + gp) * 3;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
            v[0] = __fsub_rn(q[0], ctr[0]); v[1] = __fsub_rn(q[1], ctr[1]); v[2] = __fsub_rn(q[2], ctr[2]);
            plane_store16(lane_base, a.C, v);
          }
        } else if (MODE == TC_CV1) {
          // planes: [f1(C) | f2 nbr(C)] at 0, geo(10)+6 zeros at 176
          const float* f1 = a.f_ctr + ((size_t)b * a.S + gp) * a.C;
          const float* f2 = a.f_ref + ((size_t)b * a.N + n) * a.C;
          for (int c = 0; c < 2 * a.C; c += 16, ++grp) {
            if ((grp & (TC_SLICES - 1)) != slice) continue;
            float v[16];
            load16(c < a.C ? f1 + c : f2 + (c - a.C), v);
            plane_store16(lane_base, c, v);
          }
          if ((grp & (TC_SLICES - 1)) == slice) {
            float g[16];
            geo16(g, a.xyz_ctr + ((size_t)b * a.S + gp) * 3, a.xyz_ref + ((size_t)b * a.N + n) * 3);
            plane_store16(lane_base, 176, g);
          }
        } else {  // TC_CV2
          // planes: geo(10)+6 zeros at 0 (consumed by the first layer), f1(C) at 64, e1 nbr(64) at 64+C
          if (slice == 0) {
            float g[16];
            geo16(g, a.xyz_ctr + ((size_t)b * a.S + gp) * 3, a.xyz_ref + ((size_t)b * a.S + n) * 3);
            plane_store16(lane_base, 0, g);
          }
          grp = 1;
          const float* f1 = a.f_ctr + ((size_t)b * a.S + gp) * a.C;
          for (int c = 0; c < a.C; c += 16, ++grp) {
            if ((grp & (TC_SLICES - 1)) != slice) continue;
            float v[16];
            load16(f1 + c, v);
            plane_store16(lane_base, 64 + c, v);
          }
          const float* e1 = a.f_ref + ((size_t)b * a.S + n) * 64;
          for (int c = 0; c < 64; c += 16, ++grp) {
            if ((grp & (TC_SLICES - 1)) != slice) continue;
            float v[16];
            load16(e1 + c, v);
            plane_store16(lane_base, 64 + a.C + c, v);
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(S1 + (size_t)r * a.ldS1 + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
      const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && tile == (int)gridDim.x && tid == 0;
      if (stamp) a.dbg[0] = clock64();
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive_a(aready_a);
      if (stamp) a.dbg[1] = clock64();
      if (have_prev) {      // previous tile's pooling, overlapped with this tile's first GEMM
        tc_pool<MODE>(a, S0, S1, tid, prev_tile, prev_b, prev_p0, prev_P);
        have_prev = false;
        rows_sync();        // S0 / S1 free before this tile's epilogues write them
      }

      // ---------------- layer epilogues ----------------
      for (int L = 0; L < a.nlayers; ++L) {
        const TcLayer& ly = a.l[L];
        mbar_wait_a(dready_a, d_cnt & 1);
        ++d_cnt;
        tc_fence_after();
        if (stamp) a.dbg[2 + 2 * L] = clock64();
        float* sdst = ly.out_smem == 1 ? S0 + (size_t)r * a.ldS0 : (ly.out_smem == 2 ? S1 + (size_t)r * a.ldS1 : nullptr);
        for (int c = slice * 16; c < ly.n; c += 16 * TC_SLICES) {
          float v[16];
          tmem_ld16(lane_base + COL_D + c, v);
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(ly.b + c + i));
            v[i] = fmaxf(v[i] + bv.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + bv.y, 0.f);
            v[i + 2] = fmaxf(v[i + 2] + bv.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + bv.w, 0.f);
          }
          if (ly.out_col >= 0) plane_store16(lane_base, ly.out_col + c, v);
          if (sdst != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(sdst + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
        if (L + 1 < a.nlayers) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive_a(aready_a);
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
  if (n0 % 8 != 0 || n1 % 8 != 0 || col0 % 8 != 0 || col1 % 8 != 0 || n0 <= 0) return false;
  L.w = src.w; L.b = src.b; L.n = src.cout;
  L.seg_col[0] = col0; L.seg_n[0] = n0; L.seg_col[1] = col1; L.seg_n[1] = n1;
  L.out_col = out_col; L.out_smem = out_smem;
  return true;
}

template <int MODE>
static int tc_launch(TcArgs& a, int ntiles, int tiles_per_cloud, cudaStream_t st) {
  if (a.ldS1 == 0) a.ldS1 = 4;
  if (const char* d = getenv("PWCLO_TC_DEBUG")) a.debug = atoi(d);
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
    fprintf(stderr, "[tc stamps mode %d] gather_end +%lld arrive +%lld |", MODE, h[0] - h[0], h[1] - h[0]);
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
  if (nlayers < 2 || nlayers > 3 || K > 128 || C % 16 != 0 || C + 8 > 192 || (uintptr_t)feats % 16 != 0) return PWCLO_EUNSUPPORTED;
  TcArgs a = {};
  a.xyz_ref = xyz; a.xyz_ctr = new_xyz; a.f_ref = feats; a.idx = idx; a.out = out;
  a.B = B; a.N = N; a.S = S; a.K = K; a.C = C; a.nlayers = nlayers;
  int kin = C + 8;
  for (int i = 0; i < nlayers; ++i) {
    const bool last = i == nlayers - 1;
    if (!tc_layer_set(a.l[i], layers[i], 0, kin, 0, 0, last ? -1 : 0, last ? 1 : 0)) return PWCLO_EUNSUPPORTED;
    kin = layers[i].cout;
  }
  a.ldS0 = ld_for(layers[nlayers - 1].cout);
  a.ldS1 = 0;
  const int P = TC_ROWS / K, tpc = ceil_div(S, P);
  return tc_launch<TC_SA>(a, B * tpc, tpc, (cudaStream_t)stream);
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
  return tc_launch<TC_PW>(a, ntiles, 1, (cudaStream_t)stream);
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
  bool ok = tc_layer_set(a.l[0], mlp1[0], 0, 2 * C, 176, 16, 0, 0)      // [f1 | f2 | geo] -> h1 at 0
         && tc_layer_set(a.l[1], mlp1[1], 0, 128, 0, 0, 0, 0)           // h2 at 0
         && tc_layer_set(a.l[2], mlp1[2], 0, 64, 0, 0, 64, 2)           // h3 at 64 (+ fp32 copy S1)
         && tc_layer_set(a.l[3], *enc, 176, 16, 0, 0, 0, 0)             // enc at 0
         && tc_layer_set(a.l[4], mlp2[0], 0, 128, 0, 0, 0, 0)           // a1 at 0
         && tc_layer_set(a.l[5], mlp2[1], 0, 128, 0, 0, -1, 1);         // a2 -> S0
  if (!ok) return PWCLO_EUNSUPPORTED;
  a.ldS0 = ld_for(64); a.ldS1 = ld_for(64);
  const int P = TC_ROWS / K, tpc = ceil_div(S, P);
  return tc_launch<TC_CV1>(a, B * tpc, tpc, (cudaStream_t)stream);
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
  return tc_launch<TC_CV2>(a, B * tpc, tpc, (cudaStream_t)stream);
}
