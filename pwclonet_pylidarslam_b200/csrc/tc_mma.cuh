// tcgen05 / TMEM helpers (sm_100a): 5th-generation tensor-core MMA with TF32 operands and fp32
// accumulators in tensor memory.  Used for the error-compensated 3xTF32 shared-MLP contractions:
//   D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo   with  x_hi = x & 0xffffe000 (exact in tf32), x_lo = x - x_hi.
// A operands live in TMEM (row = lane, k = column), B operands in shared memory in the canonical
// K-major no-swizzle layout of the UMMA shared-memory descriptor:
//   element (n, k) of a [N x 32] chunk at byte  (n/8)*1024 + (k/4)*128 + (n%8)*16 + (k%4)*4 .
#pragma once
#include "mlp_core.cuh"

namespace pwclo {

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major, M x N tile
__host__ __device__ constexpr uint32_t tc_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, K-major, no swizzle: LBO = 128 B (next 4 k), SBO = 1024 B (next 8 rows)
__device__ __forceinline__ uint64_t tc_smem_desc(const void* p) {
  const uint64_t a = (uint64_t)((smem_addr(p) & 0x3ffffu) >> 4);
  return a | (8ull << 16) | (64ull << 32) | (1ull << 46);
}
// D[tmem] (+)= A[tmem] * B[smem]^T : one K = 8 step (32 bytes of tf32)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective variants: called by ALL lanes of the issuing warp with warp-uniform operands; one
// elected lane issues.  Keeping the C++ control flow convergent lets ptxas hold descriptors and TMEM
// addresses in uniform registers (no per-instruction R2UR / ELECT loops).
__device__ __forceinline__ void tc_mma_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the three MMAs of one 3xTF32 k-step under a single election (fewer uniform-register moves / predicates)
__device__ __forceinline__ void tc_mma3_ts_warp(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %3, %5, p;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %5, 1;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}
// a whole 32-wide weight chunk: 4 k-steps x 3 MMAs, A columns advance by 8, descriptors by 256 B (+16)
__device__ __forceinline__ void tc_mma3x4_ts_warp(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b32 ah1, ah2, ah3, al1, al2, al3;\n"
      ".reg .b64 bh1, bh2, bh3, bl1, bl2, bl3;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "add.u32 ah1, %1, 8;  add.u32 ah2, %1, 16;  add.u32 ah3, %1, 24;\n"
      "add.u32 al1, %2, 8;  add.u32 al2, %2, 16;  add.u32 al3, %2, 24;\n"
      "add.u64 bh1, %3, 16; add.u64 bh2, %3, 32;  add.u64 bh3, %3, 48;\n"
      "add.u64 bl1, %4, 16; add.u64 bl2, %4, 32;  add.u64 bl3, %4, 48;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %3, %5, p;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], bh1, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [al1], bh1, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], bl1, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], bh2, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [al2], bh2, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah2], bl2, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], bh3, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [al3], bh3, %5, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah3], bl3, %5, 1;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_warp(uint32_t bar_addr) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar_addr)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// 32 lanes x 16 columns: thread t of the warp <-> lane (32*(warp%4) + t), 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// store 16 fp32 values of this thread's row as hi / lo tf32 planes
__device__ __forceinline__ void tmem_st16_split(uint32_t taddr_hi, uint32_t taddr_lo, const float (&v)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    hi[i] = __float_as_uint(v[i]) & 0xffffe000u;
    lo[i] = __float_as_uint(__fsub_rn(v[i], __uint_as_float(hi[i])));
  }
  tmem_st16(taddr_hi, hi);
  tmem_st16(taddr_lo, lo);
}

// ---------------------------------------------------------------------------------------------
// Hybrid split (v2): D += tf32(A)*tf32(W) + bf16(A - tf32(A))*bf16(W) + bf16(A)*bf16(W - tf32(W)).
// (activations: tf32 round-to-nearest, bf16 truncation; weights: rounded to nearest on the host)
// The two correction products carry ~2^-12 of the magnitude, so 8 mantissa bits are enough for them
// (total error ~2^-20 relative, the same class as 3xTF32) and one kind::f16 MMA covers K = 16: a
// 32-wide weight chunk costs 4 + 2 + 2 = 8 MMAs instead of 12.
// 16-bit A operands in TMEM pack two consecutive k per 32-bit column (low half = even k); 16-bit B
// operands use the canonical K-major no-swizzle layout with 8 elements per 16-byte core-matrix row:
//   element (n, k) of a [N x 32] bf16 chunk at byte (n/8)*512 + (k/8)*128 + (n%8)*16 + (k%8)*2 .
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t tc_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint64_t TC_DESC_TF32 = (8ull << 16) | (64ull << 32) | (1ull << 46);   // LBO 128 B, SBO 1024 B
constexpr uint64_t TC_DESC_BF16 = (8ull << 16) | (32ull << 32) | (1ull << 46);   // LBO 128 B, SBO  512 B
__device__ __forceinline__ uint64_t tc_desc_at(uint32_t saddr, uint64_t fixed) { return fixed | (uint64_t)((saddr & 0x3ffffu) >> 4); }

__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one 32-wide chunk of the hybrid scheme, whole warp calls, one elected lane issues 8 MMAs:
//   a_t: tf32 plane column of the chunk (32 columns), a_hb / a_lb: packed bf16 planes (16 columns each)
//   b_t / b_hb / b_lb: descriptors of the chunk's tf32(W), bf16(W), bf16(W - tf32(W)) blocks
__device__ __forceinline__ void tc_mma_hybrid_chunk_warp(uint32_t d_tmem, uint32_t a_t, uint32_t a_hb, uint32_t a_lb, uint64_t b_t,
                                                         uint64_t b_hb, uint64_t b_lb, uint32_t idesc_t, uint32_t idesc_b,
                                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b32 at1, at2, at3, ah1, al1;\n"
      ".reg .b64 bt1, bt2, bt3, bh1, bl1;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %9, 0;\n"
      "add.u32 at1, %1, 8;  add.u32 at2, %1, 16;  add.u32 at3, %1, 24;\n"
      "add.u32 ah1, %2, 8;  add.u32 al1, %3, 8;\n"
      "add.u64 bt1, %4, 16; add.u64 bt2, %4, 32;  add.u64 bt3, %4, 48;\n"
      "add.u64 bh1, %5, 16; add.u64 bl1, %6, 16;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %7, p;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [at1], bt1, %7, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [at2], bt2, %7, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [at3], bt3, %7, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], %5, %8, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [al1], bh1, %8, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %6, %8, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [ah1], bl1, %8, 1;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_t), "r"(a_hb), "r"(a_lb), "l"(b_t), "l"(b_hb), "l"(b_lb), "r"(idesc_t), "r"(idesc_b), "r"(accumulate)
      : "memory");
}
// half a chunk (16 inputs): 2 tf32 + 1 + 1 bf16 MMAs
__device__ __forceinline__ void tc_mma_hybrid_half_warp(uint32_t d_tmem, uint32_t a_t, uint32_t a_hb, uint32_t a_lb, uint64_t b_t,
                                                        uint64_t b_hb, uint64_t b_lb, uint32_t idesc_t, uint32_t idesc_b,
                                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b32 at1;\n"
      ".reg .b64 bt1;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %9, 0;\n"
      "add.u32 at1, %1, 8;\n"
      "add.u64 bt1, %4, 16;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %7, p;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [at1], bt1, %7, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], %5, %8, 1;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %6, %8, 1;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_t), "r"(a_hb), "r"(a_lb), "l"(b_t), "l"(b_hb), "l"(b_lb), "r"(idesc_t), "r"(idesc_b), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// store 16 fp32 values (inputs k .. k+15 of this thread's row): tf32 plane (16 columns at taddr_t),
// packed bf16 planes of the value and of its tf32 residual (8 columns each at taddr_hb / taddr_lb).
// Integer-pipe only (the F2F conversion instructions run at a fraction of the ALU rate): tf32 by
// round-to-nearest on the bit pattern, bf16 by truncation = a byte permute of two upper halves.
__device__ __forceinline__ void tmem_st16_hybrid(uint32_t taddr_t, uint32_t taddr_hb, uint32_t taddr_lb, const float (&v)[16]) {
  uint32_t t[16], lo[16], hb[8], lb[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    t[i] = (__float_as_uint(v[i]) + 0x1000u) & 0xffffe000u;
    lo[i] = __float_as_uint(__fsub_rn(v[i], __uint_as_float(t[i])));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hb[i] = __byte_perm(__float_as_uint(v[2 * i]), __float_as_uint(v[2 * i + 1]), 0x7632);   // low half = even k
    lb[i] = __byte_perm(lo[2 * i], lo[2 * i + 1], 0x7632);
  }
  tmem_st16(taddr_t, t);
  tmem_st8(taddr_hb, hb);
  tmem_st8(taddr_lb, lb);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
// 8-input version of tmem_st16_hybrid
__device__ __forceinline__ void tmem_st8_hybrid(uint32_t taddr_t, uint32_t taddr_hb, uint32_t taddr_lb, const float (&v)[8]) {
  uint32_t t[8], lo[8], hb[4], lb[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    t[i] = (__float_as_uint(v[i]) + 0x1000u) & 0xffffe000u;
    lo[i] = __float_as_uint(__fsub_rn(v[i], __uint_as_float(t[i])));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hb[i] = __byte_perm(__float_as_uint(v[2 * i]), __float_as_uint(v[2 * i + 1]), 0x7632);
    lb[i] = __byte_perm(lo[2 * i], lo[2 * i + 1], 0x7632);
  }
  tmem_st8(taddr_t, t);
  tmem_st4(taddr_hb, hb);
  tmem_st4(taddr_lb, lb);
}

}  // namespace pwclo
