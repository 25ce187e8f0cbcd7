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

}  // namespace pwclo
