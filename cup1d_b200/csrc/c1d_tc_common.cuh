// PTX helpers shared by the tcgen05 kernels (c1d_mlp_tc.cu, c1d_mlp_i8.cu) and tools/i8_probe.cu:
// mbarrier, bulk copies, tcgen05 fences / commit / MMA forms, TMEM loads and stores, shared-memory
// matrix descriptors and the swizzled operand layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace c1d_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp, chosen by the hardware: the predicate the compiler treats as warp-uniform control
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: every try_wait suspends for at most 10 ms; after ~20 s without the phase completing
// (a faulted bulk copy, a lost arrive) the kernel reports where it stood and traps, which the host
// sees as a CUDA error on its next call instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (int tries = 0; tries < 2048; ++tries) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("cup1d_b200: mbarrier wait timed out (block %d thread %d barrier 0x%x parity %u)\n", (int)blockIdx.x,
         (int)threadIdx.x, bar, parity);
  __trap();
}
// the same with a non-blocking test first: a phase that is already complete (the usual case for the weight ring of
// the MMA issuers) costs one check instead of the blocking form's suspend / resume path (~180 cycles measured)
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!done) mbar_wait(bar, parity);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// one slice of a weight unit, written to the same shared-memory offset of every CTA in cta_mask;
// each destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
}

// shared-memory matrix descriptors, K-major operands.  Swizzle atoms are 8 rows tall; the stride
// byte offset is the distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;           // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;                     // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
  return d;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) { return make_desc(addr, 1024u, 2u); }
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t addr) { return make_desc(addr, 512u, 4u); }
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t addr) { return make_desc(addr, 256u, 6u); }

// instruction descriptors (cute/arch/mma_sm100_desc.hpp field positions): D format at bit 4, A format at
// bit 7, B format at bit 10, N >> 3 at bit 17, M >> 4 at bit 24; A and B K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// signed 8-bit x signed 8-bit -> 32-bit integers (kind::i8, K = 32 per instruction)
__device__ __forceinline__ uint32_t make_idesc_s8(uint32_t n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

#define C1D_TC_MMA(KIND, NAME)                                                                                          \
  __device__ __forceinline__ void NAME##_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,         \
                                            uint32_t acc) {                                                             \
    asm volatile(                                                                                                       \
        "{\n\t"                                                                                                         \
        ".reg .pred p;\n\t"                                                                                             \
        "setp.ne.b32 p, %4, 0;\n\t"                                                                                     \
        "tcgen05.mma.cta_group::1.kind::" KIND " [%0], %1, %2, %3, p;\n\t"                                              \
        "}" ::"r"(d_tmem),                                                                                              \
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)                                                                  \
        : "memory");                                                                                                    \
  }                                                                                                                     \
  __device__ __forceinline__ void NAME##_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,         \
                                            uint32_t acc) {                                                             \
    asm volatile(                                                                                                       \
        "{\n\t"                                                                                                         \
        ".reg .pred p;\n\t"                                                                                             \
        "setp.ne.b32 p, %4, 0;\n\t"                                                                                     \
        "tcgen05.mma.cta_group::1.kind::" KIND " [%0], [%1], %2, %3, p;\n\t"                                            \
        "}" ::"r"(d_tmem),                                                                                              \
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)                                                                  \
        : "memory");                                                                                                    \
  }
C1D_TC_MMA("tf32", mma)
C1D_TC_MMA("i8", mma_i8)
#undef C1D_TC_MMA

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// byte offset of (row, 16-byte group g of k) inside one SWIZZLE_128B tile of 128-byte rows
__device__ __host__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t g) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((g ^ (row & 7u)) << 4);
}
// byte offset of (row, 16-byte group g in 0..3) inside a SWIZZLE_64B tile of 64-byte rows
__device__ __host__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t g) {
  return (row >> 3) * 512u + (row & 7u) * 64u + ((g ^ ((row >> 1) & 3u)) << 4);
}
// byte offset of (row, 16-byte group g in 0..1) inside a SWIZZLE_32B tile of 32-byte rows
__device__ __host__ __forceinline__ uint32_t sw32_off(uint32_t row, uint32_t g) {
  return (row >> 3) * 256u + (row & 7u) * 32u + ((g ^ ((row >> 2) & 1u)) << 4);
}

}  // namespace c1d_tc
