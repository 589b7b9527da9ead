// tc05.cuh -- PTX wrappers shared by the tcgen05 kernels of libcnc_b200 (sm_100a): mbarriers, bulk copies, shared-memory
// operand descriptors, tcgen05.mma / commit / ld / st, and the error-compensated tf32 split.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cnc {
namespace tc {

constexpr int TILE_M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// non-blocking test in a spin loop: for the two service warps, where the wake-up latency of a suspended
// try_wait would sit on the critical path of every chunk
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "SPIN_%=:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra OK_%=;\n\t"
        "bra SPIN_%=;\n\t"
        "OK_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: 8-row groups of 1024 B (SBO), rows of 128 B.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);       // start address
    d |= (uint64_t)1u << 16;                      // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;            // stride byte offset
    d |= (uint64_t)1u << 46;                      // descriptor version (sm_100)
    d |= (uint64_t)2u << 61;                      // SWIZZLE_128B
    return d;
}
template <int N>
__device__ __forceinline__ constexpr uint32_t idesc_bf16() {   // kind::f16, A/B bf16, D fp32, K = 16
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
template <int N>
__device__ __forceinline__ constexpr uint32_t idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
// The async ops of the MMA warp are executed by the whole (converged) warp with the leader election inside the
// asm block: the C++ around them stays warp-uniform straight-line code, so every operand is born in a uniform
// register (an `if (elect)` region makes ptxas shuttle each operand through R2UR + a uniformisation loop,
// ~100 cycles per MMA).  elect.sync picks the same leader every time, which tcgen05.commit relies on.
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xFFFFFFFF;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xFFFFFFFF;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xFFFFFFFF;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xFFFFFFFF;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xFFFFFFFF;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xFFFFFFFF;\n\t"
                 "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
                 "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// Error-compensated split of an fp32 operand for the tensor cores: v = hi + lo with hi = v rounded to tf32
// (nearest, ties away: add half an ulp to the magnitude, clear the low 13 bits; cvt.rna.tf32.f32 does the same but
// ptxas expands it to a ~10-instruction NaN/Inf-safe sequence).  A*B ~= Ahi*Bhi [one kind::tf32 MMA, K = 8]
//   + (Ahi*Blo + Alo*Bhi) [ONE kind::f16 MMA, K = 16: the bf16 pairs (Ahi, Alo) against (Blo, Bhi)].
// The correction terms are 2^-12 of the product, so their bf16 rounding (2^-9) costs 2^-21 relative: the same
// order as the dropped Alo*Blo term.  One 32-bit word per element holds the pair, i.e. the "lo" buffers keep
// their size and layout, while the MMA count per k-step drops from 3 to 2 and the two MMAs hit different
// accumulators (a dependent accumulate costs ~130 cycles, more than the MMA itself at N <= 160).
__device__ __forceinline__ uint32_t rna_tf32(float v) { return (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t pack_bf16(float lower, float upper) {   // element 2k = lower, 2k+1 = upper
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    return d;
}
// A side: hi word + pair (Ahi, Alo)
__device__ __forceinline__ void split_tf32(float v, uint32_t &hi, uint32_t &pair) {
    hi = rna_tf32(v);
    const float h = __uint_as_float(hi);
    pair = pack_bf16(h, __fsub_rn(v, h));
}


// MN-major operand tile of 32-bit elements (the contraction index runs over rows, 32 fp32 of the M/N index per
// 128-byte row).  tf32 has exactly one legal layout for this (CUTLASS: "for mn-major tf32 operands, SW128_32B is the
// only available smem layout"): atoms of 4 K-rows x 128 B, swizzled in 32-byte units -- byte address bits [5,7) are
// XORed with bits [7,9), i.e. the 32-byte chunk index with (row & 3) -- layout type SWIZZLE_128B_BASE32B (= 1).
// LBO = byte distance between atoms along M/N, SBO = between 4-row K groups (an MMA of K = 8 reads two of them).
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1u << 46;
    d |= (uint64_t)1u << 61;
    return d;
}
// kind::tf32 instruction descriptor with both operands MN-major (bits 15 / 16)
template <int N>
__device__ __forceinline__ constexpr uint32_t idesc_tf32_mn() {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

}  // namespace tc
}  // namespace cnc
