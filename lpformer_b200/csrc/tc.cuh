// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers and the shared-memory operand layout used by the
// tensor-core kernels (sm_100a only).
//
// Operand layout ("K-major, 128-byte swizzle", the canonical UMMA layout): an operand tile of R rows
// (R % 8 == 0) by 32 fp32 (= one 128-byte row = one "k-block") lives in R*128 bytes at a 1024-byte
// aligned shared-memory address; element (r, k) is at byte
//     (r/8)*1024 + (r%8)*128 + (((k/4) ^ (r%8)) * 16) + (k%4)*4 .
// One tcgen05.mma.kind::tf32 consumes K = 8 (32 bytes); the 4 K-slices of a k-block are addressed by
// advancing the descriptor start address by 32 bytes.
//
// fp32 accuracy on the tf32 pipe ("3xTF32"): x = hi + lo with hi = x with the low 13 mantissa bits cleared
// (exactly representable in tf32) and lo = x - hi (exact in fp32); A.W ~= Ahi.Whi + Alo.Whi + Ahi.Wlo, error
// O(2^-21) relative per product, far inside the path's 1e-4 tolerance.
#pragma once
#include "common.cuh"

namespace lpf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// byte offset of the 16-byte chunk `chunk` (0..7) of row r inside a swizzled k-block tile
__device__ __forceinline__ uint32_t swz_chunk_off(int r, int chunk) {
    return (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + (((chunk ^ (r & 7)) & 7) << 4));
}

// One lane of a converged warp (elect.sync): the warp-uniform way to issue tcgen05.mma / commit — the operands stay in
// uniform registers; code under `if (lane == 0)` makes the compiler re-derive uniformity around every instruction.
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

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// (the suspend-time hint lets the hardware park the thread until the phase flips — at most ~the hint in ns — instead of
// returning at once: a waiting warp then costs no issue slots, which the busy warps of a warp-specialised CTA need)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    return ok != 0;
}
// Bounded spin: a barrier that never completes (a descriptor / pipeline bug) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}

// ---- bulk copy global -> shared (TMA, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w of the CTA's first 4 warps owns lanes 32w..32w+31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait: the registers are valid after the next tmem_ld_wait() (which waits for every
// tcgen05.ld this thread has issued).  TMEM reads run at 64 B per cycle per SM: a 128 x 128 fp32 accumulator takes
// ~1,000 cycles to read back, worth overlapping with the arithmetic on the previous columns.
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, float (&v)[16]) {
    // (the destination registers are the asm outputs themselves: no move may sit between the load and the wait)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane, written from registers (complete after tmem_st_wait())
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp of CUTLASS 3.9+/4.x)
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start>>4 [0,14), LBO>>4 [16,30) (=1, unused for
// swizzled K-major), SBO>>4 [32,46) (= 1024 B between 8-row groups), version=1 [46,48), layout=2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major:
// c_format=F32(1) [4,6), a_format=TF32(2) [7,10), b_format=TF32(2) [10,13), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T, one 128 x N x 8 tf32 MMA issued by the calling thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T: the A operand comes from TMEM (lane = row, one 32-bit column per tf32 element of
// the K-slice: 8 consecutive columns), written there by tcgen05.st — an operand that is produced by the epilogue of the
// previous contraction never goes through shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// The three products of the split-precision contraction for one k-block (4 K-slices of 8): small terms first.
__device__ __forceinline__ void issue_kblock_3x(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                uint32_t b_lo, uint32_t idesc, bool first) {
    const uint64_t dah = make_smem_desc(a_hi), dal = make_smem_desc(a_lo);
    const uint64_t dbh = make_smem_desc(b_hi), dbl = make_smem_desc(b_lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint64_t o = (uint64_t)(j * 2);  // 32 bytes >> 4
        umma_tf32(d_tmem, dal + o, dbh + o, idesc, (first && j == 0) ? 0u : 1u);
        umma_tf32(d_tmem, dah + o, dbl + o, idesc, 1u);
        umma_tf32(d_tmem, dah + o, dbh + o, idesc, 1u);
    }
}

// the same with the A operand in TMEM: a_hi / a_lo = TMEM addresses of the k-block's 32 columns of the two halves
__device__ __forceinline__ void issue_kblock_3x_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                   uint32_t b_lo, uint32_t idesc, bool first) {
    const uint64_t dbh = make_smem_desc(b_hi), dbl = make_smem_desc(b_lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint64_t o = (uint64_t)(j * 2);  // 32 bytes >> 4
        umma_tf32_ts(d_tmem, a_lo + 8 * j, dbh + o, idesc, (first && j == 0) ? 0u : 1u);
        umma_tf32_ts(d_tmem, a_hi + 8 * j, dbl + o, idesc, 1u);
        umma_tf32_ts(d_tmem, a_hi + 8 * j, dbh + o, idesc, 1u);
    }
}

constexpr int kTileM = 128;
constexpr int kKBlock = 32;                      // fp32 per 128-byte row
constexpr int kATileBytes = kTileM * 128;        // one k-block of a 128-row operand

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ __forceinline__ uint32_t tmem_cols_for(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

}  // namespace tc
}  // namespace lpf
