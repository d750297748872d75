// Fused per-link heads on the tensor cores, fp16-split operands (d = 64): the chain of heads_tc.cu
//     xprod = X[a] * X[b]                                   (gather, reference train/testing.py:29,113)
//     h     = ReLU(LayerNorm(W1 xprod + b1))                (elementwise_lin, models/other_models.py:125-133)
//     z     = ReLU(W23 h + offset),  W23 = Ws1[:, :d] W2    (:135 folded into mlp_score layer 0, :173-177)
//     prob  = sigmoid(ws2 . z + bs2)                        (:178-179)
// with both contractions issued as tcgen05.mma.kind::f16 instead of kind::tf32.  An fp16 has the same 11-bit
// significand as a tf32, so the split x = hi + lo (hi = fp16(x), lo = fp16(x - hi)) carries the same 22 bits and the
// three products hi.hi + lo.hi + hi.lo give the same fp32-level result as 3xTF32 — but an f16 MMA covers K = 16 per
// instruction at twice the rate, and the operands take half the shared memory (weights 48 KB, operand tile 32 KB):
// the tensor-pipe time of a tile halves and the operand tile can be double-buffered.
//
// fp16 has a 5-bit exponent, so every operand is brought into range by an exact power-of-two scale that the epilogues
// undo (a multiplication folded into the bias add):
//   * weights: one scale per matrix, chosen by lpf_pack_weight_f16 callers so that max|W| lands in [2^14, 2^15);
//   * the gathered rows: one scale PER LINK (row max in [2^14, 2^15), eight lanes share a row: three shuffles), kept
//     in a four-tile shared-memory ring for epilogue 1;
//   * h: LayerNorm bounds it by sqrt(d) max|g| + max|b|; the caller folds the scale into g and b (exact).
// Elements more than 2^-10 below their row's maximum lose relative (not absolute) precision in `lo` — the same
// absolute error an fp32 sum of the row has.
//
// Pipeline as in heads_tc.cu (producers 8 warps / MMA warp / consumers 8 warps, accumulators double-buffered in TMEM,
// H written back to TMEM as the A operand of contraction 2, two packed halves per 32-bit column), plus
//   * a double-buffered operand tile: the producers run a full tile ahead of contraction 1;
//   * the gather itself by the copy engine: TMA row gathers (cp.async.bulk.tensor.2d ... tile::gather4 over a tensor map
//     of X: FOUR 256-byte X[b] rows per instruction, completion on an mbarrier) into a three-tile staging ring, issued
//     two tiles ahead — 64 KB in flight per SM without a register or a scoreboard entry (the register gather of
//     heads_tc.cu keeps 16 KB in flight and its producers wait ~3 k cycles per tile for it; one plain bulk copy per row
//     is bound by the copy engine's ~46 cycles per instruction: measured 6 k cycles per 128 rows).  The producers read
//     the staged rows from shared memory.
#include <cuda.h>
#include <cuda_fp16.h>

#include "tc.cuh"

namespace lpf {

using namespace tc;

namespace {

#ifndef LPF_HEADS_SPLIT
#define LPF_HEADS_SPLIT 2
#endif
constexpr int kSplit = LPF_HEADS_SPLIT;                  // consumer threads per link (= TMEM lane), 1/kSplit of the accumulator
                                                         // columns each: warp w reads lanes 32 (w % 4) .. + 31, column part w / 4
constexpr int kConsumers = 128 * kSplit;                 // the epilogues are latency chains (TMEM reads, exchanges): four warps
                                                         // per scheduler hide them, two do not
constexpr int kProducers = 256;                          // 8 warps: scale, split, store the staged rows
constexpr int kGatherWarps = 2;                          // warps that issue the row gathers (alternate tiles)
constexpr int kThreads = kConsumers + kProducers + 32 + 32 * kGatherWarps;   // + one warp that issues every MMA
constexpr int kStages = 3;        // staging ring of gathered X[b] rows (tiles)

struct Params {
    const int64_t* links;
    int64_t bs;
    const int32_t* idx;
    int64_t n;
    const void* X;       // node table [n_nodes, d]: fp32, or bf16 when x_bf16 (ldx in elements either way)
    int64_t ldx;
    int x_bf16;
    const void* w1p;     // lpf_pack_weight_f16 image of elementwise_lin.linears[0].weight [d, d]
    const float* b1;
    const float* ln_g;   // LayerNorm weight / bias, pre-multiplied by the h scale
    const float* ln_b;
    const void* w23p;    // lpf_pack_weight_f16 image of Ws1[:, :d] . elementwise_lin.linears[1].weight [2d, d]
    const float* c3;
    const float* zb;
    int64_t ld_zb;
    const float* ws2;
    const float* bs2;
    float* prob;
    int logits;
    const int64_t* n_dev;
    float inv_sw1;       // 1 / scale of the W1 image
    float inv_s3;        // 1 / (h scale * scale of the W23 image)
    long long* dbg;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// instruction descriptor for kind::f16 with fp16 operands, fp32 accumulate, both operands K-major:
// c_format=F32(1) [4,6), a_format=F16(0) [7,10), b_format=F16(0) [10,13), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T, one 128 x N x 16 fp16 MMA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with A in TMEM: lane = row, 8 consecutive 32-bit columns = the 16 halves of the K-slice (k even in the low half)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[16]) { tmem_st16u(taddr, v); }
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// explicit shared-space accesses by 32-bit address (through a generic pointer the compiler emits LD.E / ST.E, and splits
// the 8-byte stores whose alignment it cannot prove)
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_v2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
// four bf16 (two 32-bit words, little endian) -> four fp32
__device__ __forceinline__ float4 bf16x4(uint32_t w0, uint32_t w1) {
    return make_float4(__uint_as_float(w0 << 16), __uint_as_float(w0 & 0xffff0000u), __uint_as_float(w1 << 16),
                       __uint_as_float(w1 & 0xffff0000u));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot for two lanes' worth — the kernel is bound by instruction
// issue, not by the fp32 pipe)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 relu2(float2 a) { return make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
// (x0, x1) -> packed fp16 hi pair and packed fp16 pair of the residuals
__device__ __forceinline__ void split2(float2 x, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __float22half2_rn(x);
    const float2 f = __half22float2(h);
    const float2 r = add2(x, make_float2(-f.x, -f.y));
    hi = h2_bits(h);
    lo = h2_bits(__float22half2_rn(r));
}

// Packed fp16 weight image: per block of 64 input channels [2 (hi, lo)][NP rows][128 B swizzled]
__global__ void __launch_bounds__(256) pack_weight_f16_kernel(const float* __restrict__ W, int64_t ldw, int N, int K, int NP,
                                                              int KB, float scale, uint8_t* __restrict__ out) {
    const int64_t total = (int64_t)KB * NP * 64;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e % 64);
        const int n = (int)((e / 64) % NP);
        const int kb = (int)(e / (64 * (int64_t)NP));
        const int k = kb * 64 + kk;
        const float w = (n < N && k < K) ? W[(int64_t)n * ldw + k] * scale : 0.f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const uint32_t off = swz_chunk_off(n, kk >> 3) + (uint32_t)((kk & 7) << 1);
        uint8_t* base = out + (size_t)kb * 2 * NP * 128;
        *reinterpret_cast<__half*>(base + off) = hi;
        *reinterpret_cast<__half*>(base + (size_t)NP * 128 + off) = lo;
    }
}

template <int D, bool ZB>
__global__ void __launch_bounds__(kThreads, 1) link_heads_f16_kernel(const __grid_constant__ Params p,
                                                                     const __grid_constant__ CUtensorMap xmap) {
    static_assert(D == 64, "one 128-byte operand row = 64 halves");
    constexpr int KBX = D / 32;                      // 128-byte blocks of an fp32 X row
    constexpr int N3 = 2 * D;                        // width of mlp_score's hidden layer
    constexpr uint32_t W1_BYTES = 2 * D * 128;       // hi, lo images of [D, 64]
    constexpr uint32_t W3_BYTES = 2 * N3 * 128;      // hi, lo images of [2D, 64]
    constexpr uint32_t A_HALF = kTileM * 128;        // one image (hi or lo) of an operand tile
    constexpr uint32_t STAGE_BYTES = kTileM * D * 4; // fp32 X[b] rows of a tile
    constexpr uint32_t TMEM_COLS = 512;              // D1 x 2 (2D) | D3 x 2 (4D) | H hi (D/2) | H lo (D/2) = 7D = 448

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // bar_w: weights landed;  bar_a_ready[b]: the producers have stored operand buffer b;  bar_mma1[b]: contraction 1
    // from operand buffer b into accumulator buffer b complete;  bar_d1_free[b]: the consumers have read accumulator
    // buffer b (four arrivals: one per lane quadrant);  bar_h_ready: the consumers have stored H (four arrivals);  bar_mma3: contraction 2 complete (accumulator ready, H free).
    // bar_ids[s]: the gather warp has written id-ring slot s (32 arrivals);  bar_stage[s]: the 32 four-row gathers of staging
    // slot s have landed (32 arrivals: every issuing lane announces its own bytes);  bar_stage_free[s]: the producers have
    // converted the rows of staging slot s.
    __shared__ uint64_t bar_w, bar_mma1[2], bar_d1_free[2], bar_mma3, bar_a_ready[2], bar_h_ready, bar_stage[kStages],
        bar_stage_free[kStages], bar_ids[4];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) int32_t ids[4][2][kTileM];            // ring over tiles: [tile & 3][a, b][row] (b < 0: no such link), written by the gather warp
    __shared__ float s_rscale[4][kTileM];            // ring over tiles: 1 / (the link's operand scale)
    __shared__ __align__(16) float s_b1[D], s_g[D], s_bt[D], s_c3[N3], s_ws2[N3];
    __shared__ float s_sum[kSplit][kTileM], s_sq[kSplit][kTileM], s_dot[kSplit][kTileM];   // exchanges between the threads of a link

    const int tid = threadIdx.x, warp = tid >> 5;
    if (p.dbg && tid == 0) {
        if (blockIdx.x == 0) p.dbg[256] = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.dbg[400 + 2 * blockIdx.x]));     // (profiling: every CTA's life span)
    }
    const int64_t n_links = p.n_dev ? min(p.n, *p.n_dev) : p.n;
    const int64_t ntiles = (n_links + kTileM - 1) / kTileM;
    const int esz = p.x_bf16 ? 2 : 4;                 // bytes per element of the node table (a staged row: D * esz)
    // tile k of this CTA: blockIdx.x + k gridDim.x (every role walks the same sequence by itself)
    if ((int64_t)blockIdx.x >= ntiles) return;       // nothing to do: before any barrier / TMEM / bulk-copy state
    const uint32_t my_tiles = (uint32_t)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    auto tile_of = [&](uint32_t k) -> int64_t { return (int64_t)blockIdx.x + (int64_t)k * gridDim.x; };
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW1 = smem;
    uint8_t* sW3 = sW1 + W1_BYTES;
    uint8_t* sA = sW3 + W3_BYTES;                    // [buffer][hi, lo][128 rows][128 B]
    uint8_t* sStage = sA + 2 * 2 * A_HALF;           // [slot][128 rows][D fp32]

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_mma1[b], 1);
            mbar_init(&bar_d1_free[b], 4);
            mbar_init(&bar_a_ready[b], 1);
        }
        mbar_init(&bar_mma3, 1);
        mbar_init(&bar_h_ready, 4);
        for (int b = 0; b < kStages; ++b) {
            mbar_init(&bar_stage[b], kTileM / 4);
            mbar_init(&bar_stage_free[b], 1);
        }
        for (int b = 0; b < 4; ++b) mbar_init(&bar_ids[b], 32);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar_w, W1_BYTES + W3_BYTES);
        bulk_g2s(sW1, p.w1p, W1_BYTES, &bar_w);
        bulk_g2s(sW3, p.w23p, W3_BYTES, &bar_w);
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    for (int c = tid; c < D; c += kThreads) {
        s_b1[c] = p.b1[c];
        s_g[c] = p.ln_g[c];
        s_bt[c] = p.ln_b[c];
    }
    for (int c = tid; c < N3; c += kThreads) {
        s_c3[c] = p.c3 ? p.c3[c] : 0.f;
        s_ws2[c] = p.ws2[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tmem_d = tmem_slot;
    const uint32_t d1 = tmem_d, d3 = tmem_d + 2 * D, h_hi = tmem_d + 6 * D, h_lo = tmem_d + 6 * D + D / 2;
    const uint32_t idesc_d = make_idesc_f16(kTileM, D), idesc_3 = make_idesc_f16(kTileM, N3);
    const uint32_t aA = smem_u32(sA), aW1 = smem_u32(sW1), aW3 = smem_u32(sW3);

    if (warp > (kConsumers + kProducers) / 32) {
        // =========================== the gather warps (tiles g, g + kGatherWarps, ...): lane l brings rows 4l .. 4l+3 of a
        // tile into the staging ring with one tile::gather4 copy and writes their ids into the id ring.  An issue costs
        // the warp ~80 cycles (the operands go through uniform registers lane by lane) and blocks until the copy engine
        // accepts it, and the ids come from DRAM (the batch's links are streamed once) — which is why no other role does
        // this: here both latencies hide behind the tiles in flight (ids one own-tile ahead in registers, rows up to
        // kStages tiles ahead of the producers).  Rows without a link (partial last tile) fetch row 0; never read.
        const int lane = tid & 31, g = warp - (kConsumers + kProducers) / 32 - 1;
        auto load_ids = [&](uint32_t k, int64_t (&a)[4], int64_t (&b)[4]) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t j = tile_of(k) * kTileM + 4 * lane + r;
                const int64_t jj = j < n_links ? j : 0;
                const int64_t pos = p.idx ? (int64_t)__ldg(p.idx + jj) : jj;
                a[r] = __ldg(p.links + pos);
                b[r] = __ldg(p.links + p.bs + pos);
            }
        };
        int64_t la[4], lb[4];
        if ((uint32_t)g < my_tiles) load_ids(g, la, lb);
        for (uint32_t k = g; k < my_tiles; k += kGatherWarps) {
            const uint32_t st = k % kStages;
            // (slot k % kStages was tile k - kStages's; its id-ring slot k & 3 was tile k - 4's, converted even earlier)
            if (k >= (uint32_t)kStages) mbar_wait(&bar_stage_free[st], (k / kStages - 1) & 1);
            int32_t a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool ok = tile_of(k) * kTileM + 4 * lane + r < n_links;
                a[r] = ok ? (int32_t)la[r] : 0;
                b[r] = ok ? (int32_t)lb[r] : -1;
            }
            *reinterpret_cast<int4*>(&ids[k & 3][0][4 * lane]) = make_int4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<int4*>(&ids[k & 3][1][4 * lane]) = make_int4(b[0], b[1], b[2], b[3]);
            mbar_arrive(&bar_ids[k & 3]);
            if (k + kGatherWarps < my_tiles) load_ids(k + kGatherWarps, la, lb);
#pragma unroll
            for (int r = 0; r < 4; ++r) {                // the source rows into L2 (one row per query in an evaluation batch)
                if (r == 0 || a[r] != a[0]) {
                    const char* ra = reinterpret_cast<const char*>(p.X) + (int64_t)a[r] * p.ldx * esz;
                    for (int o = 0; o < D * esz; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ra + o));
                }
            }
            mbar_arrive_expect_tx(&bar_stage[st], 4 * D * esz);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                ::"r"(smem_u32(sStage + st * STAGE_BYTES + lane * (4 * D * esz))), "l"(&xmap), "r"(smem_u32(&bar_stage[st])),
                "r"(0), "r"(max(b[0], 0)), "r"(max(b[1], 0)), "r"(max(b[2], 0)), "r"(max(b[3], 0))
                : "memory");
        }
        return;
    }
    if (warp == (kConsumers + kProducers) / 32) {
        // =========================== the MMA warp (one elected lane issues): contraction 1 of tile it + 1, then
        // contraction 2 of tile it
        const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_d, 0);
        const uint32_t u_d1 = tbase, u_d3 = tbase + 2 * D, u_hhi = tbase + 6 * D, u_hlo = tbase + 6 * D + D / 2;
        mbar_wait(&bar_w, 0);
        auto mma1 = [&](uint32_t it) {
            const uint32_t b = it & 1;
            mbar_wait(&bar_a_ready[b], (it >> 1) & 1);
            if (it >= 2) mbar_wait(&bar_d1_free[b], ((it >> 1) - 1) & 1);     // epilogue 1 of tile it - 2 has read this buffer
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dah = make_smem_desc(aA + b * 2 * A_HALF), dal = make_smem_desc(aA + b * 2 * A_HALF + A_HALF);
                const uint64_t dbh = make_smem_desc(aW1), dbl = make_smem_desc(aW1 + D * 128);
#pragma unroll
                for (int j = 0; j < D / 16; ++j) {           // K-slices of 16 halves = 32 bytes
                    const uint64_t o = (uint64_t)(j * 2);
                    umma_f16(u_d1 + b * D, dal + o, dbh + o, idesc_d, j == 0 ? 0u : 1u);
                    umma_f16(u_d1 + b * D, dah + o, dbl + o, idesc_d, 1u);
                    umma_f16(u_d1 + b * D, dah + o, dbh + o, idesc_d, 1u);
                }
                umma_commit(&bar_mma1[b]);
            }
            __syncwarp();
        };
        const bool mstamp = p.dbg && blockIdx.x == 0 && (tid & 31) == 0;
#define LPF_MSTAMP(k) do { if (mstamp && it < 16) p.dbg[it * 16 + 8 + (k)] = clock64(); } while (0)
        mma1(0);
        for (uint32_t it = 0; it < my_tiles; ++it) {
            LPF_MSTAMP(0);
            if (it + 1 < my_tiles) {
                mbar_wait(&bar_a_ready[(it + 1) & 1], ((it + 1) >> 1) & 1);
                LPF_MSTAMP(1);
                mma1(it + 1);
            }
            LPF_MSTAMP(2);
            mbar_wait(&bar_h_ready, it & 1);
            LPF_MSTAMP(3);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dbh = make_smem_desc(aW3), dbl = make_smem_desc(aW3 + N3 * 128);
#pragma unroll
                for (int j = 0; j < D / 16; ++j) {
                    const uint64_t o = (uint64_t)(j * 2);
                    umma_f16_ts(u_d3 + (it & 1) * N3, u_hlo + 8 * j, dbh + o, idesc_3, j == 0 ? 0u : 1u);
                    umma_f16_ts(u_d3 + (it & 1) * N3, u_hhi + 8 * j, dbl + o, idesc_3, 1u);
                    umma_f16_ts(u_d3 + (it & 1) * N3, u_hhi + 8 * j, dbh + o, idesc_3, 1u);
                }
                umma_commit(&bar_mma3);
            }
            __syncwarp();
            LPF_MSTAMP(4);
        }
#undef LPF_MSTAMP
        return;
    }
    if (warp >= kConsumers / 32) {
        // =========================== producers: X[a] * staged X[b] row, scale the link's row into fp16 range, split it
        // into hi / lo halves and store them in the UMMA layout.  The X[b] rows arrive in the staging ring from the gather
        // warp.  Every global load of this loop is consumed an iteration later (ids) or at its end: nothing in it waits
        // for DRAM.
        constexpr int RPP = kProducers / 8;              // rows per pass (eight threads per row)
        constexpr int NPASS = kTileM / RPP;
        const int ptid = tid - kConsumers;
        const int chunk = ptid & 7, row_in_pass = ptid >> 3;
        uint32_t it = 0;
        // The four rows of a warp in one pass: bits 1 and 2 of the row number swapped, so that two of them have
        // (row & 4) set — their 64-byte stores land in the other half of the banks (the swizzle XORs the 16-byte chunk
        // number with row & 7): two shared-memory wavefronts per store instead of four.
        auto rowof = [&](int pass) -> int {
            const int r = pass * RPP + row_in_pass;
            return (r & ~6) | ((r & 2) << 1) | ((r & 4) >> 1);
        };
        // four consecutive channels of node row `node` from channel c
        auto ld4 = [&](int64_t node, int c) -> float4 {
            if (p.x_bf16) {
                const uint2 w = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.X) + node * p.ldx + c));
                return bf16x4(w.x, w.y);
            }
            return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.X) + node * p.ldx + c));
        };
        // ... of staged row `row`
        auto lds4 = [&](uint32_t stage, int row, int c) -> float4 {
            if (p.x_bf16) {
                const float2 w = lds_v2(stage + row * (D * 2) + c * 2);
                return bf16x4(__float_as_uint(w.x), __float_as_uint(w.y));
            }
            return lds_v4(stage + row * (D * 4) + c * 4);
        };
        // does every row of this thread in the tile of ring slot `slot` share its source?  (and that source's X row)
        auto tile_source = [&](uint32_t slot, bool& same_a, float4 (&xa)[KBX]) {
            const int32_t a0 = ids[slot][0][rowof(0)];
            same_a = true;
#pragma unroll
            for (int pass = 1; pass < NPASS; ++pass) same_a &= ids[slot][0][rowof(pass)] == a0;
#pragma unroll
            for (int kb = 0; kb < KBX; ++kb) xa[kb] = ld4(a0, kb * 32 + chunk * 4);
        };
        mbar_wait(&bar_ids[0], 0);
        float4 xa[KBX];
        bool same_a;
        tile_source(0, same_a, xa);
        const bool pstamp = p.dbg && blockIdx.x == 0 && ptid == 0;
#define LPF_PSTAMP(k) do { if (pstamp && it < 16) p.dbg[it * 16 + ((k) == 3 ? 7 : 13 + (k))] = clock64(); } while (0)
        uint32_t st = 0;                                 // staging slot of tile it = it % kStages
        for (; it < my_tiles; ++it) {
            LPF_PSTAMP(0);
            const bool more = it + 1 < my_tiles;
            // the next tile's source row: its ids come from the gather warp (tiles ahead), the row from L2; first used in
            // the next iteration
            bool same_n = true;
            float4 xa_n[KBX];
            if (more) {
                mbar_wait(&bar_ids[(it + 1) & 3], ((it + 1) >> 2) & 1);
                tile_source((it + 1) & 3, same_n, xa_n);
            }
            LPF_PSTAMP(3);
            // contraction 1 of tile it - 2 has read this operand buffer; the tile's rows have landed
            if (it >= 2) mbar_wait(&bar_mma1[it & 1], ((it - 2) >> 1) & 1);
            mbar_wait(&bar_stage[st], (it / kStages) & 1);
            LPF_PSTAMP(1);
            const uint32_t abuf = aA + (it & 1) * 2 * A_HALF;
            const uint32_t stage = smem_u32(sStage) + st * STAGE_BYTES;
            const uint32_t slot = it & 3;
            // every read of the tile first (the compiler cannot move a staging read above an operand-tile store: both
            // are shared memory), then the four row maxima with their shuffles in flight together, then the stores
            float2 x[NPASS][KBX][2];
            float mx[NPASS];
            // (a staging row without a link holds stale bytes: whatever they give stays in that row of the contraction)
            if (__all_sync(0xffffffffu, same_a)) {
#pragma unroll
                for (int q = 0; q < NPASS; ++q) {
                    const int row = rowof(q);
#pragma unroll
                    for (int kb = 0; kb < KBX; ++kb) {
                        const float4 v = lds4(stage, row, kb * 32 + chunk * 4);
                        x[q][kb][0] = mul2(make_float2(v.x, v.y), make_float2(xa[kb].x, xa[kb].y));
                        x[q][kb][1] = mul2(make_float2(v.z, v.w), make_float2(xa[kb].z, xa[kb].w));
                    }
                }
            } else {                                     // a tile with several sources: the source rows one by one (L2)
#pragma unroll
                for (int q = 0; q < NPASS; ++q) {
                    const int row = rowof(q);
#pragma unroll
                    for (int kb = 0; kb < KBX; ++kb) {
                        const float4 m = ld4(ids[slot][0][row], kb * 32 + chunk * 4);
                        const float4 v = lds4(stage, row, kb * 32 + chunk * 4);
                        x[q][kb][0] = mul2(make_float2(v.x, v.y), make_float2(m.x, m.y));
                        x[q][kb][1] = mul2(make_float2(v.z, v.w), make_float2(m.z, m.w));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < NPASS; ++q) {
                mx[q] = 0.f;
#pragma unroll
                for (int kb = 0; kb < KBX; ++kb)
                    mx[q] = fmaxf(fmaxf(mx[q], fmaxf(fabsf(x[q][kb][0].x), fabsf(x[q][kb][0].y))),
                                  fmaxf(fabsf(x[q][kb][1].x), fabsf(x[q][kb][1].y)));
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1)
#pragma unroll
                for (int q = 0; q < NPASS; ++q) mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], o));
#pragma unroll
            for (int q = 0; q < NPASS; ++q) {
                const int row = rowof(q);
                // max in [2^(E-127), 2^(E-126)) -> scale 2^(141-E) puts it in [2^14, 2^15); rows of zeros / denormals: E = 16
                // (an Inf / NaN row — E = 255 — stays Inf / NaN, as in the reference)
                const uint32_t E = max(__float_as_uint(mx[q]) >> 23, 16u);
                const float2 sc = dup2(__uint_as_float((268u - E) << 23));
                if (chunk == 0) s_rscale[slot][row] = __uint_as_float((E - 14u) << 23);
#pragma unroll
                for (int kb = 0; kb < KBX; ++kb) {
                    uint2 hi, lo;
                    split2(mul2(x[q][kb][0], sc), hi.x, lo.x);
                    split2(mul2(x[q][kb][1], sc), hi.y, lo.y);
                    // fp32 element kb*32 + chunk*4 + i = half k: 16-byte chunk k / 8 = kb*4 + chunk/2, byte (k % 8) * 2
                    const uint32_t off = swz_chunk_off(row, kb * 4 + (chunk >> 1)) + (uint32_t)((chunk & 1) << 3);
                    sts_v2(abuf + off, hi.x, hi.y);
                    sts_v2(abuf + A_HALF + off, lo.x, lo.y);
                }
            }
            same_a = same_n;
#pragma unroll
            for (int kb = 0; kb < KBX; ++kb) xa[kb] = xa_n[kb];
            fence_async_smem();
            named_bar_sync(2, kProducers);
            if (ptid == 0) {
                mbar_arrive(&bar_a_ready[it & 1]);
                mbar_arrive(&bar_stage_free[st]);            // every producer has read the slot's rows and ids
            }
            LPF_PSTAMP(2);
            __syncwarp();
            st = st + 1 == kStages ? 0 : st + 1;
        }
#undef LPF_PSTAMP
        return;
    }

    // =============================== consumers: kSplit threads per link (= TMEM lane), 1/kSplit of the columns each
    const float bs2 = p.bs2[0];
    const int row = (warp & 3) * 32 + (tid & 31), part = warp >> 2;
    // The kSplit warps that share a TMEM lane quadrant (32 links) exchange their partial sums among themselves only —
    // a named barrier per quadrant — and announce their share of H / of the accumulator reads on their own (the
    // mbarriers count four arrivals): the four groups drift apart and one group's TMEM reads run under another's
    // arithmetic instead of all eight warps meeting four times per tile.
    const int qbar = 3 + (warp & 3);
    auto xsum = [&](const float (&a)[kSplit][kTileM]) -> float {
        float t = a[0][row];
#pragma unroll
        for (int k = 1; k < kSplit; ++k) t += a[k][row];
        return t;
    };
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t it = 0;
    const bool stamp = p.dbg && blockIdx.x == 0 && tid == 0;
#define LPF_STAMP(k) do { if (stamp && it < 16) p.dbg[it * 16 + (k)] = clock64(); } while (0)

    // epilogue 2 of a tile: prob = sigmoid(ws2 . ReLU(D3 / scale + offset) + bs2)
    auto epilogue2 = [&](int64_t j, uint32_t d3b) {
        constexpr int NH = N3 / kSplit;
        const float* zrow = (ZB && j < n_links) ? p.zb + j * p.ld_zb : nullptr;
        const float2 is3 = dup2(p.inv_s3);
        float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        float buf[2][16];
        const int cb = part * NH;
        tmem_ld16_async(d3b + lane_sel + cb, buf[0]);
        tmem_ld_wait();
#pragma unroll
        for (int c0 = 0; c0 < NH; c0 += 16) {
            const int cur = (c0 >> 4) & 1;
            if (c0 + 16 < NH) tmem_ld16_async(d3b + lane_sel + cb + c0 + 16, buf[cur ^ 1]);
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 o4 = ZB ? (zrow ? __ldg(reinterpret_cast<const float4*>(zrow + cb + c0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f))
                                     : *reinterpret_cast<const float4*>(&s_c3[cb + c0 + c]);
                const float4 w4 = *reinterpret_cast<const float4*>(&s_ws2[cb + c0 + c]);
                acc[0] = fma2(relu2(fma2(make_float2(buf[cur][c], buf[cur][c + 1]), is3, make_float2(o4.x, o4.y))), make_float2(w4.x, w4.y), acc[0]);
                acc[1] = fma2(relu2(fma2(make_float2(buf[cur][c + 2], buf[cur][c + 3]), is3, make_float2(o4.z, o4.w))), make_float2(w4.z, w4.w), acc[1]);
            }
            if (c0 + 16 < NH) tmem_ld_wait();
        }
        s_dot[part][row] = (acc[0].x + acc[0].y) + (acc[1].x + acc[1].y);
        named_bar_sync(qbar, 32 * kSplit);
        if (part == 0 && j < n_links) {
            const float logit = xsum(s_dot) + bs2;
            const int64_t pos = p.idx ? (int64_t)__ldg(p.idx + j) : j;
            p.prob[pos] = p.logits ? logit : 1.0f / (1.0f + expf(-logit));
        }
    };

    int64_t j_prev = -1;
    for (; it < my_tiles; ++it) {
        const int64_t j = tile_of(it) * kTileM + row;
        const uint32_t d1b = d1 + (it & 1) * D;
        LPF_STAMP(0);
        mbar_wait(&bar_mma1[it & 1], (it >> 1) & 1);
        tc_fence_after();
        LPF_STAMP(1);

        // ---- epilogue 1: h = ReLU(LN(D1 / scale + b1)) over the link's row, half of the columns per thread
        constexpr int DH = D / kSplit;
        const int cb = part * DH;
        __align__(8) float v[DH];
        {
            const float2 rs = dup2(s_rscale[it & 3][row] * p.inv_sw1);
            float2 s2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            float (*vb)[16] = reinterpret_cast<float (*)[16]>(v);
            float2* v2 = reinterpret_cast<float2*>(v);
            tmem_ld16_async(d1b + lane_sel + cb, vb[0]);
            tmem_ld_wait();
#pragma unroll
            for (int c0 = 0; c0 < DH; c0 += 16) {
                if (c0 + 16 < DH) tmem_ld16_async(d1b + lane_sel + cb + c0 + 16, vb[(c0 >> 4) + 1]);
#pragma unroll
                for (int c = c0; c < c0 + 16; c += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&s_b1[cb + c]);
                    v2[c / 2] = fma2(v2[c / 2], rs, make_float2(b4.x, b4.y));
                    v2[c / 2 + 1] = fma2(v2[c / 2 + 1], rs, make_float2(b4.z, b4.w));
                    s2[0] = add2(s2[0], v2[c / 2]);
                    s2[1] = add2(s2[1], v2[c / 2 + 1]);
                }
                if (c0 + 16 < DH) tmem_ld_wait();
            }
            s_sum[part][row] = (s2[0].x + s2[0].y) + (s2[1].x + s2[1].y);
            tc_fence_before();
            named_bar_sync(qbar, 32 * kSplit);
            if (part == 0 && (tid & 31) == 0) mbar_arrive(&bar_d1_free[it & 1]);     // contraction 1 of tile it + 2 may overwrite the buffer
            const float mean = xsum(s_sum) * (1.0f / D);
            const float2 nmean = dup2(-mean);
            float2 q2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
            for (int c = 0; c < DH / 2; ++c) {
                const float2 dlt = add2(v2[c], nmean);
                q2[c & 1] = fma2(dlt, dlt, q2[c & 1]);
            }
            s_sq[part][row] = (q2[0].x + q2[0].y) + (q2[1].x + q2[1].y);
            named_bar_sync(qbar, 32 * kSplit);
            const float rstd = rsqrtf(xsum(s_sq) * (1.0f / D) + 1e-5f);
            const float2 rstd2 = dup2(rstd), nmr = dup2(-mean * rstd);     // (v - mean) rstd as one fma
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(&s_g[cb + c]);
                const float4 t4 = *reinterpret_cast<const float4*>(&s_bt[cb + c]);
                v2[c / 2] = relu2(fma2(fma2(v2[c / 2], rstd2, nmr), make_float2(g4.x, g4.y), make_float2(t4.x, t4.y)));
                v2[c / 2 + 1] = relu2(fma2(fma2(v2[c / 2 + 1], rstd2, nmr), make_float2(g4.z, g4.w), make_float2(t4.z, t4.w)));
            }
        }
        LPF_STAMP(2);
        // the H columns are free once contraction 2 of the previous tile has read them (its accumulator is then ready too)
        if (it > 0) {
            mbar_wait(&bar_mma3, (it - 1) & 1);
            tc_fence_after();
        }
        LPF_STAMP(3);
        {
            // (g and b carry the h scale) split into fp16 hi / lo, two consecutive k per 32-bit TMEM column
            uint32_t hi[DH / 2], lo[DH / 2];
#pragma unroll
            for (int c = 0; c < DH / 2; ++c) split2(make_float2(v[2 * c], v[2 * c + 1]), hi[c], lo[c]);
            tmem_st_cols(h_hi + lane_sel + part * (DH / 2), hi);
            tmem_st_cols(h_lo + lane_sel + part * (DH / 2), lo);
        }
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync(qbar, 32 * kSplit);
        LPF_STAMP(4);
        if (part == 0 && (tid & 31) == 0) mbar_arrive(&bar_h_ready);
        LPF_STAMP(5);
        if (it > 0) epilogue2(j_prev, d3 + ((it - 1) & 1) * N3);
        LPF_STAMP(6);
        j_prev = j;
    }
    // drain: epilogue 2 of the last tile
    mbar_wait(&bar_mma3, (it - 1) & 1);
    tc_fence_after();
    epilogue2(j_prev, d3 + ((it - 1) & 1) * N3);
#undef LPF_STAMP

    tc_fence_before();
    named_bar_sync(1, kConsumers);
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
    if (p.dbg && tid == 0) {
        if (blockIdx.x == 0) p.dbg[257] = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(p.dbg[401 + 2 * blockIdx.x]));
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query (the library links no libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// 2-D map of X [n_nodes, D] fp32 (row stride ldx) with a box of one row: a tile::gather4 copy brings four rows
int make_row_map(CUtensorMap* map, const void* X, int64_t ldx, int64_t n_nodes, int d, bool bf16) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        set_error("lpf_link_heads_f16: cuTensorMapEncodeTiled not available from this driver");
        return LPF_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)n_nodes};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * (bf16 ? 2 : 4)};
    const cuuint32_t box[2] = {(cuuint32_t)d, 1}, estr[2] = {1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("lpf_link_heads_f16: cuTensorMapEncodeTiled failed (%d) for X [%lld, %d], ldx %lld", (int)r, (long long)n_nodes, d,
                  (long long)ldx);
        return LPF_ERR_CUDA;
    }
    return LPF_OK;
}

template <int D, bool ZB>
int launch(const Params& p, int64_t n_nodes, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * D * 128 + (size_t)2 * 2 * D * 128 + (size_t)2 * 2 * kTileM * 128 +
                            (size_t)kStages * kTileM * D * 4 + 1024;
    // (the attribute is per device and a process may drive several GPUs: once per device)
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(link_heads_f16_kernel<D, ZB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("lpf_link_heads_f16: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            return LPF_ERR_CUDA;
        }
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    // the tensor map of the node table: rebuilt only when the table changes (one per calling thread)
    struct MapKey { const void* X; int64_t ldx, n; int bf; };
    thread_local MapKey key{nullptr, 0, 0, 0};
    alignas(64) thread_local CUtensorMap xmap;
    if (key.X != p.X || key.ldx != p.ldx || key.n != n_nodes || key.bf != p.x_bf16) {
        if (int rc = make_row_map(&xmap, p.X, p.ldx, n_nodes, D, p.x_bf16 != 0)) return rc;
        key = MapKey{p.X, p.ldx, n_nodes, p.x_bf16};
    }
    const int64_t ntiles = (p.n + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(ntiles < (int64_t)kNumSMs ? ntiles : (int64_t)kNumSMs);
    link_heads_f16_kernel<D, ZB><<<grid, kThreads, smem, st>>>(p, xmap);
    return check_launch("lpf_link_heads_f16");
}

}  // namespace

long long* g_heads_dbg_f16 = nullptr;

}  // namespace lpf

using namespace lpf;

extern "C" int lpf_debug_heads_f16_clocks(void* device_buffer) {
    g_heads_dbg_f16 = (long long*)device_buffer;
    return LPF_OK;
}

extern "C" int64_t lpf_pack_weight_f16_bytes(int32_t N, int32_t K) {
    if (N < 1 || K < 1) return 0;
    const int64_t KB = (K + 63) / 64, NP = tc::round_up(N, 16);
    return KB * 2 * NP * 128;
}

extern "C" int lpf_pack_weight_f16(const float* W, int64_t ldw, int32_t N, int32_t K, float scale, void* packed, void* stream) {
    LPF_REQUIRE(W && packed, "NULL argument");
    LPF_REQUIRE(N >= 1 && K >= 1 && ldw >= K, "bad shape");
    LPF_REQUIRE(scale > 0.f, "scale must be positive");
    LPF_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed image must be 16-byte aligned");
    const int KB = (K + 63) / 64, NP = tc::round_up(N, 16);
    const int64_t total = (int64_t)KB * NP * 64;
    const unsigned grid = (unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
    pack_weight_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, ldw, N, K, NP, KB, scale, (uint8_t*)packed);
    return check_launch("lpf_pack_weight_f16");
}

extern "C" int lpf_link_heads_f16(const int64_t* links, int64_t bs, const int32_t* idx, int64_t n, const void* X,
                                  int x_bf16, int64_t ldx, int64_t n_nodes, int32_t d, const void* w1_packed, float inv_scale_w1, const float* b1,
                                  const float* ln_w_scaled, const float* ln_b_scaled, const void* w23_packed,
                                  float inv_scale_h_w23, const float* c3, const float* zb, int64_t ld_zb,
                                  const float* ws2, const float* bs2, float* prob, int logits, const int64_t* n_dev,
                                  void* stream) {
    LPF_REQUIRE(bs >= 0 && n >= 0 && n_nodes >= 1, "bad size");
    if (n == 0) return LPF_OK;
    LPF_REQUIRE(links && X && w1_packed && b1 && ln_w_scaled && ln_b_scaled && w23_packed && ws2 && bs2 && prob, "NULL argument");
    LPF_REQUIRE(c3 || zb, "either the constant c3 or per-row zb must be given");
    LPF_REQUIRE(idx || n == bs, "n must equal bs when idx is NULL");
    LPF_REQUIRE(ldx >= d && (!zb || ld_zb >= 2 * d), "leading dimension too small");
    LPF_REQUIRE(!zb || ((reinterpret_cast<uintptr_t>(zb) & 15) == 0 && ld_zb % 4 == 0), "zb rows must be 16-byte aligned");
    LPF_REQUIRE(inv_scale_w1 > 0.f && inv_scale_h_w23 > 0.f, "scales must be positive");
    LPF_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx * (x_bf16 ? 2 : 4)) % 16 == 0, "X rows must be 16-byte aligned (TMA)");
    if (d != 64) {
        set_error("lpf_link_heads_f16: d = %d not supported (64)", d);
        return LPF_ERR_UNSUPPORTED;
    }
    Params p{links, bs, idx, n, X, ldx, x_bf16 ? 1 : 0, w1_packed, b1, ln_w_scaled, ln_b_scaled, w23_packed, c3, zb, ld_zb, ws2, bs2, prob,
             logits, n_dev, inv_scale_w1, inv_scale_h_w23, g_heads_dbg_f16};
    cudaStream_t st = (cudaStream_t)stream;
    return zb ? launch<64, true>(p, n_nodes, st) : launch<64, false>(p, n_nodes, st);
}
