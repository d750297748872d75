// Fused per-link heads on the tensor cores: for a 128-link tile,
//     xprod = X[a] * X[b]                                   (gather, reference train/testing.py:29,113)
//     h     = ReLU(LayerNorm(W1 xprod + b1))                (elementwise_lin, models/other_models.py:125-133)
//     el    = W2 h + b2                                     (                 :135)
//     z     = ReLU(Ws1[:, :d] el + bs1 + Ws1[:, d:] pw)     (mlp_score layer 0 on [el | pw], :173-177)
//     prob  = sigmoid(ws2 . z + bs2)                        (:178-179)
// never leaving the SM.  There is no non-linearity between el and mlp_score's first Linear, so the two are
// folded once per weight set: z = ReLU(W23 h + offset), W23 = Ws1[:, :d] W2, offset = Ws1[:, :d] b2 + bs1 +
// Ws1[:, d:] pw.  The two remaining contractions run as tcgen05.mma (kind::tf32, 3xTF32 split, accumulators in
// TMEM); their weights (pre-split, pre-swizzled images from lpf_pack_weight) are brought into shared memory
// ONCE per CTA by bulk TMA copies and stay resident while the CTA walks its tiles; two threads own one link
// (= one TMEM lane, half of the accumulator columns each), so LayerNorm and the final dot product take one
// shared-memory exchange between the two.
//
// Software pipeline across tiles: warps 8-15 (producers) gather the X[a]*X[b] rows of a tile in two halves whose
// reads and stores alternate — while one half is split into hi / lo and stored in the UMMA layout, the reads of the
// other half (and of the next tile) are in flight; link ids run three tiles ahead in a shared-memory ring and the rows
// are pulled into L2 two tiles ahead.  As soon as contraction 2 of tile t has finished reading the operand tile the
// MMA warp issues contraction 1 of tile t+1, which runs on the tensor pipe while warps 0-7 (consumers) are still in
// tile t's final epilogue.  Hand-offs are mbarriers fed by tcgen05.commit.  Measured (tools/heads_clocks.py, sustained
// load: the SM clock settles at ~1.3 GHz under the power cap): ~5.3 k cycles per tile, of which the two contractions
// (fed from shared memory, 3xTF32) hold the tensor pipe ~3.5 k and the consumers' epilogues ~2.9 k; the gather is no
// longer on the critical path (removing its reads altogether shortens the period by 10 %).  Per-row bulk TMA copies
// were tried for the gather and are slower (128 small copies per tile: ~5 k cycles from issue to completion).
//
// `offset` carries the pairwise half of the concatenated feature vector [el | pw] (models/link_transformer.py:105,
// train/testing.py:31): for a link whose selected node sets are all empty, pw is the same vector for every link
// (pairwise_lin(LayerNorm(att.bias) | 0...)), so the offset is the constant c3; links with non-empty sets
// are scored by a second call over the compacted list (`idx`) with per-row offsets zb.
//
// Supported: d in {32, 64} (weights 24 / 96 KB resident); other widths use the unfused kernels.
#include "tc.cuh"

namespace lpf {

using namespace tc;

constexpr int kHeadConsumers = 256;   // warps 0-7: TWO threads per link (= TMEM lane), half of the accumulator columns each: warp w
                                      // reads lanes 32 (w % 4) .. + 31 (the hardware's lane quadrant of a warp), column half w / 4.
                                      // The epilogues are CUDA-core work on 192 accumulator columns per link, and one warp per SM
                                      // sub-partition cannot hide its own instruction latencies; thread 0 issues the MMAs
constexpr int kHeadProducers = 256;   // warps 8-15: gather X[a]*X[b] of the NEXT tile while this one is computed
constexpr int kHeadThreads = kHeadConsumers + kHeadProducers + 32;   // + warp 16: one thread issues every MMA

struct HeadsParams {
    const int64_t* links;
    int64_t bs;
    const int32_t* idx;
    int64_t n;
    const float* X;
    int64_t ldx;
    const float* w1p;    // packed elementwise_lin.linears[0].weight            [d, d]
    const float* b1;
    const float* ln_g;
    const float* ln_b;
    const float* w23p;   // packed Ws1[:, :d] . elementwise_lin.linears[1].weight [2d, d]
    const float* c3;     // [2d] constant pre-activation offset (see header) or NULL
    const float* zb;     // [n, 2d] per-row offset or NULL
    int64_t ld_zb;
    const float* ws2;
    const float* bs2;
    float* prob;
    int logits;
    const int64_t* n_dev;   // optional device-side n
    int32_t* sched;         // optional {next tile, finished CTAs}: tiles handed out dynamically (see lpf_link_heads_tc)
    long long* dbg;         // optional: per-phase clock64 stamps of CTA 0 / thread 0 (lpf_debug_heads_clocks)
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int D, bool ZB>
__global__ void __launch_bounds__(kHeadThreads, 1) link_heads_tc_kernel(HeadsParams p) {
    constexpr int KB = D / 32;                       // k-blocks of both contractions (K = D)
    constexpr int N3 = 2 * D;                        // width of mlp_score's hidden layer
    constexpr uint32_t W1_BYTES = KB * 2 * D * 128;  // packed [D, D]
    constexpr uint32_t W3_BYTES = KB * 2 * N3 * 128; // packed [2D, D]
    constexpr uint32_t TMEM_COLS = 8 * D;            // D1 x 2 | D3 x 2 | H hi | H lo  (512 columns at D = 64: all of TMEM)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // bar_w: weights landed;  bar_mma1[b]: contraction 1 into accumulator buffer b complete (the operand tile in shared
    // memory may be overwritten);  bar_d1_free[b]: the consumers have read accumulator buffer b;  bar_mma3: contraction 2
    // of the current tile complete (its accumulator is ready, the H operand in TMEM may be overwritten).
    // One barrier per BUFFER where the waiting side can fall two completions behind (a parity wait cannot tell them apart).
    // bar_a_ready: the producers have stored the next operand tile;  bar_h_ready: the consumers have stored H.
    __shared__ uint64_t bar_w, bar_mma1[2], bar_d1_free[2], bar_mma3, bar_a_ready, bar_h_ready;
    __shared__ uint32_t tmem_slot;
    __shared__ int32_t s_tile[8];                    // ring of this CTA's tile numbers (-1: no more), written 3 ahead
    __shared__ int32_t ids[4][2][kTileM];            // ring over tiles: [tile & 3][a, b][row], written three tiles ahead
    __shared__ __align__(16) float s_b1[D], s_g[D], s_bt[D], s_c3[N3], s_ws2[N3];
    __shared__ float s_sum[2][kTileM], s_sq[2][kTileM], s_dot[2][kTileM];   // exchanges between the two threads of a link

    const int tid = threadIdx.x, warp = tid >> 5;
    if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[256] = clock64();       // (profiling: kernel start / end of CTA 0)
    if (p.n_dev) p.n = min(p.n, *p.n_dev);
    const int64_t ntiles = (p.n + kTileM - 1) / kTileM;
    // Tile k of this CTA: blockIdx.x + k gridDim.x, or — with p.sched — the next tile nobody has taken yet.  The CTAs of
    // this kernel start whenever an SM has all of its registers free, i.e. at very different times when another
    // batch's selection is still resident; with a static assignment the last CTA to start sets the duration.
    auto fetch_tile = [&](uint32_t k) -> int32_t {       // (one thread of the CTA calls this, in sequence)
        const int64_t t = p.sched ? (int64_t)atomicAdd(p.sched, 1) : (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
        return t < ntiles ? (int32_t)t : -1;
    };
    auto retire = [&]() {                                // the last CTA to finish re-arms the scheduler words
        if (p.sched && atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
        }
    };
    if (tid == 0) {
        for (uint32_t k = 0; k < 3; ++k) s_tile[k] = fetch_tile(k);
    }
    __syncthreads();
    if (s_tile[0] < 0) {                             // nothing to do: before any barrier / TMEM / bulk-copy state
        if (tid == 0) retire();
        return;
    }
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW1 = smem;
    uint8_t* sW3 = sW1 + W1_BYTES;
    uint8_t* sA = sW3 + W3_BYTES;                    // [kb][hi, lo][128 rows][128 B]

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma1[0], 1);
        mbar_init(&bar_mma1[1], 1);
        mbar_init(&bar_d1_free[0], 1);
        mbar_init(&bar_d1_free[1], 1);
        mbar_init(&bar_mma3, 1);
        mbar_init(&bar_a_ready, 1);
        mbar_init(&bar_h_ready, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar_w, W1_BYTES + W3_BYTES);
        bulk_g2s(sW1, p.w1p, W1_BYTES, &bar_w);
        bulk_g2s(sW3, p.w23p, W3_BYTES, &bar_w);
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
    for (int c = tid; c < D; c += kHeadThreads) {
        s_b1[c] = p.b1[c];
        s_g[c] = p.ln_g[c];
        s_bt[c] = p.ln_b[c];
    }
    for (int c = tid; c < N3; c += kHeadThreads) {
        s_c3[c] = p.c3 ? p.c3[c] : 0.f;
        s_ws2[c] = p.ws2[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const uint32_t tmem_d = tmem_slot;
    const uint32_t d1 = tmem_d, d3 = tmem_d + 2 * D, h_hi = tmem_d + 6 * D, h_lo = tmem_d + 7 * D;
    const uint32_t idesc_d = make_idesc_tf32(kTileM, D), idesc_3 = make_idesc_tf32(kTileM, N3);
    const uint32_t aA = smem_u32(sA), aW1 = smem_u32(sW1), aW3 = smem_u32(sW3);

    if (warp == (kHeadConsumers + kHeadProducers) / 32) {
        // =========================== the MMA thread: contraction 1 of tile it + 1, then contraction 2 of tile it —
        // the order in which their operands become ready when the tensor pipe is the bottleneck (neither the
        // consumers nor the producers ever stall behind a full MMA queue)
        // (the whole warp runs this loop; one elected lane issues: the descriptors are warp-uniform values)
        {
            const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_d, 0);
            const uint32_t u_d1 = tbase, u_d3 = tbase + 2 * D, u_hhi = tbase + 6 * D, u_hlo = tbase + 7 * D;
            mbar_wait(&bar_w, 0);
            auto mma1 = [&](uint32_t it) {
                mbar_wait(&bar_a_ready, it & 1);
                if (it >= 2) mbar_wait(&bar_d1_free[it & 1], ((it >> 1) - 1) & 1);     // epilogue 1 of tile it - 2 has read this buffer
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                        issue_kblock_3x(u_d1 + (it & 1) * D, aA + kb * 2 * kATileBytes, aA + kb * 2 * kATileBytes + kATileBytes,
                                        aW1 + kb * 2 * D * 128, aW1 + kb * 2 * D * 128 + D * 128, idesc_d, kb == 0);
                    umma_commit(&bar_mma1[it & 1]);
                }
                __syncwarp();
            };
            const bool mstamp = p.dbg && blockIdx.x == 0 && (tid & 31) == 0;
#define LPF_MSTAMP(k) do { if (mstamp && it < 16) p.dbg[it * 16 + 8 + (k)] = clock64(); } while (0)
            mma1(0);
            for (uint32_t it = 0; s_tile[it & 7] >= 0; ++it) {
                LPF_MSTAMP(0);
                if (s_tile[(it + 1) & 7] >= 0) {
                    mbar_wait(&bar_a_ready, (it + 1) & 1);
                    LPF_MSTAMP(1);
                    mma1(it + 1);
                }
                LPF_MSTAMP(2);
                mbar_wait(&bar_h_ready, it & 1);
                LPF_MSTAMP(3);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                        issue_kblock_3x_ts(u_d3 + (it & 1) * N3, u_hhi + kb * 32, u_hlo + kb * 32, aW3 + kb * 2 * N3 * 128,
                                           aW3 + kb * 2 * N3 * 128 + N3 * 128, idesc_3, kb == 0);
                    umma_commit(&bar_mma3);
                }
                __syncwarp();
                LPF_MSTAMP(4);
            }
#undef LPF_MSTAMP
        }
        return;
    }
    if (warp >= kHeadConsumers / 32) {
        // =========================== producers: gather X[a]*X[b] of the tiles, split it into hi / lo and store it in the
        // UMMA layout.  Measured: the gather of one tile (128 rows of 4D bytes through four warps' load-store path)
        // streams at ~10 B/clk and takes longer than everything else in the loop, so it must never pause: the tile is
        // handled in two HALVES (rows 0-63, 64-127) whose reads and stores alternate — while one half is split and
        // stored, the reads of the other are in flight (the same 64 data registers as one whole tile).
        constexpr int RPP = kHeadProducers / 8;          // rows per pass (eight threads per row)
        constexpr int HP = kTileM / RPP / 2;             // passes per half tile
        const int ptid = tid - kHeadConsumers;
        const int chunk = ptid & 7, row_in_pass = ptid >> 3;
        const bool vec_x = ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) && (p.ldx % 4 == 0);
        uint32_t it = 0;
        auto load_ids = [&](int32_t tile, int32_t& a, int32_t& b) {
            const int64_t j = (int64_t)tile * kTileM + ptid;
            a = 0; b = 0;
            if (tile >= 0 && j < p.n && ptid < kTileM) {
                const int64_t pos = p.idx ? (int64_t)__ldg(p.idx + j) : j;
                a = (int32_t)__ldg(p.links + pos);
                b = (int32_t)__ldg(p.links + p.bs + pos);
            }
        };
        auto prefetch_rows = [&](int32_t a, int32_t b) {
            const char* ra = reinterpret_cast<const char*>(p.X + (int64_t)a * p.ldx);
            const char* rb = reinterpret_cast<const char*>(p.X + (int64_t)b * p.ldx);
#pragma unroll
            for (int o = 0; o < D * 4; o += 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ra + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rb + o));
            }
        };
        // this thread's X[b] row of a tile into L1, without a register or a scoreboard entry (the four warps' loads alone
        // keep too few requests in flight to cover the L2 latency)
        auto prefetch_l1 = [&](int32_t b) {
            const char* rb = reinterpret_cast<const char*>(p.X + (int64_t)b * p.ldx);
#pragma unroll
            for (int o = 0; o < D * 4; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(rb + o));
        };
        auto ld4 = [&](const float* q) -> float4 {
            if (vec_x) return __ldg(reinterpret_cast<const float4*>(q));
            return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
        };
        // the X[b] reads of half h (passes 4h .. 4h+3) of the tile in ring slot `slot`
        auto load_half = [&](uint32_t slot, int h, float4 (&v)[KB][HP]) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int q = 0; q < HP; ++q)
                    v[kb][q] = ld4(p.X + (int64_t)ids[slot][1][(HP * h + q) * RPP + row_in_pass] * p.ldx + kb * 32 + chunk * 4);
        };
        // X[a] * X[b], hi / lo split, store of half h.  In an evaluation batch the links of a tile share their source
        // (train/testing.py:20-23): then one X[a] read per k-block (xa, read a tile ahead) serves the thread's rows.
        auto store_half = [&](uint32_t slot, int h, float4 (&v)[KB][HP], bool same_a, const float4 (&xa)[KB]) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int q = 0; q < HP; ++q) {
                    const int row = (HP * h + q) * RPP + row_in_pass;
                    float4 m = xa[kb];
                    if (!same_a) m = ld4(p.X + (int64_t)ids[slot][0][row] * p.ldx + kb * 32 + chunk * 4);
                    const float4 x = make_float4(v[kb][q].x * m.x, v[kb][q].y * m.y, v[kb][q].z * m.z, v[kb][q].w * m.w);
                    const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                    const float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
                    const uint32_t off = (uint32_t)kb * 2 * kATileBytes + swz_chunk_off(row, chunk);
                    *reinterpret_cast<float4*>(sA + off) = hi;
                    *reinterpret_cast<float4*>(sA + off + kATileBytes) = lo;
                }
        };
        // does every row of this thread in the tile of ring slot `slot` share its source?  (and that source's X row)
        auto tile_source = [&](uint32_t slot, bool& same_a, float4 (&xa)[KB]) {
            const int32_t a0 = ids[slot][0][row_in_pass];
            same_a = true;
#pragma unroll
            for (int pass = 1; pass < 2 * HP; ++pass) same_a &= ids[slot][0][pass * RPP + row_in_pass] == a0;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) xa[kb] = ld4(p.X + (int64_t)a0 * p.ldx + kb * 32 + chunk * 4);
        };
        // The link ids of a tile come from DRAM (the batch's links are streamed once) and its rows from L2 if they were
        // prefetched: the ids run three tiles ahead in a shared-memory ring and the rows are pulled into L2 two tiles ahead.
        {
            int32_t a, b;
            for (int k = 0; k < 3; ++k) {
                load_ids(s_tile[k], a, b);
                if (ptid < kTileM) { ids[k][0][ptid] = a; ids[k][1][ptid] = b; }
            }
        }
        named_bar_sync(2, kHeadProducers);
        if (s_tile[1] >= 0 && ptid < kTileM) prefetch_rows(ids[1][0][ptid], ids[1][1][ptid]);
        float4 v0[KB][HP], v1[KB][HP], xa[KB];
        bool same_a;
        tile_source(0, same_a, xa);
        load_half(0, 0, v0);
        load_half(0, 1, v1);
        const bool pstamp = p.dbg && blockIdx.x == 0 && ptid == 0;
#define LPF_PSTAMP(k) do { if (pstamp && it < 16) p.dbg[it * 16 + 13 + (k)] = clock64(); } while (0)
        for (; s_tile[it & 7] >= 0; ++it) {
            LPF_PSTAMP(0);
            if (ptid == 0) s_tile[(it + 3) & 7] = fetch_tile(it + 3);
            if (s_tile[(it + 2) & 7] >= 0 && ptid < kTileM) prefetch_rows(ids[(it + 2) & 3][0][ptid], ids[(it + 2) & 3][1][ptid]);
            named_bar_sync(2, kHeadProducers);
            int32_t a_n3, b_n3;
            load_ids(s_tile[(it + 3) & 7], a_n3, b_n3);      // (first used at the end of the iteration)
            const bool more = s_tile[(it + 1) & 7] >= 0;
            // the next tile's source row (a tile ahead: an L2 hit that is not waited for until the next iteration)
            bool same_n = true;
            float4 xa_n[KB];
            if (more) tile_source((it + 1) & 3, same_n, xa_n);
            if (it > 0) mbar_wait(&bar_mma1[(it - 1) & 1], ((it - 1) >> 1) & 1);     // contraction 1 of the previous tile has read the operand tile
            LPF_PSTAMP(1);
            store_half(it & 3, 0, v0, same_a, xa);
            if (more) load_half((it + 1) & 3, 0, v0);
            store_half(it & 3, 1, v1, same_a, xa);
            if (more) load_half((it + 1) & 3, 1, v1);
            same_a = same_n;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) xa[kb] = xa_n[kb];
            fence_async_smem();
            // (first use of this iteration's id reads; ring slot (it + 3) & 3 held tile it - 1, gathered an iteration ago)
            if (ptid < kTileM) { ids[(it + 3) & 3][0][ptid] = a_n3; ids[(it + 3) & 3][1][ptid] = b_n3; }
            named_bar_sync(2, kHeadProducers);
            if (ptid == 0) mbar_arrive(&bar_a_ready);
            LPF_PSTAMP(2);
            __syncwarp();
        }
        return;
    }

    // =============================== consumers: one link (= TMEM lane) per thread
    const float bs2 = p.bs2[0];
    const int row = (warp & 3) * 32 + (tid & 31), half = warp >> 2;      // this thread's link of the tile, its column half
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t it = 0;
    const bool stamp = p.dbg && blockIdx.x == 0 && tid == 0;
#define LPF_STAMP(k) do { if (stamp && it < 16) p.dbg[it * 16 + (k)] = clock64(); } while (0)

    // epilogue 2 of a tile: prob = sigmoid(ws2 . ReLU(D3 + offset) + bs2)
    auto epilogue2 = [&](int64_t j, uint32_t d3b) {
        constexpr int NH = N3 / 2;                       // columns per thread
        const float* zrow = (ZB && j < p.n) ? p.zb + j * p.ld_zb : nullptr;
        float acc4[4] = {0.f, 0.f, 0.f, 0.f};
        // the next 16 columns are on their way out of TMEM while these 16 are reduced
        float buf[2][16];
        const int cb = half * NH;
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
                acc4[0] = fmaf(fmaxf(buf[cur][c] + o4.x, 0.f), w4.x, acc4[0]);
                acc4[1] = fmaf(fmaxf(buf[cur][c + 1] + o4.y, 0.f), w4.y, acc4[1]);
                acc4[2] = fmaf(fmaxf(buf[cur][c + 2] + o4.z, 0.f), w4.z, acc4[2]);
                acc4[3] = fmaf(fmaxf(buf[cur][c + 3] + o4.w, 0.f), w4.w, acc4[3]);
            }
            if (c0 + 16 < NH) tmem_ld_wait();
        }
        s_dot[half][row] = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
        named_bar_sync(1, kHeadConsumers);
        if (half == 0 && j < p.n) {
            const float logit = (s_dot[0][row] + s_dot[1][row]) + bs2;
            const int64_t pos = p.idx ? (int64_t)__ldg(p.idx + j) : j;
            p.prob[pos] = p.logits ? logit : 1.0f / (1.0f + expf(-logit));
        }
    };

    int64_t j_prev = -1;
    for (; s_tile[it & 7] >= 0; ++it) {
        const int64_t j = (int64_t)s_tile[it & 7] * kTileM + row;
        const uint32_t d1b = d1 + (it & 1) * D;
        LPF_STAMP(0);
        mbar_wait(&bar_mma1[it & 1], (it >> 1) & 1);
        tc_fence_after();
        LPF_STAMP(1);

        // ---- epilogue 1: h = ReLU(LN(D1 + b1)) over the link's row, half of the columns per thread (the two partial
        // sums of the mean and of the variance meet in shared memory).  h is the A operand of contraction 2 and goes
        // back into TMEM (hi / lo halves of the split), not through shared memory.
        constexpr int DH = D / 2;
        const int cb = half * DH;
        float v[DH];
        {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            float (*vb)[16] = reinterpret_cast<float (*)[16]>(v);
            tmem_ld16_async(d1b + lane_sel + cb, vb[0]);
            tmem_ld_wait();
#pragma unroll
            for (int c0 = 0; c0 < DH; c0 += 16) {
                if (c0 + 16 < DH) tmem_ld16_async(d1b + lane_sel + cb + c0 + 16, vb[(c0 >> 4) + 1]);
#pragma unroll
                for (int c = c0; c < c0 + 16; c += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&s_b1[cb + c]);
                    v[c] += b4.x; v[c + 1] += b4.y; v[c + 2] += b4.z; v[c + 3] += b4.w;
                    s4[0] += v[c]; s4[1] += v[c + 1]; s4[2] += v[c + 2]; s4[3] += v[c + 3];
                }
                if (c0 + 16 < DH) tmem_ld_wait();
            }
            s_sum[half][row] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            tc_fence_before();                              // (this thread's reads of the accumulator are complete)
            named_bar_sync(1, kHeadConsumers);
            if (tid == 0) mbar_arrive(&bar_d1_free[it & 1]);     // contraction 1 of tile it + 2 may overwrite the buffer
            const float mean = (s_sum[0][row] + s_sum[1][row]) * (1.0f / D);
            float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < DH; ++c) {
                const float dlt = v[c] - mean;
                q4[c & 3] = fmaf(dlt, dlt, q4[c & 3]);
            }
            s_sq[half][row] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
            named_bar_sync(1, kHeadConsumers);
            const float rstd = rsqrtf((s_sq[0][row] + s_sq[1][row]) * (1.0f / D) + 1e-5f);
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(&s_g[cb + c]);
                const float4 t4 = *reinterpret_cast<const float4*>(&s_bt[cb + c]);
                v[c] = fmaxf(fmaf((v[c] - mean) * rstd, g4.x, t4.x), 0.f);
                v[c + 1] = fmaxf(fmaf((v[c + 1] - mean) * rstd, g4.y, t4.y), 0.f);
                v[c + 2] = fmaxf(fmaf((v[c + 2] - mean) * rstd, g4.z, t4.z), 0.f);
                v[c + 3] = fmaxf(fmaf((v[c + 3] - mean) * rstd, g4.w, t4.w), 0.f);
            }
        }
        LPF_STAMP(2);
        // the H columns are free once contraction 2 of the previous tile has read them (its accumulator is then ready too)
        if (it > 0) {
            mbar_wait(&bar_mma3, (it - 1) & 1);
            tc_fence_after();
        }
        LPF_STAMP(3);
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 16) {
            float hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                hi[c] = tf32_hi(v[c0 + c]);
                lo[c] = v[c0 + c] - hi[c];
            }
            tmem_st16(h_hi + lane_sel + cb + c0, hi);
            tmem_st16(h_lo + lane_sel + cb + c0, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        named_bar_sync(1, kHeadConsumers);
        LPF_STAMP(4);

        // ---- contraction 2: D3 = h . (Ws1[:, :d] W2)^T, A from TMEM (issued by the MMA thread); it runs while the
        // consumers reduce the previous tile's accumulator (epilogue 2), and contraction 1 of the next tile does not
        // wait for it
        if (tid == 0) mbar_arrive(&bar_h_ready);
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
    named_bar_sync(1, kHeadConsumers);
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
    if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[257] = clock64();
    if (tid == 0) retire();
}

template <int D, bool ZB>
static int launch_heads(const HeadsParams& p, cudaStream_t st) {
    constexpr int KB = D / 32;
    constexpr size_t smem = (size_t)KB * 2 * D * 128 + (size_t)KB * 2 * 2 * D * 128 + (size_t)KB * 2 * tc::kATileBytes + 1024;
    // (the attribute is per device: set on every call — a process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(link_heads_tc_kernel<D, ZB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("lpf_link_heads_tc: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
        return LPF_ERR_CUDA;
    }
    const int64_t ntiles = (p.n + tc::kTileM - 1) / tc::kTileM;
    const int per_sm = D <= 32 ? 2 : 1;
    const unsigned grid = (unsigned)(ntiles < (int64_t)kNumSMs * per_sm ? ntiles : (int64_t)kNumSMs * per_sm);
    link_heads_tc_kernel<D, ZB><<<grid, kHeadThreads, smem, st>>>(p);
    return check_launch("lpf_link_heads_tc");
}

}  // namespace lpf

using namespace lpf;

static long long* g_heads_dbg = nullptr;
// Debug hook (not part of the data path): device buffer of 16 x 16 + 16 int64 that CTA 0 of every later
// lpf_link_heads_tc launch fills with per-phase clock64() stamps of its first 16 tiles (+ kernel start / end at
// [256], [257]); NULL switches it off.
extern "C" int lpf_debug_heads_clocks(void* device_buffer) {
    g_heads_dbg = (long long*)device_buffer;
    return LPF_OK;
}

extern "C" int lpf_link_heads_tc(const int64_t* links, int64_t bs, const int32_t* idx, int64_t n, const float* X,
                                 int64_t ldx, int32_t d, const float* w1_packed, const float* b1, const float* ln_w,
                                 const float* ln_b, const float* w23_packed, const float* c3, const float* zb,
                                 int64_t ld_zb, const float* ws2, const float* bs2, float* prob, int logits,
                                 const int64_t* n_dev, int32_t* tile_sched, void* stream) {
    LPF_REQUIRE(bs >= 0 && n >= 0, "negative size");
    if (n == 0) return LPF_OK;
    LPF_REQUIRE(links && X && w1_packed && b1 && ln_w && ln_b && w23_packed && ws2 && bs2 && prob, "NULL argument");
    LPF_REQUIRE(c3 || zb, "either the constant c3 or per-row zb must be given");
    LPF_REQUIRE(idx || n == bs, "n must equal bs when idx is NULL");
    LPF_REQUIRE(ldx >= d && (!zb || ld_zb >= 2 * d), "leading dimension too small");
    HeadsParams p{links, bs, idx, n, X, ldx, w1_packed, b1, ln_w, ln_b, w23_packed, c3, zb, ld_zb, ws2, bs2, prob, logits, n_dev, tile_sched, g_heads_dbg};
    cudaStream_t st = (cudaStream_t)stream;
    LPF_REQUIRE(!zb || ((reinterpret_cast<uintptr_t>(zb) & 15) == 0 && ld_zb % 4 == 0), "zb rows must be 16-byte aligned");
    if (d == 64) return zb ? launch_heads<64, true>(p, st) : launch_heads<64, false>(p, st);
    if (d == 32) return zb ? launch_heads<32, true>(p, st) : launch_heads<32, false>(p, st);
    set_error("lpf_link_heads_tc: d = %d not supported by the fused kernel (32 or 64)", d);
    return LPF_ERR_UNSUPPORTED;
}
