// Fused path for the links that selected at least one node, small-batch regime (citation2-shaped evaluation: ~1 %
// of the links of a batch): one warp carries one link from its node sets to its score —
//     Q = lin_l(X[a]) + lin_l(X[b])                                           (modules/layers.py:208-214)
//     per selected pair: RPE MLP + its contraction, v = KV[u] + r, score, online segment softmax
//                                                                             (models/link_transformer.py:182-211,
//                                                                              modules/layers.py:193-224)
//     + bias, LayerNorm, counts, pairwise_lin                                 (:340-386, :177)
//     elementwise_lin(X[a]*X[b]), mlp_score on [el | pw], sigmoid             (models/other_models.py:125-138,173-179)
// in two launches (one warp per selected pair for the RPE contraction, then one warp per link), in fp32 FFMA
// with the vectors distributed over the lanes (adjacent channels per lane) and every weight — the TRANSPOSE W^T[k][n] —
// staged ONCE per CTA in shared memory (134 KB of matrices for the link stage at d = 64: one CTA of 24 warps per SM;
// 48 KB for the pair stage).  It replaces ~14 launches of the batched
// path (gather, 7 tensor-core contractions on a few thousand rows, 3 RPE launches, attention, LayerNorm, heads)
// whose cost at this size is launch latency; the batched path remains for batches where most links are
// non-empty (the plan picks by the fraction observed on the previous batch — both give the same numbers).
#include "common.cuh"

namespace lpf {

struct NzParams {
    const int64_t* links;
    int64_t bs;
    const int32_t* nz;
    int64_t n_cap;
    const int64_t* n_dev;
    const float* X;
    int64_t ldx;
    const float* KV;
    int64_t ld_kv;
    const int32_t* node;
    const float* pa;
    const float* pb;
    const int32_t* seg_start;
    const int32_t* counts;
    int64_t cap;
    const int64_t* hdr;   // [0..2] pairs per type (device side)
    float* R;             // [3*cap, D] RPE contraction per pair (written by nz_pairs_kernel)
    int mode, ntypes, cd;
    const float* wlT;
    const float* bl;
    const float* rpe_w1[3];
    const float* rpe_b1[3];
    const float* rpe_g[3];
    const float* rpe_b[3];
    const float* rpe_mT[3];
    const float* rpe_c[3];
    const float* att;
    const float* att_bias;
    const float* pn_w;
    const float* pn_b;
    const float* p1T;
    const float* pb1;
    const float* pln_w;
    const float* pln_b;
    const float* p2T;
    const float* pb2;
    const float* wzT;
    const float* off;
    const float* w1T;
    const float* b1;
    const float* ln_g;
    const float* ln_b;
    const float* w23T;
    const float* ws2;
    const float* bs2;
    float* prob;
    int logits;
    int tab_bf16;         // X and KV hold bf16 (ldx / ld_kv in elements)
};
// element i of a node table (fp32, or bf16 widened)
__device__ __forceinline__ float ld_tab(const float* base, int64_t i, int bf) {
    if (bf) return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const uint16_t*>(base) + i) << 16);
    return __ldg(base + i);
}

// Channel layout: a D-wide vector is spread over the lanes with KC = D / 32 ADJACENT channels per lane
// (channel KC * lane + k), a 2D-wide one with 2 KC adjacent channels per lane — so that a lane's share of a weight row
// is one 64- or 128-bit shared-memory read and one or two packed FFMA2.
template <int KC>
__device__ __forceinline__ int ch(int lane, int k) { return KC * lane + k; }

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// y[j] += sum_c x[c] * WT[c][KO * lane + j]: WT [D][ldw] in SHARED memory, x (D floats) in this warp's shared-memory
// buffer (read as 128-bit broadcasts: one read per four rows), KO outputs per lane.  Per weight row and lane: one
// shared-memory read + KO / 2 packed FMAs (the earlier form — a shuffle per row to broadcast x, channel lane + 32 j:
// KO 32-bit reads — was bound by the shared-memory / shuffle pipe, and read from L2 it paid a round trip per block of rows).
template <int D, int KO>
__device__ __forceinline__ void matvec_s(const float* WT, int ldw, const float* xs, float (&y)[KO], int lane) {
    const float* col = WT + KO * lane;
#pragma unroll 4
    for (int c4 = 0; c4 < D / 4; ++c4) {
        const float4 x4 = *reinterpret_cast<const float4*>(xs + 4 * c4);
        const float xr[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float* row = col + (4 * c4 + r) * ldw;
            if constexpr (KO == 1) {
                y[0] = fmaf(xr[r], row[0], y[0]);
            } else if constexpr (KO == 2) {
                const float2 w = *reinterpret_cast<const float2*>(row);
                const float2 t = fma2(make_float2(xr[r], xr[r]), w, make_float2(y[0], y[1]));
                y[0] = t.x; y[1] = t.y;
            } else {
                static_assert(KO == 4, "1, 2 or 4 outputs per lane");
                const float4 w = *reinterpret_cast<const float4*>(row);
                const float2 t0 = fma2(make_float2(xr[r], xr[r]), make_float2(w.x, w.y), make_float2(y[0], y[1]));
                const float2 t1 = fma2(make_float2(xr[r], xr[r]), make_float2(w.z, w.w), make_float2(y[2], y[3]));
                y[0] = t0.x; y[1] = t0.y; y[2] = t1.x; y[3] = t1.y;
            }
        }
    }
}
// The (at most four) count channels that follow the D main input channels: rows 0 .. cd - 1 of WT (shared memory), their
// inputs in lanes 0 .. cd - 1 of `xc`.
template <int KO>
__device__ __forceinline__ void matvec_counts_s(const float* WT, int ldw, float xc, int cd, float (&y)[KO], int lane) {
    for (int e = 0; e < cd; ++e) {
        const float* row = WT + e * ldw + KO * lane;
        const float xv = __shfl_sync(kFull, xc, e);
#pragma unroll
        for (int j = 0; j < KO; ++j) y[j] = fmaf(xv, row[j], y[j]);
    }
}
// this lane's channels of a D-wide vector into the warp's shared-memory buffer (read back by matvec_s)
template <int KC>
__device__ __forceinline__ void put_vec(float* xs, const float (&x)[KC], int lane) {
    __syncwarp();                // (the previous product has read the buffer)
#pragma unroll
    for (int k = 0; k < KC; ++k) xs[ch<KC>(lane, k)] = x[k];
    __syncwarp();
}

// cooperative copy of n floats into shared memory (dst 16-byte aligned; 128-bit reads when src is aligned too)
__device__ __forceinline__ void stage_floats(float* dst, const float* __restrict__ src, int n) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = n & ~3;
        for (int i = threadIdx.x * 4; i < n4; i += blockDim.x * 4)
            *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
        for (int i = n4 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}
// ... of a [rows][cols] matrix into rows of ld_dst floats
__device__ __forceinline__ void stage_rows(float* dst, int ld_dst, const float* __restrict__ src, int rows, int cols) {
    if (ld_dst == cols) { stage_floats(dst, src, rows * cols); return; }
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) dst[(i / cols) * ld_dst + i % cols] = __ldg(src + i);
}
__host__ __device__ constexpr int up4(int n) { return (n + 3) & ~3; }

template <int KC>
__device__ __forceinline__ void layer_norm(float (&x)[KC], const float* __restrict__ g, const float* __restrict__ b,
                                           int lane, bool relu) {
    constexpr float inv = 1.0f / (32 * KC);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) s += x[k];
    const float mean = warp_sum(s) * inv;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const float dlt = x[k] - mean;
        v = fmaf(dlt, dlt, v);
    }
    const float rstd = rsqrtf(warp_sum(v) * inv + 1e-5f);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = ch<KC>(lane, k);
        const float y = fmaf((x[k] - mean) * rstd, __ldg(g + c), __ldg(b + c));
        x[k] = relu ? fmaxf(y, 0.f) : y;
    }
}

constexpr int kNzPairWarps = 8;    // warps per CTA of stage 1

// Stage 1: one warp per selected pair — RPE MLP hidden vector and its folded contraction,
//   R[s] = (h(pa,pb) + h(pb,pa)) (W_pe W2_t)^T + c_t          (models/link_transformer.py:182-211)
// so that links with hundreds of pairs do not serialise that work behind one warp in stage 2.  The (at most three)
// folded matrices are staged in shared memory once per CTA.
template <int D>
__global__ void __launch_bounds__(32 * kNzPairWarps, 2) nz_pairs_kernel(const __grid_constant__ NzParams p) {
    constexpr int KC = D / 32;
    extern __shared__ __align__(16) float nz_smem[];
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    {
        int64_t most = 0;
        for (int t = 0; t < p.ntypes; ++t) most = max(most, min(p.cap, p.hdr[t]));
        if ((int64_t)blockIdx.x * kNzPairWarps >= most) return;      // no pair for this CTA: before staging
    }
    float* xs = nz_smem + 3 * D * D + (threadIdx.x >> 5) * (D + 8);
    for (int t = 0; t < p.ntypes; ++t) stage_floats(nz_smem + t * D * D, p.rpe_mT[t], D * D);
    __syncthreads();
    for (int t = 0; t < p.ntypes; ++t) {
        const int64_t rows = min(p.cap, p.hdr[t]);
        const float* w1 = p.rpe_w1[t];
        for (int64_t r = warp; r < rows; r += nwarps) {
            const int64_t s = t * p.cap + r;
            const float pa = __ldg(p.pa + s), pb = __ldg(p.pb + s);
            float z1[KC], z2[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = ch<KC>(lane, k);
                const float wx = __ldg(w1 + 2 * c), wy = __ldg(w1 + 2 * c + 1), bb = __ldg(p.rpe_b1[t] + c);
                z1[k] = fmaf(wx, pa, fmaf(wy, pb, bb));
                z2[k] = fmaf(wx, pb, fmaf(wy, pa, bb));
            }
            layer_norm<KC>(z1, p.rpe_g[t], p.rpe_b[t], lane, true);
            layer_norm<KC>(z2, p.rpe_g[t], p.rpe_b[t], lane, true);
            float hs[KC], v[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                hs[k] = z1[k] + z2[k];
                v[k] = __ldg(p.rpe_c[t] + ch<KC>(lane, k));
            }
            put_vec<KC>(xs, hs, lane);
            matvec_s<D, KC>(nz_smem + t * D * D, D, xs, v, lane);
#pragma unroll
            for (int k = 0; k < KC; ++k) p.R[s * D + ch<KC>(lane, k)] = v[k];
        }
    }
}

constexpr int kNzGroup = 8;
constexpr int kNzWarps = 24;       // warps per CTA of stage 2 (one CTA per SM: the matrices take 134 KB of its shared memory)
constexpr int kNzHeavy = 64;       // a link with more selected pairs than this is walked by its whole CTA

// the link stage's matrices in shared memory (pd = D + cd rows / columns where the count channels take part; the rows of
// p1T padded to a multiple of four floats)
struct NzW {
    const float *wlT, *p1T, *p2T, *wzT, *w1T, *w23T;
    int ldp;
};
__host__ __device__ constexpr int nz_w_floats(int D, int pd) {
    return D * D + pd * up4(pd) + up4(pd * D) + D * 2 * D + D * D + D * 2 * D;
}

// Online-softmax attention over the pairs of one link (reference modules/layers.py:193-224) — or, for a link shared
// by the `nw` warps of a CTA, over this warp's share of them (blocks of 32 pairs, round robin from `wslot`).
// The pairs are taken kNzGroup at a time: the node ids of 32 pairs come in with one coalesced read, then the
// kNzGroup gathered K/V rows (random 256-byte reads, the latency that bounds this kernel) and RPE rows are all in
// flight before the first score is reduced; one running-max update per group.
template <int D>
__device__ __forceinline__ void attend_pairs(const NzParams& p, int64_t pos, const float (&q)[D / 32],
                                             const float (&att)[D / 32], int lane, int wslot, int nw, float& mx,
                                             float& den, float (&acc)[D / 32]) {
    constexpr int KC = D / 32;
    constexpr int G = kNzGroup;
    const float* __restrict__ Rr = p.R;
    for (int t = 0; t < p.ntypes; ++t) {
        const int64_t s0 = t * p.cap + __ldg(p.seg_start + t * p.bs + pos);
        const int n_t = __ldg(p.counts + t * p.bs + pos);
        for (int base = 32 * wslot; base < n_t; base += 32 * nw) {
            const int m = min(32, n_t - base);
            const int32_t u_l = lane < m ? __ldg(p.node + s0 + base + lane) : 0;
            for (int g0 = 0; g0 < m; g0 += G) {
                float v[G][KC], sc[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int jj = g0 + j;
                    const int64_t u = __shfl_sync(kFull, u_l, jj & 31);
                    const int64_t s = s0 + base + (jj < m ? jj : g0);
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                        v[j][k] = ld_tab(p.KV, u * p.ld_kv + ch<KC>(lane, k), p.tab_bf16) + __ldg(Rr + s * D + ch<KC>(lane, k));
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    float part = 0.f;
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        float x = v[j][k] * q[k];
                        x = (x > 0.f) ? x : 0.2f * x;
                        part = fmaf(att[k], x, part);
                    }
                    sc[j] = part;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int j = 0; j < G; ++j) sc[j] += __shfl_xor_sync(kFull, sc[j], o);
                float m_new = mx;
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (g0 + j >= m) sc[j] = -INFINITY;
                    m_new = fmaxf(m_new, sc[j]);
                }
                const float scale = expf(mx - m_new);
                den *= scale;
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[k] *= scale;
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const float w = expf(sc[j] - m_new);
                    den += w;
#pragma unroll
                    for (int k = 0; k < KC; ++k) acc[k] = fmaf(w, v[j][k], acc[k]);
                }
                mx = m_new;
            }
        }
    }
}

// Everything after the attention sum of one link, by one warp: + bias, LayerNorm, counts, pairwise_lin,
// elementwise_lin, mlp_score, sigmoid.
template <int D>
__device__ __forceinline__ void finish_link(const NzParams& p, const NzW& W, float* xs, int64_t pos,
                                            const float (&xprod)[D / 32], float den, const float (&acc)[D / 32], int lane) {
    constexpr int KC = D / 32;
    const int pd = D + p.cd;
    int cnt[3] = {0, 0, 0};
#pragma unroll
    for (int t = 0; t < 3; ++t)
        if (t < p.ntypes) cnt[t] = __ldg(p.counts + t * p.bs + pos);
    // out = LN(acc / (den + 1e-16) + bias), then the counts (models/link_transformer.py:340-386)
    float f[KC];
    {
        const float inv = 1.0f / (den + 1e-16f);
#pragma unroll
        for (int k = 0; k < KC; ++k) f[k] = fmaf(acc[k], inv, __ldg(p.att_bias + ch<KC>(lane, k)));
        layer_norm<KC>(f, p.pn_w, p.pn_b, lane, false);
    }
    float fx = 0.f;     // channel D + lane of the pairwise_lin input, lane < cd
    if (p.mode == LPF_MODE_CN) {
        if (lane == 0) fx = (float)cnt[0];
    } else if (p.mode == LPF_MODE_1HOP) {
        fx = lane == 0 ? (float)cnt[0] : lane == 1 ? (float)cnt[1] : lane == 2 ? (float)(cnt[0] + cnt[1]) : 0.f;
    } else {
        fx = lane == 0 ? (float)cnt[0] : lane == 1 ? (float)cnt[1] : lane == 2 ? (float)cnt[2]
             : lane == 3 ? (float)(cnt[0] + cnt[1]) : 0.f;
    }

    // ---- pairwise_lin: Linear(pd,pd) -> LayerNorm(pd) -> ReLU -> Linear(pd,d)
    float hid[KC], hx = 0.f;      // hx: output channel D + lane (lane < cd)
#pragma unroll
    for (int k = 0; k < KC; ++k) hid[k] = __ldg(p.pb1 + ch<KC>(lane, k));
    {
        put_vec<KC>(xs, f, lane);
        matvec_s<D, KC>(W.p1T, W.ldp, xs, hid, lane);                                  // main channels in, main channels out
        matvec_counts_s<KC>(W.p1T + D * W.ldp, W.ldp, fx, p.cd, hid, lane);            // count channels in
        // the (at most four) count output channels: a dot product over all pd inputs each, reduced over the warp
        float hx_all[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (e < p.cd) {
                float part = lane < p.cd ? fx * W.p1T[(D + lane) * W.ldp + D + e] : 0.f;
#pragma unroll
                for (int k = 0; k < KC; ++k) part = fmaf(f[k], W.p1T[ch<KC>(lane, k) * W.ldp + D + e], part);
                hx_all[e] = warp_sum(part);
            }
        }
        if (lane < p.cd) hx = (lane == 0 ? hx_all[0] : lane == 1 ? hx_all[1] : lane == 2 ? hx_all[2] : hx_all[3]) + __ldg(p.pb1 + D + lane);
        // LayerNorm over the pd channels, ReLU
        float s = (lane < p.cd) ? hx : 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) s += hid[k];
        const float mean = warp_sum(s) / (float)pd;
        float vv = 0.f;
        if (lane < p.cd) vv = (hx - mean) * (hx - mean);
#pragma unroll
        for (int k = 0; k < KC; ++k) vv = fmaf(hid[k] - mean, hid[k] - mean, vv);
        const float rstd = rsqrtf(warp_sum(vv) / (float)pd + 1e-5f);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int c = ch<KC>(lane, k);
            hid[k] = fmaxf(fmaf((hid[k] - mean) * rstd, __ldg(p.pln_w + c), __ldg(p.pln_b + c)), 0.f);
        }
        if (lane < p.cd) hx = fmaxf(fmaf((hx - mean) * rstd, __ldg(p.pln_w + D + lane), __ldg(p.pln_b + D + lane)), 0.f);
    }
    float pw[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) pw[k] = __ldg(p.pb2 + ch<KC>(lane, k));
    put_vec<KC>(xs, hid, lane);
    matvec_s<D, KC>(W.p2T, D, xs, pw, lane);
    matvec_counts_s<KC>(W.p2T + D * D, D, lane < p.cd ? hx : 0.f, p.cd, pw, lane);

    // ---- mlp_score's first layer: offset from the pairwise half + folded elementwise half
    float z[2 * KC];
#pragma unroll
    for (int k = 0; k < 2 * KC; ++k) z[k] = __ldg(p.off + ch<2 * KC>(lane, k));
    put_vec<KC>(xs, pw, lane);
    matvec_s<D, 2 * KC>(W.wzT, 2 * D, xs, z, lane);
    float h[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) h[k] = __ldg(p.b1 + ch<KC>(lane, k));
    put_vec<KC>(xs, xprod, lane);
    matvec_s<D, KC>(W.w1T, D, xs, h, lane);
    layer_norm<KC>(h, p.ln_g, p.ln_b, lane, true);
    put_vec<KC>(xs, h, lane);
    matvec_s<D, 2 * KC>(W.w23T, 2 * D, xs, z, lane);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * KC; ++k) part = fmaf(fmaxf(z[k], 0.f), __ldg(p.ws2 + ch<2 * KC>(lane, k)), part);
    const float logit = warp_sum(part) + __ldg(p.bs2);
    if (lane == 0) p.prob[pos] = p.logits ? logit : 1.0f / (1.0f + expf(-logit));
}

// Stage 2: one warp per non-empty link; a link with more than kNzHeavy pairs (a positive between hubs: hundreds of
// common neighbours) is walked by all warps of its CTA, which then merge their softmax states.
template <int D>
__global__ void __launch_bounds__(32 * kNzWarps, 1) nz_fused_kernel(const __grid_constant__ NzParams p) {
    constexpr int KC = D / 32;
    extern __shared__ __align__(16) float nz_smem[];
    __shared__ float s_acc[kNzWarps][D];
    __shared__ float s_mx[kNzWarps], s_den[kNzWarps];
    __shared__ int s_heavy[kNzWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = p.n_dev ? min(p.n_cap, *p.n_dev) : p.n_cap;
    if ((int64_t)blockIdx.x * kNzWarps >= n) return;         // no link for this CTA: before staging the matrices
    const int pd = D + p.cd;
    NzW W;
    W.ldp = up4(pd);
    float* xs;
    // this warp's first link: its position, endpoints and node rows are requested before the matrices are staged (four
    // dependent DRAM trips that would otherwise start after the staging)
    {
        const int64_t j = (int64_t)blockIdx.x * kNzWarps + warp;
        if (j < n) {
            const int64_t pos = __ldg(p.nz + j);
            const int64_t a = __ldg(p.links + pos), b = __ldg(p.links + p.bs + pos);
            const int esz = p.tab_bf16 ? 2 : 4;
            const char* ra = reinterpret_cast<const char*>(p.X) + a * p.ldx * esz;
            const char* rb = reinterpret_cast<const char*>(p.X) + b * p.ldx * esz;
            for (int o = 0; o < D * esz; o += 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ra + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rb + o));
            }
            for (int t = 0; t < p.ntypes; ++t) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.counts + t * p.bs + pos));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.seg_start + t * p.bs + pos));
            }
        }
    }
    {
        float* w = nz_smem;
        W.wlT = w;  stage_floats(w, p.wlT, D * D);              w += D * D;
        W.p1T = w;  stage_rows(w, W.ldp, p.p1T, pd, pd);        w += pd * W.ldp;
        W.p2T = w;  stage_floats(w, p.p2T, pd * D);             w += up4(pd * D);
        W.wzT = w;  stage_floats(w, p.wzT, D * 2 * D);          w += D * 2 * D;
        W.w1T = w;  stage_floats(w, p.w1T, D * D);              w += D * D;
        W.w23T = w; stage_floats(w, p.w23T, D * 2 * D);         w += D * 2 * D;
        xs = w + warp * (D + 8);
    }
    __syncthreads();

    for (int64_t j0 = (int64_t)blockIdx.x * kNzWarps; j0 < n; j0 += (int64_t)gridDim.x * kNzWarps) {
        const int64_t j = j0 + warp;
        bool heavy = false;
        if (j < n) {
            const int64_t pos = __ldg(p.nz + j);
            int total = 0;
            for (int t = 0; t < p.ntypes; ++t) total += __ldg(p.counts + t * p.bs + pos);
            heavy = total > kNzHeavy;
        }
        if (lane == 0) s_heavy[warp] = heavy ? 1 : 0;
        __syncthreads();
        // this warp's own link when it is light; then link w of the CTA's slice for every heavy w, all warps together
        for (int w = -1; w < kNzWarps; ++w) {
            const bool own = w < 0;
            if (own ? (heavy || j >= n) : !s_heavy[w]) continue;       // (second test: uniform across the CTA)
            const int64_t pos = __ldg(p.nz + (own ? j : j0 + w));
            const int64_t a = __ldg(p.links + pos), b = __ldg(p.links + p.bs + pos);
            float xsum[KC], xprod[KC], q[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = ch<KC>(lane, k);
                const float xa = ld_tab(p.X, a * p.ldx + c, p.tab_bf16), xb = ld_tab(p.X, b * p.ldx + c, p.tab_bf16);
                xsum[k] = xa + xb;
                xprod[k] = xa * xb;
                q[k] = 2.0f * __ldg(p.bl + c);
            }
            put_vec<KC>(xs, xsum, lane);
            matvec_s<D, KC>(W.wlT, D, xs, q, lane);
            float att[KC], acc[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                att[k] = __ldg(p.att + ch<KC>(lane, k));
                acc[k] = 0.f;
            }
            float mx = -INFINITY, den = 0.f;
            if (own) {
                attend_pairs<D>(p, pos, q, att, lane, 0, 1, mx, den, acc);
                finish_link<D>(p, W, xs, pos, xprod, den, acc, lane);
                continue;
            }
            attend_pairs<D>(p, pos, q, att, lane, warp, kNzWarps, mx, den, acc);
#pragma unroll
            for (int k = 0; k < KC; ++k) s_acc[warp][ch<KC>(lane, k)] = acc[k];
            if (lane == 0) { s_mx[warp] = mx; s_den[warp] = den; }
            __syncthreads();
            if (warp == 0) {
                float M = -INFINITY;
#pragma unroll
                for (int x = 0; x < kNzWarps; ++x) M = fmaxf(M, s_mx[x]);
                den = 0.f;
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[k] = 0.f;
#pragma unroll
                for (int x = 0; x < kNzWarps; ++x) {
                    const float sc = expf(s_mx[x] - M);      // a warp without pairs has mx = -inf, den = 0: weight 0
                    den = fmaf(s_den[x], sc, den);
#pragma unroll
                    for (int k = 0; k < KC; ++k) acc[k] = fmaf(s_acc[x][ch<KC>(lane, k)], sc, acc[k]);
                }
                finish_link<D>(p, W, xs, pos, xprod, den, acc, lane);
            }
            __syncthreads();
        }
        __syncthreads();       // s_heavy is rewritten by the next slice
    }
}

}  // namespace lpf

using namespace lpf;

/* Pointer block of lpf_nz_links_fused, mirrored field by field by the ctypes Structure in _lib.py. */
namespace lpf { extern bool g_kernel_timing; }
static cudaEvent_t g_nz_ev[3];
static bool g_nz_ev_ready = false, g_nz_ev_valid = false;
// Profiling hook (see lpf_debug_select_timing): durations of nz_pairs_kernel and nz_fused_kernel of the last call.
extern "C" int lpf_debug_nz_timing_read(float* ms2_host) {
    if (!g_nz_ev_valid || !ms2_host) return -1;
    if (cudaEventSynchronize(g_nz_ev[2]) != cudaSuccess) return -1;
    for (int k = 0; k < 2; ++k)
        if (cudaEventElapsedTime(ms2_host + k, g_nz_ev[k], g_nz_ev[k + 1]) != cudaSuccess) return -1;
    return LPF_OK;
}

static int nz_launch(const lpf_nz_args* a, void* stream, bool pairs_only) {
    LPF_REQUIRE(a, "NULL argument block");
    LPF_REQUIRE(a->d == 32 || a->d == 64, "d must be 32 or 64");
    LPF_REQUIRE(a->n_cap >= 0 && a->bs >= 0, "negative size");
    if (a->n_cap == 0) return LPF_OK;
    LPF_REQUIRE(a->mode == LPF_MODE_CN || a->mode == LPF_MODE_1HOP || a->mode == LPF_MODE_ALL, "bad mode");
    NzParams p;
    p.links = a->links; p.bs = a->bs; p.nz = a->nz; p.n_cap = a->n_cap; p.n_dev = a->n_dev;
    p.X = a->X; p.ldx = a->ldx; p.KV = a->KV; p.ld_kv = a->ld_kv;
    p.node = a->node; p.pa = a->src_ppr; p.pb = a->tgt_ppr; p.seg_start = a->seg_start; p.counts = a->counts;
    p.cap = a->cap; p.mode = a->mode; p.hdr = a->header; p.R = a->R;
    p.ntypes = a->mode == LPF_MODE_CN ? 1 : (a->mode == LPF_MODE_1HOP ? 2 : 3);
    p.cd = a->mode == LPF_MODE_CN ? 1 : (a->mode == LPF_MODE_1HOP ? 3 : 4);
    p.wlT = a->wlT; p.bl = a->bl;
    for (int t = 0; t < 3; ++t) {
        p.rpe_w1[t] = a->rpe_w1[t]; p.rpe_b1[t] = a->rpe_b1[t]; p.rpe_g[t] = a->rpe_ln_w[t]; p.rpe_b[t] = a->rpe_ln_b[t];
        p.rpe_mT[t] = a->rpe_mT[t]; p.rpe_c[t] = a->rpe_c[t];
        LPF_REQUIRE(t >= p.ntypes || (p.rpe_w1[t] && p.rpe_b1[t] && p.rpe_g[t] && p.rpe_b[t] && p.rpe_mT[t] && p.rpe_c[t]),
                    "NULL RPE parameter");
    }
    p.att = a->att; p.att_bias = a->att_bias; p.pn_w = a->post_ln_w; p.pn_b = a->post_ln_b;
    p.p1T = a->p1T; p.pb1 = a->pb1; p.pln_w = a->pln_w; p.pln_b = a->pln_b; p.p2T = a->p2T; p.pb2 = a->pb2;
    p.wzT = a->wzT; p.off = a->off; p.w1T = a->w1T; p.b1 = a->b1; p.ln_g = a->ln_w; p.ln_b = a->ln_b;
    p.w23T = a->w23T; p.ws2 = a->ws2; p.bs2 = a->bs2; p.prob = a->prob; p.logits = a->logits;
    p.tab_bf16 = a->tab_bf16 ? 1 : 0;
    LPF_REQUIRE(p.hdr && (p.cap == 0 || p.R), "NULL header / R");
    LPF_REQUIRE(pairs_only || (p.links && p.nz && p.X && p.KV && p.seg_start && p.counts && p.wlT && p.bl && p.att && p.att_bias &&
                    p.pn_w && p.pn_b && p.p1T && p.pb1 && p.pln_w && p.pln_b && p.p2T && p.pb2 && p.wzT && p.off &&
                    p.w1T && p.b1 && p.ln_g && p.ln_b && p.w23T && p.ws2 && p.bs2 && p.prob),
                "NULL argument");
    LPF_REQUIRE(p.cap == 0 || (p.pa && p.pb), "NULL PPR pair values");
    cudaStream_t st = (cudaStream_t)stream;
    // link stage: one CTA of kNzWarps warps per SM (its matrices fill most of the SM's shared memory)
    int64_t blocks = (a->n_cap + kNzWarps - 1) / kNzWarps;
    if (blocks > kNumSMs) blocks = kNumSMs;
    // pair stage: kNzPairWarps warps per CTA, up to four CTAs per SM
    int64_t pblocks = (3 * a->cap + kNzPairWarps - 1) / kNzPairWarps;
    if (pblocks > (int64_t)kNumSMs * 4) pblocks = (int64_t)kNumSMs * 4;
    if (pblocks < 1) pblocks = 1;
    const int pd = a->d + p.cd;
    const size_t smem_links = ((size_t)nz_w_floats(a->d, pd) + (size_t)kNzWarps * (a->d + 8)) * 4;
    const size_t smem_pairs = ((size_t)3 * a->d * a->d + (size_t)kNzPairWarps * (a->d + 8)) * 4;
    {
        // (the attributes are per device: once per device and width)
        static bool configured[64][2] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        const int w = a->d == 64 ? 1 : 0;
        if (dev < 0 || dev >= 64 || !configured[dev][w]) {
            // (the largest layout: mode ALL, four count channels)
            const int smax = (nz_w_floats(a->d, a->d + 4) + kNzWarps * (a->d + 8)) * 4;
            cudaError_t e1 = a->d == 64 ? cudaFuncSetAttribute(nz_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax)
                                        : cudaFuncSetAttribute(nz_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
            cudaError_t e2 = a->d == 64 ? cudaFuncSetAttribute(nz_pairs_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pairs)
                                        : cudaFuncSetAttribute(nz_pairs_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pairs);
            if (e1 != cudaSuccess || e2 != cudaSuccess) {
                set_error("lpf_nz_links_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
                return LPF_ERR_CUDA;
            }
            if (dev >= 0 && dev < 64) configured[dev][w] = true;
        }
    }
    const bool timing = lpf::g_kernel_timing;
    if (timing && !g_nz_ev_ready) {
        for (auto& e : g_nz_ev) cudaEventCreate(&e);
        g_nz_ev_ready = true;
    }
    if (timing) cudaEventRecord(g_nz_ev[0], st);
    if (a->d == 64) nz_pairs_kernel<64><<<(unsigned)pblocks, 32 * kNzPairWarps, smem_pairs, st>>>(p);
    else nz_pairs_kernel<32><<<(unsigned)pblocks, 32 * kNzPairWarps, smem_pairs, st>>>(p);
    if (pairs_only) return check_launch("lpf_nz_pairs");
    if (timing) cudaEventRecord(g_nz_ev[1], st);
    if (a->d == 64) nz_fused_kernel<64><<<(unsigned)blocks, 32 * kNzWarps, smem_links, st>>>(p);
    else nz_fused_kernel<32><<<(unsigned)blocks, 32 * kNzWarps, smem_links, st>>>(p);
    if (timing) {
        cudaEventRecord(g_nz_ev[2], st);
        g_nz_ev_valid = true;
    }
    return check_launch("lpf_nz_links_fused");
}

extern "C" int lpf_nz_links_fused(const lpf_nz_args* a, void* stream) { return nz_launch(a, stream, false); }

/* The pair stage alone: R[s] = RPE contraction of every selected pair (all types in one launch). */
extern "C" int lpf_nz_pairs(const lpf_nz_args* a, void* stream) { return nz_launch(a, stream, true); }
