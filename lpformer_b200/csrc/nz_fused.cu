// Fused path for the links that selected at least one node, small-batch regime (citation2-shaped evaluation: ~1 %
// of the links of a batch): one warp carries one link from its node sets to its score —
//     Q = lin_l(X[a]) + lin_l(X[b])                                           (modules/layers.py:208-214)
//     per selected pair: RPE MLP + its contraction, v = KV[u] + r, score, online segment softmax
//                                                                             (models/link_transformer.py:182-211,
//                                                                              modules/layers.py:193-224)
//     + bias, LayerNorm, counts, pairwise_lin                                 (:340-386, :177)
//     elementwise_lin(X[a]*X[b]), mlp_score on [el | pw], sigmoid             (models/other_models.py:125-138,173-179)
// in two launches (one warp per selected pair for the RPE contraction, then one warp per link), in fp32 FFMA
// with the vectors distributed over the lanes (channel c = lane + 32k) and every weight read as
// coalesced 128-byte rows of its TRANSPOSE (W^T[k][n], L1/L2 resident).  It replaces ~14 launches of the batched
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
};

// y[j] += sum_c x[c] * WT[c][lane + 32 j]   for the KIN*32 lane-distributed input channels c = l + 32 kk.
// The weight rows are read in blocks of RB rows, the block after the current one already in flight while this
// one is multiplied: a row is touched once per link, so every read is an L2 round trip, and a loop that reads
// one row per shuffle pays that latency 32 * KIN times (measured: 10 us per 64 x 64 product).
// EXTRA: one more output column D + lane for lanes < n_extra (the count channels of pairwise_lin's first layer).
template <int KIN, int KOUT, bool EXTRA = false>
__device__ __forceinline__ void matvec(const float* __restrict__ WT, int ldw, const float (&x)[KIN], float (&y)[KOUT],
                                       int lane, float* y_extra = nullptr, int n_extra = 0) {
    constexpr int KO = KOUT + (EXTRA ? 1 : 0);
    constexpr int RB = KO >= 4 ? 4 : 8;
    constexpr int NB = KIN * 32 / RB;
    const bool ex = EXTRA && lane < n_extra;
    float w[2][RB][KO];
    auto load = [&](int blk, float (&dst)[RB][KO]) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const float* row = WT + (size_t)(blk * RB + r) * ldw + lane;
#pragma unroll
            for (int j = 0; j < KOUT; ++j) dst[r][j] = __ldg(row + 32 * j);
            if (EXTRA) dst[r][KO - 1] = ex ? __ldg(row + 32 * KOUT) : 0.f;
        }
    };
    load(0, w[0]);
#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
        if (blk + 1 < NB) load(blk + 1, w[(blk + 1) & 1]);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const int c = blk * RB + r;
            const float xv = __shfl_sync(kFull, x[c / 32], c % 32);
#pragma unroll
            for (int j = 0; j < KOUT; ++j) y[j] = fmaf(xv, w[blk & 1][r][j], y[j]);
            if (EXTRA) *y_extra = fmaf(xv, w[blk & 1][r][KO - 1], *y_extra);
        }
    }
}

// The (at most four) count channels that follow the D main input channels: rows D .. D + cd - 1 of WT, their
// inputs in lanes 0 .. cd - 1 of `xc`.
template <int KOUT, bool EXTRA = false>
__device__ __forceinline__ void matvec_counts(const float* __restrict__ WT, int ldw, float xc, int cd, float (&y)[KOUT],
                                              int lane, float* y_extra = nullptr) {
    constexpr int KO = KOUT + (EXTRA ? 1 : 0);
    const bool ex = EXTRA && lane < cd;
    float w[4][KO];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float* row = WT + (size_t)e * ldw + lane;
#pragma unroll
        for (int j = 0; j < KOUT; ++j) w[e][j] = e < cd ? __ldg(row + 32 * j) : 0.f;
        if (EXTRA) w[e][KO - 1] = (e < cd && ex) ? __ldg(row + 32 * KOUT) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float xv = __shfl_sync(kFull, xc, e);
#pragma unroll
        for (int j = 0; j < KOUT; ++j) y[j] = fmaf(xv, w[e][j], y[j]);
        if (EXTRA) *y_extra = fmaf(xv, w[e][KO - 1], *y_extra);
    }
}

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
        const int c = lane + 32 * k;
        const float y = fmaf((x[k] - mean) * rstd, __ldg(g + c), __ldg(b + c));
        x[k] = relu ? fmaxf(y, 0.f) : y;
    }
}

// Stage 1: one warp per selected pair — RPE MLP hidden vector and its folded contraction,
//   R[s] = (h(pa,pb) + h(pb,pa)) (W_pe W2_t)^T + c_t          (models/link_transformer.py:182-211)
// so that links with hundreds of pairs do not serialise that work behind one warp in stage 2.
template <int D>
__global__ void __launch_bounds__(256, 2) nz_pairs_kernel(const __grid_constant__ NzParams p) {
    constexpr int KC = D / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int t = 0; t < p.ntypes; ++t) {
        const int64_t rows = min(p.cap, p.hdr[t]);
        const float* w1 = p.rpe_w1[t];
        for (int64_t r = warp; r < rows; r += nwarps) {
            const int64_t s = t * p.cap + r;
            const float pa = __ldg(p.pa + s), pb = __ldg(p.pb + s);
            float z1[KC], z2[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = lane + 32 * k;
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
                v[k] = __ldg(p.rpe_c[t] + lane + 32 * k);
            }
            matvec<KC, KC>(p.rpe_mT[t], D, hs, v, lane);
#pragma unroll
            for (int k = 0; k < KC; ++k) p.R[s * D + lane + 32 * k] = v[k];
        }
    }
}

constexpr int kNzGroup = 8;
constexpr int kNzWarps = 8;        // warps per CTA of stage 2
constexpr int kNzHeavy = 64;       // a link with more selected pairs than this is walked by its whole CTA

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
                        v[j][k] = __ldg(p.KV + u * p.ld_kv + lane + 32 * k) + __ldg(Rr + s * D + lane + 32 * k);
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
__device__ __forceinline__ void finish_link(const NzParams& p, int64_t pos, const float (&xprod)[D / 32], float den,
                                            const float (&acc)[D / 32], int lane) {
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
        for (int k = 0; k < KC; ++k) f[k] = fmaf(acc[k], inv, __ldg(p.att_bias + lane + 32 * k));
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
    float hid[KC], hx = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) hid[k] = __ldg(p.pb1 + lane + 32 * k);
    if (lane < p.cd) hx = __ldg(p.pb1 + D + lane);
    {
        matvec<KC, KC, true>(p.p1T, pd, f, hid, lane, &hx, p.cd);      // main channels of the input
        matvec_counts<KC, true>(p.p1T + (size_t)D * pd, pd, fx, p.cd, hid, lane, &hx);   // the count channels
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
            const int c = lane + 32 * k;
            hid[k] = fmaxf(fmaf((hid[k] - mean) * rstd, __ldg(p.pln_w + c), __ldg(p.pln_b + c)), 0.f);
        }
        if (lane < p.cd) hx = fmaxf(fmaf((hx - mean) * rstd, __ldg(p.pln_w + D + lane), __ldg(p.pln_b + D + lane)), 0.f);
    }
    float pw[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) pw[k] = __ldg(p.pb2 + lane + 32 * k);
    matvec<KC, KC>(p.p2T, D, hid, pw, lane);
    matvec_counts<KC>(p.p2T + (size_t)D * D, D, lane < p.cd ? hx : 0.f, p.cd, pw, lane);

    // ---- mlp_score's first layer: offset from the pairwise half + folded elementwise half
    float z[2 * KC];
#pragma unroll
    for (int k = 0; k < 2 * KC; ++k) z[k] = __ldg(p.off + lane + 32 * k);
    matvec<KC, 2 * KC>(p.wzT, 2 * D, pw, z, lane);
    float h[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) h[k] = __ldg(p.b1 + lane + 32 * k);
    matvec<KC, KC>(p.w1T, D, xprod, h, lane);
    layer_norm<KC>(h, p.ln_g, p.ln_b, lane, true);
    matvec<KC, 2 * KC>(p.w23T, 2 * D, h, z, lane);
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * KC; ++k) part = fmaf(fmaxf(z[k], 0.f), __ldg(p.ws2 + lane + 32 * k), part);
    const float logit = warp_sum(part) + __ldg(p.bs2);
    if (lane == 0) p.prob[pos] = p.logits ? logit : 1.0f / (1.0f + expf(-logit));
}

// Stage 2: one warp per non-empty link; a link with more than kNzHeavy pairs (a positive between hubs: hundreds of
// common neighbours) is walked by all warps of its CTA, which then merge their softmax states.
template <int D>
__global__ void __launch_bounds__(32 * kNzWarps, 3) nz_fused_kernel(const __grid_constant__ NzParams p) {
    constexpr int KC = D / 32;
    __shared__ float s_acc[kNzWarps][D];
    __shared__ float s_mx[kNzWarps], s_den[kNzWarps];
    __shared__ int s_heavy[kNzWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = p.n_dev ? min(p.n_cap, *p.n_dev) : p.n_cap;

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
                const int c = lane + 32 * k;
                const float xa = __ldg(p.X + a * p.ldx + c), xb = __ldg(p.X + b * p.ldx + c);
                xsum[k] = xa + xb;
                xprod[k] = xa * xb;
                q[k] = 2.0f * __ldg(p.bl + c);
            }
            matvec<KC, KC>(p.wlT, D, xsum, q, lane);
            float att[KC], acc[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                att[k] = __ldg(p.att + lane + 32 * k);
                acc[k] = 0.f;
            }
            float mx = -INFINITY, den = 0.f;
            if (own) {
                attend_pairs<D>(p, pos, q, att, lane, 0, 1, mx, den, acc);
                finish_link<D>(p, pos, xprod, den, acc, lane);
                continue;
            }
            attend_pairs<D>(p, pos, q, att, lane, warp, kNzWarps, mx, den, acc);
#pragma unroll
            for (int k = 0; k < KC; ++k) s_acc[warp][lane + 32 * k] = acc[k];
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
                    for (int k = 0; k < KC; ++k) acc[k] = fmaf(s_acc[x][lane + 32 * k], sc, acc[k]);
                }
                finish_link<D>(p, pos, xprod, den, acc, lane);
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

extern "C" int lpf_nz_links_fused(const lpf_nz_args* a, void* stream) {
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
    LPF_REQUIRE(p.hdr && (p.cap == 0 || p.R), "NULL header / R");
    LPF_REQUIRE(p.links && p.nz && p.X && p.KV && p.seg_start && p.counts && p.wlT && p.bl && p.att && p.att_bias &&
                    p.pn_w && p.pn_b && p.p1T && p.pb1 && p.pln_w && p.pln_b && p.p2T && p.pb2 && p.wzT && p.off &&
                    p.w1T && p.b1 && p.ln_g && p.ln_b && p.w23T && p.ws2 && p.bs2 && p.prob,
                "NULL argument");
    int64_t blocks = (a->n_cap + 7) / 8;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t pblocks = (3 * a->cap + 7) / 8;
    if (pblocks > cap) pblocks = cap;
    if (pblocks < 1) pblocks = 1;
    const bool timing = lpf::g_kernel_timing;
    if (timing && !g_nz_ev_ready) {
        for (auto& e : g_nz_ev) cudaEventCreate(&e);
        g_nz_ev_ready = true;
    }
    if (timing) cudaEventRecord(g_nz_ev[0], st);
    if (a->d == 64) nz_pairs_kernel<64><<<(unsigned)pblocks, 256, 0, st>>>(p);
    else nz_pairs_kernel<32><<<(unsigned)pblocks, 256, 0, st>>>(p);
    if (timing) cudaEventRecord(g_nz_ev[1], st);
    if (a->d == 64) nz_fused_kernel<64><<<(unsigned)blocks, 32 * kNzWarps, 0, st>>>(p);
    else nz_fused_kernel<32><<<(unsigned)blocks, 32 * kNzWarps, 0, st>>>(p);
    if (timing) {
        cudaEventRecord(g_nz_ev[2], st);
        g_nz_ev_valid = true;
    }
    return check_launch("lpf_nz_links_fused");
}
