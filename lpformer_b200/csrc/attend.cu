// K4 — fused per-link attention: gather -> score -> segment softmax -> weighted sum ->
// +bias -> LayerNorm (+ set counts), one warp per link (all warps of the CTA for a link with hundreds of pairs), each
// link's set kept on chip, the gathers of up to eight pairs in flight per warp.
//
// Replaces LinkTransformerLayer.forward + LinkAttention.forward/message (reference
// modules/layers.py:39-82, :161-224), which materialise [S,2d] gathers, run lin_l / lin_r
// per pair and use three scatter passes for the softmax.  Algebra as in SURVEY App. B:
//   v_s = KV[node_s] + R_s   (R_s carries the RPE contraction and lin_r's bias)
//   sc_{s,h} = sum_c att[h,c] * leaky_relu(v_s[h,c] * Q_i[h,c], 0.2)
//   out_i = LN( sum_s softmax_s(sc)_h v_s[h,:] + bias )     (empty set -> LN(bias))
// The softmax is evaluated online (running max / sum), which is the same quantity as the
// reference's max-subtracted form exp(sc-max)/(sum+1e-16).
//
// lpf_attend_fused_ws adds two things for dense graphs: a link with more than kAttGiant pairs is registered in a caller-owned
// workspace by the first launch and walked by a second launch of the same kernel (phase 1), kGiantChunk pairs per CTA,
// the partial softmax states merged by the CTA that completes the link's last chunk; and an optional row map of R, through
// which the pairs whose two PPR values are 0 read one shared RPE row per node type (AttendParams::r_map / r_const).
#include "common.cuh"

namespace lpf {

struct AttendParams {
    const int64_t* ptr;
    int64_t bs;
    const int32_t* idx;   // optional: batch positions of the links to process (rows of Q / out follow the list)
    int64_t n;            // number of links to process (== bs when idx is NULL)
    const int64_t* n_dev; // optional device-side n (n is then an upper bound)
    // optional one-pass addressing (lpf_select_onepass): type t of link i owns rows
    // [t*type_stride + seg_start[t*bs+i], ... + seg_cnt[t*bs+i]) instead of [ptr[t*bs+i], ptr[t*bs+i+1])
    const int32_t* seg_start;
    const int32_t* seg_cnt;
    int64_t type_stride;
    const int32_t* node;
    const float* KV;
    int64_t ld_kv;
    const float* R;
    int64_t ld_r;
    const float* Q;
    int64_t ld_q;
    const float* att;
    const float* bias;
    const float* ln_w;
    const float* ln_b;
    int heads, ch, mode, write_counts;
    float* out;
    int64_t ld_out;
    float* alpha_out;
    int kv_bf16;          // KV holds bf16 (ld_kv in elements)
    // optional workspace for links with thousands of pairs ("giant": split over the grid in a second launch)
    // optional row map of R (pairs whose two PPR values are 0 share one RPE vector per type: dense graphs with
    // thresh_cn = 0, where most common neighbours have no PPR entry at all): pair s reads R[r_map[s]] when r_map[s] >= 0
    // and r_const[-1 - r_map[s]] (one row per node type) otherwise
    const int32_t* r_map;
    const float* r_const;
    int32_t* ws;          // [16] header (0: registered links, 1: chunk cursor), then GiantLink[kGiantCap], then the records
    int32_t pool_cap;     // chunk records the workspace holds
    int phase;            // 0: every link (giant ones registered in ws and left out); 1: the registered links, chunk by chunk
};
struct GiantLink { int32_t j, base, n, done; };
constexpr int kGiantCap = 1024;    // registered links per launch (more: walked by their CTA, as without a workspace)
constexpr int kAttGiant = 1024;    // a link with more pairs than this is split over the grid ...
constexpr int kGiantChunk = 256;   // ... in chunks of this many pairs, one CTA each
__device__ __forceinline__ float ld_kv(const AttendParams& p, int64_t u, int c) {
    if (p.kv_bf16) return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const uint16_t*>(p.KV) + u * p.ld_kv + c) << 16);
    return __ldg(p.KV + u * p.ld_kv + c);
}

constexpr int kAttWarps = 8;
constexpr int kAttHeavy = 256;     // a link with more pairs than this is walked by all warps of its CTA

// One warp over (its share of) the pairs of one link: online softmax over groups of G pairs — the node ids of 32 pairs
// come in with one coalesced read, then the G gathered K/V rows and RPE rows are all in flight before the first score
// is reduced (one pair at a time, every pair paid its own DRAM / L2 round trip), one running-max update per group.
// `wslot` / `nw`: this warp takes the blocks of 32 pairs wslot, wslot + nw, ... of every type.
template <int H, int KC, int G, bool BF, bool MAP>
__device__ __forceinline__ void attend_link(const AttendParams& p, const int64_t (&seg_lo)[3], const int64_t (&seg_hi)[3],
                                            const float (&q)[H][KC], const float (&att)[H][KC], int lane, int wslot, int nw,
                                            float (&mx)[H], float (&den)[H], float (&acc)[H][KC]) {
    const int C = p.ch;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        for (int64_t s0 = seg_lo[t] + 32 * wslot; s0 < seg_hi[t]; s0 += 32 * nw) {
            const int cnt = (int)min((int64_t)32, seg_hi[t] - s0);
            const int32_t my_node = (lane < cnt) ? __ldg(p.node + s0 + lane) : 0;
            int32_t my_map = 0;
            if constexpr (MAP) my_map = (lane < cnt) ? __ldg(p.r_map + s0 + lane) : 0;
            for (int g0 = 0; g0 < cnt; g0 += G) {
                float v[G][H][KC], sc[G][H];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int jj = g0 + j < cnt ? g0 + j : g0;          // (a short last group repeats its first pair: weight 0)
                    const int64_t u = __shfl_sync(kFull, my_node, jj);
                    const float* rr = p.R + (s0 + jj) * p.ld_r;
                    if constexpr (MAP) {
                        const int ri = __shfl_sync(kFull, my_map, jj);
                        rr = ri >= 0 ? p.R + (int64_t)ri * p.ld_r : p.r_const + (int64_t)(-1 - ri) * p.ld_r;
                    }
                    if constexpr (!BF) {
                        const float* kv = p.KV + u * p.ld_kv;
#pragma unroll
                        for (int h = 0; h < H; ++h)
#pragma unroll
                            for (int k = 0; k < KC; ++k) {
                                const int c = lane + 32 * k;
                                v[j][h][k] = (c < C) ? __ldg(kv + h * C + c) + __ldg(rr + h * C + c) : 0.f;
                            }
                    } else {
                        const uint16_t* kv = reinterpret_cast<const uint16_t*>(p.KV) + u * p.ld_kv;
#pragma unroll
                        for (int h = 0; h < H; ++h)
#pragma unroll
                            for (int k = 0; k < KC; ++k) {
                                const int c = lane + 32 * k;
                                v[j][h][k] = (c < C) ? __uint_as_float((uint32_t)__ldg(kv + h * C + c) << 16) + __ldg(rr + h * C + c) : 0.f;
                            }
                    }
                }
#pragma unroll
                for (int j = 0; j < G; ++j)
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        float part = 0.f;
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            float x = v[j][h][k] * q[h][k];
                            x = (x > 0.f) ? x : 0.2f * x;
                            part = fmaf(att[h][k], x, part);
                        }
                        sc[j][h] = part;
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int j = 0; j < G; ++j)
#pragma unroll
                        for (int h = 0; h < H; ++h) sc[j][h] += __shfl_xor_sync(kFull, sc[j][h], o);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float m_new = mx[h];
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (g0 + j >= cnt) sc[j][h] = -INFINITY;
                        m_new = fmaxf(m_new, sc[j][h]);
                    }
                    const float scale = expf(mx[h] - m_new);          // exp(-inf) = 0 on the first group
                    den[h] *= scale;
#pragma unroll
                    for (int k = 0; k < KC; ++k) acc[h][k] *= scale;
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const float w = expf(sc[j][h] - m_new);
                        den[h] += w;
#pragma unroll
                        for (int k = 0; k < KC; ++k) acc[h][k] = fmaf(w, v[j][h][k], acc[h][k]);
                    }
                    mx[h] = m_new;
                }
            }
        }
    }
}

// One warp per link; a link with more than kAttHeavy pairs (dense graphs: thousands of common neighbours) is walked by
// all warps of its CTA, which merge their softmax states in shared memory — otherwise the launch lasts as long as its
// longest link (measured on the ogbl-ppa shape: one link with 29,513 common neighbours, 13.9 ms for a 32,565-link batch).
template <int H, int KC, int G>
__global__ void __launch_bounds__(kAttWarps * 32) attend_kernel(const __grid_constant__ AttendParams p) {
    __shared__ float s_acc[kAttWarps][H * KC * 32];
    __shared__ float s_mx[kAttWarps][H], s_den[kAttWarps][H];
    __shared__ int s_heavy[kAttWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = p.ch, HC = H * C;
    float att[H][KC];
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int c = lane + 32 * k;
            att[h][k] = (c < C) ? __ldg(p.att + h * C + c) : 0.f;
        }

    // query vector and segment bounds of link j (list position) / i (batch position)
    auto load_link = [&](int64_t j, float (&q)[H][KC], int64_t (&seg_lo)[3], int64_t (&seg_hi)[3]) {
        const int64_t i = p.idx ? (int64_t)__ldg(p.idx + j) : j;
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = lane + 32 * k;
                q[h][k] = (c < C) ? __ldg(p.Q + j * p.ld_q + h * C + c) : 0.f;
            }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (p.seg_start) {
                seg_lo[t] = t * p.type_stride + __ldg(p.seg_start + t * p.bs + i);
                seg_hi[t] = seg_lo[t] + __ldg(p.seg_cnt + t * p.bs + i);
            } else {
                seg_lo[t] = __ldg(p.ptr + t * p.bs + i);
                seg_hi[t] = __ldg(p.ptr + t * p.bs + i + 1);
            }
        }
    };
    // everything after the attention sums of link j, by one warp: attention weights (debug output), + bias, LayerNorm, counts
    auto finish = [&](int64_t j, const float (&q)[H][KC], const int64_t (&seg_lo)[3], const int64_t (&seg_hi)[3],
                      const float (&mx)[H], const float (&den)[H], const float (&acc)[H][KC]) {
        // head-mean attention weights (debug output of the reference, modules/layers.py:73-75)
        if (p.alpha_out) {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                for (int64_t s = seg_lo[t]; s < seg_hi[t]; ++s) {
                    const int64_t u = __ldg(p.node + s);
                    const float* rr = p.R + s * p.ld_r;
                    if (p.r_map) {
                        const int ri = __ldg(p.r_map + s);
                        rr = ri >= 0 ? p.R + (int64_t)ri * p.ld_r : p.r_const + (int64_t)(-1 - ri) * p.ld_r;
                    }
                    float mean_alpha = 0.f;
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        float part = 0.f;
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            const int c = lane + 32 * k;
                            float val = 0.f;
                            if (c < C) val = ld_kv(p, u, h * C + c) + __ldg(rr + h * C + c);
                            float x = val * q[h][k];
                            x = (x > 0.f) ? x : 0.2f * x;
                            part = fmaf(att[h][k], x, part);
                        }
                        const float sc = warp_sum(part);
                        mean_alpha += expf(sc - mx[h]) / (den[h] + 1e-16f);
                    }
                    if (lane == 0) p.alpha_out[s] = mean_alpha / (float)H;
                }
            }
        }
        // out = LN(acc / (den + 1e-16) + bias)
        float o[H][KC];
        float sum = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float inv = 1.0f / (den[h] + 1e-16f);
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = lane + 32 * k;
                o[h][k] = (c < C) ? fmaf(acc[h][k], inv, __ldg(p.bias + h * C + c)) : 0.f;
                sum += o[h][k];
            }
        }
        const float mean = warp_sum(sum) / (float)HC;
        float var = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const float dlt = (lane + 32 * k < C) ? o[h][k] - mean : 0.f;
                var = fmaf(dlt, dlt, var);
            }
        const float rstd = rsqrtf(warp_sum(var) / (float)HC + 1e-5f);
        float* out = p.out + j * p.ld_out;
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int c = lane + 32 * k;
                if (c < C) out[h * C + c] = fmaf((o[h][k] - mean) * rstd, __ldg(p.ln_w + h * C + c), __ldg(p.ln_b + h * C + c));
            }
        if (p.write_counts && lane == 0) {
            // get_structure_cnts / get_count (reference models/link_transformer.py:340-386),
            // column order of :155, :173, :175
            const float n_cn = (float)(seg_hi[0] - seg_lo[0]);
            const float n_1h = (float)(seg_hi[1] - seg_lo[1]);
            const float n_n1 = (float)(seg_hi[2] - seg_lo[2]);
            if (p.mode == LPF_MODE_CN) {
                out[HC] = n_cn;
            } else if (p.mode == LPF_MODE_1HOP) {
                out[HC] = n_cn;
                out[HC + 1] = n_1h;
                out[HC + 2] = n_cn + n_1h;
            } else {
                out[HC] = n_cn;
                out[HC + 1] = n_1h;
                out[HC + 2] = n_n1;
                out[HC + 3] = n_cn + n_1h;
            }
        }
    };

    // the partial softmax states of the CTA's warps (s_acc / s_mx / s_den) -> one state, in warp 0
    auto merge_warps = [&](float (&mx)[H], float (&den)[H], float (&acc)[H][KC]) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float M = -INFINITY;
#pragma unroll
            for (int x = 0; x < kAttWarps; ++x) M = fmaxf(M, s_mx[x][h]);
            den[h] = 0.f;
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[h][k] = 0.f;
#pragma unroll
            for (int x = 0; x < kAttWarps; ++x) {
                const float sc = (M == -INFINITY) ? 0.f : expf(s_mx[x][h] - M);   // a warp without pairs: mx = -inf, den = 0
                den[h] = fmaf(s_den[x][h], sc, den[h]);
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[h][k] = fmaf(s_acc[x][(h * KC + k) * 32 + lane], sc, acc[h][k]);
            }
            mx[h] = M;
        }
    };

    if (p.phase == 1) {
        // ---- the registered giant links: chunk x = base + c of the launch belongs to CTA x mod gridDim.x; a chunk's
        // softmax state goes to its record, the CTA that completes a link's last chunk merges the records and finishes
        constexpr int kRec = 2 * H + H * KC * 32;
        const int count = min(p.ws[0], kGiantCap);
        const GiantLink* list = reinterpret_cast<const GiantLink*>(p.ws + 16);
        float* recs = reinterpret_cast<float*>(p.ws + 16 + 4 * kGiantCap);
        for (int sidx = 0; sidx < count; ++sidx) {
            const GiantLink gl = list[sidx];
            if (gl.j < 0) continue;
            int c = (int)(((int64_t)blockIdx.x - gl.base) % (int64_t)gridDim.x);
            if (c < 0) c += gridDim.x;
            for (; c < gl.n; c += gridDim.x) {
                float q[H][KC], acc[H][KC], mx[H], den[H];
                int64_t seg_lo[3], seg_hi[3], lo[3], hi[3];
                load_link(gl.j, q, seg_lo, seg_hi);
                int64_t off = 0;
                const int64_t fs = (int64_t)c * kGiantChunk, fe = fs + kGiantChunk;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int64_t len = seg_hi[t] - seg_lo[t];
                    lo[t] = seg_lo[t] + min(max(fs - off, (int64_t)0), len);
                    hi[t] = seg_lo[t] + min(max(fe - off, (int64_t)0), len);
                    off += len;
                }
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    mx[h] = -INFINITY;
                    den[h] = 0.f;
#pragma unroll
                    for (int k = 0; k < KC; ++k) acc[h][k] = 0.f;
                }
                if (p.r_map) { if (p.kv_bf16) attend_link<H, KC, G, true, true>(p, lo, hi, q, att, lane, warp, kAttWarps, mx, den, acc); else attend_link<H, KC, G, false, true>(p, lo, hi, q, att, lane, warp, kAttWarps, mx, den, acc); }
            else if (p.kv_bf16) attend_link<H, KC, G, true, false>(p, lo, hi, q, att, lane, warp, kAttWarps, mx, den, acc);
            else attend_link<H, KC, G, false, false>(p, lo, hi, q, att, lane, warp, kAttWarps, mx, den, acc);
#pragma unroll
                for (int h = 0; h < H; ++h) {
#pragma unroll
                    for (int k = 0; k < KC; ++k) s_acc[warp][(h * KC + k) * 32 + lane] = acc[h][k];
                    if (lane == 0) { s_mx[warp][h] = mx[h]; s_den[warp][h] = den[h]; }
                }
                __syncthreads();
                if (warp == 0) {
                    merge_warps(mx, den, acc);
                    float* rec = recs + (size_t)(gl.base + c) * kRec;
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        if (lane == 0) { rec[h] = mx[h]; rec[H + h] = den[h]; }
#pragma unroll
                        for (int k = 0; k < KC; ++k) rec[2 * H + (h * KC + k) * 32 + lane] = acc[h][k];
                    }
                    __threadfence();
                    __syncwarp();
                    int ticket = 0;
                    if (lane == 0) ticket = atomicAdd(&reinterpret_cast<GiantLink*>(p.ws + 16)[sidx].done, 1);
                    ticket = __shfl_sync(kFull, ticket, 0);
                    if (ticket == gl.n - 1) {
                        __threadfence();
                        const volatile float* vr = recs + (size_t)gl.base * kRec;
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            float M = -INFINITY;
                            for (int x = 0; x < gl.n; ++x) M = fmaxf(M, vr[(size_t)x * kRec + h]);
                            den[h] = 0.f;
#pragma unroll
                            for (int k = 0; k < KC; ++k) acc[h][k] = 0.f;
                            for (int x = 0; x < gl.n; ++x) {
                                const float m_x = vr[(size_t)x * kRec + h];
                                const float sc = (m_x == -INFINITY) ? 0.f : expf(m_x - M);
                                den[h] = fmaf(vr[(size_t)x * kRec + H + h], sc, den[h]);
#pragma unroll
                                for (int k = 0; k < KC; ++k)
                                    acc[h][k] = fmaf(vr[(size_t)x * kRec + 2 * H + (h * KC + k) * 32 + lane], sc, acc[h][k]);
                            }
                            mx[h] = M;
                        }
                        finish(gl.j, q, seg_lo, seg_hi, mx, den, acc);
                    }
                }
                __syncthreads();
            }
        }
        return;
    }

    const int64_t n_links = p.n_dev ? min(p.n, *p.n_dev) : p.n;
    for (int64_t j0 = (int64_t)blockIdx.x * kAttWarps; j0 < n_links; j0 += (int64_t)gridDim.x * kAttWarps) {
        const int64_t j = j0 + warp;
        float q[H][KC], acc[H][KC], mx[H], den[H];
        int64_t seg_lo[3] = {0, 0, 0}, seg_hi[3] = {0, 0, 0};
        bool heavy = false, j_skip = false;
        if (j < n_links) {
            load_link(j, q, seg_lo, seg_hi);
            const int64_t total = (seg_hi[0] - seg_lo[0]) + (seg_hi[1] - seg_lo[1]) + (seg_hi[2] - seg_lo[2]);
            heavy = total > kAttHeavy;
            if (p.ws && total > kAttGiant) {
                // register the link for the second launch (when the list and the record pool have room)
                int ok = 0;
                if (lane == 0) {
                    const int nch = (int)((total + kGiantChunk - 1) / kGiantChunk);
                    const int slot = atomicAdd(&p.ws[0], 1);
                    if (slot < kGiantCap) {
                        GiantLink* e = reinterpret_cast<GiantLink*>(p.ws + 16) + slot;
                        const int base = atomicAdd(&p.ws[1], nch);
                        if (base + nch <= p.pool_cap) {
                            e->base = base; e->n = nch; e->done = 0; e->j = (int32_t)j;
                            ok = 1;
                        } else {
                            e->j = -1;
                        }
                    }
                }
                if (__shfl_sync(kFull, ok, 0)) {
                    heavy = false;
                    seg_hi[0] = seg_lo[0]; seg_hi[1] = seg_lo[1]; seg_hi[2] = seg_lo[2];   // (nothing to do here)
                    j_skip = true;
                }
            }
        }
        if (lane == 0) s_heavy[warp] = heavy ? 1 : 0;
        __syncthreads();
        if (j < n_links && !heavy && !j_skip) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
                mx[h] = -INFINITY;
                den[h] = 0.f;
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[h][k] = 0.f;
            }
            if (p.r_map) { if (p.kv_bf16) attend_link<H, KC, G, true, true>(p, seg_lo, seg_hi, q, att, lane, 0, 1, mx, den, acc); else attend_link<H, KC, G, false, true>(p, seg_lo, seg_hi, q, att, lane, 0, 1, mx, den, acc); }
            else if (p.kv_bf16) attend_link<H, KC, G, true, false>(p, seg_lo, seg_hi, q, att, lane, 0, 1, mx, den, acc);
            else attend_link<H, KC, G, false, false>(p, seg_lo, seg_hi, q, att, lane, 0, 1, mx, den, acc);
            finish(j, q, seg_lo, seg_hi, mx, den, acc);
        }
        // the heavy links of the slice, one after the other, all warps together
        for (int w = 0; w < kAttWarps; ++w) {
            if (!s_heavy[w]) continue;                   // (uniform across the CTA)
            load_link(j0 + w, q, seg_lo, seg_hi);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                mx[h] = -INFINITY;
                den[h] = 0.f;
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[h][k] = 0.f;
            }
            if (p.r_map) { if (p.kv_bf16) attend_link<H, KC, G, true, true>(p, seg_lo, seg_hi, q, att, lane, warp, kAttWarps, mx, den, acc); else attend_link<H, KC, G, false, true>(p, seg_lo, seg_hi, q, att, lane, warp, kAttWarps, mx, den, acc); }
            else if (p.kv_bf16) attend_link<H, KC, G, true, false>(p, seg_lo, seg_hi, q, att, lane, warp, kAttWarps, mx, den, acc);
            else attend_link<H, KC, G, false, false>(p, seg_lo, seg_hi, q, att, lane, warp, kAttWarps, mx, den, acc);
#pragma unroll
            for (int h = 0; h < H; ++h) {
#pragma unroll
                for (int k = 0; k < KC; ++k) s_acc[warp][(h * KC + k) * 32 + lane] = acc[h][k];
                if (lane == 0) { s_mx[warp][h] = mx[h]; s_den[warp][h] = den[h]; }
            }
            __syncthreads();
            if (warp == 0) {
                merge_warps(mx, den, acc);
                finish(j0 + w, q, seg_lo, seg_hi, mx, den, acc);
            }
            __syncthreads();
        }
        __syncthreads();       // s_heavy is rewritten by the next slice
    }
}

template <int H>
static int launch_attend_h(const AttendParams& p, int64_t ws_bytes, cudaStream_t st) {
    int64_t blocks = (p.n + kAttWarps - 1) / kAttWarps;
    const int64_t cap = (int64_t)kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    const unsigned g = (unsigned)blocks;
    const int kc = (p.ch + 31) / 32;
    if (kc * H > 16) {
        set_error("lpf_attend_fused: heads*ceil(ch/32) = %d exceeds 16", kc * H);
        return LPF_ERR_UNSUPPORTED;
    }
    // pairs in flight per warp: as many as 16 registers of gathered values per lane allow, at most 8
    constexpr auto grp = [](int hk) { return hk <= 2 ? 8 : 16 / hk; };
    void (*kern)(const AttendParams) = nullptr;
    if (kc <= 1) kern = attend_kernel<H, 1, grp(H)>;
    else if (kc <= 2) { if constexpr (H * 2 <= 16) kern = attend_kernel<H, 2, grp(H * 2)>; }
    else if (kc <= 4) { if constexpr (H * 4 <= 16) kern = attend_kernel<H, 4, grp(H * 4)>; }
    else if (kc <= 8) { if constexpr (H * 8 <= 16) kern = attend_kernel<H, 8, grp(H * 8)>; }
    else { if constexpr (H * 16 <= 16) kern = attend_kernel<H, 16, 1>; }
    if (!kern) {
        set_error("lpf_attend_fused: heads = %d, ch = %d not instantiated", H, p.ch);
        return LPF_ERR_UNSUPPORTED;
    }
    AttendParams q = p;
    if (q.ws) {
        // record pool: what is left of the workspace after the header and the list, in records of this instantiation
        const int64_t rec = 4 * (2 * H + (int64_t)H * (kc <= 1 ? 1 : kc <= 2 ? 2 : kc <= 4 ? 4 : kc <= 8 ? 8 : 16) * 32);
        const int64_t pool = (ws_bytes - (int64_t)sizeof(int32_t) * (16 + 4 * kGiantCap)) / rec;
        if (pool < 8) q.ws = nullptr;
        else {
            q.pool_cap = (int32_t)(pool > (1 << 30) ? (1 << 30) : pool);
            cudaError_t e = cudaMemsetAsync(q.ws, 0, 64, st);
            if (e != cudaSuccess) { set_error("lpf_attend_fused: cudaMemsetAsync: %s", cudaGetErrorString(e)); return LPF_ERR_CUDA; }
        }
    }
    q.phase = 0;
    kern<<<g, kAttWarps * 32, 0, st>>>(q);
    if (q.ws) {
        q.phase = 1;
        kern<<<kNumSMs * 4, kAttWarps * 32, 0, st>>>(q);
    }
    return check_launch("lpf_attend_fused");
}

}  // namespace lpf

using namespace lpf;

extern "C" int lpf_attend_fused_ws(const int64_t* ptr, int64_t bs, const int32_t* idx, int64_t n, const int32_t* node,
                                   const float* KV, int64_t ld_kv,
                                   const float* R, int64_t ld_r, const float* Q, int64_t ld_q, const float* att,
                                   const float* bias, const float* ln_w, const float* ln_b, int32_t heads, int32_t ch,
                                   int mode, int write_counts, float* out, int64_t ld_out, float* alpha_out,
                                   const int64_t* n_dev, const int32_t* seg_start, const int32_t* seg_cnt,
                                   int64_t type_stride, int kv_bf16, const int32_t* r_map, const float* r_const,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    LPF_REQUIRE(bs >= 0 && n >= 0, "negative batch size");
    LPF_REQUIRE(idx || n == bs, "n must equal bs when idx is NULL");
    if (n == 0) return LPF_OK;
    LPF_REQUIRE((ptr || (seg_start && seg_cnt)) && KV && Q && att && bias && ln_w && ln_b && out, "NULL argument");
    LPF_REQUIRE(heads >= 1 && ch >= 1, "bad heads/ch");
    LPF_REQUIRE(mode == LPF_MODE_CN || mode == LPF_MODE_1HOP || mode == LPF_MODE_ALL, "bad mode");
    const int hc = heads * ch;
    const int cd = !write_counts ? 0 : (mode == LPF_MODE_CN ? 1 : (mode == LPF_MODE_1HOP ? 3 : 4));
    LPF_REQUIRE(ld_kv >= hc && ld_q >= hc && ld_out >= hc + cd, "leading dimension too small");
    LPF_REQUIRE(R == nullptr || ld_r >= hc, "ld_r too small");
    LPF_REQUIRE(r_map == nullptr || r_const != nullptr, "r_map needs r_const");      // (R may be NULL: no mapped row at all)
    LPF_REQUIRE(workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
    // node / R may be NULL only if every set is empty; the kernel never dereferences them then.
    AttendParams p{ptr, bs, idx, n, n_dev, seg_start, seg_cnt, type_stride, node, KV, ld_kv, R, ld_r, Q, ld_q, att, bias, ln_w, ln_b,
                   heads, ch, mode, write_counts, out, ld_out, alpha_out, kv_bf16 ? 1 : 0, r_map, r_const,
                   workspace_bytes >= lpf_attend_workspace_min() ? static_cast<int32_t*>(workspace) : nullptr, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    switch (heads) {
        case 1: return launch_attend_h<1>(p, workspace_bytes, st);
        case 2: return launch_attend_h<2>(p, workspace_bytes, st);
        case 4: return launch_attend_h<4>(p, workspace_bytes, st);
        case 8: return launch_attend_h<8>(p, workspace_bytes, st);
        default:
            set_error("lpf_attend_fused: heads must be 1, 2, 4 or 8 (got %d)", heads);
            return LPF_ERR_UNSUPPORTED;
    }
}

extern "C" int64_t lpf_attend_workspace_min(void) { return (int64_t)sizeof(int32_t) * (16 + 4 * kGiantCap) + (64 << 10); }

extern "C" int lpf_attend_fused(const int64_t* ptr, int64_t bs, const int32_t* idx, int64_t n, const int32_t* node,
                                const float* KV, int64_t ld_kv,
                                const float* R, int64_t ld_r, const float* Q, int64_t ld_q, const float* att,
                                const float* bias, const float* ln_w, const float* ln_b, int32_t heads, int32_t ch,
                                int mode, int write_counts, float* out, int64_t ld_out, float* alpha_out,
                                const int64_t* n_dev, const int32_t* seg_start, const int32_t* seg_cnt,
                                int64_t type_stride, int kv_bf16, void* stream) {
    return lpf_attend_fused_ws(ptr, bs, idx, n, node, KV, ld_kv, R, ld_r, Q, ld_q, att, bias, ln_w, ln_b, heads, ch, mode,
                               write_counts, out, ld_out, alpha_out, n_dev, seg_start, seg_cnt, type_stride, kv_bf16, nullptr,
                               nullptr, nullptr, 0, stream);
}
