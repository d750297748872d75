// K1 fast path — intersection-driven selection with sub-warp groups.
//
// Valid when every threshold that gates a PPR-derived set is > 0 (the reference's script
// configurations: thresh_1hop in {1e-4..1e-2}, thresh_non1hop in {1e-2, 1}) and all stored PPR
// values lie in (0, 1] (true for any PPR matrix; checked once per table on the host).  Then
//   CN      = A(a) ∩ A(b)                                     [filtered by q >= th_cn iff th_cn > 0]
//   1-hop   = { u ∈ P(a) ∩ P(b) : u in exactly one of A(a), A(b), q(P(a,u)) >= th_1hop, q(P(b,u)) >= th_1hop }
//   >1-hop  = { u ∈ P(a) ∩ P(b) : u in neither,                   q(.) >= th_non1hop both }
// because an absent PPR entry has q = 0 < th (reference models/link_transformer.py:241-250) and a
// one-sided entry of the >1-hop algebra evaluates to p - 1 <= 0 < th (:464-478).  These are the same
// sets, in the same ascending order and with the same fp32 values, as the generic merge kernel
// (select.cu) and the reference produce — the parity tests run both.
//
// Work shape: a group of G lanes (G = 8 for short rows, 32 for dense graphs) owns a link.  It walks
// the SHORTER of the two rows G elements at a time and binary-searches the longer one (galloping
// lower bound), so the cost is O(min·log max) instead of O(deg a + deg b): with citation2-style
// queries the long row is the shared source's row, which stays in L1 across the 1,000 negatives of a
// query.  All row pointers / first chunks are loaded before any dependent work (4+ loads in flight
// per lane), and the fill pass touches only links whose counts are non-zero.
#include "common.cuh"

namespace lpf {

struct SelectParams2 {
    const int64_t* links;
    int64_t bs;
    const int64_t* adj_rowptr;
    const int32_t* adj_col;
    const int64_t* ppr_rowptr;
    const int32_t* ppr_col;
    const float* ppr_val;
    float th_cn, th_1hop, th_non1hop;
    int mode;
    int32_t* counts;
    const int64_t* ptr;
    int32_t* node;
    float* pa;
    float* pb;
    int32_t* link;
    int32_t* heavy;   // workspace: [0] = number of heavy links, [4..] their batch positions (written by the count pass)
};

// A link whose shorter adjacency (or PPR) row exceeds kHeavyPerLane elements per lane of its group is deferred
// to select_heavy_kernel, where a whole CTA walks it: a few such links (hub-hub positives) would otherwise
// serialise thousands of dependent searches behind 8 lanes and set the duration of the whole launch.
constexpr int kHeavyPerLane = 16;
constexpr int kHeavyThreads = 256;

// lower_bound restricted to [lo, n)
__device__ __forceinline__ int lower_bound_from(const int32_t* __restrict__ a, int lo, int n, int32_t key) {
    int hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
    if constexpr (G == 32) return 0xffffffffu;
    else return ((1u << G) - 1u) << (lane & ~(G - 1));
}

template <int G, bool FILL>
__global__ void __launch_bounds__(256) select_fast_kernel(SelectParams2 p) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);                 // lane within the group
    const unsigned gmask = group_mask<G>(lane);
    const unsigned lt = gmask & ((1u << lane) - 1u);   // group lanes below me
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const bool cn_needs_ppr = FILL || p.th_cn > 0.0f;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;

    for (int64_t i = group; i < p.bs; i += ngroups) {
        int64_t o_cn = 0, o_1h = 0, o_n1 = 0;
        if (FILL) {
            // skip links whose sets are all empty (the vast majority for citation2-style negatives)
            const int64_t p0 = __ldg(p.ptr + i), p0e = __ldg(p.ptr + i + 1);
            const int64_t p1 = __ldg(p.ptr + p.bs + i), p1e = __ldg(p.ptr + p.bs + i + 1);
            const int64_t p2 = __ldg(p.ptr + 2 * p.bs + i), p2e = __ldg(p.ptr + 2 * p.bs + i + 1);
            if (p0 == p0e && p1 == p1e && p2 == p2e) continue;
            o_cn = p0; o_1h = p1; o_n1 = p2;
        }
        const int64_t a = __ldg(p.links + i), b = __ldg(p.links + p.bs + i);
        const int64_t a0 = __ldg(p.adj_rowptr + a), a1 = __ldg(p.adj_rowptr + a + 1);
        const int64_t b0 = __ldg(p.adj_rowptr + b), b1 = __ldg(p.adj_rowptr + b + 1);
        const int64_t pa0 = __ldg(p.ppr_rowptr + a), pa1 = __ldg(p.ppr_rowptr + a + 1);
        const int64_t pb0 = __ldg(p.ppr_rowptr + b), pb1 = __ldg(p.ppr_rowptr + b + 1);
        const int na = (int)(a1 - a0), nb = (int)(b1 - b0);
        const int npa = (int)(pa1 - pa0), npb = (int)(pb1 - pb0);
        if (max(min(na, nb), want_pi ? min(npa, npb) : 0) > kHeavyPerLane * G) {
            if (!FILL && gl == 0) p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
            continue;
        }
        const int32_t* Aa = p.adj_col + a0;
        const int32_t* Ab = p.adj_col + b0;
        const int32_t* Pac = p.ppr_col + pa0;
        const int32_t* Pbc = p.ppr_col + pb0;
        const float* Pav = p.ppr_val + pa0;
        const float* Pbv = p.ppr_val + pb0;

        // ---- CN: walk the shorter adjacency row, search the longer
        int c_cn = 0;
        {
            const bool a_short = na <= nb;
            const int32_t* S = a_short ? Aa : Ab;
            const int32_t* Lg = a_short ? Ab : Aa;
            const int ns = a_short ? na : nb, nl = a_short ? nb : na;
            int lo = 0;
            for (int k = 0; k < ns && lo < nl; k += G) {
                const bool act = k + gl < ns;
                const int32_t u = act ? __ldg(S + k + gl) : 0x7fffffff;
                int pos = nl;
                bool hit = false;
                if (act) {
                    pos = lower_bound_from(Lg, lo, nl, u);
                    hit = pos < nl && __ldg(Lg + pos) == u;
                }
                float qa = 0.f, qb = 0.f;
                if (hit && cn_needs_ppr) {
                    int t = lower_bound_from(Pac, 0, npa, u);
                    if (t < npa && __ldg(Pac + t) == u) qa = quantise(__ldg(Pav + t));
                    t = lower_bound_from(Pbc, 0, npb, u);
                    if (t < npb && __ldg(Pbc + t) == u) qb = quantise(__ldg(Pbv + t));
                    hit = qa >= p.th_cn && qb >= p.th_cn;
                }
                const unsigned m = __ballot_sync(gmask, hit);
                if (FILL && hit) {
                    const int64_t s = o_cn + c_cn + __popc(m & lt);
                    p.node[s] = u;
                    p.pa[s] = qa;
                    p.pb[s] = qb;
                    if (p.link) p.link[s] = (int32_t)i;
                }
                c_cn += __popc(m);
                // galloping: later elements are larger, so they cannot sit before the last lane's position
                lo = __shfl_sync(gmask, pos, (lane & ~(G - 1)) + G - 1);
            }
        }

        // ---- 1-hop / >1-hop from the intersection of the two PPR rows
        int c_1h = 0, c_n1 = 0;
        if (want_pi) {
            const bool a_short = npa <= npb;
            const int32_t* Sc = a_short ? Pac : Pbc;
            const float* Sv = a_short ? Pav : Pbv;
            const int32_t* Lc = a_short ? Pbc : Pac;
            const float* Lv = a_short ? Pbv : Pav;
            const int ns = a_short ? npa : npb, nl = a_short ? npb : npa;
            int lo = 0;
            for (int k = 0; k < ns && lo < nl; k += G) {
                const bool act = k + gl < ns;
                const int32_t u = act ? __ldg(Sc + k + gl) : 0x7fffffff;
                int pos = nl;
                bool k1 = false, kn = false;
                float qa = 0.f, qb = 0.f;
                if (act) {
                    pos = lower_bound_from(Lc, lo, nl, u);
                    if (pos < nl && __ldg(Lc + pos) == u) {
                        const float qs = quantise(__ldg(Sv + k + gl)), ql = quantise(__ldg(Lv + pos));
                        qa = a_short ? qs : ql;
                        qb = a_short ? ql : qs;
                        if (qa >= th_pre && qb >= th_pre) {
                            int t = lower_bound_from(Aa, 0, na, u);
                            const bool in_a = t < na && __ldg(Aa + t) == u;
                            t = lower_bound_from(Ab, 0, nb, u);
                            const bool in_b = t < nb && __ldg(Ab + t) == u;
                            k1 = (in_a != in_b) && qa >= p.th_1hop && qb >= p.th_1hop;
                            kn = want_n1 && !in_a && !in_b && qa >= p.th_non1hop && qb >= p.th_non1hop;
                        }
                    }
                }
                const unsigned m1 = __ballot_sync(gmask, k1);
                const unsigned mn = __ballot_sync(gmask, kn);
                if (FILL && (k1 || kn)) {
                    const int64_t s = k1 ? o_1h + c_1h + __popc(m1 & lt) : o_n1 + c_n1 + __popc(mn & lt);
                    p.node[s] = u;
                    p.pa[s] = qa;
                    p.pb[s] = qb;
                    if (p.link) p.link[s] = (int32_t)i;
                }
                c_1h += __popc(m1);
                c_n1 += __popc(mn);
                lo = __shfl_sync(gmask, pos, (lane & ~(G - 1)) + G - 1);
            }
        }

        if (!FILL && gl == 0) {
            p.counts[i] = c_cn;
            p.counts[p.bs + i] = c_1h;
            p.counts[2 * p.bs + i] = c_n1;
        }
    }
}


// Ordered block-wide compaction step: every thread passes its flag; returns this thread's rank among the
// flagged threads (thread order) and adds the block total to `running`.  Two __syncthreads per call.
__device__ __forceinline__ int block_rank(bool flag, int* warp_tot, int& running) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(kFull, flag);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kHeavyThreads / 32; ++w) {
        const int t = warp_tot[w];
        before += (w < warp) ? t : 0;
        total += t;
    }
    const int rank = running + before + __popc(m & ((1u << lane) - 1u));
    running += total;
    __syncthreads();
    return rank;
}

template <bool FILL>
__global__ void __launch_bounds__(kHeavyThreads) select_heavy_kernel(SelectParams2 p) {
    __shared__ int wt[kHeavyThreads / 32];
    const int nheavy = p.heavy[0];
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const bool cn_needs_ppr = FILL || p.th_cn > 0.0f;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    for (int q = blockIdx.x; q < nheavy; q += gridDim.x) {
        const int64_t i = p.heavy[4 + q];
        int64_t o_cn = 0, o_1h = 0, o_n1 = 0;
        if (FILL) {
            o_cn = __ldg(p.ptr + i);
            o_1h = __ldg(p.ptr + p.bs + i);
            o_n1 = __ldg(p.ptr + 2 * p.bs + i);
            if (o_cn == __ldg(p.ptr + i + 1) && o_1h == __ldg(p.ptr + p.bs + i + 1) &&
                o_n1 == __ldg(p.ptr + 2 * p.bs + i + 1))
                continue;   // nothing selected for this link
        }
        const int64_t a = __ldg(p.links + i), b = __ldg(p.links + p.bs + i);
        const int64_t a0 = __ldg(p.adj_rowptr + a), b0 = __ldg(p.adj_rowptr + b);
        const int na = (int)(__ldg(p.adj_rowptr + a + 1) - a0), nb = (int)(__ldg(p.adj_rowptr + b + 1) - b0);
        const int64_t pa0 = __ldg(p.ppr_rowptr + a), pb0 = __ldg(p.ppr_rowptr + b);
        const int npa = (int)(__ldg(p.ppr_rowptr + a + 1) - pa0), npb = (int)(__ldg(p.ppr_rowptr + b + 1) - pb0);
        const int32_t* Aa = p.adj_col + a0;
        const int32_t* Ab = p.adj_col + b0;
        const int32_t* Pac = p.ppr_col + pa0;
        const int32_t* Pbc = p.ppr_col + pb0;
        const float* Pav = p.ppr_val + pa0;
        const float* Pbv = p.ppr_val + pb0;

        int c_cn = 0, c_1h = 0, c_n1 = 0;
        {   // CN: every thread probes one element of the shorter adjacency row per step
            const bool a_short = na <= nb;
            const int32_t* S = a_short ? Aa : Ab;
            const int32_t* Lg = a_short ? Ab : Aa;
            const int ns = a_short ? na : nb, nl = a_short ? nb : na;
            int lo = 0;   // this thread's elements ascend, so its search window only moves right
            for (int k = 0; k < ns; k += kHeavyThreads) {
                const bool act = k + (int)threadIdx.x < ns;
                const int32_t u = act ? __ldg(S + k + threadIdx.x) : 0x7fffffff;
                bool hit = false;
                float qa = 0.f, qb = 0.f;
                if (act) {
                    lo = lower_bound_from(Lg, lo, nl, u);
                    hit = lo < nl && __ldg(Lg + lo) == u;
                    if (hit && cn_needs_ppr) {
                        int t = lower_bound_from(Pac, 0, npa, u);
                        if (t < npa && __ldg(Pac + t) == u) qa = quantise(__ldg(Pav + t));
                        t = lower_bound_from(Pbc, 0, npb, u);
                        if (t < npb && __ldg(Pbc + t) == u) qb = quantise(__ldg(Pbv + t));
                        hit = qa >= p.th_cn && qb >= p.th_cn;
                    }
                }
                const int r = block_rank(hit, wt, c_cn);
                if (FILL && hit) {
                    const int64_t s = o_cn + r;
                    p.node[s] = u;
                    p.pa[s] = qa;
                    p.pb[s] = qb;
                    if (p.link) p.link[s] = (int32_t)i;
                }
            }
        }
        if (want_pi) {   // 1-hop / >1-hop from the intersection of the two PPR rows
            const bool a_short = npa <= npb;
            const int32_t* Sc = a_short ? Pac : Pbc;
            const float* Sv = a_short ? Pav : Pbv;
            const int32_t* Lc = a_short ? Pbc : Pac;
            const float* Lv = a_short ? Pbv : Pav;
            const int ns = a_short ? npa : npb, nl = a_short ? npb : npa;
            int lo = 0;
            for (int k = 0; k < ns; k += kHeavyThreads) {
                const bool act = k + (int)threadIdx.x < ns;
                const int32_t u = act ? __ldg(Sc + k + threadIdx.x) : 0x7fffffff;
                bool k1 = false, kn = false;
                float qa = 0.f, qb = 0.f;
                if (act) {
                    lo = lower_bound_from(Lc, lo, nl, u);
                    if (lo < nl && __ldg(Lc + lo) == u) {
                        const float qs = quantise(__ldg(Sv + k + threadIdx.x)), ql = quantise(__ldg(Lv + lo));
                        qa = a_short ? qs : ql;
                        qb = a_short ? ql : qs;
                        if (qa >= th_pre && qb >= th_pre) {
                            int t = lower_bound_from(Aa, 0, na, u);
                            const bool in_a = t < na && __ldg(Aa + t) == u;
                            t = lower_bound_from(Ab, 0, nb, u);
                            const bool in_b = t < nb && __ldg(Ab + t) == u;
                            k1 = (in_a != in_b) && qa >= p.th_1hop && qb >= p.th_1hop;
                            kn = want_n1 && !in_a && !in_b && qa >= p.th_non1hop && qb >= p.th_non1hop;
                        }
                    }
                }
                const int r1 = block_rank(k1, wt, c_1h);
                const int rn = block_rank(kn, wt, c_n1);
                if (FILL && (k1 || kn)) {
                    const int64_t s = k1 ? o_1h + r1 : o_n1 + rn;
                    p.node[s] = u;
                    p.pa[s] = qa;
                    p.pb[s] = qb;
                    if (p.link) p.link[s] = (int32_t)i;
                }
            }
        }
        if (!FILL && threadIdx.x == 0) {
            p.counts[i] = c_cn;
            p.counts[p.bs + i] = c_1h;
            p.counts[2 * p.bs + i] = c_n1;
        }
    }
}

__global__ void select_reset_heavy(int32_t* heavy) { heavy[0] = 0; }

template <int G>
static int launch_fast(bool fill, const SelectParams2& p, cudaStream_t st) {
    const int64_t groups_per_block = 256 / G;
    int64_t blocks = (p.bs + groups_per_block - 1) / groups_per_block;
    const int64_t cap = (int64_t)kNumSMs * 8 * 8;
    if (blocks > cap) blocks = cap;
    const unsigned hgrid = kNumSMs * 2;
    if (fill) {
        select_fast_kernel<G, true><<<(unsigned)blocks, 256, 0, st>>>(p);
        select_heavy_kernel<true><<<hgrid, kHeavyThreads, 0, st>>>(p);
    } else {
        select_reset_heavy<<<1, 1, 0, st>>>(p.heavy);
        select_fast_kernel<G, false><<<(unsigned)blocks, 256, 0, st>>>(p);
        select_heavy_kernel<false><<<hgrid, kHeavyThreads, 0, st>>>(p);
    }
    return check_launch(fill ? "lpf_select_fill" : "lpf_select_count");
}

int select_fast(bool fill, int group, const int64_t* links, int64_t bs, const int64_t* adj_rowptr,
                const int32_t* adj_col, const int64_t* ppr_rowptr, const int32_t* ppr_col, const float* ppr_val,
                float th_cn, float th_1hop, float th_non1hop, int mode, int32_t* counts, const int64_t* ptr,
                int32_t* node, float* pa, float* pb, int32_t* link, int32_t* heavy, cudaStream_t st) {
    SelectParams2 p{links, bs, adj_rowptr, adj_col, ppr_rowptr, ppr_col, ppr_val, th_cn, th_1hop, th_non1hop,
                    mode, counts, ptr, node, pa, pb, link, heavy};
    if (group == 32) return launch_fast<32>(fill, p, st);
    return launch_fast<8>(fill, p, st);
}

}  // namespace lpf
