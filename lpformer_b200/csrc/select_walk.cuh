// Shared device code of the K1 intersection kernels (select_fast.cu, select_packed.cu): parameters, row access,
// the group walk of one link and the one-pass segment allocation.
#pragma once
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
    // one-pass mode (select_onepass_kernel): pairs of type t go to rows [t*cap, t*cap + hdr[t]) of the pair arrays
    int64_t cap;       // rows reserved per type
    int64_t* hdr;      // [0..2] pairs per type, [3] non-empty links, [4] overflow flag
    int32_t* seg_start;  // [3*bs] first row of link i's type-t segment, relative to t*cap
    int32_t* nz_list;    // [bs] batch positions of the non-empty links (any order)
    long long* dbg;      // optional profiling buffer (lpf_debug_select_clocks): per-phase clock64() totals
    int32_t* hub = nullptr;   // packed kernel: candidate list ([0] = count, [4..] batch positions of the links the screening flagged)
    int slot_limit = 0;       // packed kernel, tests only (lpf_debug_select_slots): usable hash slots, 0 = all
};

// Workspace of the one-pass launch sequences, in int32 words: [0, bs + 4) deferred (heavy) links; then the candidate
// list of the packed kernel (count in word 0, entries from word 4) and, another ws_list_words further, the list of
// the candidates whose shorter row takes a whole CTA (count in word 1 of the candidate list, entries from word 0).
__host__ __device__ inline int64_t ws_list_words(int64_t bs) { return (bs + 4 + 3) / 4 * 4; }
__host__ __device__ inline int64_t ws_hub_words(int64_t bs) { return 2 * ws_list_words(bs); }
__host__ __device__ inline int64_t ws_total_words(int64_t bs) { return (bs + 4) + ws_hub_words(bs); }

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


// Rows of one link.
struct LinkRows {
    const int32_t* Aa; const int32_t* Ab; const int32_t* Pac; const int32_t* Pbc; const float* Pav; const float* Pbv;
    int na, nb, npa, npb;
};
__device__ __forceinline__ LinkRows load_rows(const SelectParams2& p, int64_t i) {
    const int64_t a = __ldg(p.links + i), b = __ldg(p.links + p.bs + i);
    const int64_t a0 = __ldg(p.adj_rowptr + a), a1 = __ldg(p.adj_rowptr + a + 1);
    const int64_t b0 = __ldg(p.adj_rowptr + b), b1 = __ldg(p.adj_rowptr + b + 1);
    const int64_t pa0 = __ldg(p.ppr_rowptr + a), pa1 = __ldg(p.ppr_rowptr + a + 1);
    const int64_t pb0 = __ldg(p.ppr_rowptr + b), pb1 = __ldg(p.ppr_rowptr + b + 1);
    LinkRows r;
    r.na = (int)(a1 - a0); r.nb = (int)(b1 - b0); r.npa = (int)(pa1 - pa0); r.npb = (int)(pb1 - pb0);
    r.Aa = p.adj_col + a0; r.Ab = p.adj_col + b0;
    r.Pac = p.ppr_col + pa0; r.Pbc = p.ppr_col + pb0; r.Pav = p.ppr_val + pa0; r.Pbv = p.ppr_val + pb0;
    return r;
}
__device__ __forceinline__ bool is_heavy(const LinkRows& r, bool want_pi, int lanes) {
    return max(min(r.na, r.nb), want_pi ? min(r.npa, r.npb) : 0) > kHeavyPerLane * lanes;
}

// One group of G lanes walks one link: counts its three sets and, with WRITE, stores the pairs at rows
// o_cn / o_1h / o_n1 (ascending node id within each set).
template <int G, bool WRITE>
__device__ __forceinline__ void walk_link(const SelectParams2& p, const LinkRows& r, int64_t i, int lane, int64_t o_cn,
                                          int64_t o_1h, int64_t o_n1, int& c_cn, int& c_1h, int& c_n1) {
    const int gl = lane & (G - 1);
    const unsigned gmask = group_mask<G>(lane);
    const unsigned lt = gmask & ((1u << lane) - 1u);
    const int last = (lane & ~(G - 1)) + G - 1;
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const bool cn_needs_ppr = WRITE || p.th_cn > 0.0f;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    c_cn = c_1h = c_n1 = 0;
    {   // ---- CN: walk the shorter adjacency row, search the longer
        const bool a_short = r.na <= r.nb;
        const int32_t* S = a_short ? r.Aa : r.Ab;
        const int32_t* Lg = a_short ? r.Ab : r.Aa;
        const int ns = a_short ? r.na : r.nb, nl = a_short ? r.nb : r.na;
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
                int t = lower_bound_from(r.Pac, 0, r.npa, u);
                if (t < r.npa && __ldg(r.Pac + t) == u) qa = quantise(__ldg(r.Pav + t));
                t = lower_bound_from(r.Pbc, 0, r.npb, u);
                if (t < r.npb && __ldg(r.Pbc + t) == u) qb = quantise(__ldg(r.Pbv + t));
                hit = qa >= p.th_cn && qb >= p.th_cn;
            }
            const unsigned m = __ballot_sync(gmask, hit);
            if (WRITE && hit) {
                const int64_t s = o_cn + c_cn + __popc(m & lt);
                p.node[s] = u;
                p.pa[s] = qa;
                p.pb[s] = qb;
                if (p.link) p.link[s] = (int32_t)i;
            }
            c_cn += __popc(m);
            // galloping: later elements are larger, so they cannot sit before the last lane's position
            lo = __shfl_sync(gmask, pos, last);
        }
    }
    if (want_pi) {   // ---- 1-hop / >1-hop from the intersection of the two PPR rows
        const bool a_short = r.npa <= r.npb;
        const int32_t* Sc = a_short ? r.Pac : r.Pbc;
        const float* Sv = a_short ? r.Pav : r.Pbv;
        const int32_t* Lc = a_short ? r.Pbc : r.Pac;
        const float* Lv = a_short ? r.Pbv : r.Pav;
        const int ns = a_short ? r.npa : r.npb, nl = a_short ? r.npb : r.npa;
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
                        int t = lower_bound_from(r.Aa, 0, r.na, u);
                        const bool in_a = t < r.na && __ldg(r.Aa + t) == u;
                        t = lower_bound_from(r.Ab, 0, r.nb, u);
                        const bool in_b = t < r.nb && __ldg(r.Ab + t) == u;
                        k1 = (in_a != in_b) && qa >= p.th_1hop && qb >= p.th_1hop;
                        kn = want_n1 && !in_a && !in_b && qa >= p.th_non1hop && qb >= p.th_non1hop;
                    }
                }
            }
            const unsigned m1 = __ballot_sync(gmask, k1);
            const unsigned mn = __ballot_sync(gmask, kn);
            if (WRITE && (k1 || kn)) {
                const int64_t s = k1 ? o_1h + c_1h + __popc(m1 & lt) : o_n1 + c_n1 + __popc(mn & lt);
                p.node[s] = u;
                p.pa[s] = qa;
                p.pb[s] = qb;
                if (p.link) p.link[s] = (int32_t)i;
            }
            c_1h += __popc(m1);
            c_n1 += __popc(mn);
            lo = __shfl_sync(gmask, pos, last);
        }
    }
}

// The same allocation by the first four lanes of a converged warp (lane t < 3: the pool of type t, lane 3: the slot
// in the list of non-empty links): the four atomics are in flight together — one round trip to L2 instead of up
// to four on the critical path of every resolved link.  (The list slot is taken before the capacity check: when a
// pool overflows the batch is redone anyway and select_finalize_onepass zeroes the sizes.)  Every lane of the warp
// calls this with the same counts and gets the three segment starts; returns false if a pool is full.
__device__ __forceinline__ bool alloc_segments_warp(const SelectParams2& p, int64_t i, int c_cn, int c_1h, int c_n1,
                                                    int lane, int64_t& s_cn, int64_t& s_1h, int64_t& s_n1) {
    const int total = c_cn + c_1h + c_n1;
    const int mine = lane == 0 ? c_cn : (lane == 1 ? c_1h : (lane == 2 ? c_n1 : (total > 0 ? 1 : 0)));
    long long got = 0;
    if (lane < 4 && mine > 0)
        got = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(p.hdr + lane), (unsigned long long)mine);
    s_cn = __shfl_sync(kFull, got, 0);
    s_1h = __shfl_sync(kFull, got, 1);
    s_n1 = __shfl_sync(kFull, got, 2);
    const long long nzpos = __shfl_sync(kFull, got, 3);
    const bool fits = s_cn + c_cn <= p.cap && s_1h + c_1h <= p.cap && s_n1 + c_n1 <= p.cap;
    if (lane < 3) {
        const int c = lane == 0 ? c_cn : (lane == 1 ? c_1h : c_n1);
        p.counts[lane * p.bs + i] = c;
        p.seg_start[lane * p.bs + i] = (int32_t)got;
    } else if (lane == 3 && total > 0) {
        if (fits) p.nz_list[nzpos] = (int32_t)i;
        else p.hdr[4] = 1;
    }
    return fits;
}

// Segment allocation of the one-pass mode: rows for (c_cn, c_1h, c_n1) pairs in the three per-type pools.
// Called by ONE thread; returns false (and raises the overflow flag) if a pool is full.
__device__ __forceinline__ bool alloc_segments(const SelectParams2& p, int64_t i, int c_cn, int c_1h, int c_n1,
                                               int64_t& s_cn, int64_t& s_1h, int64_t& s_n1) {
    auto take = [&](int t, int c) -> int64_t {
        return c > 0 ? (int64_t)atomicAdd(reinterpret_cast<unsigned long long*>(p.hdr + t), (unsigned long long)c) : 0;
    };
    s_cn = take(0, c_cn);
    s_1h = take(1, c_1h);
    s_n1 = take(2, c_n1);
    p.counts[i] = c_cn;
    p.counts[p.bs + i] = c_1h;
    p.counts[2 * p.bs + i] = c_n1;
    p.seg_start[i] = (int32_t)s_cn;
    p.seg_start[p.bs + i] = (int32_t)s_1h;
    p.seg_start[2 * p.bs + i] = (int32_t)s_n1;
    if (c_cn + c_1h + c_n1 == 0) return true;
    const bool fits = s_cn + c_cn <= p.cap && s_1h + c_1h <= p.cap && s_n1 + c_n1 <= p.cap;
    if (!fits) {
        p.hdr[4] = 1;
        return false;
    }
    p.nz_list[atomicAdd(reinterpret_cast<unsigned long long*>(p.hdr + 3), 1ull)] = (int32_t)i;
    return true;
}

}  // namespace lpf
