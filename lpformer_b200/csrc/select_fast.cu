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
#include "select_walk.cuh"

namespace lpf {


// ---- two-pass interface (count / fill with caller-side scan): output ordered by (type, link, node)
template <int G, bool FILL>
__global__ void __launch_bounds__(256) select_fast_kernel(SelectParams2 p) {
    const int lane = threadIdx.x & 31;
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
    const bool want_pi = p.mode != LPF_MODE_CN;
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
        const LinkRows r = load_rows(p, i);
        if (is_heavy(r, want_pi, G)) {
            if (!FILL && (lane & (G - 1)) == 0) p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
            continue;
        }
        int c_cn, c_1h, c_n1;
        walk_link<G, FILL>(p, r, i, lane, o_cn, o_1h, o_n1, c_cn, c_1h, c_n1);
        if (!FILL && (lane & (G - 1)) == 0) {
            p.counts[i] = c_cn;
            p.counts[p.bs + i] = c_1h;
            p.counts[2 * p.bs + i] = c_n1;
        }
    }
}

// ---- one-pass interface: count, allocate, write in the same launch; a link's pairs are contiguous and
// ascending within each type pool, links appear in arbitrary order.  No scan, no second launch, no host sync.
template <int G>
__global__ void __launch_bounds__(256) select_onepass_kernel(SelectParams2 p);

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

// A whole CTA walks one (heavy) link.
template <bool WRITE>
__device__ __forceinline__ void walk_link_cta(const SelectParams2& p, const LinkRows& r, int64_t i, int* wt, int64_t o_cn,
                                              int64_t o_1h, int64_t o_n1, int& c_cn, int& c_1h, int& c_n1) {
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const bool cn_needs_ppr = WRITE || p.th_cn > 0.0f;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    c_cn = c_1h = c_n1 = 0;
    {   // CN: every thread probes one element of the shorter adjacency row per step
        const bool a_short = r.na <= r.nb;
        const int32_t* S = a_short ? r.Aa : r.Ab;
        const int32_t* Lg = a_short ? r.Ab : r.Aa;
        const int ns = a_short ? r.na : r.nb, nl = a_short ? r.nb : r.na;
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
                    int t = lower_bound_from(r.Pac, 0, r.npa, u);
                    if (t < r.npa && __ldg(r.Pac + t) == u) qa = quantise(__ldg(r.Pav + t));
                    t = lower_bound_from(r.Pbc, 0, r.npb, u);
                    if (t < r.npb && __ldg(r.Pbc + t) == u) qb = quantise(__ldg(r.Pbv + t));
                    hit = qa >= p.th_cn && qb >= p.th_cn;
                }
            }
            const int rk = block_rank(hit, wt, c_cn);
            if (WRITE && hit) {
                const int64_t s = o_cn + rk;
                p.node[s] = u;
                p.pa[s] = qa;
                p.pb[s] = qb;
                if (p.link) p.link[s] = (int32_t)i;
            }
        }
    }
    if (want_pi) {   // 1-hop / >1-hop from the intersection of the two PPR rows
        const bool a_short = r.npa <= r.npb;
        const int32_t* Sc = a_short ? r.Pac : r.Pbc;
        const float* Sv = a_short ? r.Pav : r.Pbv;
        const int32_t* Lc = a_short ? r.Pbc : r.Pac;
        const float* Lv = a_short ? r.Pbv : r.Pav;
        const int ns = a_short ? r.npa : r.npb, nl = a_short ? r.npb : r.npa;
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
                        int t = lower_bound_from(r.Aa, 0, r.na, u);
                        const bool in_a = t < r.na && __ldg(r.Aa + t) == u;
                        t = lower_bound_from(r.Ab, 0, r.nb, u);
                        const bool in_b = t < r.nb && __ldg(r.Ab + t) == u;
                        k1 = (in_a != in_b) && qa >= p.th_1hop && qb >= p.th_1hop;
                        kn = want_n1 && !in_a && !in_b && qa >= p.th_non1hop && qb >= p.th_non1hop;
                    }
                }
            }
            const int r1 = block_rank(k1, wt, c_1h);
            const int rn = block_rank(kn, wt, c_n1);
            if (WRITE && (k1 || kn)) {
                const int64_t s = k1 ? o_1h + r1 : o_n1 + rn;
                p.node[s] = u;
                p.pa[s] = qa;
                p.pb[s] = qb;
                if (p.link) p.link[s] = (int32_t)i;
            }
        }
    }
}

template <bool FILL>
__global__ void __launch_bounds__(kHeavyThreads) select_heavy_kernel(SelectParams2 p) {
    __shared__ int wt[kHeavyThreads / 32];
    const int nheavy = p.heavy[0];
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
        const LinkRows r = load_rows(p, i);
        int c_cn, c_1h, c_n1;
        walk_link_cta<FILL>(p, r, i, wt, o_cn, o_1h, o_n1, c_cn, c_1h, c_n1);
        if (!FILL && threadIdx.x == 0) {
            p.counts[i] = c_cn;
            p.counts[p.bs + i] = c_1h;
            p.counts[2 * p.bs + i] = c_n1;
        }
    }
}

// count -> allocate -> write of one link by one group of G lanes (generic walk)
template <int G>
__device__ __forceinline__ void onepass_link_generic(const SelectParams2& p, const LinkRows& r, int64_t i, int lane) {
    const int leader = lane & ~(G - 1);
    const unsigned gmask = group_mask<G>(lane);
    int c_cn, c_1h, c_n1;
    walk_link<G, false>(p, r, i, lane, 0, 0, 0, c_cn, c_1h, c_n1);
    int64_t s_cn = 0, s_1h = 0, s_n1 = 0;
    int ok = 1;
    if (lane == leader) ok = alloc_segments(p, i, c_cn, c_1h, c_n1, s_cn, s_1h, s_n1) ? 1 : 0;
    if (c_cn + c_1h + c_n1 == 0) return;       // uniform within the group
    ok = __shfl_sync(gmask, ok, leader);
    if (!ok) return;
    s_cn = __shfl_sync(gmask, s_cn, leader);
    s_1h = __shfl_sync(gmask, s_1h, leader);
    s_n1 = __shfl_sync(gmask, s_n1, leader);
    walk_link<G, true>(p, r, i, lane, s_cn, p.cap + s_1h, 2 * p.cap + s_n1, c_cn, c_1h, c_n1);
}

// Deferred links, two tiers: a full warp walks a link whose shorter row has up to kHugeRow elements (hundreds per
// citation2-shaped batch: hub-hub pairs); a whole CTA walks the few that are longer still.
constexpr int kHugeRow = 256;
__device__ __forceinline__ bool is_huge(const LinkRows& r, bool want_pi) {
    return max(min(r.na, r.nb), want_pi ? min(r.npa, r.npb) : 0) > kHugeRow;
}

__global__ void __launch_bounds__(kHeavyThreads) select_heavy_onepass_kernel(const __grid_constant__ SelectParams2 p) {
    __shared__ int wt[kHeavyThreads / 32];
    __shared__ int64_t seg[3];
    __shared__ int ok_s;
    const int nheavy = p.heavy[0];
    const bool want_pi = p.mode != LPF_MODE_CN;
    const int lane = threadIdx.x & 31;
    // tier 1: one warp per deferred link
    const int warp_global = (blockIdx.x * kHeavyThreads + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kHeavyThreads) >> 5;
    for (int q = warp_global; q < nheavy; q += nwarps) {
        const int64_t i = p.heavy[4 + q];
        const LinkRows r = load_rows(p, i);
        if (is_huge(r, want_pi)) continue;
        onepass_link_generic<32>(p, r, i, lane);
    }
    __syncthreads();
    // tier 2: one CTA per huge link
    for (int q = blockIdx.x; q < nheavy; q += gridDim.x) {
        const int64_t i = p.heavy[4 + q];
        const LinkRows r = load_rows(p, i);
        if (!is_huge(r, want_pi)) continue;      // uniform across the CTA
        int c_cn, c_1h, c_n1;
        walk_link_cta<false>(p, r, i, wt, 0, 0, 0, c_cn, c_1h, c_n1);
        if (threadIdx.x == 0) {
            int64_t s0, s1, s2;
            ok_s = alloc_segments(p, i, c_cn, c_1h, c_n1, s0, s1, s2) ? 1 : 0;
            seg[0] = s0; seg[1] = s1; seg[2] = s2;
        }
        __syncthreads();
        const bool go = ok_s && (c_cn + c_1h + c_n1 > 0);
        const int64_t s0 = seg[0], s1 = seg[1], s2 = seg[2];
        __syncthreads();
        if (go) walk_link_cta<true>(p, r, i, wt, s0, p.cap + s1, 2 * p.cap + s2, c_cn, c_1h, c_n1);
    }
}

template <int G>
__global__ void __launch_bounds__(256) select_onepass_kernel(SelectParams2 p) {
    const int lane = threadIdx.x & 31;
    const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
    const bool want_pi = p.mode != LPF_MODE_CN;
    for (int64_t i = group; i < p.bs; i += ngroups) {
        const LinkRows r = load_rows(p, i);
        if (is_heavy(r, want_pi, G)) {
            if ((lane & (G - 1)) == 0) p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
            continue;
        }
        onepass_link_generic<G>(p, r, i, lane);
    }
}

__global__ void select_reset_heavy(int32_t* heavy) { heavy[0] = 0; }
// a full pool invalidates the batch: later launches of the same stream / graph see empty sizes and do nothing,
// the host finds header[4] != 0 and re-runs the batch through the two-pass interface
__global__ void select_finalize_onepass(int64_t* hdr) {
    if (hdr[4]) hdr[0] = hdr[1] = hdr[2] = hdr[3] = 0;
}
__global__ void select_reset_onepass(int32_t* heavy, int64_t* hdr) {
    heavy[0] = 0;
    hdr[0] = hdr[1] = hdr[2] = hdr[3] = hdr[4] = 0;
}

// prologue / epilogue of a one-pass launch sequence, for the kernels of other translation units
__global__ void select_reset_packed(int32_t* heavy, int64_t* hdr, int32_t* hub) {
    heavy[0] = 0;
    heavy[1] = 0;        // piece counter of the screening launch
    hdr[0] = hdr[1] = hdr[2] = hdr[3] = hdr[4] = 0;
    hub[0] = 0;
    hub[1] = 0;          // links for the CTA-wide packed resolution
}
void launch_onepass_reset(const SelectParams2& p, cudaStream_t st) {
    if (p.hub) select_reset_packed<<<1, 1, 0, st>>>(p.heavy, p.hdr, p.hub);
    else select_reset_onepass<<<1, 1, 0, st>>>(p.heavy, p.hdr);
}
void launch_onepass_heavy(const SelectParams2& p, cudaStream_t st) {
    select_heavy_onepass_kernel<<<kNumSMs * 2, kHeavyThreads, 0, st>>>(p);
}
void launch_onepass_finalize(const SelectParams2& p, cudaStream_t st) { select_finalize_onepass<<<1, 1, 0, st>>>(p.hdr); }
void launch_onepass_tail(const SelectParams2& p, cudaStream_t st) {
    launch_onepass_heavy(p, st);
    launch_onepass_finalize(p, st);
}

template <int G>
static int launch_fast(bool fill, const SelectParams2& p, cudaStream_t st) {
    const int64_t groups_per_block = 256 / G;
    int64_t blocks = (p.bs + groups_per_block - 1) / groups_per_block;
    const int64_t cap = (int64_t)kNumSMs * 8 * 8;
    if (blocks > cap) blocks = cap;
    const unsigned hgrid = kNumSMs * 2;
    if (p.hdr) {
        select_reset_onepass<<<1, 1, 0, st>>>(p.heavy, p.hdr);
        select_onepass_kernel<G><<<(unsigned)blocks, 256, 0, st>>>(p);
        select_heavy_onepass_kernel<<<hgrid, kHeavyThreads, 0, st>>>(p);
        select_finalize_onepass<<<1, 1, 0, st>>>(p.hdr);
        return check_launch("lpf_select_onepass");
    }
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
                    mode, counts, ptr, node, pa, pb, link, heavy, 0, nullptr, nullptr, nullptr, nullptr};
    if (group == 32) return launch_fast<32>(fill, p, st);
    return launch_fast<8>(fill, p, st);
}

long long* g_select_dbg = nullptr;

int select_onepass(int group, const int64_t* links, int64_t bs, const int64_t* adj_rowptr, const int32_t* adj_col,
                   const int64_t* ppr_rowptr, const int32_t* ppr_col, const float* ppr_val, float th_cn, float th_1hop,
                   float th_non1hop, int mode, int64_t cap, int32_t* counts, int32_t* seg_start, int32_t* nz_list,
                   int64_t* hdr, int32_t* node, float* pa, float* pb, int32_t* heavy, cudaStream_t st) {
    SelectParams2 p{links, bs, adj_rowptr, adj_col, ppr_rowptr, ppr_col, ppr_val, th_cn, th_1hop, th_non1hop,
                    mode, counts, nullptr, node, pa, pb, nullptr, heavy, cap, hdr, seg_start, nz_list, g_select_dbg};
    if (group == 32) return launch_fast<32>(false, p, st);
    return launch_fast<8>(false, p, st);
}

}  // namespace lpf
