// K1, one-pass selection over PACKED LINK ROWS: the HBM layout built for the per-link walk.
//
// The CSR tables cost a link four dependent random reads per endpoint (two rowptr pairs, then the adjacency row,
// the PPR columns and — on a match — the PPR values, all in different arrays).  lpf_pack_link_rows rewrites them
// once per graph as
//   locator[x] (uint32) = (first 64-byte unit of row x) << 6 | min(units, 63)       4 B/node: the whole array stays
//                                                                                   in the 126 MB L2
//   row x (64-byte aligned, `units` x 64 B) = header (deg, nP, 0, 0), then 8-byte SLOTS: the nP PPR entries as
//           (col | 0x80000000, value bits), then the neighbour ids two per slot, padded with 0x7fffffff
// so a target is ONE DRAM round trip: the locator read hits L2, and it tells where every 16-byte chunk of the row
// lies, so all of them are fetched at once.  A slot says what it is (bit 31 of its first word), hence a chunk can be
// screened without the row's header.
//
// Work shape (citation2-style evaluation: runs of links sharing their source, reference train/testing.py:20-23):
// the batch is cut evenly over one resident wave of CTAs (296 x 512 threads on a B200); a CTA's piece is at most
// 1,024 consecutive links = at most kPkMaxRuns runs of equal source.  The sources' adjacency rows become bucketed
// hash sets in shared memory and their PPR rows small shared-memory tables (all runs of the piece side by side).
// The piece's target rows are then flattened into 64-byte UNITS; FOUR LANES take one unit (one coalesced 64-byte
// read), every unit of every link is independent of every other, and each lane keeps four reads in flight.
// A unit only answers "does this link select anything?" (a common neighbour, or a node in both PPR rows above the
// smaller PPR threshold); 99 % of a citation2-shaped batch selects nothing and is finished there.  The links that
// do are resolved — count -> allocate -> ordered write — from their packed row against the staged source: by a
// warp each (phase B; the four header atomics of a link issued by four lanes at once), by the whole CTA for long
// target rows (phase C: one block-wide scan per 2,048 slots, the hits kept in registers across the allocation).
// Runs whose source does not fit next to the others of its piece go to a second launch of the same kernel with a
// 32K-slot table (in pieces of 256 links); sources beyond even that are searched in global memory and resolved over
// the CSR tables; pieces that are not run-shaped take the generic group walk of select_walk.cuh.  Selected sets,
// their order inside a link and the fp32 values are those of every other K1 variant (tests compare all of them
// with the oracle).  Measured (tools/select_clocks.py): a piece is a ~45 us chain of dependent phases of which only
// the screening (~18 us, ~3 TB/s of random 64-byte reads) is bandwidth-like; the launch lasts as long as its slowest
// piece (~100 us: a hub source selects something with every tenth target).
#include <stdlib.h>

#include "select_hashed.cuh"

namespace lpf {

constexpr int kPkThreads = 512;       // threads per CTA
constexpr int kPkChunk = 1024;        // most links of one chunk: the batch is cut evenly over the resident CTAs
constexpr int kPkMaxRuns = 3;
constexpr int kPkHashSlots = 16384;   // int32 slots shared by the runs of a chunk (load <= 0.5): sources up to 8,192
constexpr int kPkHubSlots = 32768;    // second launch, one CTA per SM: hub sources up to 16,384 neighbours
constexpr int kPkMaxPprRow = 128;
constexpr int kPkMaxUnits = 62;       // target rows of up to 62 units (4 KB: ~980 neighbours) are screened unit-wise (the locator's
                                      // unit count saturates at 63: the few longer rows are resolved unscreened by the whole CTA)
constexpr int kPkMaxItems = 6144;     // units of one chunk that are screened unit-wise (the rest get a warp)
constexpr int kPkWarpUnits = 32;      // resolution by a warp of the screening CTA: target rows up to 32 units (~500 neighbours);
                                      // longer rows that select take the whole CTA
constexpr int kPkHubDeg = 1 << 30;    // sources with more neighbours would be cut into pieces for the hub launch (a hub source selects
                                      // something with every tenth target: ~100 links to resolve in its piece); off: with 1,500 the hub
                                      // launch took 59 us and the step was no shorter
constexpr int kPkSplit = 1;           // pieces per resident CTA (2 measured: the slowest CTA is no faster, more total work)
constexpr int kPkMinPiece = 512;
constexpr int kPkInflight = 4;        // 16-byte reads a lane of the screening keeps in flight (8 spill at 64 registers: slower)
constexpr int kPkHubPiece = 256;      // links per entry of the hub list
static_assert(kPkChunk == 1024 && kPkMaxUnits < 65536, "items[] packs (position:16 | unit:16)");
constexpr uint32_t kPkPprTag = 0x80000000u;
constexpr uint32_t kPkPad = 0x7fffffffu;

struct PkRunTab {
    int32_t pac[kPkMaxPprRow];
    float pav[kPkMaxPprRow];
    int32_t ppos[kPprHashSlots];
};

template <int SLOTS>
struct PkSmemT {
    int32_t tab[SLOTS];
    PkRunTab run[kPkMaxRuns];
    uint32_t l_loc[kPkChunk];            // locator of every link's target row
    uint32_t items[kPkMaxItems];         // (chunk position << 16 | unit) of every unit to screen
    uint16_t q_slow[kPkChunk];           // chunk positions of the links a warp resolves here (they select something)
    uint8_t l_any[kPkChunk];             // "selects something" flags raised by the screening
    int32_t warp_tot[kPkThreads / 32];
    int32_t run_start[kPkMaxRuns + 1];
    int32_t r_tab0[kPkMaxRuns], r_lg[kPkMaxRuns], r_na[kPkMaxRuns], r_npa[kPkMaxRuns], r_hashed[kPkMaxRuns];
    uint32_t r_loc[kPkMaxRuns];
    int32_t r_slots[kPkMaxRuns];
    int4 r_ctx[kPkMaxRuns];              // (first table slot, bucket mask, 32 - log2(buckets), 0): one read per unit
    int n_runs, n_slow, n_items, tab_used, items_full, n_cta, cta_ok, cur_chunk, n_mid, b_next;
    uint16_t q_mid[kPkChunk];             // ... of the links a warp resolves whose target row has more than four units
    uint16_t q_cta[kPkChunk];             // chunk positions of the links the whole CTA walks (long target rows that select)
    uint32_t scan_tot[2 * 4 * (kPkThreads / 32)];
    int64_t cta_seg[3];
    uint2 cta_ppr[kPkMaxPprRow];         // PPR slots of the row the whole CTA is walking
    int dbg_ph[16];                      // profiling: this chunk's cycles per phase
};

// 64-byte units of a packed row: header (16 B) + 8-byte slots (PPR entries, then the neighbour ids two per slot)
__host__ __device__ __forceinline__ int64_t row_units(int64_t deg, int64_t npp) {
    return (16 + 8 * (npp + ((deg + 1) >> 1)) + 63) >> 6;
}

__device__ __forceinline__ uint4 ldg16(const uint4* p) { return __ldg(p); }
__device__ __forceinline__ const uint4* row_of(const uint4* __restrict__ blob, uint32_t loc) {
    return blob + (size_t)(loc >> 6) * 4;          // 64-byte units -> 16-byte chunks
}

template <class SM>
__device__ __forceinline__ RunCtx make_ctx(const SM& sm, int r) {
    RunCtx h;
    const int lgb = sm.r_lg[r] - 2;             // buckets of four slots
    h.tab = sm.tab + sm.r_tab0[r];
    h.mask = (1u << lgb) - 1u;
    h.shift = 32 - lgb;
    h.pac = sm.run[r].pac;
    h.pav = sm.run[r].pav;
    h.ppos = sm.run[r].ppos;
    h.npa = sm.r_npa[r];
    return h;
}

// One 8-byte slot of a target row against the staged source: a PPR entry (both values above the smaller PPR
// threshold = a candidate 1-hop / >1-hop node) or two neighbour ids (either one in A(a) = a common neighbour).
// HASHED = false: the source's adjacency row was not staged (a source beyond even the hub table); its ids in its
// packed row in global memory (ascending) are searched instead.
template <bool HASHED>
__device__ __forceinline__ bool screen_slot(const RunCtx& h, const int32_t* __restrict__ arow_ids, int na, uint32_t w0,
                                            uint32_t w1, bool want_pi, float th_pre) {
    if (w0 & kPkPprTag) {
        float qa;
        return want_pi && smem_ppr_lookup(h, (int32_t)(w0 & ~kPkPprTag), qa) && qa >= th_pre &&
               quantise(__uint_as_float(w1)) >= th_pre;
    }
    if (w0 == kPkPad) return false;          // ids ascend: a pad in front means the slot is all padding
    if (HASHED) return hash_contains_any2(h.tab, h.mask, h.shift, (int32_t)w0, (int32_t)w1);
    int t = lower_bound_from(arow_ids, 0, na, (int32_t)w0);
    if (t < na && __ldg(arow_ids + t) == (int32_t)w0) return true;
    t = lower_bound_from(arow_ids, 0, na, (int32_t)w1);
    return t < na && __ldg(arow_ids + t) == (int32_t)w1;
}

// Rare, register-hungry paths kept out of line so that they do not set the register budget of the screening.
__device__ __noinline__ void generic_link8(const SelectParams2& p, int64_t i, int lane) {
    const LinkRows r = load_rows(p, i);
    if (is_heavy(r, p.mode != LPF_MODE_CN, 8)) {
        if ((lane & 7) == 0) p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
        return;
    }
    onepass_link<8>(p, r, i, lane);
}
// What slot s of a PACKED target row contributes against the staged source: a PPR entry is a candidate 1-hop /
// >1-hop node (k1 / kn, node u, values qa, qb), a pair of neighbour ids up to two common neighbours (h0, h1 with
// values (qa, qb) and (qa1, qb1)).  Slots ascend by node id within each kind, so slot order is the output order.
struct SlotHit {
    bool k1, kn, h0, h1;
    int32_t u, w0, w1;
    float qa, qb, qa1, qb1;
};
template <bool WRITE>
__device__ __forceinline__ SlotHit eval_slot(const SelectParams2& p, const RunCtx& h, const int32_t* __restrict__ words,
                                             const int32_t* __restrict__ ids, int deg, int npp, uint2 slot,
                                             bool want_pi, bool want_n1, float th_pre,
                                             const uint2* ppr_sm = nullptr) {
    const bool cn_needs_ppr = WRITE || p.th_cn > 0.0f;
    SlotHit r;
    r.k1 = r.kn = r.h0 = r.h1 = false;
    r.qa = r.qb = r.qa1 = r.qb1 = 0.f;
    const uint32_t w0 = slot.x, w1 = slot.y;
    r.w0 = (int32_t)w0; r.w1 = (int32_t)w1;
    r.u = (int32_t)w0;
    if (w0 & kPkPprTag) {
        r.u = (int32_t)(w0 & ~kPkPprTag);
        if (want_pi && smem_ppr_lookup(h, r.u, r.qa)) {
            r.qb = quantise(__uint_as_float(w1));
            if (r.qa >= th_pre && r.qb >= th_pre) {
                const bool in_a = hash_contains(h.tab, h.mask, h.shift, r.u);
                const int t = lower_bound_from(ids, 0, deg, r.u);
                const bool in_b = t < deg && __ldg(ids + t) == r.u;
                r.k1 = (in_a != in_b) && r.qa >= p.th_1hop && r.qb >= p.th_1hop;
                r.kn = want_n1 && !in_a && !in_b && r.qa >= p.th_non1hop && r.qb >= p.th_non1hop;
            }
        }
    } else {
        r.h0 = hash_contains(h.tab, h.mask, h.shift, (int32_t)w0);
        r.h1 = hash_contains(h.tab, h.mask, h.shift, (int32_t)w1);
        if (cn_needs_ppr && (r.h0 || r.h1)) {
            // PPR values of a common neighbour: P(a) from shared memory, P(b) by search over the row's PPR slots
            // (ppr_sm: the row's PPR slots staged in shared memory by the CTA-wide walk, else the row in global memory)
            auto pb_of = [&](int32_t x) -> float {
                int lo = 0, hi = npp;
                if (ppr_sm) {
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if ((int32_t)(ppr_sm[mid].x & 0x7fffffffu) < x) lo = mid + 1; else hi = mid;
                    }
                    return (lo < npp && (int32_t)(ppr_sm[lo].x & 0x7fffffffu) == x) ? quantise(__uint_as_float(ppr_sm[lo].y)) : 0.f;
                }
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((__ldg(words + 2 * mid) & 0x7fffffff) < x) lo = mid + 1; else hi = mid;
                }
                return (lo < npp && (__ldg(words + 2 * lo) & 0x7fffffff) == x) ? quantise(__int_as_float(__ldg(words + 2 * lo + 1))) : 0.f;
            };
            if (r.h0) {
                smem_ppr_lookup(h, (int32_t)w0, r.qa);
                r.qb = pb_of((int32_t)w0);
                r.h0 = r.qa >= p.th_cn && r.qb >= p.th_cn;
            }
            if (r.h1) {
                smem_ppr_lookup(h, (int32_t)w1, r.qa1);
                r.qb1 = pb_of((int32_t)w1);
                r.h1 = r.qa1 >= p.th_cn && r.qb1 >= p.th_cn;
            }
        }
    }
    return r;
}
// slot s of a packed row (padding beyond the row; S may be the upper bound the locator gives: the row's own padding
// evaluates to nothing)
__device__ __forceinline__ uint2 load_slot(const int32_t* __restrict__ words, int s, int S) {
    if (s < S) return __ldg(reinterpret_cast<const uint2*>(words) + s);
    return make_uint2(kPkPad, kPkPad);
}
__device__ __forceinline__ void write_hits(const SelectParams2& p, const SlotHit& r, int64_t r_pi, int64_t r_cn) {
    if (r.k1 || r.kn) { p.node[r_pi] = r.u; p.pa[r_pi] = r.qa; p.pb[r_pi] = r.qb; }
    if (r.h0) { p.node[r_cn] = r.w0; p.pa[r_cn] = r.qa; p.pb[r_cn] = r.qb; }
    if (r.h1) { const int64_t r1 = r_cn + (r.h0 ? 1 : 0); p.node[r1] = r.w1; p.pa[r1] = r.qa1; p.pb[r1] = r.qb1; }
}

// A group of G lanes (8 for the usual short row, a whole warp for rows of hundreds of ids) walks one link's PACKED
// target row (L2-hot: the screening just read it) against the staged source: lane l of the group takes slot l,
// l + G, ... so ascending node order within each set is lane order, and the ordered write needs only ballots.
// Same sets, order and values as the generic walk (select_walk.cuh).
template <int G, bool WRITE>
__device__ __forceinline__ void walk_packed_group(const SelectParams2& p, const RunCtx& h, const uint4* __restrict__ row,
                                                  int units, int lane, int64_t o_cn, int64_t o_1h, int64_t o_n1, int& c_cn,
                                                  int& c_1h, int& c_n1) {
    constexpr unsigned GM = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    const unsigned gmask = group_mask<G>(lane);
    const int gsh = lane & ~(G - 1), gl = lane & (G - 1);
    // the header is read now but only NEEDED where something is found (and for the length of a row whose locator
    // saturates): the slots are fetched without waiting for it — `units` (from the locator) bounds the row
    const uint4 hd = ldg16(row);
    const int deg = (int)hd.x, npp = (int)hd.y;
    const int32_t* words = reinterpret_cast<const int32_t*>(row) + 4;
    const int32_t* ids = words + 2 * npp;
    const int S = (units > 0 && units < 63) ? units * 8 - 2 : npp + ((deg + 1) >> 1);
    const unsigned lt = (1u << gl) - 1u;
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    c_cn = c_1h = c_n1 = 0;
    for (int s0 = 0; s0 < S; s0 += G) {
        const SlotHit r = eval_slot<WRITE>(p, h, words, ids, deg, npp, load_slot(words, s0 + gl, S), want_pi, want_n1, th_pre);
        const unsigned m1 = (__ballot_sync(gmask, r.k1) >> gsh) & GM, mn = (__ballot_sync(gmask, r.kn) >> gsh) & GM;
        const unsigned mh0 = (__ballot_sync(gmask, r.h0) >> gsh) & GM, mh1 = (__ballot_sync(gmask, r.h1) >> gsh) & GM;
        if (WRITE)
            write_hits(p, r, r.k1 ? o_1h + c_1h + __popc(m1 & lt) : o_n1 + c_n1 + __popc(mn & lt),
                       o_cn + c_cn + __popc(mh0 & lt) + __popc(mh1 & lt));
        c_1h += __popc(m1);
        c_n1 += __popc(mn);
        c_cn += __popc(mh0) + __popc(mh1);
    }
}

// count -> allocate -> ordered write of one link of a staged source by a group of G lanes
template <int G>
__device__ __forceinline__ void resolve_packed_group(const SelectParams2& p, const RunCtx& h, const uint4* __restrict__ row,
                                                     int units, int64_t i, int lane) {
    const unsigned gmask = group_mask<G>(lane);
    const int leader = lane & ~(G - 1);
    int c_cn, c_1h, c_n1;
    if constexpr (G == 32) {
        if (units > 0 && units <= 4) {
            // a row of at most 30 slots is ONE step of the warp: the hits stay in registers across the allocation
            const uint4 hd = ldg16(row);
            const int deg = (int)hd.x, npp = (int)hd.y;
            const int32_t* words = reinterpret_cast<const int32_t*>(row) + 4;
            const bool want_pi = p.mode != LPF_MODE_CN;
            const bool want_n1 = p.mode == LPF_MODE_ALL;
            const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
            const SlotHit r = eval_slot<true>(p, h, words, words + 2 * npp, deg, npp, load_slot(words, lane, units * 8 - 2),
                                              want_pi, want_n1, th_pre);
            const unsigned lt = (1u << lane) - 1u;
            const unsigned m1 = __ballot_sync(kFull, r.k1), mn = __ballot_sync(kFull, r.kn);
            const unsigned mh0 = __ballot_sync(kFull, r.h0), mh1 = __ballot_sync(kFull, r.h1);
            c_cn = __popc(mh0) + __popc(mh1); c_1h = __popc(m1); c_n1 = __popc(mn);
            int64_t s_cn, s_1h, s_n1;
            const bool ok = alloc_segments_warp(p, i, c_cn, c_1h, c_n1, lane, s_cn, s_1h, s_n1);
            if (c_cn + c_1h + c_n1 == 0 || !ok) return;
            write_hits(p, r, r.k1 ? p.cap + s_1h + __popc(m1 & lt) : 2 * p.cap + s_n1 + __popc(mn & lt),
                       s_cn + __popc(mh0 & lt) + __popc(mh1 & lt));
            return;
        }
    }
    walk_packed_group<G, false>(p, h, row, units, lane, 0, 0, 0, c_cn, c_1h, c_n1);
    int64_t s_cn = 0, s_1h = 0, s_n1 = 0;
    if constexpr (G == 32) {
        const bool ok = alloc_segments_warp(p, i, c_cn, c_1h, c_n1, lane, s_cn, s_1h, s_n1);
        if (c_cn + c_1h + c_n1 == 0 || !ok) return;
    } else {
        int ok = 1;
        if (lane == leader) ok = alloc_segments(p, i, c_cn, c_1h, c_n1, s_cn, s_1h, s_n1) ? 1 : 0;
        if (c_cn + c_1h + c_n1 == 0) return;       // uniform within the group
        ok = __shfl_sync(gmask, ok, leader);
        if (!ok) return;
        s_cn = __shfl_sync(gmask, s_cn, leader);
        s_1h = __shfl_sync(gmask, s_1h, leader);
        s_n1 = __shfl_sync(gmask, s_n1, leader);
    }
    walk_packed_group<G, true>(p, h, row, units, lane, s_cn, p.cap + s_1h, 2 * p.cap + s_n1, c_cn, c_1h, c_n1);
}

// The whole CTA walks one link's packed target row (a hub target: thousands of slots): per step thread t takes the
// slots t, t + 512, t + 1024, t + 1536 of the next 2,048 (four reads in flight), and the ordered positions come
// from one block-wide scan per step of the three counts packed into one word per 512 slots (common neighbours: up
// to two per slot | 1-hop | >1-hop).  `tot` = 2 x 4 x (kPkThreads / 32) words.
constexpr int kPkCtaUnroll = 4;
template <bool WRITE>
__device__ __forceinline__ void walk_packed_cta(const SelectParams2& p, const RunCtx& h, const uint4* __restrict__ row,
                                                uint32_t* tot, int64_t o_cn, int64_t o_1h, int64_t o_n1, int& c_cn,
                                                int& c_1h, int& c_n1) {
    constexpr int U = kPkCtaUnroll, NW = kPkThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 hd = ldg16(row);
    const int deg = (int)hd.x, npp = (int)hd.y;
    const int32_t* words = reinterpret_cast<const int32_t*>(row) + 4;
    const int32_t* ids = words + 2 * npp;
    const int S = npp + ((deg + 1) >> 1);
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    c_cn = c_1h = c_n1 = 0;
    int par = 0;
    for (int s0 = 0; s0 < S; s0 += U * kPkThreads, par ^= 1) {
        uint2 slot[U];
#pragma unroll
        for (int u = 0; u < U; ++u) slot[u] = load_slot(words, s0 + u * kPkThreads + tid, S);
        SlotHit r[U];
        uint32_t mine[U], inc[U];
        uint32_t* tt = tot + par * (U * NW);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            r[u] = eval_slot<WRITE>(p, h, words, ids, deg, npp, slot[u], want_pi, want_n1, th_pre);
            mine[u] = (uint32_t)((r[u].h0 ? 1 : 0) + (r[u].h1 ? 1 : 0)) | (r[u].k1 ? 1u << 12 : 0u) | (r[u].kn ? 1u << 22 : 0u);
            inc[u] = mine[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(kFull, inc[u], o);
                if (lane >= o) inc[u] += x;
            }
            if (lane == 31) tt[u * NW + warp] = inc[u];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t before = inc[u] - mine[u], total = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t x = tt[u * NW + w];
                before += (w < warp) ? x : 0u;
                total += x;
            }
            if (WRITE)
                write_hits(p, r[u], r[u].k1 ? o_1h + c_1h + (int)((before >> 12) & 1023u) : o_n1 + c_n1 + (int)(before >> 22),
                           o_cn + c_cn + (int)(before & 4095u));
            c_cn += (int)(total & 4095u);
            c_1h += (int)((total >> 12) & 1023u);
            c_n1 += (int)(total >> 22);
        }
    }
}

// count -> allocate -> ordered write of one link by the whole CTA.  A row of up to 2,048 slots (~4,000 neighbours)
// is ONE step of the walk: the hits stay in registers while thread 0 allocates, and are written without a second
// walk; longer rows are walked twice.  `seg` = 3 x int64 and `ok` in shared memory.
__device__ __forceinline__ void resolve_packed_cta(const SelectParams2& p, const RunCtx& h, const uint4* __restrict__ row,
                                                   int units, uint2* ppr_sm, uint32_t* tot, int64_t* seg, int* ok, int64_t i) {
    constexpr int U = kPkCtaUnroll, NW = kPkThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (the header is needed for the row's length only when the locator saturates: the slots of a shorter row are
    // fetched without waiting for it)
    const uint4 hd = ldg16(row);
    const int deg = (int)hd.x, npp = (int)hd.y;
    const int S = (units > 0 && units < 63) ? units * 8 - 2 : npp + ((deg + 1) >> 1);
    int c_cn, c_1h, c_n1;
    if (S <= U * kPkThreads) {
        const int32_t* words = reinterpret_cast<const int32_t*>(row) + 4;
        const int32_t* ids = words + 2 * npp;
        const bool want_pi = p.mode != LPF_MODE_CN;
        const bool want_n1 = p.mode == LPF_MODE_ALL;
        const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
        uint2 slot[U];
#pragma unroll
        for (int u = 0; u < U; ++u) slot[u] = load_slot(words, u * kPkThreads + tid, S);
        // the row's PPR slots (the first npp) into shared memory: a common neighbour's P(b) is then a search there
        // instead of a chain of dependent global reads
        const bool staged = npp <= kPkMaxPprRow;
        if (staged && tid < npp) ppr_sm[tid] = slot[0];
        __syncthreads();
        SlotHit r[U];
        uint32_t mine[U], inc[U], before[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            r[u] = eval_slot<true>(p, h, words, ids, deg, npp, slot[u], want_pi, want_n1, th_pre, staged ? ppr_sm : nullptr);
            mine[u] = (uint32_t)((r[u].h0 ? 1 : 0) + (r[u].h1 ? 1 : 0)) | (r[u].k1 ? 1u << 12 : 0u) | (r[u].kn ? 1u << 22 : 0u);
            inc[u] = mine[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(kFull, inc[u], o);
                if (lane >= o) inc[u] += x;
            }
            if (lane == 31) tot[u * NW + warp] = inc[u];
        }
        __syncthreads();
        c_cn = c_1h = c_n1 = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t bf = inc[u] - mine[u], total = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t x = tot[u * NW + w];
                bf += (w < warp) ? x : 0u;
                total += x;
            }
            // position of this thread's hits within the link's three sets
            before[u] = (uint32_t)(c_cn + (int)(bf & 4095u)) | (uint32_t)(c_1h + (int)((bf >> 12) & 1023u)) << 12 |
                        (uint32_t)(c_n1 + (int)(bf >> 22)) << 22;
            c_cn += (int)(total & 4095u);
            c_1h += (int)((total >> 12) & 1023u);
            c_n1 += (int)(total >> 22);
        }
        if (warp == 0) {
            int64_t s0, s1, s2;
            const bool fits = alloc_segments_warp(p, i, c_cn, c_1h, c_n1, lane, s0, s1, s2);
            if (lane == 0) { *ok = fits ? 1 : 0; seg[0] = s0; seg[1] = s1; seg[2] = s2; }
        }
        __syncthreads();
        if (*ok && c_cn + c_1h + c_n1 > 0) {
            const int64_t o_cn = seg[0], o_1h = p.cap + seg[1], o_n1 = 2 * p.cap + seg[2];
#pragma unroll
            for (int u = 0; u < U; ++u)
                write_hits(p, r[u], r[u].k1 ? o_1h + (int)((before[u] >> 12) & 1023u) : o_n1 + (int)(before[u] >> 22),
                           o_cn + (int)(before[u] & 4095u));
        }
        __syncthreads();       // tot / seg / ok are reused by the next link
        return;
    }
    walk_packed_cta<false>(p, h, row, tot, 0, 0, 0, c_cn, c_1h, c_n1);
    if (tid == 0) {
        int64_t s0, s1, s2;
        *ok = alloc_segments(p, i, c_cn, c_1h, c_n1, s0, s1, s2) ? 1 : 0;
        seg[0] = s0; seg[1] = s1; seg[2] = s2;
    }
    __syncthreads();
    const bool go = *ok && (c_cn + c_1h + c_n1 > 0);
    const int64_t s0 = seg[0], s1 = seg[1], s2 = seg[2];
    __syncthreads();
    if (go) walk_packed_cta<true>(p, h, row, tot, s0, p.cap + s1, 2 * p.cap + s2, c_cn, c_1h, c_n1);
}

// A link of a source that is not staged in shared memory (searched in global memory: beyond even the hub table),
// by one warp over the CSR tables, or — both rows long — handed to the deferred-link kernel.
__device__ __noinline__ void resolve_unstaged32(const SelectParams2& p, int64_t i, int lane) {
    const LinkRows rows = load_rows(p, i);
    if (!is_heavy(rows, p.mode != LPF_MODE_CN, 8)) {
        onepass_link<32>(p, rows, i, lane);
    } else if (lane == 0) {
        p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
    }
}

template <class SM>
__device__ __forceinline__ int run_of(const SM& sm, int t) {
    return (t >= sm.run_start[1] ? 1 : 0) + (t >= sm.run_start[2] ? 1 : 0);
}

// HUB = false: the chunks are the batch cut into pieces of 512 links; runs whose source row does not fit the hash
// are appended to the hub list.  HUB = true (second launch, one CTA per SM, a 32,768-slot table): the chunks are the
// entries of that list.
template <int SLOTS, bool HUB, int MINB>
__global__ void __launch_bounds__(kPkThreads, MINB)
select_onepass_packed_kernel(const __grid_constant__ SelectParams2 p, const uint32_t* __restrict__ locator,
                             const uint4* __restrict__ blob) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    using SM = PkSmemT<SLOTS>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = tid >> 3;                    // 64 groups of 8 lanes (generic fallback)
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    // !HUB: the batch in kPkSplit x `gridDim.x` (or more) even pieces of at most kPkChunk links, handed out through a
    // counter: the launch lasts as long as its slowest CTA, and a piece with a hub source takes twice the average
    const int64_t per = HUB ? 0 : min((int64_t)kPkChunk, max((int64_t)kPkMinPiece, (p.bs + kPkSplit * gridDim.x - 1) / (kPkSplit * gridDim.x)));
    const int64_t nchunks = HUB ? (int64_t)p.hub[0] : (p.bs + per - 1) / per;
    const int slots_cap = p.slot_limit > 0 ? min(SLOTS, HUB ? 4 * p.slot_limit : p.slot_limit) : SLOTS;
    const int hub_cap = p.slot_limit > 0 ? min(kPkHubSlots, 4 * p.slot_limit) : kPkHubSlots;

    long long t_mark = clock64();
#define LPF_PHASE(k)                                                              \
    do {                                                                          \
        if (p.dbg && tid == 0) {                                                  \
            const long long now = clock64();                                      \
            atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (k)), (unsigned long long)(now - t_mark)); \
            sm.dbg_ph[(k)] += (int)(now - t_mark);                                \
            t_mark = now;                                                         \
        }                                                                         \
    } while (0)

    const long long t_cta = t_mark;
    for (int64_t round = 0;; ++round) {
        int64_t chunk;
        if (HUB) {
            chunk = blockIdx.x + round * gridDim.x;
        } else {
            if (tid == 0) sm.cur_chunk = atomicAdd(p.heavy + 1, 1);       // (workspace word 1: reset with the counters)
            __syncthreads();
            chunk = sm.cur_chunk;
        }
        if (chunk >= nchunks) break;
        const long long t_chunk = clock64();
        const int64_t i0 = HUB ? (int64_t)p.hub[1 + 2 * chunk] : chunk * per;
        const int len = HUB ? p.hub[2 + 2 * chunk] : (int)min(per, p.bs - i0);
        if (p.dbg && tid == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 5), 1ull);
            for (int k = 0; k < 16; ++k) sm.dbg_ph[k] = 0;
        }
        if (tid == 0) { sm.n_runs = 0; sm.n_slow = 0; sm.n_items = 0; sm.items_full = 0; sm.n_cta = 0; sm.n_mid = 0; sm.b_next = 0; }
        __syncthreads();
        // ---- the chunk's links (positions tid, tid + 512): locator of the target (an L2-resident array), run
        // boundaries with the locator of their source
        for (int t = tid; t < len; t += kPkThreads) {
            const int64_t i = i0 + t;
            const int64_t a_me = __ldg(p.links + i);
            const int64_t a_prev = t > 0 ? __ldg(p.links + i - 1) : -1;
            sm.l_loc[t] = __ldg(locator + __ldg(p.links + p.bs + i));
            sm.l_any[t] = 0;
            if (!HUB) {
                // every link starts as "nothing selected"; the links that are resolved later overwrite their entries
                p.counts[i] = 0; p.counts[p.bs + i] = 0; p.counts[2 * p.bs + i] = 0;
                p.seg_start[i] = 0; p.seg_start[p.bs + i] = 0; p.seg_start[2 * p.bs + i] = 0;
            }
            if (t == 0 || a_me != a_prev) {
                const int k = atomicAdd(&sm.n_runs, 1);
                if (k < kPkMaxRuns) {
                    sm.run_start[k] = t;
                    sm.r_loc[k] = __ldg(locator + a_me);
                }
            }
        }
        __syncthreads();
        LPF_PHASE(10);
        const int n_runs = sm.n_runs;
        if (n_runs > kPkMaxRuns) {
            // not run-shaped: generic one-pass walk, 8 lanes per link
            for (int t = group; t < len; t += kPkThreads / 8) generic_link8(p, i0 + t, lane);
            __syncthreads();
            LPF_PHASE(4);
            continue;
        }
        if (tid == 0) {
            // sort the (at most kPkMaxRuns) boundaries, close the list
            for (int x = 1; x < n_runs; ++x)
                for (int y = x; y > 0 && sm.run_start[y] < sm.run_start[y - 1]; --y) {
                    const int tmp = sm.run_start[y]; sm.run_start[y] = sm.run_start[y - 1]; sm.run_start[y - 1] = tmp;
                    const uint32_t tl = sm.r_loc[y]; sm.r_loc[y] = sm.r_loc[y - 1]; sm.r_loc[y - 1] = tl;
                }
            for (int x = n_runs; x <= kPkMaxRuns; ++x) sm.run_start[x] = len;
            // hash regions per run, side by side
            int used = 0;
            for (int r = 0; r < n_runs; ++r) {
                // the source's size: from its locator when that says it all (an upper bound of the degree: every slot
                // taken as two ids), from its header for rows of 63 units and more (one more read, big sources only)
                const int units_a = (int)(sm.r_loc[r] & 63u);
                int na = 2 * (8 * units_a - 2), slots_a = 8 * units_a - 2;
                if (units_a == 63) {
                    const uint4 hd = ldg16(row_of(blob, sm.r_loc[r]));
                    na = (int)hd.x;
                    slots_a = (int)hd.y + (((int)hd.x + 1) >> 1);
                }
                sm.r_slots[r] = slots_a;
                // slots: a power of two >= 4*na (load <= 0.25: almost every probe ends in its home bucket) while
                // the table has room, never less than 2*na
                int lg = 6;
                while ((1 << lg) < 2 * na && lg < 30) ++lg;
                if ((1 << lg) < 4 * na && used + (2 << lg) <= slots_cap / 2) ++lg;
                // r_hashed: 1 = adjacency row hashed in shared memory; 2 = searched in global memory (a source beyond
                // even the hub table); 3 = handed to the hub launch; 0 = no screening (source PPR row too long for the
                // shared table): generic walk
                const bool fits = used + (1 << lg) <= slots_cap;
                int m = fits ? 1 : 2;
                if (!HUB && (m == 2 || (na > kPkHubDeg && p.slot_limit == 0)) && (1 << lg) <= hub_cap) {
                    // in pieces of kPkHubPiece links: the hub launch has a CTA (and an SM) for each of them
                    m = 3;
                    const int first = sm.run_start[r], n_links = sm.run_start[r + 1] - first;
                    const int pieces = (n_links + kPkHubPiece - 1) / kPkHubPiece;
                    const int e = atomicAdd(p.hub, pieces);
                    for (int k = 0; k < pieces; ++k) {
                        p.hub[1 + 2 * (e + k)] = (int32_t)(i0 + first + k * kPkHubPiece);
                        p.hub[2 + 2 * (e + k)] = min(kPkHubPiece, n_links - k * kPkHubPiece);
                    }
                }
                sm.r_hashed[r] = m;
                sm.r_lg[r] = lg;
                sm.r_tab0[r] = used;
                sm.r_ctx[r] = make_int4(used, (1 << (lg - 2)) - 1, 32 - (lg - 2), 0);
                if (m == 1) used += 1 << lg;
            }
            sm.tab_used = used;
        }
        __syncthreads();
        LPF_PHASE(12);
        {
            const int used = sm.tab_used;
            for (int s = tid; s < used; s += kPkThreads) sm.tab[s] = -1;
            for (int s = tid; s < n_runs * kPprHashSlots; s += kPkThreads)
                sm.run[s / kPprHashSlots].ppos[s % kPprHashSlots] = -1;
        }
        __syncthreads();     // tables cleared
        // ---- flatten the chunk's target rows into 64-byte units: items[] = (chunk position << 16 | unit), in link
        // order, 512 positions per pass.  Every row is listed, whatever its length (a row of 63 units and more says
        // so in its header); the list holds a prefix of the chunk's units, what does not fit is resolved unscreened.
        for (int t0 = 0; t0 < len; t0 += kPkThreads) {
            const int t = t0 + tid;
            const int r = run_of(sm, min(t, len - 1));
            const bool screened = t < len && (sm.r_hashed[r] == 1 || sm.r_hashed[r] == 2);
            const uint32_t loc_b = t < len ? sm.l_loc[t] : 0u;
            int units = screened ? (int)(loc_b & 63u) : 0;
            if (units == 63) {
                const uint4 hd = ldg16(row_of(blob, loc_b));
                units = (int)row_units((int64_t)hd.x, (int64_t)hd.y);
            }
            const int mine = units > kPkMaxUnits ? 0 : units;
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += x;
            }
            if (lane == 31) sm.warp_tot[warp] = inc;
            const int base0 = sm.n_items;        // units listed by the previous pass
            const bool full_before = sm.items_full != 0;
            __syncthreads();
            int ex = base0 + inc - mine;
#pragma unroll
            for (int w = 0; w < kPkThreads / 32; ++w) ex += (w < warp) ? sm.warp_tot[w] : 0;
            const bool fits = !full_before && ex + mine <= kPkMaxItems;
            if (screened) {
                if (units > kPkMaxUnits || !fits) {
                    if (units > kPkWarpUnits) sm.q_cta[atomicAdd(&sm.n_cta, 1)] = (uint16_t)t;
                    else sm.q_slow[atomicAdd(&sm.n_slow, 1)] = (uint16_t)t;
                } else {
                    for (int u = 0; u < units; ++u) sm.items[ex + u] = ((uint32_t)t << 16) | (uint32_t)u;
                }
            }
            __syncthreads();     // everyone has read n_items / items_full of the previous pass
            // the listed units end at the first link that does not fit (later, shorter rows must not leave holes)
            if (!full_before && mine > 0 && !fits && ex <= kPkMaxItems) { sm.n_items = ex; sm.items_full = 1; }
            if (tid == kPkThreads - 1 && fits) sm.n_items = ex + mine;
            __syncthreads();
        }
        LPF_PHASE(13);
        // ---- stage the sources from their packed rows: the slots say what they are, so every thread just takes
        // slot tid, tid + 512, ...: a PPR entry goes to the run's table, two ids go to its hash set
        for (int r = 0; r < n_runs; ++r) {
            const int mode_r = sm.r_hashed[r];
            if (mode_r == 3) continue;
            const int lg = sm.r_lg[r], n_slots = sm.r_slots[r];
            int32_t* tab = sm.tab + sm.r_tab0[r];
            const uint4* row = row_of(blob, sm.r_loc[r]);
            if (tid == 0) {
                const uint4 hd = ldg16(row);
                sm.r_na[r] = (int)hd.x;
                sm.r_npa[r] = (int)hd.y;
                if ((int)hd.y > kPkMaxPprRow) sm.r_hashed[r] = 0;     // PPR row beyond the shared table: generic walk
            }
            const uint2* slots = reinterpret_cast<const uint2*>(row) + 2;      // after the header
            for (int s = tid; s < n_slots; s += kPkThreads) {
                const uint2 v = __ldg(slots + s);
                if (v.x & kPkPprTag) {
                    if (s < kPkMaxPprRow) {
                        const int32_t u = (int32_t)(v.x & ~kPkPprTag);
                        sm.run[r].pac[s] = u;
                        sm.run[r].pav[s] = __uint_as_float(v.y);
                        uint32_t slot = hash_slot(u, 32 - 8);
                        while (atomicCAS(&sm.run[r].ppos[slot], -1, s) != -1) slot = (slot + 1) & (kPprHashSlots - 1);
                    }
                } else if (mode_r == 1) {
                    if (v.x != kPkPad) hash_insert(tab, (1u << (lg - 2)) - 1u, 32 - (lg - 2), (int32_t)v.x);
                    if (v.y != kPkPad) hash_insert(tab, (1u << (lg - 2)) - 1u, 32 - (lg - 2), (int32_t)v.y);
                }
            }
        }
        __syncthreads();
        LPF_PHASE(0);

        // ---- phase A: four lanes per 64-byte unit, kPkInflight units in flight per lane
        {
            constexpr int K = kPkInflight;
            const int n_items = sm.n_items;
            const int ql = tid & 3;
            for (int q0 = tid >> 2; q0 < n_items; q0 += K * (kPkThreads / 4)) {
                uint4 v[K];
                int tt[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int q = q0 + k * (kPkThreads / 4);
                    tt[k] = -1;
                    v[k] = make_uint4(kPkPad, kPkPad, kPkPad, kPkPad);
                    if (q < n_items) {
                        const uint32_t it = sm.items[q];
                        const int t = (int)(it >> 16), u = (int)(it & 0xffffu);
                        tt[k] = (u == 0 && ql == 0) ? -1 : t;              // chunk 0 of unit 0 is the row's header
                        v[k] = ldg16(row_of(blob, sm.l_loc[t]) + 4 * u + ql);
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (tt[k] < 0) continue;
                    const int t = tt[k], r = run_of(sm, t);
                    const int mode_r = sm.r_hashed[r];
                    if (mode_r == 0) continue;               // the run turned out not to be screenable
                    const int4 cx = sm.r_ctx[r];
                    RunCtx h;
                    h.tab = sm.tab + cx.x; h.mask = (uint32_t)cx.y; h.shift = cx.z;
                    h.pac = sm.run[r].pac; h.pav = sm.run[r].pav; h.ppos = sm.run[r].ppos; h.npa = 0;
                    bool any;
                    if (mode_r == 1) {
                        any = screen_slot<true>(h, nullptr, 0, v[k].x, v[k].y, want_pi, th_pre) |
                              screen_slot<true>(h, nullptr, 0, v[k].z, v[k].w, want_pi, th_pre);
                    } else {
                        const int32_t* ids = reinterpret_cast<const int32_t*>(row_of(blob, sm.r_loc[r])) + 4 + 2 * sm.r_npa[r];
                        any = screen_slot<false>(h, ids, sm.r_na[r], v[k].x, v[k].y, want_pi, th_pre) |
                              screen_slot<false>(h, ids, sm.r_na[r], v[k].z, v[k].w, want_pi, th_pre);
                    }
                    if (any) sm.l_any[t] = 1;
                }
            }
        }
        __syncthreads();
        for (int t = tid; t < len; t += kPkThreads)
            if (sm.l_any[t]) {
                // a long target row that selects something: the whole CTA walks it
                if ((int)(sm.l_loc[t] & 63u) > kPkWarpUnits) sm.q_cta[atomicAdd(&sm.n_cta, 1)] = (uint16_t)t;
                else if ((int)(sm.l_loc[t] & 63u) > 4) sm.q_mid[atomicAdd(&sm.n_mid, 1)] = (uint16_t)t;
                else sm.q_slow[atomicAdd(&sm.n_slow, 1)] = (uint16_t)t;
            }
        __syncthreads();
        LPF_PHASE(1);

        // ---- runs that cannot be screened (source PPR row beyond the shared table): generic group walk
        for (int r = 0; r < n_runs; ++r) {
            if (sm.r_hashed[r] != 0) continue;
            for (int t = sm.run_start[r] + group; t < sm.run_start[r + 1]; t += kPkThreads / 8)
                generic_link8(p, i0 + t, lane);
        }
        // ---- phase B: the links that select something (count -> allocate -> ordered write), one warp per link against
        // the staged source (a handful per chunk: the runs of hub sources, which select something with every tenth
        // target, were cut into pieces for the hub launch)
        // The warps take the next link from a counter, the longer rows (5 .. 32 units: up to eight steps of the warp,
        // twice) first: a fixed round robin left the warp that drew two of them working long after the others.
        const int nm = sm.n_mid, ns = nm + sm.n_slow;
        for (;;) {
            int q = 0;
            if (lane == 0) q = atomicAdd(&sm.b_next, 1);
            q = __shfl_sync(kFull, q, 0);
            if (q >= ns) break;
            const int t = q < nm ? sm.q_mid[q] : sm.q_slow[q - nm];
            const int r = run_of(sm, t);
            if (sm.r_hashed[r] == 1) resolve_packed_group<32>(p, make_ctx(sm, r), row_of(blob, sm.l_loc[t]), (int)(sm.l_loc[t] & 63u), i0 + t, lane);
            else if (sm.r_hashed[r] == 2) resolve_unstaged32(p, i0 + t, lane);       // (searched in global memory: rare)
        }
        if (p.dbg && tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 6), (unsigned long long)ns);
        LPF_PHASE(2);
        __syncthreads();
        // ---- phase C: long target rows that select something, the whole CTA per link
        const int nc = sm.n_cta;
        for (int q = 0; q < nc; ++q) {
            const int t = sm.q_cta[q];
            const int r = run_of(sm, t);
            if (sm.r_hashed[r] == 0) continue;
            if (sm.r_hashed[r] != 1) {
                if (warp == 0) resolve_unstaged32(p, i0 + t, lane);
                continue;
            }
            resolve_packed_cta(p, make_ctx(sm, r), row_of(blob, sm.l_loc[t]), (int)(sm.l_loc[t] & 63u), sm.cta_ppr, sm.scan_tot,
                               sm.cta_seg, &sm.cta_ok, i0 + t);
        }
        if (p.dbg && tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 7), (unsigned long long)nc);
        LPF_PHASE(3);
        __syncthreads();     // the shared tables are rebuilt for the next chunk
        if (p.dbg && tid == 0) {
            const long long dt = clock64() - t_chunk;
            if (atomicMax(p.dbg + 8, dt) < dt) {          // the slowest chunk so far: its phases, size and queues
                for (int k = 0; k < 16; ++k) p.dbg[16 + k] = sm.dbg_ph[k];
                p.dbg[32] = len; p.dbg[33] = sm.n_runs; p.dbg[34] = sm.n_items; p.dbg[35] = sm.n_slow;
                p.dbg[36] = sm.n_cta; p.dbg[37] = sm.tab_used; p.dbg[38] = HUB ? 1 : 0;
            }
            atomicMax(p.dbg + 9, clock64() - t_cta);
        }
    }
}
#undef LPF_PHASE

// ---------------------------------------------------------------------------------------------------------
// Building the packed rows (once per graph).
// ---------------------------------------------------------------------------------------------------------

__global__ void pack_count_kernel(const int64_t* __restrict__ arp, const int64_t* __restrict__ prp, int64_t n,
                                  int32_t* __restrict__ units) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    units[x] = (int32_t)row_units(arp[x + 1] - arp[x], prp[x + 1] - prp[x]);
}

__global__ void __launch_bounds__(256) pack_fill_kernel(const int64_t* __restrict__ arp, const int32_t* __restrict__ ac,
                                                        const int64_t* __restrict__ prp, const int32_t* __restrict__ pc,
                                                        const float* __restrict__ pv, int64_t n,
                                                        const int64_t* __restrict__ off, uint32_t* __restrict__ locator,
                                                        uint4* __restrict__ blob) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t x = warp0; x < n; x += nwarps) {
        const int64_t a0 = arp[x], p0 = prp[x];
        const int deg = (int)(arp[x + 1] - a0), npp = (int)(prp[x + 1] - p0);
        const int64_t o = off[x];
        const int units = (int)row_units(deg, npp);
        if (lane == 0) locator[x] = ((uint32_t)o << 6) | (uint32_t)min(units, 63);
        uint32_t* w = reinterpret_cast<uint32_t*>(blob + 4 * o);
        if (lane < 4) w[lane] = lane == 0 ? (uint32_t)deg : (lane == 1 ? (uint32_t)npp : 0u);
        for (int k = lane; k < npp; k += 32) {
            w[4 + 2 * k] = (uint32_t)pc[p0 + k] | kPkPprTag;
            w[4 + 2 * k + 1] = __float_as_uint(pv[p0 + k]);
        }
        const int ids0 = 4 + 2 * npp, words = units * 16;
        for (int k = lane; ids0 + k < words; k += 32) w[ids0 + k] = k < deg ? (uint32_t)ac[a0 + k] : kPkPad;
    }
}

extern long long* g_select_dbg;
// profiling hook (lpf_debug_select_timing): CUDA events around the three kernels of the packed launch sequence
bool g_kernel_timing = false;     // shared with nz_fused.cu
static cudaEvent_t g_pk_ev[4];
static bool g_pk_ev_ready = false, g_pk_ev_valid = false;
static constexpr int g_pk_ctas = 2;
static int g_pk_slot_limit = 0;
void launch_onepass_reset(const SelectParams2& p, cudaStream_t st);
void launch_onepass_tail(const SelectParams2& p, cudaStream_t st);

}  // namespace lpf

using namespace lpf;

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int64_t lpf_link_rows_bytes(int64_t n, int64_t adj_nnz, int64_t ppr_nnz) {
    if (n < 0 || adj_nnz < 0 || ppr_nnz < 0) return -1;
    // per node ceil((16 + 8 nP + 8 ceil(deg/2)) / 64) units <= (16 + 63 + 4 + 8 nP + 4 deg) / 64
    const int64_t units = (83 * n + 8 * ppr_nnz + 4 * adj_nnz) / 64 + 1;
    if (units >= ((int64_t)1 << 26)) return -1;    // 26-bit unit index in the locator
    return units * 64;
}

extern "C" int64_t lpf_link_rows_scratch_bytes(int64_t n) {
    if (n < 0) return -1;
    return align_up(n * 4, 16) + align_up((n + 1) * 8, 16) + align_up(lpf_scan_scratch_bytes(n), 16);
}

extern "C" int lpf_pack_link_rows(const int64_t* adj_rowptr, const int32_t* adj_col, const int64_t* ppr_rowptr,
                                  const int32_t* ppr_col, const float* ppr_val, int64_t n, uint32_t* locator,
                                  void* row_blob, void* scratch, void* stream) {
    LPF_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) - 1, "bad node count");
    LPF_REQUIRE(adj_rowptr && ppr_rowptr, "rowptr is NULL");
    LPF_REQUIRE(n == 0 || (locator && row_blob && scratch), "NULL output");
    LPF_REQUIRE((reinterpret_cast<uintptr_t>(row_blob) & 63) == 0, "row_blob must be 64-byte aligned");
    if (n == 0) return LPF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* s = static_cast<uint8_t*>(scratch);
    int32_t* units = reinterpret_cast<int32_t*>(s);
    int64_t* off = reinterpret_cast<int64_t*>(s + align_up(n * 4, 16));
    void* scan_scratch = s + align_up(n * 4, 16) + align_up((n + 1) * 8, 16);
    pack_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(adj_rowptr, ppr_rowptr, n, units);
    int rc = lpf_scan_counts(units, n, off, scan_scratch, stream);
    if (rc) return rc;
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    pack_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(adj_rowptr, adj_col, ppr_rowptr, ppr_col, ppr_val, n, off,
                                                       locator, static_cast<uint4*>(row_blob));
    return check_launch("lpf_pack_link_rows");
}

extern "C" int lpf_select_onepass_packed(const int64_t* links, int64_t bs, const int64_t* adj_rowptr,
                                         const int32_t* adj_col, const int64_t* ppr_rowptr, const int32_t* ppr_col,
                                         const float* ppr_val, const uint32_t* locator, const void* row_blob,
                                         float th_cn, float th_1hop, float th_non1hop, int mode, int64_t cap,
                                         int32_t* counts, int32_t* seg_start, int32_t* nz_list, int64_t* header,
                                         int32_t* node, float* src_ppr, float* tgt_ppr, void* workspace, void* stream) {
    LPF_REQUIRE(bs >= 0, "negative batch size");
    LPF_REQUIRE(bs == 0 || links, "links is NULL");
    LPF_REQUIRE(adj_rowptr && ppr_rowptr, "rowptr is NULL");
    LPF_REQUIRE(mode == LPF_MODE_CN || mode == LPF_MODE_1HOP || mode == LPF_MODE_ALL, "bad mode");
    const bool ok = (mode == LPF_MODE_CN) || (th_1hop > 0.0f && (mode != LPF_MODE_ALL || th_non1hop > 0.0f));
    if (!ok) {
        lpf::set_error("lpf_select_onepass_packed needs th_1hop > 0 (and th_non1hop > 0 in mode ALL)");
        return LPF_ERR_UNSUPPORTED;
    }
    LPF_REQUIRE(locator && row_blob, "locator / row_blob is NULL (lpf_pack_link_rows)");
    LPF_REQUIRE(cap >= 0 && 3 * cap < ((int64_t)1 << 31), "bad pair capacity");
    LPF_REQUIRE(header && workspace, "header/workspace is NULL");
    LPF_REQUIRE(bs == 0 || (counts && seg_start && nz_list), "NULL output");
    LPF_REQUIRE(cap == 0 || (node && src_ppr && tgt_ppr), "NULL pair arrays");
    cudaStream_t st = (cudaStream_t)stream;
    SelectParams2 p{links, bs, adj_rowptr, adj_col, ppr_rowptr, ppr_col, ppr_val, th_cn, th_1hop, th_non1hop,
                    mode, counts, nullptr, node, src_ppr, tgt_ppr, nullptr, (int32_t*)workspace, cap, header,
                    seg_start, nz_list, g_select_dbg, (int32_t*)workspace + bs + 4, g_pk_slot_limit};
    using SmMain = PkSmemT<kPkHashSlots>;
    using SmHub = PkSmemT<kPkHubSlots>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(select_onepass_packed_kernel<kPkHashSlots, false, 2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmMain));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(select_onepass_packed_kernel<kPkHubSlots, true, 1>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmHub));
        if (e != cudaSuccess) {
            set_error("lpf_select_onepass_packed: cudaFuncSetAttribute(%zu / %zu B): %s", sizeof(SmMain), sizeof(SmHub),
                      cudaGetErrorString(e));
            return LPF_ERR_CUDA;
        }
        configured = true;
    }
    launch_onepass_reset(p, st);
    const bool timing = g_kernel_timing && bs > 0;
    if (timing && !g_pk_ev_ready) {
        for (auto& e : g_pk_ev) cudaEventCreate(&e);
        g_pk_ev_ready = true;
    }
    if (timing) cudaEventRecord(g_pk_ev[0], st);
    if (bs > 0) {
        // one resident wave (g_pk_ctas CTAs per SM), the batch cut evenly over it: pieces of 512 .. 1,024 links
        int64_t blocks = (bs + kPkThreads - 1) / kPkThreads;
        const int64_t cap_blocks = (int64_t)kNumSMs * g_pk_ctas;
        if (blocks > cap_blocks) blocks = cap_blocks;
        select_onepass_packed_kernel<kPkHashSlots, false, 2><<<(unsigned)blocks, kPkThreads, sizeof(SmMain), st>>>(
            p, locator, static_cast<const uint4*>(row_blob));
        if (timing) cudaEventRecord(g_pk_ev[1], st);
        select_onepass_packed_kernel<kPkHubSlots, true, 1><<<kNumSMs, kPkThreads, sizeof(SmHub), st>>>(
            p, locator, static_cast<const uint4*>(row_blob));
        if (timing) cudaEventRecord(g_pk_ev[2], st);
    }
    launch_onepass_tail(p, st);
    if (timing) {
        cudaEventRecord(g_pk_ev[3], st);
        g_pk_ev_valid = true;
    }
    return check_launch("lpf_select_onepass_packed");
}

// Test hook: caps the hash slots the screening kernel may use (the hub launch gets four times as many), so that
// small graphs reach the hub launch and the global-memory search of sources beyond it; 0 restores the full tables.
extern "C" int lpf_debug_select_slots(int limit) {
    lpf::g_pk_slot_limit = limit > 0 ? limit : 0;
    return LPF_OK;
}

// Profiling hook: with enable != 0 later lpf_select_onepass_packed calls record CUDA events on their stream around
// the screening kernel, the hub-source kernel and the deferred-link tail; lpf_debug_select_timing_read waits for the
// last such call and returns the three durations in milliseconds (0 on success, -1 if nothing was recorded).
extern "C" int lpf_debug_select_timing(int enable) {
    lpf::g_kernel_timing = enable != 0;
    if (!enable) lpf::g_pk_ev_valid = false;
    return LPF_OK;
}
extern "C" int lpf_debug_select_timing_read(float* ms3_host) {
    if (!lpf::g_pk_ev_valid || !ms3_host) return -1;
    if (cudaEventSynchronize(lpf::g_pk_ev[3]) != cudaSuccess) return -1;
    for (int k = 0; k < 3; ++k)
        if (cudaEventElapsedTime(ms3_host + k, lpf::g_pk_ev[k], lpf::g_pk_ev[k + 1]) != cudaSuccess) return -1;
    return LPF_OK;
}
