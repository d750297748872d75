// K1, one-pass selection over PACKED LINK ROWS: the HBM layout built for the per-link walk.
//
// The CSR tables cost a link four dependent random reads per endpoint (two rowptr pairs, then the adjacency row,
// the PPR columns and — on a match — the PPR values, all in different arrays).  Measured on a B200
// (tools/gather_probe.cu): DRAM serves random reads in 128-byte LINES at ~35-43 G lines/s whatever part of the line
// is used, so the cost of a link is the number of lines it touches and the number of DEPENDENT trips it makes.
// lpf_pack_link_rows therefore rewrites (adjacency CSR, PPR CSR) once per graph as
//   slab[x]  one 128-byte line per node, at x * 128 (no locator, no indirection): chunk 0 = header (deg, nP, first
//            128-byte unit of the row's overflow, PPR chunks | row chunks << 16 (both saturating at 65,535)), chunks
//            1..7 = the first seven 16-byte chunks of the row
//   overflow the remaining chunks of rows longer than seven chunks, 128-byte aligned
//   a row    = ceil(nP/2) PPR chunks: two entries (col | 0x80000000, value bits), the odd one out padded with
//              (0xffffffff, 0); then ceil(deg/4) id chunks: four ascending neighbour ids, padded with 0x7fffffff
// so the median target (deg <= 16) is ONE line and one trip (link -> slab line), the others one more trip.
//
// Two launches (citation2-style evaluation: runs of links sharing their source, reference train/testing.py:20-23):
//
// SCREEN — "does this link select anything?" — one CTA of 256 threads per piece of 256 consecutive links (at most
// kPkMaxRuns runs of equal source).  The CTA works together only to stage the sources (run boundaries by ballot,
// the sources' slab lines, their overflow: three barriers) as BLOOM FILTERS in shared memory: one of the source's
// neighbour ids (256 bits per id while the piece's 16 KB last, two bits of one word per id: one shared-memory read
// per probe), one of the PPR columns it holds above the smaller PPR threshold.  The slab lines of the targets are in
// flight meanwhile (eight 16-byte reads per lane: lane 8g+j holds chunk j of four links per read).  Then every WARP
// is on its own with its 32 links: probe the slab lines, flatten the overflow chunks of the links not flagged yet
// into a list (one lane per chunk, four reads in flight), probe those, and append the flagged links — a common
// neighbour, or a node in both PPR rows above the threshold; 1-2 % of a citation2-shaped batch, false positives of
// the filters included — to the candidate list.  A filter has no capacity limit (a hub source only makes it
// denser), so there is no second launch for big sources and nothing is resolved here: the screening streams.
//
// RESOLVE — count -> allocate -> ordered write — one warp per candidate, grid-wide (the ~100 selecting links of a
// hub source spread over the whole GPU instead of serialising in the CTA that screened them).  The warp walks the
// SHORTER of the link's two packed rows (L2-hot: just screened, or shared by every candidate of the run), lane per
// 8-byte slot, and searches the other one in global memory (sorted ids, sorted PPR columns): exact, so the filters'
// false positives end here with three zero counts.  The hits of the (at most four) steps stay in registers across
// the allocation.  Only links whose shorter row exceeds 128 slots (hub-hub pairs) go on to the deferred-link kernel
// of the launch sequence, where a whole CTA walks them over the CSR tables (select_fast.cu).
// Pieces that are not run-shaped are appended to the candidate list whole.  Selected sets, their order inside a
// link and the fp32 values are those of every other K1 variant (tests compare all of them with the oracle).
#include <limits.h>
#include <stdlib.h>

#include "select_walk.cuh"

namespace lpf {

constexpr int kPkThreads = 256;        // screening launch: threads per CTA = links per piece
constexpr int kPkCtas = 4;             // CTAs per SM the screening is compiled for
constexpr int kPkBloomWords = 4096;    // 32-bit words of a piece's id filters (16 KB)
constexpr int kPkPprBloomWords = 256;  // ... of one run's PPR-column filter
constexpr int kPkMaxRuns = 3;
constexpr int kPkFirst = 7;            // row chunks in the slab line
constexpr int kPkMaxRowChunks = 1024;  // target rows with more 16-byte chunks are not screened (deferred: CSR walk of the short source row)
constexpr int kPkWarpItems = 512;      // overflow chunks of a warp's 32 links that are screened (the rest: candidates unscreened)
constexpr int kPkInflight = 4;         // overflow reads a lane keeps in flight
constexpr int kPkResolveThreads = 256;
#ifndef LPF_RESOLVE_CTAS
#define LPF_RESOLVE_CTAS 3
#endif
constexpr int kPkResolveCtas = LPF_RESOLVE_CTAS;   // resolve CTAs per SM (one warp per candidate: ~3,500 candidates per citation2-shaped batch)
constexpr int kPkBigSlots = 2048;       // links whose shorter row has up to this many slots take a whole CTA (select_resolve_big_kernel)
constexpr uint32_t kPkPprTag = 0x80000000u;
constexpr uint32_t kPkPad = 0x7fffffffu;
static_assert(kPkMaxRowChunks <= 2048, "items[] packs (link:5 | chunk:11)");

struct PkWarpTab {
    uint4 hdr[32];                   // slab header of every link's target: (deg, nP, overflow unit, 0)
    uint16_t items[kPkWarpItems];    // (link << 11 | row chunk) of the overflow chunks to screen
    uint32_t any;                    // links flagged by the overflow screening
    uint32_t pad[3];
};

struct PkSmem {
    uint32_t bloom[kPkBloomWords];                       // id filters of the piece's runs, side by side
    uint32_t pbloom[kPkMaxRuns][kPkPprBloomWords];       // PPR-column filter of every run
    PkWarpTab warp[kPkThreads / 32];
    uint32_t bnd[kPkThreads / 32];       // ballots of "a run starts here"
    int32_t l_a[kPkThreads];             // source node of the links that start a run
    uint4 r_hdr[kPkMaxRuns];             // slab header of every run's source
    int dbg_ph[16];                      // profiling: this piece's cycles per phase
};

__host__ __device__ __forceinline__ int64_t row_chunks(int64_t deg, int64_t npp) { return ((npp + 1) >> 1) + ((deg + 3) >> 2); }
// 128-byte units of a row's overflow
__host__ __device__ __forceinline__ int64_t ovf_units(int64_t deg, int64_t npp) {
    const int64_t rc = row_chunks(deg, npp);
    return rc > kPkFirst ? (rc - kPkFirst + 7) >> 3 : 0;
}

__device__ __forceinline__ uint4 ldg16(const uint4* p) { return __ldg(p); }

// A packed row as the resolution walks it: chunk c < 7 in the slab line, the others in the overflow.
struct RowView {
    const uint4* blk;        // slab line (chunk 0 = header)
    const uint4* ovf;        // overflow of this row (chunk kPkFirst onwards)
    int deg, npp, pc, S;     // pc = PPR chunks; S = 8-byte slots of the row (padding included)
    int nb;                  // neighbour ids in the slab line; id k lives at idb[k] (k < nb) or ido[k] (k >= nb)
    const int32_t* idb;
    const int32_t* ido;
    __device__ __forceinline__ const uint4* chunk(int c) const { return c < kPkFirst ? blk + 1 + c : ovf + (c - kPkFirst); }
    __device__ __forceinline__ uint2 slot(int s) const { return __ldg(reinterpret_cast<const uint2*>(chunk(s >> 1)) + (s & 1)); }
    __device__ __forceinline__ int32_t id_at(int k) const { return __ldg((k < nb ? idb : ido) + k); }
    __device__ __forceinline__ uint2 ppr_at(int e) const { return slot(e); }      // PPR entry e = slot e
    // node u among the row's neighbour ids?
    __device__ __forceinline__ bool has_id(int32_t u) const {
        int lo = 0, hi = deg;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (id_at(mid) < u) lo = mid + 1; else hi = mid;
        }
        return lo < deg && id_at(lo) == u;
    }
    // (present, q) of node u in the row's PPR entries
    __device__ __forceinline__ bool ppr(int32_t u, float& q) const {
        int lo = 0, hi = npp;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int32_t)(ppr_at(mid).x & 0x7fffffffu) < u) lo = mid + 1; else hi = mid;
        }
        q = 0.f;
        if (lo >= npp) return false;
        const uint2 e = ppr_at(lo);
        if ((int32_t)(e.x & 0x7fffffffu) != u) return false;
        q = quantise(__uint_as_float(e.y));
        return true;
    }
};
__device__ __forceinline__ RowView view_row(const uint4* __restrict__ slab, const uint4* __restrict__ ovf, int64_t node, uint4 hdr) {
    RowView v;
    v.deg = (int)hdr.x;
    v.npp = (int)hdr.y;
    v.pc = (v.npp + 1) >> 1;
    v.blk = slab + (size_t)node * 8;
    v.ovf = ovf + (size_t)hdr.z * 8;
    v.S = 2 * (v.pc + ((v.deg + 3) >> 2));
    v.nb = v.pc < kPkFirst ? 4 * (kPkFirst - v.pc) : 0;
    v.idb = reinterpret_cast<const int32_t*>(v.blk + 1 + min(v.pc, kPkFirst));
    v.ido = reinterpret_cast<const int32_t*>(v.ovf) + (v.pc > kPkFirst ? 4 * (v.pc - kPkFirst) : 0) - v.nb;
    return v;
}

// Blocked Bloom filter over node ids: two bits of ONE 32-bit word per id (word from the top bits of a multiplicative
// hash, the two bit positions from its low bits), so a probe is one shared-memory read.
__device__ __forceinline__ uint32_t bloom_hash(int32_t u) { return (uint32_t)u * 0x9E3779B1u; }
// (1 << (s & 31) as one funnel shift: the hardware shift clamps, the funnel shift wraps)
__device__ __forceinline__ uint32_t bloom_bits(uint32_t h) { return __funnelshift_l(0u, 1u, h) | __funnelshift_l(0u, 1u, h >> 5); }
__device__ __forceinline__ void bloom_insert(uint32_t* bl, int lgw, int32_t u) {
    const uint32_t h = bloom_hash(u);
    atomicOr(bl + (h >> (32 - lgw)), bloom_bits(h));
}
__device__ __forceinline__ bool bloom_has(const uint32_t* bl, int lgw, int32_t u) {
    const uint32_t h = bloom_hash(u), m = bloom_bits(h);
    return (bl[h >> (32 - lgw)] & m) == m;
}
// One 16-byte chunk of a target row against the source's filters, without a branch on the chunk's kind (the lanes of
// a warp hold both kinds).  An id chunk (x, y, z, w = four ids): any of them (maybe) a neighbour of the source.  A PPR
// chunk ((x, y), (z, w) = two (col | tag, value) entries): an entry above the smaller PPR threshold whose column the
// source (maybe) holds above it too.  Words x and z are probed in the filter of the chunk's kind, y and w in the id
// filter (ignored for a PPR chunk, whose y and w are compared as values instead).
__device__ __forceinline__ bool chunk_hit(const uint32_t* bl, int lgw, const uint32_t* pbl, bool is_ppr, uint4 v, float th_pre) {
    const uint32_t* f = is_ppr ? pbl : bl;
    const int lgf = is_ppr ? 8 : lgw;
    const uint32_t hx = bloom_hash((int32_t)(v.x & 0x7fffffffu)), hz = bloom_hash((int32_t)(v.z & 0x7fffffffu));
    const uint32_t hy = bloom_hash((int32_t)v.y), hw = bloom_hash((int32_t)v.w);
    const uint32_t wx = f[hx >> (32 - lgf)], wz = f[hz >> (32 - lgf)];
    const uint32_t wy = bl[hy >> (32 - lgw)], ww = bl[hw >> (32 - lgw)];
    const uint32_t mx = bloom_bits(hx), mz = bloom_bits(hz), my = bloom_bits(hy), mw = bloom_bits(hw);
    // (ids ascend and the padding 0x7fffffff comes last; the padding entry of a PPR chunk is (0xffffffff, 0): value 0
    // fails the threshold)
    const bool tx = (wx & mx) == mx, tz = (wz & mz) == mz;
    const bool ty = (wy & my) == my && v.y != kPkPad, tw = (ww & mw) == mw && v.w != kPkPad;
    const bool qy = quantise(__uint_as_float(v.y)) >= th_pre, qw = quantise(__uint_as_float(v.w)) >= th_pre;
    return is_ppr ? ((tx && qy) || (tz && qw)) : ((tx && v.x != kPkPad) || ty || (tz && v.z != kPkPad) || tw);
}
static_assert(kPkPprBloomWords == 256, "ppr_chunk_hit / staging use lgw = 8");

// What slot s of a PACKED target row contributes against the source row: a PPR entry is a candidate 1-hop /
// >1-hop node (k1 / kn, node u, values qa, qb), a pair of neighbour ids up to two common neighbours (h0, h1 with
// values (qa, qb) and (qa1, qb1)).  Slots ascend by node id within each kind, so slot order is the output order.
struct SlotHit {
    bool k1, kn, h0, h1;
    int32_t u, w0, w1;
    float qa, qb, qa1, qb1;
};
// in0 / in1: the slot's two ids are among the searched row's ids (found by resolve_packed_warp's lockstep search)
// has_src / has_row: "node u is a neighbour in the searched / the walked row" (global-memory search, or a staged copy)
template <class HasSrc, class HasRow>
__device__ __forceinline__ SlotHit eval_slot_t(const SelectParams2& p, const RowView& src, const RowView& row, uint2 slot,
                                               bool in0, bool in1, bool want_pi, bool want_n1, float th_pre, HasSrc has_src,
                                               HasRow has_row) {
    SlotHit r;
    r.k1 = r.kn = r.h0 = r.h1 = false;
    r.qa = r.qb = r.qa1 = r.qb1 = 0.f;
    const uint32_t w0 = slot.x, w1 = slot.y;
    r.w0 = (int32_t)w0; r.w1 = (int32_t)w1;
    r.u = (int32_t)w0;
    if (w0 & kPkPprTag) {
        r.u = (int32_t)(w0 & ~kPkPprTag);
        if (want_pi && r.u != (int32_t)kPkPad && src.ppr(r.u, r.qa)) {
            r.qb = quantise(__uint_as_float(w1));
            if (r.qa >= th_pre && r.qb >= th_pre) {
                const bool in_a = has_src(r.u);
                const bool in_b = has_row(r.u);
                r.k1 = (in_a != in_b) && r.qa >= p.th_1hop && r.qb >= p.th_1hop;
                r.kn = want_n1 && !in_a && !in_b && r.qa >= p.th_non1hop && r.qb >= p.th_non1hop;
            }
        }
    } else {
        // PPR values of a common neighbour: by search over the PPR entries of both rows
        if (in0) {
            src.ppr((int32_t)w0, r.qa);
            row.ppr((int32_t)w0, r.qb);
            r.h0 = r.qa >= p.th_cn && r.qb >= p.th_cn;
        }
        if (in1) {
            src.ppr((int32_t)w1, r.qa1);
            row.ppr((int32_t)w1, r.qb1);
            r.h1 = r.qa1 >= p.th_cn && r.qb1 >= p.th_cn;
        }
    }
    return r;
}
__device__ __forceinline__ SlotHit eval_slot(const SelectParams2& p, const RowView& src, const RowView& row, uint2 slot,
                                             bool in0, bool in1, bool want_pi, bool want_n1, float th_pre) {
    return eval_slot_t(p, src, row, slot, in0, in1, want_pi, want_n1, th_pre, [&](int32_t u) { return src.has_id(u); },
                       [&](int32_t u) { return row.has_id(u); });
}
// node u among n ascending ids in shared memory?
__device__ __forceinline__ bool has_id_s(const int32_t* ids, int n, int32_t u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ids[mid] < u) lo = mid + 1; else hi = mid;
    }
    return lo < n && ids[lo] == u;
}
__device__ __forceinline__ uint2 load_slot(const RowView& row, int s) {
    if (s < row.S) return row.slot(s);
    return make_uint2(kPkPad, kPkPad);
}
// (swapped: the walked row is the SOURCE's and the searched one the target's, so qa / qb change places)
__device__ __forceinline__ void write_hits(const SelectParams2& p, const SlotHit& r, bool swapped, int64_t r_pi, int64_t r_cn) {
    if (r.k1 || r.kn) { p.node[r_pi] = r.u; p.pa[r_pi] = swapped ? r.qb : r.qa; p.pb[r_pi] = swapped ? r.qa : r.qb; }
    if (r.h0) { p.node[r_cn] = r.w0; p.pa[r_cn] = swapped ? r.qb : r.qa; p.pb[r_cn] = swapped ? r.qa : r.qb; }
    if (r.h1) {
        const int64_t r1 = r_cn + (r.h0 ? 1 : 0);
        p.node[r1] = r.w1; p.pa[r1] = swapped ? r.qb1 : r.qa1; p.pb[r1] = swapped ? r.qa1 : r.qb1;
    }
}

// count -> allocate -> ordered write of one link by a warp: the warp walks `row` (the shorter of the link's two packed
// rows, at most kPkResolveSlots slots), lane l taking slot l, l + 32, ... so that ascending node order within each
// set is lane order and the ordered write needs only ballots, and searches `src` (the other row).  The hits of all
// (at most four) steps stay in registers across the allocation: one walk.  Same sets, order and values as the
// generic walk (select_walk.cuh).
constexpr int kPkResolveSlots = 128;
__device__ __forceinline__ void resolve_packed_warp(const SelectParams2& p, const RowView& src, const RowView& row, bool swapped,
                                                    int64_t i, int lane) {
    constexpr int STEPS = kPkResolveSlots / 32;
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    const unsigned lt = (1u << lane) - 1u;
    // the lane's slots of all steps, and its (up to 2 * STEPS) id keys searched in the other row's ids in LOCKSTEP: one
    // dependent read per halving for all keys together instead of one binary search after the other
    uint2 sl[STEPS];
#pragma unroll
    for (int s = 0; s < STEPS; ++s) sl[s] = 32 * s < row.S ? load_slot(row, 32 * s + lane) : make_uint2(kPkPad, kPkPad);
    int lo[2 * STEPS], hi[2 * STEPS];
#pragma unroll
    for (int x = 0; x < 2 * STEPS; ++x) {
        const uint32_t key = (x & 1) ? sl[x >> 1].y : sl[x >> 1].x;
        const bool active = !(sl[x >> 1].x & kPkPprTag) && key != kPkPad;
        lo[x] = 0;
        hi[x] = active ? src.deg : 0;
    }
    const int iters = 32 - __clz(src.deg);          // halvings that empty [0, deg)
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int x = 0; x < 2 * STEPS; ++x) {
            if (lo[x] < hi[x]) {
                const int32_t key = (int32_t)((x & 1) ? sl[x >> 1].y : sl[x >> 1].x);
                const int mid = (lo[x] + hi[x]) >> 1;
                if (src.id_at(mid) < key) lo[x] = mid + 1; else hi[x] = mid;
            }
        }
    }
    bool found[2 * STEPS];
#pragma unroll
    for (int x = 0; x < 2 * STEPS; ++x) {
        const uint32_t key = (x & 1) ? sl[x >> 1].y : sl[x >> 1].x;
        const bool active = !(sl[x >> 1].x & kPkPprTag) && key != kPkPad;
        found[x] = active && lo[x] < src.deg && src.id_at(lo[x]) == (int32_t)key;
    }
    SlotHit r[STEPS];
    unsigned m1[STEPS], mn[STEPS], mh0[STEPS], mh1[STEPS];
    int c_cn = 0, c_1h = 0, c_n1 = 0;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
        m1[s] = mn[s] = mh0[s] = mh1[s] = 0u;
        r[s].k1 = r[s].kn = r[s].h0 = r[s].h1 = false;
        if (32 * s < row.S) {       // uniform
            r[s] = eval_slot(p, src, row, sl[s], found[2 * s], found[2 * s + 1], want_pi, want_n1, th_pre);
            m1[s] = __ballot_sync(kFull, r[s].k1); mn[s] = __ballot_sync(kFull, r[s].kn);
            mh0[s] = __ballot_sync(kFull, r[s].h0); mh1[s] = __ballot_sync(kFull, r[s].h1);
            c_1h += __popc(m1[s]); c_n1 += __popc(mn[s]); c_cn += __popc(mh0[s]) + __popc(mh1[s]);
        }
    }
    int64_t s_cn, s_1h, s_n1;
    const bool ok = alloc_segments_warp(p, i, c_cn, c_1h, c_n1, lane, s_cn, s_1h, s_n1);
    if (c_cn + c_1h + c_n1 == 0 || !ok) return;
    int64_t o_cn = s_cn, o_1h = p.cap + s_1h, o_n1 = 2 * p.cap + s_n1;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
        write_hits(p, r[s], swapped, r[s].k1 ? o_1h + __popc(m1[s] & lt) : o_n1 + __popc(mn[s] & lt),
                   o_cn + __popc(mh0[s] & lt) + __popc(mh1[s] & lt));
        o_1h += __popc(m1[s]); o_n1 += __popc(mn[s]); o_cn += __popc(mh0[s]) + __popc(mh1[s]);
    }
}

// The common candidate — the walked row is one step of the warp (at most 32 slots) and both PPR rows have at most 32
// entries — with two dependent reads instead of ~35: the walked row's slots and the searched row's PPR entries sit
// in registers (one per lane) and are looked up by shuffle loops; the searched row's neighbour ids are staged in
// shared memory (up to kPkStageIds of them; a longer row is searched in global memory, both keys in lockstep).
constexpr int kPkStageIds = 256;
__device__ __forceinline__ void resolve_packed_fast(const SelectParams2& p, const RowView& src, const RowView& row, bool swapped,
                                                    int64_t i, int lane, int32_t* ids_sm) {
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    const uint2 sl = load_slot(row, lane);
    const uint2 sp = lane < src.npp ? src.ppr_at(lane) : make_uint2(0xffffffffu, 0u);
    const bool staged = src.deg <= kPkStageIds;
    if (staged)
        for (int k = lane; k < src.deg; k += 32) ids_sm[k] = src.id_at(k);
    __syncwarp();
    const bool is_ppr = (sl.x & kPkPprTag) != 0u;
    // keys: a PPR slot asks for its column u, an id slot for its two ids (-1: nothing to ask)
    const int32_t u = (is_ppr && sl.x != 0xffffffffu) ? (int32_t)(sl.x & ~kPkPprTag) : -1;
    const int32_t k0 = (!is_ppr && sl.x != kPkPad) ? (int32_t)sl.x : -1;
    const int32_t k1 = (!is_ppr && sl.y != kPkPad) ? (int32_t)sl.y : -1;
    // ---- the two ids among the searched row's ids?
    bool in0 = false, in1 = false;
    {
        int lo0 = 0, hi0 = k0 >= 0 ? src.deg : 0, lo1 = 0, hi1 = k1 >= 0 ? src.deg : 0;
        const int iters = 32 - __clz(src.deg);
        if (staged) {
            for (int it = 0; it < iters; ++it) {
                if (lo0 < hi0) { const int mid = (lo0 + hi0) >> 1; if (ids_sm[mid] < k0) lo0 = mid + 1; else hi0 = mid; }
                if (lo1 < hi1) { const int mid = (lo1 + hi1) >> 1; if (ids_sm[mid] < k1) lo1 = mid + 1; else hi1 = mid; }
            }
            in0 = k0 >= 0 && lo0 < src.deg && ids_sm[lo0] == k0;
            in1 = k1 >= 0 && lo1 < src.deg && ids_sm[lo1] == k1;
        } else {
            for (int it = 0; it < iters; ++it) {
                const int m0 = (lo0 + hi0) >> 1, m1 = (lo1 + hi1) >> 1;
                const int32_t v0 = lo0 < hi0 ? src.id_at(m0) : 0, v1 = lo1 < hi1 ? src.id_at(m1) : 0;
                if (lo0 < hi0) { if (v0 < k0) lo0 = m0 + 1; else hi0 = m0; }
                if (lo1 < hi1) { if (v1 < k1) lo1 = m1 + 1; else hi1 = m1; }
            }
            const int32_t f0 = (k0 >= 0 && lo0 < src.deg) ? src.id_at(lo0) : -2, f1 = (k1 >= 0 && lo1 < src.deg) ? src.id_at(lo1) : -2;
            in0 = f0 == k0;
            in1 = f1 == k1;
        }
    }
    // ---- PPR values by shuffle loops: q(P(src, .)) of u, k0, k1 from the searched row's entries; q(P(row, .)) of k0, k1
    // from the walked row's own PPR slots (its first npp slots = the lanes below npp)
    float qa_u = 0.f, qa0 = 0.f, qa1 = 0.f, qb0 = 0.f, qb1 = 0.f;
    bool has_u = false;
    for (int e = 0; e < src.npp; ++e) {
        const int32_t c = (int32_t)(__shfl_sync(kFull, sp.x, e) & 0x7fffffffu);
        const float q = quantise(__uint_as_float(__shfl_sync(kFull, sp.y, e)));
        if (c == u) { has_u = true; qa_u = q; }
        if (c == k0) qa0 = q;
        if (c == k1) qa1 = q;
    }
    for (int e = 0; e < row.npp; ++e) {
        const int32_t c = (int32_t)(__shfl_sync(kFull, sl.x, e) & 0x7fffffffu);
        const float q = quantise(__uint_as_float(__shfl_sync(kFull, sl.y, e)));
        if (c == k0) qb0 = q;
        if (c == k1) qb1 = q;
    }
    SlotHit r;
    r.k1 = r.kn = false;
    r.u = u; r.w0 = k0; r.w1 = k1;
    r.qa = is_ppr ? qa_u : qa0; r.qb = is_ppr ? quantise(__uint_as_float(sl.y)) : qb0;
    r.qa1 = qa1; r.qb1 = qb1;
    r.h0 = in0 && qa0 >= p.th_cn && qb0 >= p.th_cn;
    r.h1 = in1 && qa1 >= p.th_cn && qb1 >= p.th_cn;
    if (want_pi && has_u && r.qa >= th_pre && r.qb >= th_pre) {      // (rare: a node in both PPR rows above the threshold)
        const bool in_a = src.has_id(u), in_b = row.has_id(u);
        r.k1 = (in_a != in_b) && r.qa >= p.th_1hop && r.qb >= p.th_1hop;
        r.kn = want_n1 && !in_a && !in_b && r.qa >= p.th_non1hop && r.qb >= p.th_non1hop;
    }
    const unsigned lt = (1u << lane) - 1u;
    const unsigned m1 = __ballot_sync(kFull, r.k1), mn = __ballot_sync(kFull, r.kn);
    const unsigned mh0 = __ballot_sync(kFull, r.h0), mh1 = __ballot_sync(kFull, r.h1);
    const int c_cn = __popc(mh0) + __popc(mh1), c_1h = __popc(m1), c_n1 = __popc(mn);
    int64_t s_cn, s_1h, s_n1;
    const bool ok = alloc_segments_warp(p, i, c_cn, c_1h, c_n1, lane, s_cn, s_1h, s_n1);
    if (c_cn + c_1h + c_n1 == 0 || !ok) return;
    write_hits(p, r, swapped, r.k1 ? p.cap + s_1h + __popc(m1 & lt) : 2 * p.cap + s_n1 + __popc(mn & lt),
               s_cn + __popc(mh0 & lt) + __popc(mh1 & lt));
}

// RESOLVE launch: one warp per candidate of the screening.  The warp walks the SHORTER of the link's two rows and
// searches the other; a link whose shorter row has more than kPkResolveSlots slots (a hub-hub pair) goes to the
// deferred-link kernel, where a whole CTA walks it.
__global__ void __launch_bounds__(kPkResolveThreads, kPkResolveCtas)
select_resolve_packed_kernel(const __grid_constant__ SelectParams2 p, const uint4* __restrict__ slab,
                             const uint4* __restrict__ ovf) {
    __shared__ int32_t ids_sm[kPkResolveThreads / 32][kPkStageIds];
    const int lane = threadIdx.x & 31;
    const int n = p.hub[0];
    const int warp0 = (blockIdx.x * kPkResolveThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kPkResolveThreads) >> 5;
    // Two passes over this warp's candidates: the common ones (one step of the warp, short PPR rows: resolve_packed_fast)
    // first, the others (bit k of `slow` = the warp's k-th candidate) after — the two paths are long straight-line code,
    // and warps that alternate between them keep missing the instruction cache (measured: "no instruction" was the
    // second stall reason of this kernel).
    unsigned slow = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int k = 0;
        for (int q = warp0; q < n; q += nwarps, ++k) {
            if (pass == 1 && !(k < 32 && ((slow >> k) & 1u))) continue;
            const int64_t i = p.hub[4 + q];
            const int64_t a = __ldg(p.links + i), b = __ldg(p.links + p.bs + i);
            const uint4 ha = ldg16(slab + (size_t)a * 8), hb = ldg16(slab + (size_t)b * 8);
            const RowView A = view_row(slab, ovf, a, ha), B = view_row(slab, ovf, b, hb);
            const bool swapped = A.S < B.S;
            if (min(A.S, B.S) > kPkResolveSlots) {
                // a hub-hub pair: a whole CTA walks the shorter row (select_resolve_big_kernel), or — beyond kPkBigSlots —
                // the deferred-link kernel over the CSR tables
                if (lane == 0) {
                    if (min(A.S, B.S) <= kPkBigSlots) p.hub[ws_list_words(p.bs) + atomicAdd(p.hub + 1, 1)] = (int32_t)i;
                    else p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
                }
                continue;
            }
            const RowView& row = swapped ? A : B;
            const RowView& src = swapped ? B : A;
            const bool fast = row.S <= 32 && row.npp <= 32 && src.npp <= 32;
            if (pass == 0 && !fast && k < 32) {          // (beyond 32 candidates per warp: in place)
                slow |= 1u << k;
                continue;
            }
            __syncwarp();       // (ids_sm of the previous candidate is no longer read)
            if (fast) resolve_packed_fast(p, src, row, swapped, i, lane, ids_sm[threadIdx.x >> 5]);
            else resolve_packed_warp(p, src, row, swapped, i, lane);
        }
        if (!slow) break;
    }
}

// RESOLVE, hub-hub pairs: one CTA per link whose shorter row has 129 .. kPkBigSlots slots.  Thread t takes the slots
// t, t + 256, ... (kPkBigPer of them), searches their (up to 2 kPkBigPer) ids in the other row in lockstep — one
// dependent read per halving for all of them — and keeps its hits in registers; the ordered positions come from one
// block-wide scan per group of 256 slots; one allocation, one write.  Same sets, order and values as every other
// variant.
constexpr int kPkBigThreads = 256;
constexpr int kPkBigPer = 8;
constexpr int kPkBigStage = 8192;      // neighbour ids of the searched row staged in shared memory (32 KB)
constexpr int kPkBigSmem = (kPkBigStage + 2 * kPkBigSlots) * 4;    // + the walked row's ids (16 KB)
static_assert(kPkBigSlots == kPkBigThreads * kPkBigPer, "2,048 slots: a row of ~4,000 neighbours");
__global__ void __launch_bounds__(kPkBigThreads, 2)
select_resolve_big_kernel(const __grid_constant__ SelectParams2 p, const uint4* __restrict__ slab, const uint4* __restrict__ ovf) {
    constexpr int NW = kPkBigThreads / 32, M = kPkBigPer;
    __shared__ uint32_t wtot[M][NW];
    __shared__ int64_t seg[3];
    __shared__ int ok_s;
    // the searched row's neighbour ids: one coalesced pass over the row instead of ~13 dependent L2 / DRAM trips per
    // search (the halvings then cost a shared-memory read each)
    // ... and the walked row's (<= 2 kPkBigSlots of them, from the slots this CTA holds): the neighbour tests of the PPR
    // slots are then two searches in shared memory instead of ~26 dependent trips
    extern __shared__ __align__(16) int32_t big_smem[];
    int32_t* ids_s = big_smem;
    int32_t* ids_r = big_smem + kPkBigStage;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.hub[1];
    const int32_t* list = p.hub + ws_list_words(p.bs);
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int64_t i = list[q];
        const int64_t a = __ldg(p.links + i), b = __ldg(p.links + p.bs + i);
        const uint4 ha = ldg16(slab + (size_t)a * 8), hb = ldg16(slab + (size_t)b * 8);
        const RowView A = view_row(slab, ovf, a, ha), B = view_row(slab, ovf, b, hb);
        const bool swapped = A.S < B.S;
        const RowView& row = swapped ? A : B;
        const RowView& src = swapped ? B : A;
        uint2 sl[M];
#pragma unroll
        for (int m = 0; m < M; ++m) sl[m] = load_slot(row, m * kPkBigThreads + tid);
        const bool staged = src.deg <= kPkBigStage;
        if (staged)
            for (int k = tid; k < src.deg; k += kPkBigThreads) ids_s[k] = src.id_at(k);
#pragma unroll
        for (int m = 0; m < M; ++m) {                    // id slot s holds ids 2 (s - PPR slots), + 1 (padded with the maximum)
            const int k = 2 * (m * kPkBigThreads + tid - 2 * row.pc);
            if (k >= 0 && m * kPkBigThreads + tid < row.S) { ids_r[k] = (int32_t)sl[m].x; ids_r[k + 1] = (int32_t)sl[m].y; }
        }
        __syncthreads();
        auto sid = [&](int k) -> int32_t { return staged ? ids_s[k] : src.id_at(k); };
        int lo[2 * M], hi[2 * M];
#pragma unroll
        for (int x = 0; x < 2 * M; ++x) {
            const uint32_t key = (x & 1) ? sl[x >> 1].y : sl[x >> 1].x;
            const bool active = !(sl[x >> 1].x & kPkPprTag) && key != kPkPad;
            lo[x] = 0;
            hi[x] = active ? src.deg : 0;
        }
        const int iters = 32 - __clz(src.deg);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int x = 0; x < 2 * M; ++x) {
                if (lo[x] < hi[x]) {
                    const int32_t key = (int32_t)((x & 1) ? sl[x >> 1].y : sl[x >> 1].x);
                    const int mid = (lo[x] + hi[x]) >> 1;
                    if (sid(mid) < key) lo[x] = mid + 1; else hi[x] = mid;
                }
            }
        }
        SlotHit r[M];
        uint32_t mine[M], inc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t k0 = sl[m].x, k1 = sl[m].y;
            const bool idslot = !(k0 & kPkPprTag);
            const bool in0 = idslot && k0 != kPkPad && lo[2 * m] < src.deg && sid(lo[2 * m]) == (int32_t)k0;
            const bool in1 = idslot && k1 != kPkPad && lo[2 * m + 1] < src.deg && sid(lo[2 * m + 1]) == (int32_t)k1;
            r[m] = eval_slot_t(p, src, row, sl[m], in0, in1, want_pi, want_n1, th_pre,
                               [&](int32_t u) { return staged ? has_id_s(ids_s, src.deg, u) : src.has_id(u); },
                               [&](int32_t u) { return has_id_s(ids_r, row.deg, u); });
            // counts of this slot packed in one word: common neighbours (0..2) | 1-hop << 12 | >1-hop << 22
            mine[m] = (uint32_t)((r[m].h0 ? 1 : 0) + (r[m].h1 ? 1 : 0)) | (r[m].k1 ? 1u << 12 : 0u) | (r[m].kn ? 1u << 22 : 0u);
            inc[m] = mine[m];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(kFull, inc[m], o);
                if (lane >= o) inc[m] += x;
            }
            if (lane == 31) wtot[m][warp] = inc[m];
        }
        __syncthreads();
        // exclusive position of every slot's hits within the link: groups of 256 slots in order, warps in order
        int c_cn = 0, c_1h = 0, c_n1 = 0;
        int before[M];          // position of this slot's first hit within its set (a slot feeds one set only)
#pragma unroll
        for (int m = 0; m < M; ++m) {
            uint32_t bf = inc[m] - mine[m], total = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t x = wtot[m][w];
                bf += (w < warp) ? x : 0u;
                total += x;
            }
            before[m] = r[m].k1 ? c_1h + (int)((bf >> 12) & 1023u) : (r[m].kn ? c_n1 + (int)(bf >> 22) : c_cn + (int)(bf & 4095u));
            c_cn += (int)(total & 4095u);
            c_1h += (int)((total >> 12) & 1023u);
            c_n1 += (int)(total >> 22);
        }
        if (warp == 0) {
            int64_t s0, s1, s2;
            const bool fits = alloc_segments_warp(p, i, c_cn, c_1h, c_n1, lane, s0, s1, s2);
            if (lane == 0) { ok_s = fits ? 1 : 0; seg[0] = s0; seg[1] = s1; seg[2] = s2; }
        }
        __syncthreads();
        if (ok_s && c_cn + c_1h + c_n1 > 0) {
            const int64_t o_cn = seg[0], o_1h = p.cap + seg[1], o_n1 = 2 * p.cap + seg[2];
#pragma unroll
            for (int m = 0; m < M; ++m)
                write_hits(p, r[m], swapped, (r[m].k1 ? o_1h : o_n1) + before[m], o_cn + before[m]);
        }
        __syncthreads();       // wtot / seg / ok_s are reused by the next link
    }
}

// the links of `mask` (lane l = link i_first + l) appended to a list (count in word 0, entries from word 4)
__device__ __forceinline__ void push_links(int32_t* list, unsigned mask, int64_t i_first, int lane) {
    if (mask == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(list, __popc(mask));
    base = __shfl_sync(kFull, base, 0);
    if ((mask >> lane) & 1u) list[4 + base + __popc(mask & ((1u << lane) - 1u))] = (int32_t)(i_first + lane);
}

// SCREEN launch: one CTA per piece of kPkThreads links
__global__ void __launch_bounds__(kPkThreads, kPkCtas)
select_screen_packed_kernel(const __grid_constant__ SelectParams2 p, const uint4* __restrict__ slab,
                            const uint4* __restrict__ ovf) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PkSmem& sm = *reinterpret_cast<PkSmem*>(smem_raw);
    constexpr int NT = kPkThreads, NW = NT / 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, j = lane & 7;       // lane 8g + j holds chunk j of a slab line
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    PkWarpTab& W = sm.warp[warp];

    long long t_mark = clock64();
#define LPF_PHASE(k)                                                              \
    do {                                                                          \
        if (p.dbg && tid == 0) {                                                  \
            const long long now = clock64();                                      \
            atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (k)), (unsigned long long)(now - t_mark)); \
            t_mark = now;                                                         \
        }                                                                         \
    } while (0)

    const int64_t i0 = (int64_t)blockIdx.x * NT;
    const int len = (int)min((int64_t)NT, p.bs - i0);
    const int t0 = 32 * warp;
    if (p.dbg && tid == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 5), 1ull);
        if (blockIdx.x < 2048) {       // (profiling: start time and SM of every piece)
            unsigned long long gt; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.dbg[64 + 3 * blockIdx.x] = (long long)gt;
            p.dbg[64 + 3 * blockIdx.x + 2] = (long long)smid;
        }
    }
    // ---- the piece's links, one per thread; run boundaries by ballot
    const bool valid = tid < len;
    int64_t b_me = 0;
    {
        const int64_t i = i0 + tid;
        int64_t a_me = -1, a_prev = -1;
        if (valid) {
            a_me = __ldg(p.links + i);
            a_prev = tid > 0 ? __ldg(p.links + i - 1) : -1;
            b_me = __ldg(p.links + p.bs + i);
            // every link starts as "nothing selected"; the links that are resolved later overwrite their entries
            p.counts[i] = 0; p.counts[p.bs + i] = 0; p.counts[2 * p.bs + i] = 0;
            p.seg_start[i] = 0; p.seg_start[p.bs + i] = 0; p.seg_start[2 * p.bs + i] = 0;
        }
        const bool starts = valid && (tid == 0 || a_me != a_prev);
        const unsigned bm = __ballot_sync(kFull, starts);
        if (lane == 0) sm.bnd[warp] = bm;
        if (starts) sm.l_a[tid] = (int32_t)a_me;
    }
    // ---- the slab lines of this warp's 32 targets: 8 reads of 16 bytes per lane, in flight across the staging
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int l = 4 * k + grp;
        const int64_t b_l = __shfl_sync(kFull, b_me, l);
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (t0 + l < len) v[k] = ldg16(slab + (size_t)b_l * 8 + j);
    }
    // filters cleared
    for (int s = tid; s < kPkBloomWords / 4; s += NT) reinterpret_cast<uint4*>(sm.bloom)[s] = make_uint4(0u, 0u, 0u, 0u);
    for (int s = tid; s < kPkMaxRuns * kPkPprBloomWords; s += NT) (&sm.pbloom[0][0])[s] = 0u;
    if (lane == 0) W.any = 0;
    __syncthreads();                                   // (1) boundaries known, filters cleared
    LPF_PHASE(10);
    // every thread: the runs of the piece (first kPkMaxRuns boundaries)
    int n_runs = 0, rs1 = len, rs2 = len, rs0 = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        unsigned bm = sm.bnd[w];
        while (bm) {
            const int pos = 32 * w + __ffs(bm) - 1;
            bm &= bm - 1;
            if (n_runs == 0) rs0 = pos;
            else if (n_runs == 1) rs1 = pos;
            else if (n_runs == 2) rs2 = pos;
            ++n_runs;
        }
    }
    if (n_runs > kPkMaxRuns) {
        // not run-shaped: every link is a candidate (resolved warp by warp in the next launch)
        push_links(p.hub, __ballot_sync(kFull, valid), i0 + t0, lane);
        return;
    }
    auto run_of = [&](int tt) -> int { return (tt >= rs1 ? 1 : 0) + (tt >= rs2 ? 1 : 0); };
    // ---- stage the sources: eight threads per run read its slab line ...
    uint4 sv = make_uint4(kPkPad, kPkPad, kPkPad, kPkPad);
    const int my_run = tid >> 3;
    if (my_run < n_runs) {
        const int32_t a_r = sm.l_a[my_run == 0 ? rs0 : (my_run == 1 ? rs1 : rs2)];
        sv = ldg16(slab + (size_t)a_r * 8 + (tid & 7));
        if ((tid & 7) == 0) sm.r_hdr[my_run] = sv;
    }
    __syncthreads();                                   // (2) source headers known
    LPF_PHASE(12);
    // filter geometry of run r: a fixed share of the words (all for one run, half for two, a quarter each for three),
    // 256 bits per neighbour id while that fits
    const int share_lg = n_runs == 1 ? 12 : (n_runs == 2 ? 11 : 10);
    int lgw[kPkMaxRuns];
#pragma unroll
    for (int r = 0; r < kPkMaxRuns; ++r) {
        lgw[r] = 5;
        if (r < n_runs) {
            const uint32_t na = min(sm.r_hdr[r].x, 1u << 20);
            lgw[r] = min(share_lg, max(5, 32 - __clz((int)(8u * na) - 1)));       // words: a power of two >= 8 * deg
        }
    }
    static_assert(kPkBloomWords == 4096, "share_lg");
    // ... and every chunk of every source goes to its filters (the slab chunks by the threads that read them, the
    // overflow by everybody, four reads in flight)
    auto stage_chunk = [&](int r, int lg, int c, uint4 cv, int npa, int pca) {
        if (c < pca) {
            if (2 * c < npa && quantise(__uint_as_float(cv.y)) >= th_pre) bloom_insert(sm.pbloom[r], 8, (int32_t)(cv.x & ~kPkPprTag));
            if (2 * c + 1 < npa && quantise(__uint_as_float(cv.w)) >= th_pre) bloom_insert(sm.pbloom[r], 8, (int32_t)(cv.z & ~kPkPprTag));
        } else {
            uint32_t* bl = sm.bloom + (r << share_lg);
            if (cv.x != kPkPad) bloom_insert(bl, lg, (int32_t)cv.x);
            if (cv.y != kPkPad) bloom_insert(bl, lg, (int32_t)cv.y);
            if (cv.z != kPkPad) bloom_insert(bl, lg, (int32_t)cv.z);
            if (cv.w != kPkPad) bloom_insert(bl, lg, (int32_t)cv.w);
        }
    };
#pragma unroll
    for (int r = 0; r < kPkMaxRuns; ++r) {
        if (r >= n_runs) continue;
        const uint4 hd = sm.r_hdr[r];
        const int na = (int)min(hd.x, 1u << 28), npa = (int)min(hd.y, 1u << 28);
        const int pca = (npa + 1) >> 1, rca = pca + ((na + 3) >> 2);
        if (my_run == r && (tid & 7) >= 1 && (tid & 7) - 1 < rca) stage_chunk(r, lgw[r], (tid & 7) - 1, sv, npa, pca);
        const uint4* orow = ovf + (size_t)hd.z * 8;
        for (int c0 = kPkFirst + tid; c0 < rca; c0 += 4 * NT) {
            uint4 cv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c0 + k * NT < rca) cv[k] = ldg16(orow + (c0 + k * NT - kPkFirst));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c0 + k * NT < rca) stage_chunk(r, lgw[r], c0 + k * NT, cv[k], npa, pca);
        }
    }
    __syncthreads();                                   // (3) sources staged: from here every warp is on its own
    LPF_PHASE(0);

    // ---- this warp's 32 links: the slab lines.  Lane 8g+j probes chunk j of link 4k+g; its hits of the eight reads
    // are collected in a lane-local byte and OR-ed over the eight lanes of the group at the end.
    unsigned lane_hits = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int l = 4 * k + grp;
        if (j == 0) W.hdr[l] = v[k];
        const uint32_t geo = __shfl_sync(kFull, v[k].w, 8 * grp);       // header word 3: PPR chunks | row chunks << 16
        const int pc = (int)(geo & 0xffffu), rc = (int)(geo >> 16);
        const int c = j - 1;
        if (j >= 1 && c < rc && (want_pi || c >= pc)) {
            const int r = run_of(t0 + l);
            if (chunk_hit(sm.bloom + (r << share_lg), r == 0 ? lgw[0] : (r == 1 ? lgw[1] : lgw[2]), sm.pbloom[r], c < pc, v[k], th_pre))
                lane_hits |= 1u << k;
        }
    }
    lane_hits |= __shfl_xor_sync(kFull, lane_hits, 1);
    lane_hits |= __shfl_xor_sync(kFull, lane_hits, 2);
    lane_hits |= __shfl_xor_sync(kFull, lane_hits, 4);
    // bit l of any_mask: link l = 4k+g of this warp (maybe) selects something (lane l looks at group l&3, read l>>2)
    unsigned any_mask = __ballot_sync(kFull, (__shfl_sync(kFull, lane_hits, 8 * (lane & 3)) >> (lane >> 2)) & 1u);
    __syncwarp();
    // ---- flatten the overflow chunks of the links not flagged yet: items[] = (link << 11 | row chunk)
    int n_items;
    unsigned long_mask;         // rows too long to screen: candidates unscreened
    {
        const int l = lane;
        const uint32_t geo = W.hdr[l].w;
        const int rc_me = (int)(geo >> 16);
        const bool live = t0 + l < len;
        const bool too_long = live && rc_me > kPkMaxRowChunks;
        const bool listed = live && !too_long && !((any_mask >> l) & 1u) && rc_me > kPkFirst;
        const int mine = listed ? rc_me - kPkFirst : 0;
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += x;
        }
        const int ex = inc - mine;
        const bool fits = ex + mine <= kPkWarpItems;
        // the list ends at the first link that does not fit (a suffix: the prefix sum is monotone); those links
        // become candidates unscreened
        const unsigned nofit = __ballot_sync(kFull, listed && !fits);
        // a link with up to kOwn overflow chunks (most of them: deg <= ~60) is written by its own lane — eight predicated
        // stores for the whole warp; only the longer rows take the warp-wide loop below (one iteration per link: with every
        // listed link in it, that loop was 18 % of the kernel's instructions)
        constexpr int kOwn = 8;
        const bool own = listed && fits && mine <= kOwn;
#pragma unroll
        for (int e = 0; e < kOwn; ++e)
            if (own && e < mine) W.items[ex + e] = (uint16_t)((l << 11) | (kPkFirst + e));
        unsigned todo = __ballot_sync(kFull, listed && fits && mine > kOwn);
        long_mask = __ballot_sync(kFull, too_long);
        any_mask = (any_mask | nofit) & ~long_mask;
        const int first_nofit = nofit ? __ffs(nofit) - 1 : 31;
        n_items = __shfl_sync(kFull, nofit ? ex : inc, first_nofit);
        // the whole warp writes the items of one link at a time
        while (todo) {
            const int ll = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ex_l = __shfl_sync(kFull, ex, ll), n_l = __shfl_sync(kFull, mine, ll);
            for (int e = lane; e < n_l; e += 32) W.items[ex_l + e] = (uint16_t)((ll << 11) | (kPkFirst + e));
        }
    }
    __syncwarp();
    // ---- the overflow chunks: one lane per chunk, kPkInflight reads in flight per lane
    {
        constexpr int K = kPkInflight;
        for (int q0 = lane; q0 < n_items; q0 += K * 32) {
            uint4 cv[K];
            int li[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int q = q0 + 32 * k;
                li[k] = -1;
                if (q < n_items) {
                    li[k] = (int)W.items[q];
                    cv[k] = ldg16(ovf + (size_t)W.hdr[li[k] >> 11].z * 8 + ((li[k] & 2047) - kPkFirst));
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (li[k] < 0) continue;
                const int l = li[k] >> 11, c = li[k] & 2047;
                const int r = run_of(t0 + l);
                const int pc = (int)(W.hdr[l].w & 0xffffu);
                if (!want_pi && c < pc) continue;
                if (chunk_hit(sm.bloom + (r << share_lg), r == 0 ? lgw[0] : (r == 1 ? lgw[1] : lgw[2]), sm.pbloom[r], c < pc, cv[k], th_pre))
                    atomicOr(&W.any, 1u << l);
            }
        }
    }
    __syncwarp();
    any_mask |= W.any;
    LPF_PHASE(1);
    // ---- the flagged links, and the rows too long to screen, to the candidate list
    push_links(p.hub, any_mask | long_mask, i0 + t0, lane);
    LPF_PHASE(2);
    if (p.dbg) {
        __syncthreads();
        if (tid == 0 && blockIdx.x < 2048) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            p.dbg[64 + 3 * blockIdx.x + 1] = (long long)gt;
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
    units[x] = (int32_t)ovf_units(arp[x + 1] - arp[x], prp[x + 1] - prp[x]);
}

__global__ void __launch_bounds__(256) pack_fill_kernel(const int64_t* __restrict__ arp, const int32_t* __restrict__ ac,
                                                        const int64_t* __restrict__ prp, const int32_t* __restrict__ pc,
                                                        const float* __restrict__ pv, int64_t n,
                                                        const int64_t* __restrict__ off, uint32_t* __restrict__ slab,
                                                        uint32_t* __restrict__ ovf) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t x = warp0; x < n; x += nwarps) {
        const int64_t a0 = arp[x], p0 = prp[x];
        const int64_t deg = arp[x + 1] - a0, npp = prp[x + 1] - p0;
        const int64_t o = off[x], units = ovf_units(deg, npp);
        const int64_t pcn = (npp + 1) >> 1, rc = row_chunks(deg, npp);
        // word wi of row chunk c
        auto word = [&](int64_t c, int wi) -> uint32_t {
            if (c < pcn) {
                const int64_t e = 2 * c + (wi >> 1);
                if (e < npp) return (wi & 1) ? __float_as_uint(pv[p0 + e]) : ((uint32_t)pc[p0 + e] | kPkPprTag);
                return (wi & 1) ? 0u : 0xffffffffu;
            }
            if (c < rc) {
                const int64_t k = 4 * (c - pcn) + wi;
                return k < deg ? (uint32_t)ac[a0 + k] : kPkPad;
            }
            return kPkPad;
        };
        uint32_t* line = slab + x * 32;
        const uint32_t geo = (uint32_t)min(pcn, (int64_t)65535) | ((uint32_t)min(rc, (int64_t)65535) << 16);
        if (lane < 4) line[lane] = lane == 0 ? (uint32_t)deg : (lane == 1 ? (uint32_t)npp : (lane == 2 ? (uint32_t)o : geo));
        else line[lane] = word((lane >> 2) - 1, lane & 3);
        uint32_t* w = ovf + o * 32;
        for (int64_t k = lane; k < units * 32; k += 32) w[k] = word(kPkFirst + (k >> 2), (int)(k & 3));
    }
}

extern long long* g_select_dbg;
// profiling hook (lpf_debug_select_timing): CUDA events around the three kernels of the packed launch sequence
bool g_kernel_timing = false;     // shared with nz_fused.cu
static cudaEvent_t g_pk_ev[5];
static bool g_pk_ev_ready = false, g_pk_ev_valid = false;
static int g_pk_slot_limit = 0;
void launch_onepass_reset(const SelectParams2& p, cudaStream_t st);
void launch_onepass_tail(const SelectParams2& p, cudaStream_t st);
void launch_onepass_heavy(const SelectParams2& p, cudaStream_t st);
void launch_onepass_finalize(const SelectParams2& p, cudaStream_t st);

}  // namespace lpf

using namespace lpf;

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int64_t lpf_link_rows_slab_bytes(int64_t n) { return n < 0 ? -1 : n * 128; }

extern "C" int64_t lpf_link_rows_bytes(int64_t n, int64_t adj_nnz, int64_t ppr_nnz) {
    if (n < 0 || adj_nnz < 0 || ppr_nnz < 0) return -1;
    // a row overflows with its chunks beyond the seventh, rounded up to 128 B: at most 16 B per row chunk, and the row
    // chunks hold the PPR entries padded to an even count (8 nP + 8) and the ids padded to four (4 deg + 12)
    const int64_t bytes = align_up(8 * ppr_nnz + 4 * adj_nnz + 20 * n + 128, 128);
    if (bytes / 128 >= ((int64_t)1 << 32)) return -1;    // 32-bit unit index in the header
    return bytes;
}

extern "C" int64_t lpf_link_rows_scratch_bytes(int64_t n) {
    if (n < 0) return -1;
    return align_up(n * 4, 16) + align_up((n + 1) * 8, 16) + align_up(lpf_scan_scratch_bytes(n), 16);
}

extern "C" int lpf_pack_link_rows(const int64_t* adj_rowptr, const int32_t* adj_col, const int64_t* ppr_rowptr,
                                  const int32_t* ppr_col, const float* ppr_val, int64_t n, void* slab,
                                  void* overflow, void* scratch, void* stream) {
    LPF_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) - 1, "bad node count");
    LPF_REQUIRE(adj_rowptr && ppr_rowptr, "rowptr is NULL");
    LPF_REQUIRE(n == 0 || (slab && overflow && scratch), "NULL output");
    LPF_REQUIRE((reinterpret_cast<uintptr_t>(slab) & 127) == 0 && (reinterpret_cast<uintptr_t>(overflow) & 127) == 0,
                "slab / overflow must be 128-byte aligned");
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
                                                       static_cast<uint32_t*>(slab), static_cast<uint32_t*>(overflow));
    return check_launch("lpf_pack_link_rows");
}

extern "C" int lpf_select_onepass_packed(const int64_t* links, int64_t bs, const int64_t* adj_rowptr,
                                         const int32_t* adj_col, const int64_t* ppr_rowptr, const int32_t* ppr_col,
                                         const float* ppr_val, const void* slab, const void* overflow,
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
    LPF_REQUIRE(slab && overflow, "slab / overflow is NULL (lpf_pack_link_rows)");
    LPF_REQUIRE(cap >= 0 && 3 * cap < ((int64_t)1 << 31), "bad pair capacity");
    LPF_REQUIRE(header && workspace, "header/workspace is NULL");
    LPF_REQUIRE(bs == 0 || (counts && seg_start && nz_list), "NULL output");
    LPF_REQUIRE(cap == 0 || (node && src_ppr && tgt_ppr), "NULL pair arrays");
    cudaStream_t st = (cudaStream_t)stream;
    SelectParams2 p{links, bs, adj_rowptr, adj_col, ppr_rowptr, ppr_col, ppr_val, th_cn, th_1hop, th_non1hop,
                    mode, counts, nullptr, node, src_ppr, tgt_ppr, nullptr, (int32_t*)workspace, cap, header,
                    seg_start, nz_list, g_select_dbg, (int32_t*)workspace + bs + 4, g_pk_slot_limit};
    // (the attribute is per device: set on every call — a process may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(select_screen_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PkSmem));
    if (e != cudaSuccess) {
        set_error("lpf_select_onepass_packed: cudaFuncSetAttribute(%zu B): %s", sizeof(PkSmem), cudaGetErrorString(e));
        return LPF_ERR_CUDA;
    }
    e = cudaFuncSetAttribute(select_resolve_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPkBigSmem);
    if (e != cudaSuccess) {
        set_error("lpf_select_onepass_packed: cudaFuncSetAttribute(%d B): %s", kPkBigSmem, cudaGetErrorString(e));
        return LPF_ERR_CUDA;
    }
    launch_onepass_reset(p, st);
    const bool timing = g_kernel_timing && bs > 0;
    if (timing && !g_pk_ev_ready) {
        for (auto& ev : g_pk_ev) cudaEventCreate(&ev);
        g_pk_ev_ready = true;
    }
    if (timing) cudaEventRecord(g_pk_ev[0], st);
    if (bs > 0) {
        // SCREEN: one CTA per piece of kPkThreads links (the hardware hands them to the SMs as slots free up)
        const int64_t blocks = (bs + kPkThreads - 1) / kPkThreads;
        select_screen_packed_kernel<<<(unsigned)blocks, kPkThreads, sizeof(PkSmem), st>>>(
            p, static_cast<const uint4*>(slab), static_cast<const uint4*>(overflow));
        if (timing) cudaEventRecord(g_pk_ev[1], st);
        // RESOLVE: one warp per candidate (their number is on the device: a resident grid strides over the list)
        select_resolve_packed_kernel<<<kNumSMs * kPkResolveCtas, kPkResolveThreads, 0, st>>>(
            p, static_cast<const uint4*>(slab), static_cast<const uint4*>(overflow));
        if (timing) {
            // (profiling: the kernels one after the other, each between two events)
            cudaEventRecord(g_pk_ev[2], st);
            select_resolve_big_kernel<<<kNumSMs * 2, kPkBigThreads, kPkBigSmem, st>>>(
                p, static_cast<const uint4*>(slab), static_cast<const uint4*>(overflow));
            cudaEventRecord(g_pk_ev[3], st);
            launch_onepass_tail(p, st);
            cudaEventRecord(g_pk_ev[4], st);
            g_pk_ev_valid = true;
            return check_launch("lpf_select_onepass_packed");
        }
        // The hub-hub resolve and the deferred-link kernel both take lists that are final once the resolve kernel has
        // run, and both are a latency chain on a few CTAs: they run side by side (fork / join through a side stream
        // per device; under stream capture the fork and the join become graph edges).
        static cudaStream_t side[64] = {};
        static cudaEvent_t ev_fork[64] = {}, ev_join[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) dev = 0;
        if (!side[dev]) {
            cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming);
        }
        cudaEventRecord(ev_fork[dev], st);
        cudaStreamWaitEvent(side[dev], ev_fork[dev], 0);
        select_resolve_big_kernel<<<kNumSMs * 2, kPkBigThreads, kPkBigSmem, side[dev]>>>(
            p, static_cast<const uint4*>(slab), static_cast<const uint4*>(overflow));
        cudaEventRecord(ev_join[dev], side[dev]);
        launch_onepass_heavy(p, st);
        cudaStreamWaitEvent(st, ev_join[dev], 0);
        launch_onepass_finalize(p, st);
        return check_launch("lpf_select_onepass_packed");
    }
    launch_onepass_tail(p, st);
    return check_launch("lpf_select_onepass_packed");
}

// Test hook of the earlier hash-table kernels (capped their tables); the Bloom-filter screening has no capacity
// limit, so the value is accepted and ignored.
extern "C" int lpf_debug_select_slots(int limit) {
    lpf::g_pk_slot_limit = limit > 0 ? limit : 0;
    return LPF_OK;
}

// Profiling hook: with enable != 0 later lpf_select_onepass_packed calls record CUDA events on their stream around
// the screening kernel, the resolve kernel, the hub-hub resolve kernel and the deferred-link tail;
// lpf_debug_select_timing_read waits for the last such call and returns the four durations in milliseconds (0 on
// success, -1 if nothing was recorded).
extern "C" int lpf_debug_select_timing(int enable) {
    lpf::g_kernel_timing = enable != 0;
    if (!enable) lpf::g_pk_ev_valid = false;
    return LPF_OK;
}
extern "C" int lpf_debug_select_timing_read(float* ms4_host) {
    if (!lpf::g_pk_ev_valid || !ms4_host) return -1;
    if (cudaEventSynchronize(lpf::g_pk_ev[4]) != cudaSuccess) return -1;
    for (int k = 0; k < 4; ++k)
        if (cudaEventElapsedTime(ms4_host + k, lpf::g_pk_ev[k], lpf::g_pk_ev[k + 1]) != cudaSuccess) return -1;
    return LPF_OK;
}
