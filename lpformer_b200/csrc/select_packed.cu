// K1, one-pass selection over PACKED LINK ROWS: the HBM layout built for the per-link walk.
//
// The CSR tables cost a link four dependent random reads per endpoint (two rowptr pairs, then the adjacency row,
// the PPR columns and — on a match — the PPR values, all in different arrays).  lpf_pack_link_rows rewrites them
// once per graph as
//   node_desc[x] = (first 16-byte chunk of row x, deg x, nP x, 0)                      one 16-byte read
//   row x        = ceil(deg/4) chunks of adjacency ids (pad -2), then ceil(nP/2) chunks of (PPR col, PPR value)
//                  pairs (pad col -2), contiguous                                        one contiguous region
// so a link's target is TWO dependent reads (descriptor, then its row region), and the PPR value arrives with
// its column.
//
// Work shape (citation2-style evaluation: runs of links sharing their source, reference train/testing.py:20-23):
// a CTA takes 512 consecutive links = at most kPkMaxRuns runs of equal source.  The sources' adjacency rows become
// open-addressing hash sets in shared memory and their PPR rows small shared-memory tables (all runs of the chunk
// side by side, so every thread works at the same time), then ONE THREAD OWNS ONE LINK: it streams its target's
// row region with 16-byte loads and probes shared memory.  All 2,048 x 148 links in flight are independent, which
// is what hides the DRAM latency.  Links that select something (about 1 % of a citation2-shaped batch) are queued
// and written by a compacted second walk; target rows too long for one thread are queued and walked by a warp
// each; sources that do not fit the hash, short runs and chunks that are not run-shaped take the generic group
// walk of select_walk.cuh.  Selected sets, their order inside a link and the fp32 values are those of every other
// K1 variant (tests compare all of them with the oracle).
#include <stdlib.h>

#include "select_hashed.cuh"

namespace lpf {

constexpr int kPkThreads = 512;       // threads per CTA = links per chunk
constexpr int kPkMaxRuns = 3;
constexpr int kPkHashSlots = 8192;    // int32 slots shared by the runs of a chunk (load <= 0.5): sources up to 4,096
constexpr int kPkHubSlots = 32768;    // second launch, one CTA per SM: hub sources up to 16,384 neighbours
constexpr int kPkMaxPprRow = 128;
constexpr int kPkLaneAdj = 32;        // target rows of up to 32 adjacency chunks (128 neighbours) ...
constexpr int kPkLanePpr = 32;        // ... and 32 PPR chunks (64 entries) are screened by the link's own thread
constexpr int kPkWarpRow = 512;       // resolution: target rows up to this length take the hashed warp walk
constexpr int kPkHubPiece = 128;      // links per entry of the hub list
constexpr int kPkBins = 8;            // row-length bins of the in-chunk ordering
constexpr int kPkPrefetchLines = 32;  // 128-byte lines of a target row prefetched into L2 by its thread

struct PkRunTab {
    int32_t pac[kPkMaxPprRow];
    float pav[kPkMaxPprRow];
    int32_t ppos[kPprHashSlots];
};

template <int SLOTS>
struct PkSmemT {
    int32_t tab[SLOTS];
    PkRunTab run[kPkMaxRuns];
    int32_t q_slow[kPkThreads];          // chunk positions of the links that get a warp (long row / selects something)
    int16_t order[kPkThreads];           // chunk positions ordered by row length
    uint32_t l_off[kPkThreads];          // per link of the chunk: first chunk of the target's packed row,
    uint16_t l_ca[kPkThreads], l_nc[kPkThreads];   // its adjacency chunks and the chunks the screening walks
    int32_t bin_cnt[kPkBins];
    int32_t run_start[kPkMaxRuns + 1];
    int32_t r_tab0[kPkMaxRuns], r_lg[kPkMaxRuns], r_na[kPkMaxRuns], r_npa[kPkMaxRuns], r_hashed[kPkMaxRuns];
    uint32_t r_off[kPkMaxRuns];
    int n_runs, n_slow, tab_used;
};

__device__ __forceinline__ uint4 ldg16(const uint4* p) { return __ldg(p); }

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

// Membership of u in the source's adjacency row: the shared-memory hash set when the row was staged, else a
// binary search over the source's packed row in global memory (hub sources: the row stays L1/L2-resident because
// every link of the run probes it).
template <bool HASHED>
__device__ __forceinline__ bool src_contains(const RunCtx& h, const int32_t* __restrict__ arow, int na, int32_t u) {
    if (HASHED) return hash_contains(h.tab, h.mask, h.shift, u);
    const int t = lower_bound_from(arow, 0, na, u);
    return t < na && __ldg(arow + t) == u;
}

// One thread screens one link: does its packed target row select ANYTHING against the staged source?  (A common
// neighbour, or a node in both PPR rows above the smaller PPR threshold — a superset test for the 1-hop / >1-hop
// sets, exact for "nothing selected".)  Three 16-byte loads stay in flight; the rows were prefetched into L2.
template <bool HASHED>
__device__ __forceinline__ bool screen_packed(const RunCtx& h, const int32_t* __restrict__ arow, int na,
                                              const uint4* __restrict__ row, int ca, int nc, float th_pre) {
    if (nc == 0) return false;
    const uint4 z = make_uint4(0xfffffffeu, 0, 0xfffffffeu, 0);
    uint4 v0 = ldg16(row);
    uint4 v1 = nc > 1 ? ldg16(row + 1) : z;
    bool any = false;
    int c = 0;
    for (; c < ca; ++c) {
        const uint4 v2 = (c + 2 < nc) ? ldg16(row + c + 2) : z;
        if (HASHED) {
            any |= hash_contains_any2(h.tab, h.mask, h.shift, (int32_t)v0.x, (int32_t)v0.y);
            any |= hash_contains_any2(h.tab, h.mask, h.shift, (int32_t)v0.z, (int32_t)v0.w);
        } else {
            any |= src_contains<false>(h, arow, na, (int32_t)v0.x);
            any |= src_contains<false>(h, arow, na, (int32_t)v0.y);
            any |= src_contains<false>(h, arow, na, (int32_t)v0.z);
            any |= src_contains<false>(h, arow, na, (int32_t)v0.w);
        }
        v0 = v1;
        v1 = v2;
    }
    for (; c < nc; ++c) {
        const uint4 v2 = (c + 2 < nc) ? ldg16(row + c + 2) : z;
        float qa;
        if (smem_ppr_lookup(h, (int32_t)v0.x, qa)) any |= qa >= th_pre && quantise(__uint_as_float(v0.y)) >= th_pre;
        if (smem_ppr_lookup(h, (int32_t)v0.z, qa)) any |= qa >= th_pre && quantise(__uint_as_float(v0.w)) >= th_pre;
        v0 = v1;
        v1 = v2;
    }
    return any;
}

// Rare, register-hungry paths kept out of line so that they do not set the register budget of the screening walk.
__device__ __noinline__ void generic_link8(const SelectParams2& p, int64_t i, int lane) {
    const LinkRows r = load_rows(p, i);
    if (is_heavy(r, p.mode != LPF_MODE_CN, 8)) {
        if ((lane & 7) == 0) p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
        return;
    }
    onepass_link<8>(p, nullptr, r, i, lane);
}
// A link that needs resolving (it selects something, or its target row is long), by one warp: the hashed walk of
// the target row when that row is short enough, else the generic walk (shorter row against the longer), else —
// both rows long — the CTA-wide kernel that runs afterwards.
template <class SM>
__device__ __noinline__ void resolve_link32(const SelectParams2& p, const SM& sm, int r, int64_t i, int lane) {
    const LinkRows rows = load_rows(p, i);
    const bool want_pi = p.mode != LPF_MODE_CN;
    if (sm.r_hashed[r] == 1 && rows.nb <= kPkWarpRow && rows.npb <= kPkWarpRow) {
        const RunCtx h = make_ctx(sm, r);
        onepass_link<32>(p, &h, rows, i, lane);
    } else if (!is_heavy(rows, want_pi, 8)) {
        onepass_link<32>(p, nullptr, rows, i, lane);
    } else if (lane == 0) {
        p.heavy[4 + atomicAdd(p.heavy, 1)] = (int32_t)i;
    }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
// one instruction brings a whole contiguous row region (16-byte aligned, a multiple of 16 bytes) into L2
__device__ __forceinline__ void prefetch_l2_bulk(const void* ptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

template <class SM>
__device__ __forceinline__ int run_of(const SM& sm, int t) {
    return (t >= sm.run_start[1] ? 1 : 0) + (t >= sm.run_start[2] ? 1 : 0);
}

// HUB = false: the chunks are the batch cut into pieces of 512 links; runs whose source row does not fit the hash
// are appended to the hub list.  HUB = true (second launch, one CTA per SM, a 32,768-slot table): the chunks are the
// entries of that list.
template <int SLOTS, bool HUB>
__global__ void __launch_bounds__(kPkThreads, HUB ? 1 : 3)
select_onepass_packed_kernel(SelectParams2 p, const int4* __restrict__ desc, const uint4* __restrict__ blob, int variant) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    using SM = PkSmemT<SLOTS>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = tid >> 3;                    // 64 groups of 8 lanes (generic fallback)
    const bool want_pi = p.mode != LPF_MODE_CN;
    const bool want_n1 = p.mode == LPF_MODE_ALL;
    const float th_pre = want_n1 ? fminf(p.th_1hop, p.th_non1hop) : p.th_1hop;
    const int64_t nchunks = HUB ? (int64_t)p.hub[0] : (p.bs + kPkThreads - 1) / kPkThreads;

    long long t_mark = clock64();
#define LPF_PHASE(k)                                                              \
    do {                                                                          \
        if (p.dbg && tid == 0) {                                                  \
            const long long now = clock64();                                      \
            atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (k)), (unsigned long long)(now - t_mark)); \
            t_mark = now;                                                         \
        }                                                                         \
    } while (0)

    const long long t_cta = t_mark;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const long long t_chunk = clock64();
        const int64_t i0 = HUB ? (int64_t)p.hub[1 + 2 * chunk] : chunk * kPkThreads;
        const int len = HUB ? p.hub[2 + 2 * chunk] : (int)min((int64_t)kPkThreads, p.bs - i0);
        const int64_t i = i0 + tid;
        if (p.dbg && tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 5), 1ull);
        // ---- this thread's link: its target descriptor is fetched now and its row region prefetched into L2, so
        // that both reads overlap the staging of the sources
        const int64_t a_me = (tid < len) ? __ldg(p.links + i) : -1;
        const int64_t a_prev = (tid > 0 && tid < len) ? __ldg(p.links + i - 1) : -2;
        int4 db = make_int4(0, 0, 0, 0), da = make_int4(0, 0, 0, 0);
        if (tid < len) {
            db = __ldg(desc + __ldg(p.links + p.bs + i));
            da = __ldg(desc + a_me);             // one address per run: a broadcast read
            const int nc = ((db.y + 3) >> 2) + (want_pi ? ((db.z + 1) >> 1) : 0);
            sm.l_off[tid] = (uint32_t)db.x;
            sm.l_ca[tid] = (uint16_t)min((db.y + 3) >> 2, 65535);
            sm.l_nc[tid] = (uint16_t)min(nc, 65535);
            const uint8_t* r0 = reinterpret_cast<const uint8_t*>(blob + (uint32_t)db.x);
            if (variant == 1) {
                const uint8_t* r1 = r0 + 16 * nc;
                const uint8_t* line = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(r0) & ~(uintptr_t)127);
                for (int k = 0; k < kPkPrefetchLines && line < r1; ++k, line += 128) prefetch_l2(line);
            } else if (variant == 2) {
                if (nc > 0) prefetch_l2_bulk(r0, 16u * (uint32_t)min(nc, 8 * kPkPrefetchLines));
            }
            if (!HUB) {
                // every link starts as "nothing selected"; the links that are resolved later overwrite their entries
                p.counts[i] = 0; p.counts[p.bs + i] = 0; p.counts[2 * p.bs + i] = 0;
                p.seg_start[i] = 0; p.seg_start[p.bs + i] = 0; p.seg_start[2 * p.bs + i] = 0;
            }
        }
        const bool boundary = tid < len && (tid == 0 || a_me != a_prev);
        if (tid == 0) sm.n_runs = 0;
        if (tid < kPkBins) sm.bin_cnt[tid] = 0;
        __syncthreads();
        if (boundary) {
            const int k = atomicAdd(&sm.n_runs, 1);
            if (k < kPkMaxRuns) sm.run_start[k] = tid;
            // the source's row: its first lines towards L2 while the run structure is sorted out
            const uint8_t* r0 = reinterpret_cast<const uint8_t*>(blob + (uint32_t)da.x);
            const uint8_t* r1 = r0 + 16 * (((da.y + 3) >> 2) + ((da.z + 1) >> 1));
            for (int kk = 0; kk < kPkPrefetchLines && r0 < r1; ++kk, r0 += 128) prefetch_l2(r0);
        }
        __syncthreads();
        const int n_runs = sm.n_runs;
        if (n_runs > kPkMaxRuns) {
            // not run-shaped: generic one-pass walk, 8 lanes per link
            for (int t = group; t < len; t += kPkThreads / 8) generic_link8(p, i0 + t, lane);
            __syncthreads();
            LPF_PHASE(4);
            continue;
        }
        if (tid == 0) {      // sort the (at most kPkMaxRuns) boundaries, close the list
            for (int x = 1; x < n_runs; ++x)
                for (int y = x; y > 0 && sm.run_start[y] < sm.run_start[y - 1]; --y) {
                    const int tmp = sm.run_start[y]; sm.run_start[y] = sm.run_start[y - 1]; sm.run_start[y - 1] = tmp;
                }
            for (int x = n_runs; x <= kPkMaxRuns; ++x) sm.run_start[x] = len;
            sm.n_slow = 0;
        }
        __syncthreads();
        if (boundary) {
            const int r = run_of(sm, tid);
            sm.r_off[r] = (uint32_t)da.x;
            sm.r_na[r] = da.y;
            sm.r_npa[r] = da.z;
        }
        __syncthreads();
        if (tid == 0) {      // hash regions: power of two >= 2*na (>= 64) per run, side by side
            int used = 0;
            for (int r = 0; r < n_runs; ++r) {
                const int na = sm.r_na[r];
                // slots: a power of two >= 4*na (load <= 0.25: almost every probe ends in its home bucket) while
                // the table has room, never less than 2*na
                int lg = 6;
                while ((1 << lg) < 2 * na && lg < 30) ++lg;
                if ((1 << lg) < 4 * na && used + (2 << lg) <= SLOTS / 2) ++lg;
                // r_hashed: 1 = adjacency row hashed in shared memory; 2 = searched in global memory (a source beyond
                // even the hub table); 3 = handed to the hub launch; 0 = no screening (source PPR row too long for the
                // shared table): generic walk
                const bool fits = used + (1 << lg) <= SLOTS;
                int m = sm.r_npa[r] > kPkMaxPprRow ? 0 : (fits ? 1 : 2);
                if (!HUB && m == 2 && (1 << lg) <= kPkHubSlots) {
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
                if (m == 1) used += 1 << lg;
            }
            sm.tab_used = used;
        }
        __syncthreads();
        {
            const int used = sm.tab_used;
            for (int s = tid; s < used; s += kPkThreads) sm.tab[s] = -1;
            for (int s = tid; s < n_runs * kPprHashSlots; s += kPkThreads)
                sm.run[s / kPprHashSlots].ppos[s % kPprHashSlots] = -1;
        }
        __syncthreads();
        for (int r = 0; r < n_runs; ++r) {
            const int mode_r = sm.r_hashed[r];
            if (mode_r == 0 || mode_r == 3) continue;
            const int na = sm.r_na[r], npa = sm.r_npa[r], lg = sm.r_lg[r];
            int32_t* tab = sm.tab + sm.r_tab0[r];
            const int32_t* arow = reinterpret_cast<const int32_t*>(blob + sm.r_off[r]);
            if (mode_r == 1)
                for (int s = tid; s < na; s += kPkThreads)
                    hash_insert(tab, (1u << (lg - 2)) - 1u, 32 - (lg - 2), __ldg(arow + s));
            const int32_t* prow = arow + 4 * ((na + 3) >> 2);
            for (int s = tid; s < npa; s += kPkThreads) {
                const int32_t u = __ldg(prow + 2 * s);
                sm.run[r].pac[s] = u;
                sm.run[r].pav[s] = __int_as_float(__ldg(prow + 2 * s + 1));
                uint32_t slot = hash_slot(u, 32 - 8);
                while (atomicCAS(&sm.run[r].ppos[slot], -1, s) != -1) slot = (slot + 1) & (kPprHashSlots - 1);
            }
        }
        // ---- order the chunk's links by row length (descending, 8 bins) so that the lanes of a warp finish together
        const int my_run = run_of(sm, tid);
        const int ca_me = (db.y + 3) >> 2, cp_me = (db.z + 1) >> 1;
        const int nc_me = ca_me + (want_pi ? cp_me : 0);
        const bool screened = tid < len && (sm.r_hashed[my_run] == 1 || sm.r_hashed[my_run] == 2);
        const bool too_long = ca_me > kPkLaneAdj || cp_me > kPkLanePpr;
        int bin = 0, rank_in_bin = 0;
        if (screened) {
            if (too_long) {
                sm.q_slow[atomicAdd(&sm.n_slow, 1)] = tid;
            } else {
                bin = nc_me <= 1 ? 0 : min(kPkBins - 1, 32 - __clz(nc_me - 1));
                rank_in_bin = atomicAdd(&sm.bin_cnt[bin], 1);
            }
        }
        __syncthreads();
        if (screened && !too_long) {
            int base = 0;
#pragma unroll
            for (int k = kPkBins - 1; k > 0; --k) base += (k > bin) ? sm.bin_cnt[k] : 0;
            sm.order[base + rank_in_bin] = (int16_t)tid;
        }
        int n_sorted = 0;
#pragma unroll
        for (int k = 0; k < kPkBins; ++k) n_sorted += sm.bin_cnt[k];
        __syncthreads();
        LPF_PHASE(0);

        // ---- phase A: one thread, one link.  A link that selects nothing (99 % of a citation2-shaped batch) is
        // finished here; the others are queued for a warp each.
        if (tid < n_sorted) {
            const int t = sm.order[tid];
            const int r = run_of(sm, t);
            const RunCtx h = make_ctx(sm, r);
            const int32_t* arow = reinterpret_cast<const int32_t*>(blob + sm.r_off[r]);
            const uint4* row = blob + sm.l_off[t];
            const int ca = sm.l_ca[t], nc = sm.l_nc[t];
            const bool any = sm.r_hashed[r] == 1 ? screen_packed<true>(h, arow, sm.r_na[r], row, ca, nc, th_pre)
                                                 : screen_packed<false>(h, arow, sm.r_na[r], row, ca, nc, th_pre);
            if (any) sm.q_slow[atomicAdd(&sm.n_slow, 1)] = t;
        }
        __syncthreads();
        LPF_PHASE(1);

        // ---- runs that cannot be screened (source PPR row beyond the shared table): generic group walk
        for (int r = 0; r < n_runs; ++r) {
            if (sm.r_hashed[r] != 0) continue;
            for (int t = sm.run_start[r] + group; t < sm.run_start[r + 1]; t += kPkThreads / 8)
                generic_link8(p, i0 + t, lane);
        }
        // ---- phase B: the queued links, one warp per link (count -> allocate -> write)
        const int ns = sm.n_slow;
        for (int q = warp; q < ns; q += kPkThreads / 32) {
            const int t = sm.q_slow[q];
            resolve_link32(p, sm, run_of(sm, t), i0 + t, lane);
        }
        if (p.dbg && tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 6), (unsigned long long)ns);
        LPF_PHASE(2);
        __syncthreads();     // the shared tables are rebuilt for the next chunk
        if (p.dbg && tid == 0) {
            atomicMax(p.dbg + 8, clock64() - t_chunk);
            atomicMax(p.dbg + 9, clock64() - t_cta);
        }
    }
}
#undef LPF_PHASE

// ---------------------------------------------------------------------------------------------------------
// Building the packed rows (once per graph).
// ---------------------------------------------------------------------------------------------------------
__global__ void pack_count_kernel(const int64_t* __restrict__ arp, const int64_t* __restrict__ prp, int64_t n,
                                  int32_t* __restrict__ chunks) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const int64_t deg = arp[x + 1] - arp[x], npp = prp[x + 1] - prp[x];
    chunks[x] = (int32_t)(((deg + 3) >> 2) + ((npp + 1) >> 1));
}

__global__ void __launch_bounds__(256) pack_fill_kernel(const int64_t* __restrict__ arp, const int32_t* __restrict__ ac,
                                                        const int64_t* __restrict__ prp, const int32_t* __restrict__ pc,
                                                        const float* __restrict__ pv, int64_t n,
                                                        const int64_t* __restrict__ off, int4* __restrict__ desc,
                                                        uint4* __restrict__ blob) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t x = warp0; x < n; x += nwarps) {
        const int64_t a0 = arp[x], p0 = prp[x];
        const int deg = (int)(arp[x + 1] - a0), npp = (int)(prp[x + 1] - p0);
        const int64_t o = off[x];
        if (lane == 0) desc[x] = make_int4((int)(uint32_t)o, deg, npp, 0);
        int32_t* w = reinterpret_cast<int32_t*>(blob + o);
        const int ca4 = ((deg + 3) >> 2) * 4;
        for (int k = lane; k < ca4; k += 32) w[k] = k < deg ? ac[a0 + k] : -2;
        const int cp2 = ((npp + 1) >> 1) * 2;
        for (int k = lane; k < cp2; k += 32) {
            w[ca4 + 2 * k] = k < npp ? pc[p0 + k] : -2;
            w[ca4 + 2 * k + 1] = k < npp ? __float_as_int(pv[p0 + k]) : 0;
        }
    }
}

extern long long* g_select_dbg;
void launch_onepass_reset(const SelectParams2& p, cudaStream_t st);
void launch_onepass_tail(const SelectParams2& p, cudaStream_t st);

}  // namespace lpf

using namespace lpf;

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int64_t lpf_link_rows_bytes(int64_t n, int64_t adj_nnz, int64_t ppr_nnz) {
    if (n < 0 || adj_nnz < 0 || ppr_nnz < 0) return -1;
    // sum ceil(deg/4) <= (adj_nnz + 3n)/4, sum ceil(nP/2) <= (ppr_nnz + n)/2
    const int64_t chunks = (adj_nnz + 3 * n) / 4 + (ppr_nnz + n) / 2 + 1;
    if (chunks >= ((int64_t)1 << 32)) return -1;    // chunk indices are 32-bit
    return chunks * 16;
}

extern "C" int64_t lpf_link_rows_scratch_bytes(int64_t n) {
    if (n < 0) return -1;
    return align_up(n * 4, 16) + align_up((n + 1) * 8, 16) + align_up(lpf_scan_scratch_bytes(n), 16);
}

extern "C" int lpf_pack_link_rows(const int64_t* adj_rowptr, const int32_t* adj_col, const int64_t* ppr_rowptr,
                                  const int32_t* ppr_col, const float* ppr_val, int64_t n, int32_t* node_desc,
                                  void* row_blob, void* scratch, void* stream) {
    LPF_REQUIRE(n >= 0, "negative node count");
    LPF_REQUIRE(adj_rowptr && ppr_rowptr, "rowptr is NULL");
    LPF_REQUIRE(n == 0 || (node_desc && row_blob && scratch), "NULL output");
    LPF_REQUIRE((reinterpret_cast<uintptr_t>(row_blob) & 15) == 0 && (reinterpret_cast<uintptr_t>(node_desc) & 15) == 0,
                "node_desc / row_blob must be 16-byte aligned");
    if (n == 0) return LPF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* s = static_cast<uint8_t*>(scratch);
    int32_t* chunks = reinterpret_cast<int32_t*>(s);
    int64_t* off = reinterpret_cast<int64_t*>(s + align_up(n * 4, 16));
    void* scan_scratch = s + align_up(n * 4, 16) + align_up((n + 1) * 8, 16);
    pack_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(adj_rowptr, ppr_rowptr, n, chunks);
    int rc = lpf_scan_counts(chunks, n, off, scan_scratch, stream);
    if (rc) return rc;
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    pack_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(adj_rowptr, adj_col, ppr_rowptr, ppr_col, ppr_val, n, off,
                                                       reinterpret_cast<int4*>(node_desc),
                                                       static_cast<uint4*>(row_blob));
    return check_launch("lpf_pack_link_rows");
}

extern "C" int lpf_select_onepass_packed(const int64_t* links, int64_t bs, const int64_t* adj_rowptr,
                                         const int32_t* adj_col, const int64_t* ppr_rowptr, const int32_t* ppr_col,
                                         const float* ppr_val, const int32_t* node_desc, const void* row_blob,
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
    LPF_REQUIRE(node_desc && row_blob, "node_desc / row_blob is NULL (lpf_pack_link_rows)");
    LPF_REQUIRE(cap >= 0 && 3 * cap < ((int64_t)1 << 31), "bad pair capacity");
    LPF_REQUIRE(header && workspace, "header/workspace is NULL");
    LPF_REQUIRE(bs == 0 || (counts && seg_start && nz_list), "NULL output");
    LPF_REQUIRE(cap == 0 || (node && src_ppr && tgt_ppr), "NULL pair arrays");
    cudaStream_t st = (cudaStream_t)stream;
    SelectParams2 p{links, bs, adj_rowptr, adj_col, ppr_rowptr, ppr_col, ppr_val, th_cn, th_1hop, th_non1hop,
                    mode, counts, nullptr, node, src_ppr, tgt_ppr, nullptr, (int32_t*)workspace, cap, header,
                    seg_start, nz_list, g_select_dbg, (int32_t*)workspace + bs + 4};
    static const int variant = getenv("LPF_PK_VARIANT") ? atoi(getenv("LPF_PK_VARIANT")) : 2;
    using SmMain = PkSmemT<kPkHashSlots>;
    using SmHub = PkSmemT<kPkHubSlots>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(select_onepass_packed_kernel<kPkHashSlots, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmMain));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(select_onepass_packed_kernel<kPkHubSlots, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmHub));
        if (e != cudaSuccess) {
            set_error("lpf_select_onepass_packed: cudaFuncSetAttribute(%zu / %zu B): %s", sizeof(SmMain), sizeof(SmHub),
                      cudaGetErrorString(e));
            return LPF_ERR_CUDA;
        }
        configured = true;
    }
    launch_onepass_reset(p, st);
    if (bs > 0) {
        int64_t blocks = (bs + kPkThreads - 1) / kPkThreads;
        const int64_t cap_blocks = (int64_t)kNumSMs * 3;
        if (blocks > cap_blocks) blocks = cap_blocks;
        select_onepass_packed_kernel<kPkHashSlots, false><<<(unsigned)blocks, kPkThreads, sizeof(SmMain), st>>>(
            p, reinterpret_cast<const int4*>(node_desc), static_cast<const uint4*>(row_blob), variant);
        select_onepass_packed_kernel<kPkHubSlots, true><<<kNumSMs, kPkThreads, sizeof(SmHub), st>>>(
            p, reinterpret_cast<const int4*>(node_desc), static_cast<const uint4*>(row_blob), variant);
    }
    launch_onepass_tail(p, st);
    return check_launch("lpf_select_onepass_packed");
}
