// Device code of the packed-row selection kernel (select_packed.cu): the shared-memory hash set of a source
// adjacency row, the staged source PPR row, and the CSR fallback for one link.
#pragma once
#include "select_walk.cuh"

namespace lpf {

constexpr int kPprHashSlots = 256;   // position hash of a staged PPR row (<= 128 entries)

__device__ __forceinline__ uint32_t hash_slot(int32_t u, int shift) { return ((uint32_t)u * 0x9E3779B1u) >> shift; }

// Hash set of node ids in shared memory: open addressing over BUCKETS of four int32 slots, so that a probe is one
// 16-byte shared-memory read and four compares.  Slots of a bucket are filled in order and never emptied, hence
// "slot 3 is empty" = "no key ever overflowed out of this bucket": a miss ends at the first bucket that is not full.
// `mask` is the bucket mask, `shift` = 32 - log2(buckets); empty slots hold -1 (node ids are >= 0, row pads -2).
__device__ __forceinline__ void hash_insert(int32_t* tab, uint32_t mask, int shift, int32_t u) {
    uint32_t b = hash_slot(u, shift);
    while (true) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (atomicCAS(&tab[4 * b + j], -1, u) == -1) return;
        b = (b + 1) & mask;
    }
}
__device__ __forceinline__ bool bucket_has(const int4& q, int32_t u) {
    return (q.x == u) | (q.y == u) | (q.z == u) | (q.w == u);
}
__device__ __forceinline__ bool hash_contains_from(const int32_t* tab, uint32_t mask, uint32_t b, int32_t u) {
    while (true) {
        const int4 q = *reinterpret_cast<const int4*>(tab + 4 * b);
        if (bucket_has(q, u)) return true;
        if (q.w < 0) return false;
        b = (b + 1) & mask;
    }
}
__device__ __forceinline__ bool hash_contains(const int32_t* tab, uint32_t mask, int shift, int32_t u) {
    return hash_contains_from(tab, mask, hash_slot(u, shift), u);
}
// either of two ids in the set?  The two home buckets are read back to back (one shared-memory latency).
__device__ __forceinline__ bool hash_contains_any2(const int32_t* tab, uint32_t mask, int shift, int32_t u0, int32_t u1) {
    const uint32_t b0 = hash_slot(u0, shift), b1 = hash_slot(u1, shift);
    const int4 q0 = *reinterpret_cast<const int4*>(tab + 4 * b0), q1 = *reinterpret_cast<const int4*>(tab + 4 * b1);
    bool hit = bucket_has(q0, u0) | bucket_has(q1, u1);
    if (!hit && (q0.w >= 0 || q1.w >= 0)) {
        // a full home bucket: follow the overflow chain of that id (rare at load <= 0.5)
        if (q0.w >= 0) hit |= hash_contains_from(tab, mask, (b0 + 1) & mask, u0);
        if (q1.w >= 0) hit |= hash_contains_from(tab, mask, (b1 + 1) & mask, u1);
    }
    return hit;
}

struct RunCtx {
    const int32_t* tab;      // hash set of A(a): buckets of four slots
    uint32_t mask;           // bucket mask
    int shift;               // 32 - log2(buckets)
    const int32_t* pac;      // smem copy of P(a): cols, values, and a hash col -> position
    const float* pav;
    const int32_t* ppos;     // kPprHashSlots entries: position in pac or -1
    int npa;
};

// (present, q) of node u in the shared-memory copy of the source's PPR row
__device__ __forceinline__ bool smem_ppr_lookup(const RunCtx& h, int32_t u, float& q) {
    uint32_t s = hash_slot(u, 32 - 8);
    while (true) {
        const int32_t pos = h.ppos[s];
        if (pos < 0) {
            q = 0.f;
            return false;
        }
        if (h.pac[pos] == u) {
            q = quantise(h.pav[pos]);
            return true;
        }
        s = (s + 1) & (kPprHashSlots - 1);
    }
}

// count -> allocate -> write of one link by a group of G lanes over the CSR tables (the generic walk of
// select_walk.cuh: the shorter row is walked, the longer searched) — the fallback of the packed kernel for links
// whose source is not staged in shared memory.
template <int G>
__device__ __forceinline__ void onepass_link(const SelectParams2& p, const LinkRows& r, int64_t i, int lane) {
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

}  // namespace lpf
