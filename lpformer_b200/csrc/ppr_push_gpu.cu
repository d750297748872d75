// PPR precompute on the GPU: the Andersen push of the reference's offline tool
// (util/calc_ppr_scores.py:137-192: numba, one source per prange iteration) for every source, emitting the
// (row, col, fp32 value) entries that become the sorted PPR table the selection kernel reads
// (util/calc_ppr_scores.py:221-241, util/read_datasets.py:122-129).
//
// The values have to be those of the reference, so the arithmetic is the reference's, operation for operation:
// float64 residuals, LIFO work list (`q.pop()`), `val = (1 - alpha) * res / out_degree[u]`, the push test
// `r[v] >= alpha * eps * out_degree[v]` with "not already queued", one `p[u] += res` per pop.  What is parallel:
//   * sources are independent: one WARP per source, taken from a global counter;
//   * inside one push the neighbours of u are distinct nodes, so their residual updates are independent; the only
//     order that matters is the order in which they are appended to the work list — ascending neighbour order,
//     kept by taking the neighbours 32 at a time and compacting the appends by ballot in lane order.
// The numba dicts (p, r) become one open-addressing hash table per warp in global memory (key, r, p, flags; the
// touched slots are listed so that only they are cleared); the table never holds more than 1 + sum of deg(u) over the
// pushes <= 1 + 1 / (alpha * eps) keys (each push moves res >= alpha * eps * deg(u) of a total mass <= 1), which is what
// sizes it — as an upper bound: almost every source touches far fewer nodes, so the caller (ppr.py) runs all sources
// with small tables first (they stay L2-resident) and only the sources whose table filled up again with the full
// size.  Entries leave through a warp-aggregated global cursor in arbitrary row order; the caller sorts them by
// (row, col) once (ppr.py).
#include "common.cuh"

namespace lpf {

struct PprParams {
    const int64_t* indptr;
    const int32_t* indices;
    int64_t n;
    int64_t src0, nsrc;
    const int32_t* src_list;    // optional: the sources are src_list[0 .. nsrc) instead of src0 + k
    int32_t* ovf_list;          // optional: sources whose table filled up are listed here (count in status[1])
    double alpha, alpha_eps;
    int32_t slots;              // power of two
    // per-warp scratch, `slots` entries each
    int32_t* keys;              // node id or -1
    double* r;
    double* p;
    int32_t* flags;             // bit 0: has a p entry, bit 1: queued
    int32_t* queue;             // LIFO work list
    int32_t* touched;           // occupied slots, in insertion order
    // output pool
    int32_t* out_row;
    int32_t* out_col;
    float* out_val;
    int64_t cap;
    unsigned long long* cursor;  // [0] entries written, [1] next source (work counter)
    int32_t* status;             // [0] != 0: pool overflow, [1] != 0: a table filled up
};

__device__ __forceinline__ uint32_t ppr_hash(int32_t u, int32_t slots) {
    return ((uint32_t)u * 0x9E3779B1u) & (uint32_t)(slots - 1);
}

// Slot of key u in this warp's table, inserting it (r = 0, no flags) if absent.  Lanes of one warp call this with
// DISTINCT keys concurrently: the claim of an empty slot is an atomicCAS.  Returns -1 when the table is full.
__device__ __forceinline__ int ppr_find_or_insert(const PprParams& P, int32_t* keys, double* r, int32_t* flags,
                                                  int32_t* touched, int* n_touched, int32_t u) {
    uint32_t s = ppr_hash(u, P.slots);
    for (int probes = 0; probes < P.slots; ++probes) {
        const int32_t k = __ldcg(keys + s);         // (L2: the claims are atomics, an L1 copy could be stale)
        if (k == u) return (int)s;
        if (k == -1) {
            const int32_t old = atomicCAS(keys + s, -1, u);
            if (old == -1) {
                __stcg(r + s, 0.0);
                __stcg(flags + s, 0);
                __stcg(touched + atomicAdd(n_touched, 1), (int32_t)s);
                return (int)s;
            }
            if (old == u) return (int)s;
        }
        s = (s + 1) & (uint32_t)(P.slots - 1);
    }
    return -1;
}

__global__ void __launch_bounds__(128) ppr_push_kernel(const __grid_constant__ PprParams P) {
    __shared__ int s_ntouched[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int32_t* keys = P.keys + warp * P.slots;
    double* r = P.r + warp * P.slots;
    double* p = P.p + warp * P.slots;
    int32_t* flags = P.flags + warp * P.slots;
    int32_t* queue = P.queue + warp * P.slots;
    int32_t* touched = P.touched + warp * P.slots;
    int* n_touched = &s_ntouched[wib];
    const unsigned lt = (1u << lane) - 1u;
    const double one_minus_alpha = 1.0 - P.alpha;

    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(P.cursor + 1, 1ull);
        k = __shfl_sync(kFull, k, 0);
        if ((int64_t)k >= P.nsrc) break;
        const int32_t src = P.src_list ? __ldg(P.src_list + k) : (int32_t)(P.src0 + (int64_t)k);
        if (lane == 0) *n_touched = 0;
        __syncwarp();
        // p = {src: 0.0}; r = {src: alpha}; q = [src]
        int qn = 1;
        bool full = false;
        if (lane == 0) {
            const int s = ppr_find_or_insert(P, keys, r, flags, touched, n_touched, src);
            __stcg(r + s, P.alpha);
            __stcg(p + s, 0.0);
            __stcg(flags + s, 1 | 2);
            __stcg(queue, src);
        }
        __syncwarp();
        while (qn > 0 && !full) {
            const int32_t u = __ldcg(queue + qn - 1);   // (same address for every lane)
            --qn;
            // the slot of u: it is in the table (it was queued)
            uint32_t su = ppr_hash(u, P.slots);
            while (__ldcg(keys + su) != u) su = (su + 1) & (uint32_t)(P.slots - 1);
            const double res = __ldcg(r + su);
            __syncwarp();
            if (lane == 0) {
                const int32_t f = __ldcg(flags + su);
                __stcg(p + su, (f & 1) ? __ldcg(p + su) + res : res);
                __stcg(flags + su, (f | 1) & ~2);
                __stcg(r + su, 0.0);
            }
            __syncwarp();
            const int64_t e0 = P.indptr[u], e1 = P.indptr[u + 1];
            const double val = one_minus_alpha * res / (double)(e1 - e0);
            for (int64_t eb = e0; eb < e1; eb += 32) {
                const int64_t e = eb + lane;
                bool want = false;
                int32_t v = -1;
                int sv = -1;
                if (e < e1) {
                    v = __ldg(P.indices + e);
                    sv = ppr_find_or_insert(P, keys, r, flags, touched, n_touched, v);
                    if (sv >= 0) {
                        const double rv = __ldcg(r + sv) + val;
                        __stcg(r + sv, rv);
                        const double deg_v = (double)(P.indptr[v + 1] - P.indptr[v]);
                        want = rv >= P.alpha_eps * deg_v && !(__ldcg(flags + sv) & 2);
                    }
                }
                full = __any_sync(kFull, e < e1 && sv < 0);
                if (full) break;
                const unsigned m = __ballot_sync(kFull, want);
                if (want) {
                    __stcg(queue + qn + __popc(m & lt), v);
                    __stcg(flags + sv, __ldcg(flags + sv) | 2);
                }
                qn += __popc(m);
                __syncwarp();
            }
        }
        // ---- emit the p entries (every node that was popped), clear the touched slots
        const int nt = *n_touched;
        if (full) {
            if (lane == 0) {
                const int idx = atomicAdd(P.status + 1, 1);
                if (P.ovf_list) P.ovf_list[idx] = src;
            }
        } else {
            int mine = 0;
            for (int t = lane; t < nt; t += 32) mine += __ldcg(flags + __ldcg(touched + t)) & 1;
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += x;
            }
            const int total = __shfl_sync(kFull, inc, 31);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.cursor, (unsigned long long)total);
            base = __shfl_sync(kFull, base, 0);
            if ((int64_t)(base + total) > P.cap) {
                if (lane == 0) P.status[0] = 1;
            } else {
                // lane l writes its entries at base + (exclusive prefix): order inside a row is irrelevant (sorted later)
                int64_t w = (int64_t)base + inc - mine;
                for (int t = lane; t < nt; t += 32) {
                    const int s = __ldcg(touched + t);
                    if (__ldcg(flags + s) & 1) {
                        P.out_row[w] = src;
                        P.out_col[w] = __ldcg(keys + s);
                        P.out_val[w] = (float)__ldcg(p + s);
                        ++w;
                    }
                }
            }
        }
        __syncwarp();
        for (int t = lane; t < nt; t += 32) __stcg(keys + __ldcg(touched + t), -1);
        __syncwarp();
    }
}

__global__ void ppr_fill_kernel(int32_t* keys, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = -1;
}

}  // namespace lpf

using namespace lpf;

static inline int64_t ppr_align(int64_t v) { return (v + 255) / 256 * 256; }

extern "C" int32_t lpf_ppr_push_slots(double alpha, double eps) {
    if (!(alpha > 0.0 && alpha < 1.0) || !(eps > 0.0)) return -1;
    const double keys = 1.0 + 1.0 / (alpha * eps);
    if (keys > (double)(1 << 24)) return -1;
    int32_t slots = 64;
    while ((double)slots < 2.0 * keys) slots <<= 1;      // load <= 0.5
    return slots;
}

extern "C" int64_t lpf_ppr_push_scratch_bytes(int32_t slots, int32_t nwarps) {
    if (slots <= 0 || nwarps <= 0) return -1;
    const int64_t e = (int64_t)slots * nwarps;
    return ppr_align(e * 4) * 4 + ppr_align(e * 8) * 2 + 256;
}

extern "C" int lpf_ppr_push(const int64_t* indptr, const int32_t* indices, int64_t n, double alpha, double eps,
                            int64_t src0, int64_t nsrc, const int32_t* src_list, int32_t* ovf_list, int32_t slots,
                            int32_t nwarps, void* scratch,
                            int32_t* out_row, int32_t* out_col, float* out_val, int64_t cap, int64_t* cursor,
                            int32_t* status, void* stream) {
    LPF_REQUIRE(indptr && (indices || n == 0), "NULL graph");
    LPF_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "bad node count");
    LPF_REQUIRE(nsrc >= 0 && (src_list || (src0 >= 0 && src0 + nsrc <= n)), "bad source range");
    LPF_REQUIRE(alpha > 0.0 && alpha < 1.0 && eps > 0.0, "alpha must be in (0, 1) and eps > 0");
    LPF_REQUIRE(slots >= 64 && (slots & (slots - 1)) == 0, "slots must be a power of two >= 64");
    LPF_REQUIRE(nwarps > 0 && nwarps % 4 == 0, "nwarps must be a positive multiple of 4");
    LPF_REQUIRE(scratch && cursor && status, "NULL scratch / cursor / status");
    LPF_REQUIRE(cap >= 0 && (cap == 0 || (out_row && out_col && out_val)), "NULL output pool");
    if (nsrc == 0) return LPF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t e = (int64_t)slots * nwarps;
    uint8_t* s = static_cast<uint8_t*>(scratch);
    PprParams P;
    P.indptr = indptr; P.indices = indices; P.n = n; P.src0 = src0; P.nsrc = nsrc;
    P.src_list = src_list; P.ovf_list = ovf_list;
    P.alpha = alpha; P.alpha_eps = alpha * eps; P.slots = slots;
    P.keys = reinterpret_cast<int32_t*>(s); s += ppr_align(e * 4);
    P.flags = reinterpret_cast<int32_t*>(s); s += ppr_align(e * 4);
    P.queue = reinterpret_cast<int32_t*>(s); s += ppr_align(e * 4);
    P.touched = reinterpret_cast<int32_t*>(s); s += ppr_align(e * 4);
    P.r = reinterpret_cast<double*>(s); s += ppr_align(e * 8);
    P.p = reinterpret_cast<double*>(s);
    P.out_row = out_row; P.out_col = out_col; P.out_val = out_val; P.cap = cap;
    P.cursor = reinterpret_cast<unsigned long long*>(cursor); P.status = status;
    ppr_fill_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(P.keys, e);
    ppr_push_kernel<<<(unsigned)(nwarps / 4), 128, 0, st>>>(P);
    return check_launch("lpf_ppr_push");
}
