// Host-side approximate Personalised PageRank (Andersen push), multi-threaded over sources.
//
// Data-preparation tool for the path's input tables: the reference computes the N x N PPR
// matrix offline with a numba kernel (util/calc_ppr_scores.py:137-192) and stores it sorted
// by (row, col) as fp32 (:221-241).  This is the same algorithm — LIFO work list, float64
// residual arithmetic in the same operation order, push test r[v] >= alpha*eps*deg[v] with a
// "not already queued" check — so values are bit-identical to the reference's, emitted
// directly as the sorted CSR (rowptr int64, col int32, val fp32) the selection kernel reads.
// Per-thread dense scratch (p, r, queued flag) with touched-lists replaces the numba dicts.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/lpformer_b200.h"

namespace {

struct PprResult {
    int64_t n = 0;
    std::vector<std::vector<int32_t>> cols;  // per source, ascending
    std::vector<std::vector<float>> vals;
};

void worker(const int64_t* indptr, const int32_t* indices, int64_t n, double alpha, double eps,
            std::atomic<int64_t>* next, PprResult* out) {
    const double alpha_eps = alpha * eps;
    std::vector<double> p(n, 0.0), r(n, 0.0);
    std::vector<uint8_t> in_p(n, 0), in_q(n, 0), in_r(n, 0);
    std::vector<int32_t> q, touched_p, touched_r;
    std::vector<std::pair<int32_t, float>> row;
    constexpr int64_t kChunk = 64;
    for (;;) {
        const int64_t s0 = next->fetch_add(kChunk);
        if (s0 >= n) break;
        const int64_t s1 = std::min(n, s0 + kChunk);
        for (int64_t s = s0; s < s1; ++s) {
            q.clear(); touched_p.clear(); touched_r.clear();
            p[s] = 0.0; in_p[s] = 1; touched_p.push_back((int32_t)s);
            r[s] = alpha; in_r[s] = 1; touched_r.push_back((int32_t)s);
            q.push_back((int32_t)s); in_q[s] = 1;
            while (!q.empty()) {
                const int32_t u = q.back();
                q.pop_back();
                in_q[u] = 0;
                const double res = in_r[u] ? r[u] : 0.0;
                if (in_p[u]) p[u] += res;
                else { p[u] = res; in_p[u] = 1; touched_p.push_back(u); }
                r[u] = 0.0;
                if (!in_r[u]) { in_r[u] = 1; touched_r.push_back(u); }
                const int64_t e0 = indptr[u], e1 = indptr[u + 1];
                const double deg_u = (double)(e1 - e0);
                for (int64_t e = e0; e < e1; ++e) {
                    const int32_t v = indices[e];
                    const double val = (1.0 - alpha) * res / deg_u;
                    if (in_r[v]) r[v] += val;
                    else { r[v] = val; in_r[v] = 1; touched_r.push_back(v); }
                    const double deg_v = (double)(indptr[v + 1] - indptr[v]);
                    if (r[v] >= alpha_eps * deg_v && !in_q[v]) { q.push_back(v); in_q[v] = 1; }
                }
            }
            row.clear();
            for (int32_t u : touched_p) row.emplace_back(u, (float)p[u]);
            std::sort(row.begin(), row.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            auto& oc = out->cols[s];
            auto& ov = out->vals[s];
            oc.resize(row.size());
            ov.resize(row.size());
            for (size_t k = 0; k < row.size(); ++k) { oc[k] = row[k].first; ov[k] = row[k].second; }
            for (int32_t u : touched_p) { in_p[u] = 0; p[u] = 0.0; }
            for (int32_t u : touched_r) { in_r[u] = 0; r[u] = 0.0; }
        }
    }
}

}  // namespace

// Runs the push for every source.  Returns an opaque handle (NULL on bad arguments) and
// writes the total number of entries to *nnz.  indptr/indices: HOST CSR of the (symmetric)
// graph, sorted columns.  nthreads <= 0 -> hardware concurrency.
extern "C" void* lpf_ppr_push_host(const int64_t* indptr_host, const int32_t* indices_host, int64_t n, double alpha,
                                   double eps, int nthreads, int64_t* nnz) {
    if (!indptr_host || n < 0 || !nnz) return nullptr;
    auto* res = new PprResult();
    res->n = n;
    res->cols.resize(n);
    res->vals.resize(n);
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(1, n / 64));
    std::atomic<int64_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, indptr_host, indices_host, n, alpha, eps, &next, res);
    for (auto& t : th) t.join();
    int64_t total = 0;
    for (auto& c : res->cols) total += (int64_t)c.size();
    *nnz = total;
    return res;
}

// Copies the result into caller-owned HOST arrays (rowptr [n+1], col [nnz], val [nnz]) and frees the handle.
extern "C" int lpf_ppr_push_host_fetch(void* handle, int64_t* rowptr_host, int32_t* col_host, float* val_host) {
    if (!handle) return LPF_ERR_INVALID;
    auto* res = static_cast<PprResult*>(handle);
    int64_t off = 0;
    for (int64_t s = 0; s < res->n; ++s) {
        rowptr_host[s] = off;
        const size_t k = res->cols[s].size();
        if (k) {
            std::memcpy(col_host + off, res->cols[s].data(), k * sizeof(int32_t));
            std::memcpy(val_host + off, res->vals[s].data(), k * sizeof(float));
        }
        off += (int64_t)k;
    }
    rowptr_host[res->n] = off;
    delete res;
    return LPF_OK;
}
