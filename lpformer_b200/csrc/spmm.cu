// K2 — GCN message passing as a vectorised CSR SpMM: Y[r,:] = sum_k val[k] * XW[col[k],:] + bias.
//
// Replaces GCNConv's torch_sparse.matmul (reference models/other_models.py:66; PyG 2.2.0
// gcn_norm semantics are applied when the normalised CSR is built).  One
// warp per row; the warp first loads up to 32 (col, val) pairs with one coalesced load each
// and broadcasts them by shuffle, then every lane accumulates its float4 slice of the
// gathered feature rows (128-bit loads, 4 neighbour rows in flight per lane).
#include "common.cuh"

namespace lpf {

template <int VEC>  // VEC = 4: float4 slices (d % 4 == 0, 16-byte aligned rows); VEC = 1: scalar
__global__ void __launch_bounds__(256) gcn_spmm_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                       const float* __restrict__ val, int64_t row0, int64_t rows,
                                                       const float* __restrict__ XW, int64_t ld_xw,
                                                       const float* __restrict__ bias, int d, float* __restrict__ Y,
                                                       int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int MAXC = 4;  // channel chunks per lane: d <= 32*VEC*MAXC (512 for VEC=4)
    for (int64_t r = warp; r < rows; r += nwarps) {
        const int64_t row = row0 + r;
        const int64_t k0 = __ldg(rowptr + row), k1 = __ldg(rowptr + row + 1);
        float acc[MAXC][VEC];
#pragma unroll
        for (int q = 0; q < MAXC; ++q)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[q][e] = 0.f;
        for (int64_t kb = k0; kb < k1; kb += 32) {
            const int cnt = (int)min((int64_t)32, k1 - kb);
            const int32_t my_c = (lane < cnt) ? __ldg(col + kb + lane) : 0;
            const float my_v = (lane < cnt) ? __ldg(val + kb + lane) : 0.f;
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const int64_t c = __shfl_sync(kFull, my_c, j);
                const float w = __shfl_sync(kFull, my_v, j);
                const float* x = XW + c * ld_xw;
#pragma unroll
                for (int q = 0; q < MAXC; ++q) {
                    const int ch = (lane + 32 * q) * VEC;
                    if (ch < d) {
                        if constexpr (VEC == 4) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(x + ch));
                            acc[q][0] = fmaf(w, t.x, acc[q][0]);
                            acc[q][1] = fmaf(w, t.y, acc[q][1]);
                            acc[q][2] = fmaf(w, t.z, acc[q][2]);
                            acc[q][3] = fmaf(w, t.w, acc[q][3]);
                        } else {
                            acc[q][0] = fmaf(w, __ldg(x + ch), acc[q][0]);
                        }
                    }
                }
            }
        }
        float* y = Y + row * ldy;
#pragma unroll
        for (int q = 0; q < MAXC; ++q) {
            const int ch = (lane + 32 * q) * VEC;
            if (ch < d) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) y[ch + e] = acc[q][e] + (bias ? __ldg(bias + ch + e) : 0.f);
            }
        }
    }
}

}  // namespace lpf

using namespace lpf;

extern "C" int lpf_gcn_spmm(const int64_t* rowptr, const int32_t* col, const float* val, int64_t row0, int64_t rows,
                            const float* XW, int64_t ld_xw, const float* bias, int32_t d, float* Y, int64_t ldy,
                            void* stream) {
    LPF_REQUIRE(rows >= 0 && row0 >= 0, "negative row range");
    if (rows == 0) return LPF_OK;
    LPF_REQUIRE(rowptr && XW && Y, "NULL argument");
    LPF_REQUIRE(d >= 1 && ld_xw >= d && ldy >= d, "bad d / leading dimension");
    const bool vec = (d % 4 == 0) && (ld_xw % 4 == 0) && ((reinterpret_cast<uintptr_t>(XW) & 15) == 0);
    LPF_REQUIRE(vec ? d <= 512 : d <= 128, "d too large (<=512 when d%4==0 and rows are 16B aligned, else <=128)");
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = (int64_t)kNumSMs * 8 * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) gcn_spmm_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, val, row0, rows, XW, ld_xw, bias, d, Y, ldy);
    else gcn_spmm_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, val, row0, rows, XW, ld_xw, bias, d, Y, ldy);
    return check_launch("lpf_gcn_spmm");
}
