// K3 — dense contraction C = epi(A W^T + bias) on the 5th-generation tensor cores (tcgen05.mma, kind::tf32,
// accumulator in TMEM), with fp32-level accuracy by the 3xTF32 split (tc.cuh).
//
// Every nn.Linear of the path whose weight is reused across many rows goes through here: GCNConv.lin and the
// per-node K/V projection (N rows per eval), the RPE contraction (S rows), lin_l and the MLP heads (BS rows)
// (reference modules/layers.py:130-131,208-214; models/other_models.py:61-76,125-138,173-179).
//
// Work shape: one CTA = one 128-row tile of A and the whole (padded) N <= 256; the k-blocks (32 fp32 = one
// 128-byte swizzled row) stream through a 2-stage shared-memory ring:
//   * the weight's pre-split, pre-swizzled image (lpf_pack_weight, built once per weight) arrives by one bulk
//     TMA copy (cp.async.bulk) per k-block, completion on an mbarrier;
//   * the A tile is loaded with coalesced 128-bit global loads, split into hi/lo in registers and stored in
//     the UMMA canonical layout;
//   * thread 0 issues the 12 MMAs of the k-block (4 K-slices x 3 split products) and commits them to the
//     stage's "free" barrier, so loading k-block i+1 overlaps the tensor pipe working on k-block i;
//   * epilogue by four warps of their own: tcgen05.ld of the thread's own row (TMEM lane = row), bias / ReLU /
//     sigmoid, 128-bit stores — from one of TWO accumulators in TMEM, so the read-back and the stores of tile t run
//     under the loads and MMAs of tile t + 1 (full / empty mbarriers per accumulator).
// A persistent grid (a few CTAs per SM where shared memory allows) walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
// Measured on the ogbl-ddi shape (M = 2.3 M pairs, N = K = 256): 8.8 ms per launch with A pieces requested one at a
// time (the compiler kept one destination register: eight serial DRAM round trips per k-block), one CTA per tile
// and the epilogue in line; see DESIGN section 4 for the numbers after each step.
#include <type_traits>

#include "tc.cuh"

namespace lpf {

using namespace tc;

constexpr int kGemmProducers = 128;           // warps 0-3: operand staging and the MMA-issuing thread
constexpr int kGemmThreads = 256;             // warps 4-7: epilogue (TMEM -> registers -> global), one accumulator behind

// Packed weight image: [KB][2 (hi, lo)][NP rows][128 B swizzled];  KB = ceil(K/32), NP = round_up(N,16).
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ W, int64_t ldw, int N, int K,
                                                          int NP, int KB, float* __restrict__ out) {
    const int64_t total = (int64_t)KB * NP * 32;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e % 32);
        const int n = (int)((e / 32) % NP);
        const int kb = (int)(e / (32 * (int64_t)NP));
        const int k = kb * 32 + kk;
        const float w = (n < N && k < K) ? W[(int64_t)n * ldw + k] : 0.f;
        const float hi = tf32_hi(w), lo = w - hi;
        const uint32_t off = swz_chunk_off(n, kk >> 2) + (uint32_t)((kk & 3) << 2);
        char* base = reinterpret_cast<char*>(out) + (size_t)kb * 2 * NP * 128;
        *reinterpret_cast<float*>(base + off) = hi;
        *reinterpret_cast<float*>(base + (size_t)NP * 128 + off) = lo;
    }
}

struct GemmTcParams {
    const float* A;
    int64_t lda;
    const float* Wp;   // packed image
    const float* bias;
    float bias_scale;
    float* C;
    int64_t ldc;
    int64_t M;
    int N, K, NP, KB, epi;
    uint32_t tmem_cols;
    const int64_t* m_dev;   // optional device-side row count (M is then the capacity the grid was sized for)
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads) gemm_tc_kernel(GemmTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_w[2], bar_free[2], bar_full[2], bar_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_bias[256];

    const int tid = threadIdx.x, warp = tid >> 5;
    if (p.m_dev) p.M = min(p.M, *p.m_dev);
    const int64_t ntiles = (p.M + kTileM - 1) / kTileM;
    if ((int64_t)blockIdx.x >= ntiles) return;   // whole CTA, before any barrier / TMEM state exists
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t w_tile = (uint32_t)p.NP * 128;                 // bytes of one W part (hi or lo) per k-block
    const uint32_t stage_bytes = 2 * kATileBytes + 2 * w_tile;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_w[i], 1);
            mbar_init(&bar_free[i], 1);
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], kGemmProducers);
        }
        fence_mbar_init();
    }
    s_bias[tid] = (p.bias && tid < p.N) ? __ldg(p.bias + tid) : 0.f;      // kGemmThreads == 256 >= NP
    __syncwarp();
    if (warp == 0) tmem_alloc(&tmem_slot, 2 * p.tmem_cols);      // two accumulators: tile t + 1 is contracted while tile t is read back
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (tid >= kGemmProducers) {
        // ---------------- epilogue warps: thread e owns row m0 + e (TMEM lane e) of every tile of the CTA ----------------
        const int e = tid - kGemmProducers;
        const bool vec_c = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.ldc % 4 == 0);
        const uint32_t lane_base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        // (the epilogue kind is a compile-time constant of the loop body and the bias comes from shared memory, four columns
        // per read: with a per-element `__ldg(bias + n)` and a per-element branch on the kind this loop was half of the
        // kernel's instructions and the accumulators were handed back late — ncu: producers waiting for bar_empty)
        auto run = [&](auto kind) {
            constexpr int EPI = decltype(kind)::value;
            uint32_t tile_it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
                const int b = tile_it & 1;
                mbar_wait(&bar_full[b], (tile_it >> 1) & 1);
                tc_fence_after();
                const int64_t m = tile * kTileM + e;
                const uint32_t acc = lane_base + (uint32_t)b * p.tmem_cols;
                float* row = p.C + m * p.ldc;
                for (int c0 = 0; c0 < p.NP; c0 += 16) {
                    float v[16];
                    tmem_ld16(acc + (uint32_t)c0, v);    // warp-collective: every lane executes it
                    if (m < p.M) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 sb = *reinterpret_cast<const float4*>(s_bias + c0 + j);
                            v[j] = fmaf(p.bias_scale, sb.x, v[j]);
                            v[j + 1] = fmaf(p.bias_scale, sb.y, v[j + 1]);
                            v[j + 2] = fmaf(p.bias_scale, sb.z, v[j + 2]);
                            v[j + 3] = fmaf(p.bias_scale, sb.w, v[j + 3]);
                        }
                        if constexpr (EPI == LPF_EPI_RELU) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                        } else if constexpr (EPI == LPF_EPI_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
                        }
                        float* dst = row + c0;
                        if (vec_c && c0 + 16 <= p.N) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < p.N) dst[j] = v[j];
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bar_empty[b]);              // the accumulator may be overwritten (by the CTA's tile after next)
            }
        };
        if (p.epi == LPF_EPI_RELU) run(std::integral_constant<int, LPF_EPI_RELU>{});
        else if (p.epi == LPF_EPI_SIGMOID) run(std::integral_constant<int, LPF_EPI_SIGMOID>{});
        else run(std::integral_constant<int, LPF_EPI_NONE>{});
    } else {
        // ---------------- producer warps: A k-blocks (split, UMMA layout), weight k-blocks (bulk copies), MMAs ----------------
        const uint32_t idesc = make_idesc_tf32(kTileM, p.NP);
        const bool vec_a = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && (p.lda % 4 == 0);
        const int chunk = tid & 7;          // 16-byte chunk of the 128-byte row
        const int row_in_pass = tid >> 3;   // 16 rows per pass

        // the thread's share of one A k-block: eight independent 16-byte requests, all in flight together
        float4 areg[8];
        auto load_a = [&](int64_t m0, int kb) {
            const int k0 = kb * 32 + chunk * 4;
#pragma unroll
            for (int pass = 0; pass < 8; ++pass) {
                const int64_t m = m0 + pass * 16 + row_in_pass;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < p.M) {
                    const float* src = p.A + m * p.lda + k0;
                    if (vec_a && k0 + 3 < p.K) {
                        v = __ldg(reinterpret_cast<const float4*>(src));
                    } else {
                        if (k0 + 0 < p.K) v.x = __ldg(src + 0);
                        if (k0 + 1 < p.K) v.y = __ldg(src + 1);
                        if (k0 + 2 < p.K) v.z = __ldg(src + 2);
                        if (k0 + 3 < p.K) v.w = __ldg(src + 3);
                    }
                }
                areg[pass] = v;
            }
        };
        load_a((int64_t)blockIdx.x * kTileM, 0);

        uint32_t it = 0;        // k-blocks issued by this CTA so far (ring position / barrier phases)
        uint32_t tile_it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
            const int64_t m0 = tile * kTileM;
            const int b = tile_it & 1;
            const uint32_t acc = tmem_d + (uint32_t)b * p.tmem_cols;
            for (int kb = 0; kb < p.KB; ++kb, ++it) {
                const int s = it & 1;
                const uint32_t use = it >> 1;
                uint8_t* st = smem + (size_t)s * stage_bytes;
                if (it >= 2) mbar_wait(&bar_free[s], (use - 1) & 1);      // MMAs that read this stage have completed
                if (it == 0 && tid == 0) {          // (every later k-block's weights are requested one k-block ahead, below)
                    mbar_arrive_expect_tx(&bar_w[s], 2 * w_tile);
                    bulk_g2s(st + 2 * kATileBytes, reinterpret_cast<const uint8_t*>(p.Wp) + (size_t)kb * 2 * w_tile,
                             2 * w_tile, &bar_w[s]);
                }
                __syncwarp();
                // A k-block: rows m0..m0+127, columns kb*32 .. +31 — the thread's eight 16-byte pieces are in registers
                // already (requested one k-block ahead); split, store in the UMMA layout, then request the next k-block's
                // so that their DRAM round trip runs under this k-block's barrier, MMAs and the next wait
#pragma unroll
                for (int pass = 0; pass < 8; ++pass) {
                    const int r = pass * 16 + row_in_pass;
                    const float4 v = areg[pass];
                    const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    const uint32_t off = swz_chunk_off(r, chunk);
                    *reinterpret_cast<float4*>(st + off) = hi;
                    *reinterpret_cast<float4*>(st + kATileBytes + off) = lo;
                }
                if (kb + 1 < p.KB) load_a(m0, kb + 1);
                else if (tile + gridDim.x < ntiles) load_a((tile + gridDim.x) * kTileM, 0);
                fence_async_smem();
                named_bar_sync(1, kGemmProducers);
                if (tid == 0) {
                    mbar_wait(&bar_w[s], use & 1);
                    // the tile's first MMA overwrites the accumulator: the epilogue warps must have read the tile before last
                    if (kb == 0 && tile_it >= 2) mbar_wait(&bar_empty[b], ((tile_it >> 1) - 1) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(st), a_lo = a_hi + kATileBytes;
                    const uint32_t b_hi = a_hi + 2 * kATileBytes, b_lo = b_hi + w_tile;
                    issue_kblock_3x(acc, a_hi, a_lo, b_hi, b_lo, idesc, kb == 0);
                    umma_commit(&bar_free[s]);
                    if (kb == p.KB - 1) umma_commit(&bar_full[b]);
                    // the next k-block's weights (of this tile, or the first of the CTA's next tile) into the other stage as
                    // soon as the MMAs issued one k-block ago have drained it: a whole k-block of lead for the bulk copy
                    // (issued at the top of its own k-block its 2 NP x 128 bytes arrived ~2 us after the request, every k-block)
                    const bool more = kb + 1 < p.KB || tile + gridDim.x < ntiles;
                    if (more) {
                        const uint32_t nit = it + 1;
                        const int ns = nit & 1;
                        if (nit >= 2) mbar_wait(&bar_free[ns], ((nit >> 1) - 1) & 1);
                        const int nkb = kb + 1 < p.KB ? kb + 1 : 0;
                        mbar_arrive_expect_tx(&bar_w[ns], 2 * w_tile);
                        bulk_g2s(smem + (size_t)ns * stage_bytes + 2 * kATileBytes,
                                 reinterpret_cast<const uint8_t*>(p.Wp) + (size_t)nkb * 2 * w_tile, 2 * w_tile, &bar_w[ns]);
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 2 * p.tmem_cols);
}

}  // namespace lpf

using namespace lpf;

extern "C" int64_t lpf_pack_weight_bytes(int32_t N, int32_t K) {
    if (N < 1 || K < 1) return 0;
    const int64_t KB = (K + 31) / 32, NP = tc::round_up(N, 16);
    return KB * 2 * NP * 128;
}

extern "C" int lpf_pack_weight(const float* W, int64_t ldw, int32_t N, int32_t K, float* packed, void* stream) {
    LPF_REQUIRE(W && packed, "NULL argument");
    LPF_REQUIRE(N >= 1 && K >= 1 && ldw >= K, "bad shape");
    LPF_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed image must be 16-byte aligned");
    const int KB = (K + 31) / 32, NP = tc::round_up(N, 16);
    const int64_t total = (int64_t)KB * NP * 32;
    const unsigned grid = (unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
    pack_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, ldw, N, K, NP, KB, packed);
    return check_launch("lpf_pack_weight");
}

extern "C" int lpf_gemm_tc(const float* A, int64_t lda, const float* Wpacked, const float* bias, float bias_scale,
                           float* C, int64_t ldc, int64_t M, int32_t N, int32_t K, int epilogue, const int64_t* m_dev,
                           void* stream) {
    LPF_REQUIRE(M >= 0 && N >= 1 && K >= 1, "bad shape");
    if (M == 0) return LPF_OK;
    LPF_REQUIRE(A && Wpacked && C, "NULL argument");
    LPF_REQUIRE(lda >= K && ldc >= N, "leading dimension too small");
    LPF_REQUIRE(epilogue >= LPF_EPI_NONE && epilogue <= LPF_EPI_SIGMOID, "bad epilogue");
    LPF_REQUIRE((reinterpret_cast<uintptr_t>(Wpacked) & 15) == 0, "packed image must be 16-byte aligned");
    if (N > 256) {
        set_error("lpf_gemm_tc: N = %d exceeds one UMMA tile (256); split the weight by rows", N);
        return LPF_ERR_UNSUPPORTED;
    }
    GemmTcParams p;
    p.A = A; p.lda = lda; p.Wp = Wpacked; p.bias = bias; p.bias_scale = bias_scale; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.NP = tc::round_up(N, 16); p.KB = (K + 31) / 32; p.epi = epilogue;
    p.tmem_cols = tc::tmem_cols_for(p.NP);
    p.m_dev = m_dev;
    // always two ring stages: the kernel indexes the ring with a k-block counter that keeps running across the tiles
    // of a persistent CTA, so a single-k-block contraction (K <= 32) uses stage 1 for its second tile
    const int stages = 2;
    const size_t smem = (size_t)stages * (2 * tc::kATileBytes + 2 * (size_t)p.NP * 128) + 1024;
    // (the attribute is per device and a process may drive several GPUs: set it on every call, at its maximum)
    {
        const size_t smem_max = (size_t)stages * (2 * tc::kATileBytes + 2 * (size_t)256 * 128) + 1024;
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
        if (e != cudaSuccess) {
            set_error("lpf_gemm_tc: cudaFuncSetAttribute(%zu B): %s", smem_max, cudaGetErrorString(e));
            return LPF_ERR_CUDA;
        }
    }
    // a persistent grid of a few CTAs per SM walks the tiles (the A pieces of a CTA's next tile are requested during
    // its current tile's last k-block); with a device-side row count the host M is only a capacity
    int64_t grid = (M + tc::kTileM - 1) / tc::kTileM;
    const int64_t per_sm = (int64_t)(227 * 1024) / (int64_t)(smem + 1024);
    const int64_t persistent = (int64_t)kNumSMs * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    if (grid > persistent) grid = persistent;
    gemm_tc_kernel<<<(unsigned)grid, kGemmThreads, smem, (cudaStream_t)stream>>>(p);
    return check_launch("lpf_gemm_tc");
}
