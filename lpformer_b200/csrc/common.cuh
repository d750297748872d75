// Shared helpers for the lpformer_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lpformer_b200.h"

namespace lpf {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSMs = 148;  // B200

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return LPF_ERR_CUDA;
    }
    return LPF_OK;
}

#define LPF_REQUIRE(cond, msg)                                  \
    do {                                                        \
        if (!(cond)) {                                          \
            lpf::set_error("%s: %s", __func__, msg);            \
            return LPF_ERR_INVALID;                             \
        }                                                       \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// q(p) = fl32(p + 1) - 1: the fp32 round trip the reference applies to every PPR value
// before thresholding (models/link_transformer.py:290-291,316-317,464-476).  Must not be
// contracted or reassociated, hence the explicit round-to-nearest intrinsics.
__device__ __forceinline__ float quantise(float p) { return __fsub_rn(__fadd_rn(p, 1.0f), 1.0f); }

// lower_bound over a sorted int32 row in global memory.
__device__ __forceinline__ int lower_bound(const int32_t* __restrict__ a, int n, int32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace lpf
