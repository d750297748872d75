// Micro-benchmark: throughput of random row gathers (the access pattern of the screening kernel) on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu && ./gather_probe
// Each group of G lanes reads one random ROW of G*16 bytes (64-byte aligned when G=4), K rows in flight per lane.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int G, int K>
__global__ void gather(const uint4* __restrict__ blob, const uint32_t* __restrict__ idx, size_t nrows_total, uint32_t* out) {
    const size_t gid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const size_t ngroups = ((size_t)gridDim.x * blockDim.x) / G;
    uint32_t acc = 0;
    for (size_t r0 = gid; r0 < nrows_total; r0 += K * ngroups) {
        uint4 v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const size_t r = r0 + k * ngroups;
            v[k] = make_uint4(0, 0, 0, 0);
            if (r < nrows_total) v[k] = __ldg(blob + (size_t)__ldg(idx + r) * G + gl);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int G, int K>
float run(const uint4* blob, const uint32_t* idx, size_t n, uint32_t* out, int blocks, int threads) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<G, K><<<blocks, threads>>>(blob, idx, n, out);
    cudaEventRecord(e0);
    gather<G, K><<<blocks, threads>>>(blob, idx, n, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    const size_t nacc = 4u << 20;      // rows gathered per launch
    uint32_t* out; cudaMalloc(&out, 4);
    for (size_t mb : {64, 256, 600, 2048, 8192}) {
        const size_t bytes = mb << 20;
        uint4* blob; cudaMalloc(&blob, bytes); cudaMemset(blob, 1, bytes);
        for (int G : {2, 4, 8, 16}) {
            const size_t rows = bytes / (16 * G);
            uint32_t* h = (uint32_t*)malloc(nacc * 4);
            uint64_t s = 88172645463325252ull;
            for (size_t i = 0; i < nacc; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s % rows); }
            uint32_t* idx; cudaMalloc(&idx, nacc * 4); cudaMemcpy(idx, h, nacc * 4, cudaMemcpyHostToDevice); free(h);
            for (int occ : {1, 2, 4, 8}) {
                const int threads = 256, blocks = 148 * occ;
                float ms1 = 0, ms4 = 0, ms8 = 0;
                if (G == 2) { ms1 = run<2, 1>(blob, idx, nacc, out, blocks, threads); ms4 = run<2, 4>(blob, idx, nacc, out, blocks, threads); ms8 = run<2, 8>(blob, idx, nacc, out, blocks, threads); }
                if (G == 4) { ms1 = run<4, 1>(blob, idx, nacc, out, blocks, threads); ms4 = run<4, 4>(blob, idx, nacc, out, blocks, threads); ms8 = run<4, 8>(blob, idx, nacc, out, blocks, threads); }
                if (G == 8) { ms1 = run<8, 1>(blob, idx, nacc, out, blocks, threads); ms4 = run<8, 4>(blob, idx, nacc, out, blocks, threads); ms8 = run<8, 8>(blob, idx, nacc, out, blocks, threads); }
                if (G == 16) { ms1 = run<16, 1>(blob, idx, nacc, out, blocks, threads); ms4 = run<16, 4>(blob, idx, nacc, out, blocks, threads); ms8 = run<16, 8>(blob, idx, nacc, out, blocks, threads); }
                const double gb = (double)nacc * 16 * G / 1e9;
                printf("footprint %5zu MB row %3d B  ctas/SM %d x256thr: K=1 %7.0f GB/s  K=4 %7.0f GB/s  K=8 %7.0f GB/s   (rows/us K=4: %.0f)\n", mb, 16 * G, occ,
                       gb / (ms1 * 1e-3), gb / (ms4 * 1e-3), gb / (ms8 * 1e-3), nacc / (ms4 * 1e3));
            }
            cudaFree(idx);
        }
        cudaFree(blob);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
