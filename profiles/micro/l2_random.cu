// Microbenchmark: effective L2 capacity for random 32-byte-sector reads on this GPU.
// Each thread reads `iters` pseudo-random 8-byte words (one per 32-byte sector) from a buffer of `mb` megabytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_random l2_random.cu ; run: ./l2_random
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(const uint64_t* __restrict__ buf, uint64_t nsec, uint32_t iters, uint64_t* out, uint32_t seed) {
    uint64_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + seed;
    uint64_t acc = 0;
    for (uint32_t i = 0; i < iters; i += 4) {
        uint64_t a[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            a[j] = (x % nsec) * 4;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc += __ldg(buf + a[j]);
    }
    if (acc == 0x1234567) out[0] = acc;
}
int main() {
    const size_t maxb = 1024ull << 20;
    uint64_t *buf, *out;
    cudaMalloc(&buf, maxb); cudaMalloc(&out, 8);
    cudaMemset(buf, 1, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mb : {8, 16, 32, 48, 64, 96, 128, 192, 256, 512, 1024}) {
        uint64_t nsec = (uint64_t)mb * (1 << 20) / 32;
        const int grid = 148 * 8, block = 256; const uint32_t iters = 2048;
        k<<<grid, block>>>(buf, nsec, iters, out, 1);  // warm
        cudaEventRecord(e0);
        k<<<grid, block>>>(buf, nsec, iters, out, 2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double acc = (double)grid * block * iters;
        printf("buffer %4d MB: %.1f G sector reads/s, %.2f TB/s of 32B sectors\n", mb, acc / ms / 1e6, acc * 32 / ms / 1e9);
    }
    return 0;
}
