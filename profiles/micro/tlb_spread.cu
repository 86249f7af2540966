// Microbenchmark: does spreading an L2-resident working set over many 2 MB pages cost throughput (TLB reach)?
// Random 8-byte reads (one 32-byte sector each) from `pages` regions of `kb` KB each, regions `stride_mb` apart.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(const uint64_t* __restrict__ buf, uint32_t pages, uint32_t sec_per_page, uint64_t stride_words, uint32_t iters,
                  uint64_t* out, uint32_t seed) {
    uint64_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + seed;
    uint64_t acc = 0;
    for (uint32_t i = 0; i < iters; i += 4) {
        uint64_t a[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            uint32_t pg = (uint32_t)(x >> 32) % pages, sc = (uint32_t)x % sec_per_page;
            a[j] = (uint64_t)pg * stride_words + (uint64_t)sc * 4;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc += __ldg(buf + a[j]);
    }
    if (acc == 0x1234567) out[0] = acc;
}
int main() {
    const size_t maxb = 12ull << 30;
    uint64_t *buf, *out;
    if (cudaMalloc(&buf, maxb) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 8);
    cudaMemset(buf, 1, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { uint32_t pages, kb, stride_kb; } cfgs[] = {
        {4400, 10, 10}, {4400, 10, 2048}, {1100, 40, 40}, {1100, 40, 2048}, {200, 220, 220}, {200, 220, 2048}, {4400, 10, 2560}, {64, 700, 2048}};
    for (auto c : cfgs) {
        const int grid = 148 * 8, block = 256; const uint32_t iters = 2048;
        uint32_t spp = c.kb * 1024 / 32; uint64_t sw = (uint64_t)c.stride_kb * 1024 / 8;
        k<<<grid, block>>>(buf, c.pages, spp, sw, iters, out, 1);
        cudaEventRecord(e0);
        k<<<grid, block>>>(buf, c.pages, spp, sw, iters, out, 2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double acc = (double)grid * block * iters;
        printf("%5u regions x %4u KB, stride %5u KB (footprint %.0f MB): %.1f G sector reads/s  err=%d\n", c.pages, c.kb, c.stride_kb,
               c.pages * (double)c.kb / 1024, acc / ms / 1e6, (int)cudaGetLastError());
    }
    return 0;
}
