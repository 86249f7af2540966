// Microbenchmark: is the TLB limit per SM or shared? Every CTA reads random sectors from ITS OWN `ppc` pages (2 MB apart,
// `kb` KB used in each), so the chip touches grid*ppc pages in total but each SM only a few.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(const uint64_t* __restrict__ buf, uint32_t ppc, uint32_t sec_per_page, uint32_t iters, uint64_t* out, uint32_t seed,
                  uint32_t share) {
    uint64_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + seed;
    const uint64_t page_words = (2ull << 20) / 8;
    const uint32_t group = share ? blockIdx.x / share : blockIdx.x;  // `share` consecutive CTAs use the same pages
    uint64_t acc = 0;
    for (uint32_t i = 0; i < iters; i += 4) {
        uint64_t a[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            uint32_t pg = (uint32_t)(x >> 32) % ppc, sc = (uint32_t)x % sec_per_page;
            a[j] = ((uint64_t)group * ppc + pg) * page_words + (uint64_t)sc * 4;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc += __ldg(buf + a[j]);
    }
    if (acc == 0x1234567) out[0] = acc;
}
int main() {
    const int grid = 148 * 8, block = 256;
    const size_t maxb = (size_t)grid * 16 * (2ull << 20);  // up to 16 pages per CTA = 37 GB
    uint64_t *buf, *out;
    if (cudaMalloc(&buf, maxb) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 8);
    cudaMemset(buf, 1, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { uint32_t ppc, kb, share; } cfgs[] = {{1, 32, 0}, {2, 16, 0}, {8, 4, 0}, {16, 2, 0}, {8, 4, 8}, {16, 2, 8}, {8, 256, 8}, {8, 2048, 8}};
    for (auto c : cfgs) {
        const uint32_t iters = 2048;
        uint32_t spp = c.kb * 1024 / 32;
        k<<<grid, block>>>(buf, c.ppc, spp, iters, out, 1, c.share);
        cudaEventRecord(e0);
        k<<<grid, block>>>(buf, c.ppc, spp, iters, out, 2, c.share);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double acc = (double)grid * block * iters;
        uint32_t groups = c.share ? grid / c.share : grid;
        printf("%2u pages per CTA x %4u KB, %u CTAs share a page set: %5u pages chip-wide, footprint %6.0f MB: %.1f G sector reads/s err=%d\n",
               c.ppc, c.kb, c.share ? c.share : 1, groups * c.ppc, groups * c.ppc * (double)c.kb / 1024, acc / ms / 1e6, (int)cudaGetLastError());
    }
    return 0;
}
