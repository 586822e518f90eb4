// micro-benchmark: shared-memory atomicAdd throughput (random bins vs bank-distinct bins)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t *out, int iters, int n_bins_log2, uint32_t one) {
    extern __shared__ uint32_t hist[];
    const uint32_t nb = 1u << n_bins_log2;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t mm = (nb - 1) << 2;
    const uint32_t lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x = x * 1664525u + 1013904223u;
            uint32_t off = (x >> 10) & mm;
            if (MODE == 1) off = (off & ~0x7Cu) | (lane << 2);   // every lane its own bank
            if (MODE == 2) off = (off & ~0x7Cu) | ((lane & 15) << 2);   // 2-way
            if (MODE == 3) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(hist) + off), "r"(one) : "memory");
            else atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(hist) + off), 1u);
        }
    }
    __syncthreads();
    uint32_t s = 0;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s += hist[i];
    if (s == 0xFFFFFFFFu) out[0] = s;
}
template <int MODE>
void run(const char *name, int threads, int blocks_per_sm, int nbl) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; cudaMalloc(&out, 4);
    size_t smem = (size_t)4 << nbl;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * blocks_per_sm, threads, smem>>>(out, 10, nbl, 1u);
    cudaEventRecord(e0);
    k<MODE><<<sms * blocks_per_sm, threads, smem>>>(out, iters, nbl, 1u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)sms * blocks_per_sm * threads * iters * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s thr=%4d x%d bins=2^%d: %.3f ms, %.1f G atomics/s, %.2f atomics/clk/SM (at %d MHz max)  %s\n", name, threads, blocks_per_sm, nbl, ms, n / ms / 1e6,
           n / sms / (ms * 1e-3 * clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("random", 768, 1, 14); run<0>("random", 1024, 1, 14); run<0>("random", 288, 3, 14); run<0>("random", 256, 1, 14);
    run<1>("bank-distinct", 768, 1, 14); run<1>("bank-distinct", 288, 3, 14);
    run<2>("2-way", 768, 1, 14);
    run<0>("random 2^10", 768, 1, 10);
    run<3>("random, ATOMS.ADD reg", 768, 1, 14);
    return 0;
}
