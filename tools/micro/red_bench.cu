// micro-benchmark: global histogram updates, random bins: RED.64 into int64[2^24] (128 MiB) vs RED.32 into u32[2^24] (64 MiB)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int W>
__global__ void k(void *tab, uint32_t nb_log2, int iters) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t m = (1u << nb_log2) - 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x = x * 1664525u + 1013904223u;
            const uint32_t b = (x >> 8) & m;
            if (W == 64) atomicAdd(reinterpret_cast<unsigned long long *>(tab) + b, 1ull);
            else atomicAdd(reinterpret_cast<uint32_t *>(tab) + b, 1u);
        }
    }
}
template <int W>
void run(int nbl) {
    void *tab; size_t bytes = ((size_t)W / 8) << nbl; cudaMalloc(&tab, bytes); cudaMemset(tab, 0, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, blocks = sms * 2, iters = 1000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<W><<<blocks, threads>>>(tab, nbl, 10);
    cudaEventRecord(e0);
    k<W><<<blocks, threads>>>(tab, nbl, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)blocks * threads * iters * 8;
    printf("RED.%d into 2^%d bins (%zu MiB): %.3f ms for %.2e updates = %.1f G updates/s  %s\n", W, nbl, bytes >> 20, ms, n, n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(tab);
}
int main() { run<64>(24); run<32>(24); run<64>(22); run<32>(22); run<64>(20); run<32>(26); return 0; }
