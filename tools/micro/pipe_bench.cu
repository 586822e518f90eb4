// micro-benchmark: issue rates of the integer instructions this path is made of (per SM, per clock)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t *out, int iters, uint32_t c1, uint32_t c2) {
    uint32_t a = threadIdx.x * 2654435761u + c1, b = a ^ 0x9E3779B9u, c = a + 77u, d = b + 99u;
    uint32_t e = a * 3u, f = b * 5u, g = c * 7u, h = d * 9u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) {        // SHF.R funnel (alu)
                a = __funnelshift_r(a, b, 7); b = __funnelshift_r(b, c, 9); c = __funnelshift_r(c, d, 11); d = __funnelshift_r(d, a, 13);
                e = __funnelshift_r(e, f, 7); f = __funnelshift_r(f, g, 9); g = __funnelshift_r(g, h, 11); h = __funnelshift_r(h, e, 13);
            } else if (MODE == 1) { // LOP3 (alu)
                a = (a & b) ^ c; b = (b | c) ^ d; c = (c & d) ^ a; d = (d | a) ^ b;
                e = (e & f) ^ g; f = (f | g) ^ h; g = (g & h) ^ e; h = (h | e) ^ f;
            } else if (MODE == 2) { // IMAD lo (fma)
                a = a * c1 + b; b = b * c2 + c; c = c * c1 + d; d = d * c2 + a;
                e = e * c1 + f; f = f * c2 + g; g = g * c1 + h; h = h * c2 + e;
            } else if (MODE == 3) { // IMAD.HI (fma?)
                a = __umulhi(a, c1) + 1u; b = __umulhi(b, c2) + 1u; c = __umulhi(c, c1) + 1u; d = __umulhi(d, c2) + 1u;
                e = __umulhi(e, c1) + 1u; f = __umulhi(f, c2) + 1u; g = __umulhi(g, c1) + 1u; h = __umulhi(h, c2) + 1u;
            } else if (MODE == 4) { // half LOP3, half IMAD
                a = (a & b) ^ c; b = b * c2 + c; c = (c & d) ^ a; d = d * c2 + a;
                e = (e & f) ^ g; f = f * c2 + g; g = (g & h) ^ e; h = h * c2 + e;
            } else if (MODE == 5) { // half LOP3, half IMAD.HI
                a = (a & b) ^ c; b = __umulhi(b, 0x40000000u); c = (c & d) ^ a; d = __umulhi(d, 0x10000000u);
                e = (e & f) ^ g; f = __umulhi(f, 0x04000000u); g = (g & h) ^ e; h = __umulhi(h, 0x40000000u);
                b += a; d += c; f += e; h += g;   // keep values alive (IADD: alu or fma)
            } else if (MODE == 6) { // PRMT
                a = __byte_perm(a, b, c1); b = __byte_perm(b, c, c2); c = __byte_perm(c, d, c1); d = __byte_perm(d, a, c2);
                e = __byte_perm(e, f, c1); f = __byte_perm(f, g, c2); g = __byte_perm(g, h, c1); h = __byte_perm(h, e, c2);
            } else if (MODE == 7) { // IADD3
                a = a + b + c; b = b + c + d; c = c + d + a; d = d + a + b;
                e = e + f + g; f = f + g + h; g = g + h + e; h = h + e + f;
            } else if (MODE == 9) { // IDP.4A
                a = __dp4a(a, c1, b); b = __dp4a(b, c2, c); c = __dp4a(c, c1, d); d = __dp4a(d, c2, a);
                e = __dp4a(e, c1, f); f = __dp4a(f, c2, g); g = __dp4a(g, c1, h); h = __dp4a(h, c2, e);
            } else if (MODE == 10) { // half LOP3, half IDP.4A
                a = (a & b) ^ c; b = __dp4a(b, c2, c); c = (c & d) ^ a; d = __dp4a(d, c2, a);
                e = (e & f) ^ g; f = __dp4a(f, c2, g); g = (g & h) ^ e; h = __dp4a(h, c2, e);
            } else if (MODE == 11) { // POPC (+IADD)
                a = __popc(a) + b; b = __popc(b) + c; c = __popc(c) + d; d = __popc(d) + a;
                e = __popc(e) + f; f = __popc(f) + g; g = __popc(g) + h; h = __popc(h) + e;
            } else if (MODE == 12) { // 2 LOP3 : 1 IMAD : 1 IDP.4A (the newline test mix)
                a = (a & b) ^ c; b = (b | c) ^ d; c = c * c1 + d; d = __dp4a(d, c2, a);
                e = (e & f) ^ g; f = (f | g) ^ h; g = g * c1 + h; h = __dp4a(h, c2, e);
            } else if (MODE == 13) { // FLO/ffs (+IADD)
                a = __ffs(a) + b; b = __ffs(b) + c; c = __ffs(c) + d; d = __ffs(d) + a;
                e = __ffs(e) + f; f = __ffs(f) + g; g = __ffs(g) + h; h = __ffs(h) + e;
            } else if (MODE == 8) { // SHF.R.U32 plain shift
                a = (a >> 3) ^ 0; b = b >> 5; c = c >> 7; d = d >> 9; e = e >> 3; f = f >> 5; g = g >> 7; h = h >> 9;
                a += c1; b += c1; c += c1; d += c1; e += c2; f += c2; g += c2; h += c2;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
template <int MODE>
void run(const char *name, int ops_per_unroll) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; cudaMalloc(&out, sms * 1024 * 4 * 2);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, 1024>>>(out, 10, 0x01041041u, 0x02040811u);
    cudaEventRecord(e0);
    k<MODE><<<sms, 1024>>>(out, iters, 0x01041041u, 0x02040811u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double winst = (double)1024 / 32 * iters * 16 * ops_per_unroll;   // warp-instructions per SM (source-level ops)
    printf("%-26s %.3f ms  %.2f source-ops/clk/SM (%.2f per sub-partition) at %d MHz max  %s\n", name, ms, winst / (ms * 1e-3 * clk * 1e3),
           winst / (ms * 1e-3 * clk * 1e3) / 4, clk / 1000, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<0>("SHF funnel", 8); run<1>("LOP3", 8); run<2>("IMAD lo", 8); run<3>("IMAD.HI (+IADD)", 8); run<4>("LOP3 + IMAD", 8);
    run<5>("LOP3 + IMAD.HI (+IADD)", 8); run<6>("PRMT", 8); run<7>("IADD3", 8); run<8>("SHF.R shift (+IADD)", 8);
    run<9>("IDP.4A", 8); run<10>("LOP3 + IDP.4A", 8); run<11>("POPC (+IADD)", 8); run<12>("2 LOP3 + IMAD + IDP.4A", 8); run<13>("ffs (+IADD)", 8);
    return 0;
}
