// Latencies that shape the packed-inverse primitives: dependent DFMA chain, shared-memory load-to-use, and DFMA issue
// rate for 1..16 warps of one CTA (one CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, int warps, long long* cyc, double* sink) {
    __shared__ double sm[4096];
    __shared__ int nxt[1024];
    for (int t = threadIdx.x; t < 4096; t += blockDim.x) sm[t] = 1.0 + 1e-9 * t;
    for (int t = threadIdx.x; t < 1024; t += blockDim.x) nxt[t] = (t * 37 + 11) & 1023;
    __syncthreads();
    if ((threadIdx.x >> 5) >= warps) return;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1.0000001, c = 1e-9;
    double a2 = a, a3 = a, a4 = a, a5 = a, a6 = a, a7 = a, a8 = a;
    int p = threadIdx.x & 1023;
    const long long t0 = clock64();
    if (mode == 0) {            // one dependent DFMA chain
#pragma unroll 16
        for (int i = 0; i < 1024; ++i) a = a * b + c;
    } else if (mode == 1) {     // 8 independent chains
#pragma unroll 4
        for (int i = 0; i < 1024; ++i) { a = a * b + c; a2 = a2 * b + c; a3 = a3 * b + c; a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c; a8 = a8 * b + c; }
    } else if (mode == 2) {     // shared-memory pointer chase
#pragma unroll 16
        for (int i = 0; i < 1024; ++i) p = nxt[p];
    } else if (mode == 3) {     // load -> DFMA dependent chain (a feeds the address)
#pragma unroll 16
        for (int i = 0; i < 1024; ++i) { a = a * sm[p] + c; p = (p + 33) & 4095; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a + a2 + a3 + a4 + a5 + a6 + a7 + a8 + p;
}
int main() {
    long long* cyc; double* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 8);
    const char* names[] = {"1 dependent DFMA chain (cycles per DFMA)", "8 independent DFMA chains (cycles per 8 DFMA)", "LDS pointer chase (cycles per load)", "LDS + DFMA, independent addresses (cycles per step)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            k<<<148, 512>>>(mode, warps, cyc, sink);
            k<<<148, 512>>>(mode, warps, cyc, sink);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            double a = 0; for (int i = 0; i < 148; ++i) a += h[i];
            printf("%-52s warps=%2d: %.1f\n", names[mode], warps, a / 148.0 / 1024.0);
        }
    return 0;
}
