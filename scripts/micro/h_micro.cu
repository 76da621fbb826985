// Microbenchmark of the packed-inverse primitives of csrc/ssqp_kernel.cuh (symv_leaf, syr_leaf), one 512-thread CTA
// per SM, H entirely in shared memory.
#include <cstdio>
#include <cuda_runtime.h>
#define SSQP_ONLY_VW4 1
#include "../../statusswitchingqp.jl_b200/csrc/ssqp_kernel.cuh"
using namespace ssqp;
__global__ void __launch_bounds__(512, 1) k(int n, int reps, long long* cyc, double* sink) {
    double* Hs = smem_d + 4096; double* x = smem_d; double* y = smem_d + 1024; double* buf = smem_d + 2048;
    for (int t = threadIdx.x; t < n * (n + 1) / 2; t += 512) Hs[t] = 1e-3 * (t % 17);
    for (int t = threadIdx.x; t < n; t += 512) x[t] = 1.0 + t;
    __syncthreads();
    HView h{Hs, nullptr, 1000, buf};
    long long t1 = 0, t2 = 0;
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        long long t0 = clock64();
        symv_leaf<512>(h, n, x, y);
        long long tm = clock64();
        syr_leaf<512>(h, n, y, 1e-9);
        t1 += tm - t0; t2 += clock64() - tm;
    }
    if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t1; cyc[2 * blockIdx.x + 1] = t2; }
    sink[blockIdx.x * 512 + threadIdx.x] = y[threadIdx.x % n] + Hs[threadIdx.x];
}
int main() {
    long long* cyc; double* sink; cudaMalloc(&cyc, 148 * 16); cudaMalloc(&sink, 148 * 512 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int n : {64, 100, 140, 170, 200}) {
        const int reps = 20;
        k<<<148, 512, 220 * 1024>>>(n, 2, cyc, sink);
        k<<<148, 512, 220 * 1024>>>(n, reps, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[296]; cudaMemcpy(h, cyc, 296 * 8, cudaMemcpyDeviceToHost);
        double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
        a /= 148.0 * reps; b /= 148.0 * reps;
        const double bytes = 4.0 * n * (n + 1);
        printf("n=%3d: symv %6.0f cyc (%.1f B/clk of 2x%.0f KB)   syr %6.0f cyc (%.1f B/clk r+w)  %s\n", n, a, 2 * bytes / a, bytes / 1e3, b, 2 * bytes / b, e ? cudaGetErrorString(e) : "");
    }
    return 0;
}
