// Microbenchmark of the solver's streaming GEMV *as written in csrc/ssqp_kernel.cuh* (in-warp slices, column list and
// weights in shared memory, predicated 256-bit loads, shuffle epilogue), one 512-thread CTA per SM, hot and "cold"
// (a large dummy code block executed between passes to evict the instruction cache).
#include <cstdio>
#include <cuda_runtime.h>
#define SSQP_ONLY_VW4 1
#include "../../statusswitchingqp.jl_b200/csrc/ssqp_kernel.cuh"
using namespace ssqp;

template <int COLD>
__global__ void __launch_bounds__(512, 1) k(const double* V, int N, int cnt, int reps, long long* cyc, double* sink, int ncols) {
    int* list = reinterpret_cast<int*>(smem_d + 4096);
    double* w = smem_d; double* out = smem_d + 1024; double* init = smem_d + 2048;
    for (int t = threadIdx.x; t < cnt; t += 512) list[t] = (int)(((unsigned)(t * 37 + blockIdx.x * 101 + 11) * 2654435761u) % (unsigned)ncols);
    for (int t = threadIdx.x; t < N; t += 512) { w[t] = 1.0 + t; init[t] = 0.5; }
    __syncthreads();
    long long tot = 0;
    double junk = threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        if (COLD) {      // ~200 KB of straight-line code (distinct instructions) between passes
#pragma unroll 1
            for (int it = 0; it < 1; ++it) {
#pragma unroll
                for (int u = 0; u < 12000; ++u) junk = junk * 1.0000001 + (double)u;
            }
        }
        __syncthreads();
        const long long t0 = clock64();
        gemv_cols<512>(GemvArgs{V, N, ioff(list), soff(w), cnt, N, nullptr, soff(init), soff(out), soff(smem_d + 5120), 1536, -1});
        tot += clock64() - t0;
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = tot;
    sink[blockIdx.x * 512 + threadIdx.x] = out[threadIdx.x % N] + junk;
}
template <int COLD> void run(const char* nm, const double* V, int N, int cnt, int ncols) {
    long long* cyc; double* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 8);
    cudaFuncSetAttribute(k<COLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 20;
    k<COLD><<<148, 512, 64 * 1024>>>(V, N, cnt, 2, cyc, sink, ncols);
    k<COLD><<<148, 512, 64 * 1024>>>(V, N, cnt, reps, cyc, sink, ncols);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    printf("%-10s rows=%d cnt=%d of %d cols: %.0f cycles/pass (%.1f B/clk/SM) %s\n", nm, N, cnt, ncols, m / reps, 8.0 * N * cnt / (m / reps), cudaGetErrorString(e));
}
int main() {
    double* V; cudaMalloc(&V, 500 * 500 * 8); cudaMemset(V, 0, 500 * 500 * 8);
    run<0>("hot", V, 500, 85, 500); run<0>("hot", V, 500, 65, 100);
    run<0>("hot", V, 100, 70, 100); run<0>("hot", V, 100, 66, 500); run<0>("hot", V, 100, 20, 500); run<0>("hot", V, 100, 8, 500);
    run<0>("hot", V, 100, 130, 500);
    return 0;
}
