// Calibration: cost (SM cycles, one 512-thread CTA per SM) of the solver's basic building blocks.
#include <cstdio>
#include <cuda_runtime.h>
#define SSQP_ONLY_VW4 1
#include "../../statusswitchingqp.jl_b200/csrc/ssqp_kernel.cuh"
using namespace ssqp;
#define TIME(slot, ...) { __syncthreads(); long long t0_ = clock64(); for (int r_ = 0; r_ < REPS; ++r_) { __VA_ARGS__; } __syncthreads(); if (threadIdx.x == 0) cyc[slot] = (clock64() - t0_) / REPS; }
__global__ void __launch_bounds__(512, 1) k(long long* cyc, double* sink, int M0, int div) {
    constexpr int REPS = 50;
    Ctx c; c.buf = smem_d + 14000; c.red = smem_d + 15000;
    double* A = smem_d; double* x = smem_d + 12000; double* y = smem_d + 13000;
    const int ldB = M0 | 1;
    for (int t = threadIdx.x; t < ldB * M0; t += 512) A[t] = 1e-3 * (t % 13);
    for (int t = threadIdx.x; t < 1024; t += 512) { x[t] = t; y[t] = 0; }
    double v = threadIdx.x;
    int iv = threadIdx.x + 7;
    TIME(0, __syncthreads());
    TIME(1, v = block_sum<512>(c, v * 0.5));
    TIME(2, v = block_max<512>(c, v));
    { Cand q; TIME(3, q.offer(v, threadIdx.x); block_argmin<512>(c, q); v += q.id); }
    TIME(4, small_reduce<512>(c, M0, M0, [=](int j, int i) { return A[j + (size_t)i * ldB] * x[i]; }, y));
    TIME(5, small_reduce<512>(c, M0, M0, [=](int i, int j) { return A[j + (size_t)i * ldB] * x[j]; }, y));
    TIME(6, iv = iv / div + threadIdx.x);
    TIME(7, { double a0 = 0, a1 = 0; for (int i = 0; i < 24; i += 2) { a0 += A[threadIdx.x + i * 101] * x[i]; a1 += A[threadIdx.x + (i + 1) * 101] * x[i + 1]; } y[threadIdx.x] = a0 + a1; });
    TIME(8, { int n = compact_nonzero<512>(c, x, 500, reinterpret_cast<int*>(smem_d + 16000)); iv += n; });
    TIME(9, { for (int t = threadIdx.x; t < ldB * M0; t += 512) A[t] = A[t] * 1.0000001 + 1e-9; });
    sink[blockIdx.x * 512 + threadIdx.x] = v + iv + y[threadIdx.x];
}
int main() {
    long long* cyc; double* sink; cudaMalloc(&cyc, 16 * 8); cudaMalloc(&sink, 148 * 512 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    k<<<148, 512, 220 * 1024>>>(cyc, sink, 100, 7);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    const char* nm[] = {"__syncthreads", "block_sum", "block_max", "block_argmin", "small_reduce A*x (100x100)", "small_reduce A'*x (100x100)", "int division", "24 LDS+FMA per thread", "compact_nonzero(500)", "elementwise update 100x101"};
    for (int i = 0; i < 10; ++i) printf("%-30s %6lld cycles\n", nm[i], h[i]);
    printf("%s\n", e ? cudaGetErrorString(e) : "ok");
    return 0;
}
