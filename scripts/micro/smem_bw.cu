// Calibration: shared-memory read bandwidth of one 512-thread CTA per SM, contiguous 64-bit vs 128-bit loads,
// with and without a dependent FMA chain / a broadcast operand.
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ double sm[];
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int ndbl, int reps, long long* cyc, double* sink) {
    for (int t = threadIdx.x; t < ndbl + 1024; t += 512) sm[t] = 1e-3 * (t % 7);
    __syncthreads();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const double* x = sm + ndbl;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {            // LDS.64, 4 accumulators
            for (int i = threadIdx.x; i + 1536 < ndbl; i += 2048) { a0 += sm[i]; a1 += sm[i + 512]; a2 += sm[i + 1024]; a3 += sm[i + 1536]; }
        } else if (MODE == 1) {     // LDS.128
            const double2* s2 = reinterpret_cast<const double2*>(sm);
            for (int i = threadIdx.x; i + 512 < ndbl / 2; i += 1024) { double2 u = s2[i], v = s2[i + 512]; a0 += u.x; a1 += u.y; a2 += v.x; a3 += v.y; }
        } else if (MODE == 2) {     // LDS.64 + broadcast LDS.64 + FMA (the GEMV inner loop)
            for (int i = threadIdx.x, m = 0; i + 1536 < ndbl; i += 2048, m += 4) {
                a0 += sm[i] * x[m]; a1 += sm[i + 512] * x[m + 1]; a2 += sm[i + 1024] * x[m + 2]; a3 += sm[i + 1536] * x[m + 3]; }
        } else {                    // LDS.128 + broadcast LDS.128 + FMA
            const double2* s2 = reinterpret_cast<const double2*>(sm); const double2* x2 = reinterpret_cast<const double2*>(x);
            for (int i = threadIdx.x, m = 0; i + 512 < ndbl / 2; i += 1024, m += 2) { double2 u = s2[i], v = s2[i + 512], p = x2[m], q = x2[m + 1];
                a0 += u.x * p.x; a1 += u.y * p.y; a2 += v.x * q.x; a3 += v.y * q.y; }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 512 + threadIdx.x] = a0 + a1 + a2 + a3;
}
template <int MODE> void run(const char* nm, int ndbl) {
    long long* cyc; double* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 8);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 50;
    k<MODE><<<148, 512, 200 * 1024>>>(ndbl, reps, cyc, sink);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148.0 * reps;
    printf("%-36s %6d doubles: %6.0f cycles/pass, %.1f B/clk\n", nm, ndbl, m, 8.0 * ndbl / m);
}
int main() {
    for (int nd : {10240, 20480}) {
        run<0>("LDS.64 sum", nd); run<1>("LDS.128 sum", nd);
        run<2>("LDS.64 * bcast LDS.64 (FMA)", nd); run<3>("LDS.128 * bcast LDS.128 (FMA)", nd);
    }
    return 0;
}
