// ssqp_helpers.cuh — one-off helper kernels of the host runtime (set_shared preprocessing, roofline microbenchmarks).
#pragma once
#include <cuda_runtime.h>
namespace ssqp {
// ---- set_shared helpers ----------------------------------------------------------------------------
// Crow = Ccol' ; cA[k] = ||Ccol[:,k]||_2
__global__ void ssqp_prep_kernel(int N, int M0, const double* Ccol, double* Crow, double* cA) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int i = 0; i < M0; ++i) {
            const double v = Ccol[i + (size_t)k * M0];
            Crow[k + (size_t)i * N] = v;
            s += v * v;
        }
        cA[k] = sqrt(s);
    }
}
// Ccol = [A;G] from separate column-major A (M x N) and G (J x N)
__global__ void ssqp_stack_kernel(int N, int M, int J, const double* A, const double* G, double* Ccol) {
    const int M0 = M + J;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)N * M0; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / M0), r = (int)(t % M0);
        Ccol[t] = (r < M) ? A[r + (size_t)k * M] : G[(r - M) + (size_t)k * J];
    }
}

// ---- roofline microbenchmarks -------------------------------------------------------------------
__global__ void ssqp_dfma_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
__global__ void ssqp_readbw_kernel(const double2* __restrict__ in, long long n2, int reps, double* out) {
    double s = 0.0;
    for (int r = 0; r < reps; ++r)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
            double2 v = in[i];
            s += v.x + v.y;
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


}  // namespace ssqp
