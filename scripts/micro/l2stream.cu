// Microbenchmark: per-SM L2->SM streaming rate with ONE 512-thread CTA per SM (the solve kernel's regime).
// Each thread streams 32-byte (or 16-byte) vectors from an L2-resident buffer with NB loads in flight.
#include <cstdio>
#include <cuda_runtime.h>
template <int VW> __device__ __forceinline__ void ld(const double* p, double* v);
template <> __device__ __forceinline__ void ld<4>(const double* p, double* v) {
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
template <> __device__ __forceinline__ void ld<2>(const double* p, double* v) {
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
// buffer = ncols columns of `rows` doubles; every CTA reads ALL columns (like the gradient pass): thread (g, sl)
template <int VW, int NB, bool ROLL>
__global__ void __launch_bounds__(512, 1) stream(const double* buf, int rows, int ncols, int reps, double* out, long long* cyc) {
    const int G = rows / VW, SL = 512 / G > 0 ? 512 / G : 1;
    const int g = threadIdx.x % G, sl = threadIdx.x / G;
    double acc[VW] = {0};
    __syncthreads();
    long long t0 = clock64();
    if (sl < SL)
    for (int r = 0; r < reps; ++r) {
        if (ROLL) {
            double v[NB][VW];
#pragma unroll
            for (int e = 0; e < NB; ++e) { int t = sl + e * SL; if (t < ncols) ld<VW>(buf + (size_t)t * rows + VW * g, v[e]); else for (int q = 0; q < VW; ++q) v[e][q] = 0; }
            for (int t0_ = sl; t0_ < ncols; t0_ += NB * SL) {
#pragma unroll
                for (int e = 0; e < NB; ++e) {
#pragma unroll
                    for (int q = 0; q < VW; ++q) acc[q] += v[e][q];
                    int t = t0_ + (NB + e) * SL;
                    if (t < ncols) ld<VW>(buf + (size_t)t * rows + VW * g, v[e]); else for (int q = 0; q < VW; ++q) v[e][q] = 0;
                }
            }
        } else {
            for (int t0_ = sl; t0_ < ncols; t0_ += NB * SL) {
                double v[NB][VW];
#pragma unroll
                for (int e = 0; e < NB; ++e) { int t = t0_ + e * SL; if (t < ncols) ld<VW>(buf + (size_t)t * rows + VW * g, v[e]); else for (int q = 0; q < VW; ++q) v[e][q] = 0; }
#pragma unroll
                for (int e = 0; e < NB; ++e)
#pragma unroll
                    for (int q = 0; q < VW; ++q) acc[q] += v[e][q];
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int q = 0; q < VW; ++q) s += acc[q];
    out[blockIdx.x * 512 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int VW, int NB, bool ROLL>
void run(const char* name, const double* buf, int rows, int ncols, int grid) {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
    const int reps = 50;
    stream<VW, NB, ROLL><<<grid, 512>>>(buf, rows, ncols, 2, out, cyc);
    stream<VW, NB, ROLL><<<grid, 512>>>(buf, rows, ncols, reps, out, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    double bytes = (double)rows * ncols * 8 * reps;
    printf("%-28s grid %3d rows %4d cols %3d: %.0f cycles/pass, %.1f B/clk/SM\n", name, grid, rows, ncols, mean / reps, bytes / mean);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    double* buf; cudaMalloc(&buf, 8 << 20); cudaMemset(buf, 0, 8 << 20);
    for (int grid : {1, 148}) {
        run<4, 6, false>("v4 NB6 drain", buf, 500, 85, grid);
        run<4, 6, true>("v4 NB6 roll", buf, 500, 85, grid);
        run<4, 3, true>("v4 NB3 roll", buf, 500, 85, grid);
        run<4, 8, true>("v4 NB8 roll", buf, 500, 85, grid);
        run<2, 8, false>("v2 NB8 drain", buf, 500, 85, grid);
        run<2, 8, true>("v2 NB8 roll", buf, 500, 85, grid);
        run<2, 12, true>("v2 NB12 roll", buf, 500, 85, grid);
        run<4, 6, true>("v4 NB6 roll cols500", buf, 500, 500, grid);
        run<4, 6, true>("v4 NB6 roll rows100 c70", buf, 100, 70, grid);
    }
    return 0;
}
