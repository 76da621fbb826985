// Variants of the streaming GEMV, to find the fast formulation (one 512-thread CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ double smem_d[];
__device__ __forceinline__ void ld4(const double* p, double* v) {
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld4p(const double* p, double* v, int pred) {
    asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.s32 pp, %5, 0;\n\t@pp ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]) : "l"(p), "r"(pred));
}
// LAYOUT 0: slices inside the warp (shuffle epilogue); 1: slices across warps (smem partials + barrier)
// LOADK 0: predicated asm; 1: clamped unconditional; 2: if-guarded plain; 3: packed {offset, weight} 16-byte entries +
// predicated asm; 4: predicated asm, no weights / no list (upper bound of the structure)
template <int LAYOUT, int LOADK, int NB, bool USELIST>
__device__ void gemv(const double* base, int ld, const int* list, const double* w, int cnt, int rows, const double* init, double* out, double* buf) {
    constexpr int VW = 4, NT = 512, NW = 16;
    const int G = rows / VW;
    int SL, sl, g;
    if (LAYOUT == 0) {
        SL = 32; while (SL > 1 && (G * SL > NT || 2 * SL > cnt)) SL >>= 1;
        const int GPW = 32 / SL, wv = threadIdx.x >> 5, l = threadIdx.x & 31;
        sl = l / GPW; g = wv * GPW + (l - sl * GPW);
    } else {
        SL = NT / G; if (SL < 1) SL = 1; if (SL > 8) SL = 8;
        sl = threadIdx.x / G; g = threadIdx.x - sl * G;
        if (sl >= SL) g = G;      // idle
    }
    double acc[VW] = {0, 0, 0, 0}, acc2[VW] = {0, 0, 0, 0};
    if (g < G) {
        const double* bg = base + VW * g;
        for (int t0 = sl; t0 < cnt; t0 += NB * SL) {
            double v[NB][VW], wv[NB]; int kk[NB];
            if (LOADK == 3) {
                const double2* cw = reinterpret_cast<const double2*>(buf + 4096);    // {offset as double bits, weight}
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int te = t0 + e * SL; const int tc = te < cnt ? te : cnt - 1;
                    const double2 ent = cw[tc];
                    kk[e] = (int)__double_as_longlong(ent.x); wv[e] = (te < cnt) ? ent.y : 0.0;
                    for (int q = 0; q < VW; ++q) v[e][q] = 0.0;
                }
#pragma unroll
                for (int e = 0; e < NB; ++e) ld4p(bg + kk[e], v[e], (t0 + e * SL < cnt) ? 1 : 0);
            } else if (LOADK == 4) {
#pragma unroll
                for (int e = 0; e < NB; ++e) { wv[e] = 1.0; for (int q = 0; q < VW; ++q) v[e][q] = 0.0; }
#pragma unroll
                for (int e = 0; e < NB; ++e) ld4p(bg + (size_t)(t0 + e * SL) * ld, v[e], (t0 + e * SL < cnt) ? 1 : 0);
            } else {
#pragma unroll
            for (int e = 0; e < NB; ++e) { const int te = t0 + e * SL; const int tc = te < cnt ? te : cnt - 1; kk[e] = USELIST ? list[tc] : tc; }
#pragma unroll
            for (int e = 0; e < NB; ++e) { const double wk = w[kk[e]]; wv[e] = (t0 + e * SL < cnt) ? wk : 0.0; for (int q = 0; q < VW; ++q) v[e][q] = 0.0; }
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                const double* p = bg + (size_t)kk[e] * ld;
                if (LOADK == 0) ld4p(p, v[e], (t0 + e * SL < cnt) ? 1 : 0);
                else if (LOADK == 1) ld4(p, v[e]);
                else { if (t0 + e * SL < cnt) ld4(p, v[e]); }
            }
            }
#pragma unroll
            for (int e = 0; e < NB; ++e)
#pragma unroll
                for (int q = 0; q < VW; ++q) { if (e & 1) acc2[q] += v[e][q] * wv[e]; else acc[q] += v[e][q] * wv[e]; }
        }
    }
    if (LAYOUT == 0) {
        const int GPW = 32 / SL;
#pragma unroll
        for (int q = 0; q < VW; ++q) {
            double s2 = acc[q] + acc2[q];
            for (int o = GPW; o < 32; o <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (sl == 0 && g < G) out[VW * g + q] = (init ? init[VW * g + q] : 0.0) + s2;
        }
        __syncthreads();
    } else {
        if (g < G) for (int q = 0; q < VW; ++q) buf[sl * rows + VW * g + q] = acc[q] + acc2[q];
        __syncthreads();
        for (int r = threadIdx.x; r < rows; r += NT) { double s = init ? init[r] : 0.0; for (int s2 = 0; s2 < SL; ++s2) s += buf[s2 * rows + r]; out[r] = s; }
        __syncthreads();
    }
}
template <int LAYOUT, int LOADK, int NB, bool USELIST>
__global__ void __launch_bounds__(512, 1) k(const double* V, int N, int cnt, int reps, long long* cyc, double* sink) {
    int* list = reinterpret_cast<int*>(smem_d + 8192);
    double* w = smem_d; double* out = smem_d + 1024; double* init = smem_d + 2048; double* buf = smem_d + 3072;
    for (int t = threadIdx.x; t < cnt; t += 512) list[t] = (t * 5) % N;
    for (int t = threadIdx.x; t < N; t += 512) { w[t] = 1.0 + t; init[t] = 0.5; }
    { double2* cw = reinterpret_cast<double2*>(buf + 4096);
      for (int t = threadIdx.x; t < cnt; t += 512) { const int kx = (t * 5) % N; cw[t] = make_double2(__longlong_as_double((long long)kx * N), 1.0 + kx); } }
    __syncthreads();
    long long tot = 0;
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        const long long t0 = clock64();
        gemv<LAYOUT, LOADK, NB, USELIST>(V, N, list, w, cnt, N, init, out, buf);
        tot += clock64() - t0;
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = tot;
    sink[blockIdx.x * 512 + threadIdx.x] = out[threadIdx.x % N];
}
template <int LAYOUT, int LOADK, int NB, bool USELIST> void run(const char* nm, const double* V, int N, int cnt) {
    long long* cyc; double* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 8);
    auto kk = k<LAYOUT, LOADK, NB, USELIST>;
    cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int reps = 20;
    kk<<<148, 512, 96 * 1024>>>(V, N, cnt, 2, cyc, sink);
    kk<<<148, 512, 96 * 1024>>>(V, N, cnt, reps, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    printf("%-34s N=%d cnt=%d: %6.0f cycles/pass (%5.1f B/clk/SM) %s\n", nm, N, cnt, m / reps, 8.0 * N * cnt / (m / reps), e ? cudaGetErrorString(e) : "");
    cudaFree(cyc); cudaFree(sink);
}
int main() {
    double* V; cudaMalloc(&V, 500 * 500 * 8); cudaMemset(V, 0, 500 * 500 * 8);
    for (int pass = 0; pass < 2; ++pass) {
        const int N = pass ? 100 : 500, cnt = pass ? 70 : 85;
        run<1, 0, 6, true>("across predasm NB6 list", V, N, cnt);
        run<1, 3, 6, true>("across packed  NB6", V, N, cnt);
        run<1, 3, 4, true>("across packed  NB4", V, N, cnt);
        run<1, 3, 8, true>("across packed  NB8", V, N, cnt);
        run<1, 4, 6, true>("across noweights NB6 (bound)", V, N, cnt);
        run<1, 4, 8, true>("across noweights NB8 (bound)", V, N, cnt);
        run<0, 3, 6, true>("inwarp packed  NB6", V, N, cnt);
        run<0, 4, 6, true>("inwarp noweights NB6 (bound)", V, N, cnt);
    }
    return 0;
}
