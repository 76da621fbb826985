// Candidate rewrites of the packed-inverse primitives (symv_leaf / syr_leaf of csrc/ssqp_kernel.cuh), timed against
// the current ones on a shared-memory resident packed inverse, one 512-thread CTA per SM.  Results are compared too.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#define SSQP_ONLY_VW4 1
#include "../../statusswitchingqp.jl_b200/csrc/ssqp_kernel.cuh"
using namespace ssqp;

// ---- syr: groups of 8 rows x 32 columns per warp step, loads first, round-robin over the warps ----------------------
template <int NT>
static __device__ __noinline__ void syr_v2(const HView h, int n, const double* __restrict__ v, double sigma) {
    constexpr int NW = NT / 32;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int ns = n < h.R ? n : h.R;
    double* __restrict__ Hs = h.Hs;
    int gbase = 0;
    for (int q0 = 0; q0 < ns; q0 += 32) {
        const int k = q0 + l;
        const double vk = (k < ns) ? v[k] : 0.0;
        const int ngr = (ns - q0 + 7) >> 3;
        for (int g = (w - gbase) & (NW - 1); g < ngr; g += NW) {
            const int i0 = q0 + 8 * g;
            int offs[8];
            {
                int off = tri(i0) + k;
#pragma unroll
                for (int e = 0; e < 8; ++e) { offs[e] = off; off += i0 + e + 1; }
            }
            double a[8], cc[8];
            if (g >= 4 && i0 + 7 < ns) {          // every lane's column is <= every row of the group
#pragma unroll
                for (int e = 0; e < 8; ++e) cc[e] = sigma * v[i0 + e];
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = Hs[offs[e]];
#pragma unroll
                for (int e = 0; e < 8; ++e) Hs[offs[e]] = a[e] + cc[e] * vk;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) cc[e] = (i0 + e < ns) ? sigma * v[i0 + e] : 0.0;
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = (i0 + e < ns && k <= i0 + e) ? Hs[offs[e]] : 0.0;
#pragma unroll
                for (int e = 0; e < 8; ++e) if (i0 + e < ns && k <= i0 + e) Hs[offs[e]] = a[e] + cc[e] * vk;
            }
        }
        gbase += ngr;
    }
    __syncthreads();
}

// ---- symv: same thread layout as symv_leaf, 8 loads in flight and 8 accumulators ------------------------------------
template <int NT>
static __device__ __noinline__ void symv_v2(const HView h, int n, const double* x, double* y) {
    const int ns = n < h.R ? n : h.R;
    const double* Hs = h.Hs;
    const int tid = threadIdx.x;
    const int Wd = rup(ns, 32);
    const int S = NT / Wd;
    const int chunk = rup((ns + S - 1) / S, 2);
    const int s = tid / Wd, j = tid - s * Wd;
    if (s < S) {
        double acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0;
        const int m0 = s * chunk;
        const int m1 = (m0 + chunk < ns) ? m0 + chunk : ns;
        if (j < ns && m0 < m1) {
            const int jlo = j & ~31;
            const int jhi = (jlo + 31 < ns - 1) ? jlo + 31 : ns - 1;
            const int tj = tri(j);
            int m = m0;
            {
                const int mA = (jlo < m1) ? jlo : m1;
                const double* rowj = Hs + tj;
                for (; m + 7 < mA; m += 8) {
                    double hv[8], xv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { hv[e] = rowj[m + e]; xv[e] = x[m + e]; }
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] += hv[e] * xv[e];
                }
                for (; m < mA; ++m) acc[0] += rowj[m] * x[m];
            }
            int off = tri(m) + j;
            {
                const int mD = (jhi + 1 < m1) ? jhi + 1 : m1;
                for (; m < mD; ++m) {
                    acc[m & 7] += Hs[(m <= j) ? tj + m : off] * x[m];
                    off += m + 1;
                }
            }
            {
                for (; m + 7 < m1; m += 8) {
                    double hv[8], xv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { hv[e] = Hs[off]; xv[e] = x[m + e]; off += m + e + 1; }
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] += hv[e] * xv[e];
                }
                for (; m < m1; ++m) { acc[0] += Hs[off] * x[m]; off += m + 1; }
            }
        }
        h.buf[s * Wd + j] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    }
    __syncthreads();
    for (int o2 = tid; o2 < ns; o2 += NT) {
        double sum = 0.0;
        for (int g = 0; g < S; ++g) sum += h.buf[g * Wd + o2];
        y[o2] = sum;
    }
    __syncthreads();
}

// ---- symv v3: warp per 32-column strip x row range, every element of the strip's sub-diagonal part read ONCE:
// lane l owns column k = q0 + l: colacc += H[i][k] * x[i]  (private), and the row sums  sum_k H[i][k] x[k]  are formed
// by a butterfly over the lanes, 32 rows at a time (transpose-reduce: 31 shuffles per 32 rows instead of 5 per row).
template <int NT>
static __device__ __noinline__ void symv_v3(const HView h, int n, const double* x, double* y, double* part /* >= NW * Wd doubles */) {
    constexpr int NW = NT / 32;
    const int ns = n < h.R ? n : h.R;
    const double* Hs = h.Hs;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int Wd = rup(ns, 32);
    // per-warp partial results: part[w * Wd + i], zeroed first
    for (int t = l; t < Wd; t += 32) part[w * Wd + t] = 0.0;
    __syncwarp();
    // tasks: (column strip q, block of 32 rows r) with r >= q; round-robin over the warps
    const int nq = Wd >> 5;
    int tix = 0;
    for (int q = 0; q < nq; ++q) {
        const int k = 32 * q + l;
        const double xk = (k < ns) ? x[k] : 0.0;
        double colacc = 0.0;
        bool any = false;
        for (int r = q; r < nq; ++r, ++tix) {
            if ((tix & (NW - 1)) != w) continue;
            any = true;
            const int i0 = 32 * r;
            // lane l keeps rowv = H[i][k] * x[k] for the 32 rows of the block and reduces them by a butterfly transpose
            double rv[32];
            int off = tri(i0) + k;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int i = i0 + e;
                const bool ok = (i < ns) && (k <= i) && (k < ns);
                const double hv = ok ? Hs[off] : 0.0;
                off += i + 1;
                rv[e] = hv * xk;
                if (k != i) colacc += hv * ((i < ns) ? x[i] : 0.0);      // the diagonal element counts once (in the row sum)
            }
            // transpose-reduce: after the 5 steps lane l holds the sum over lanes of rv[l]
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
#pragma unroll
                for (int e = 0; e < st; ++e) {
                    const bool up = (l & st) != 0;
                    const double send = up ? rv[e] : rv[e + st];
                    const double keep = up ? rv[e + st] : rv[e];
                    rv[e] = keep + __shfl_xor_sync(0xffffffffu, send, st);
                }
            }
            part[w * Wd + i0 + l] += rv[0];
        }
        if (any) part[w * Wd + k] += colacc;
        __syncwarp();
    }
    __syncthreads();
    for (int o = threadIdx.x; o < ns; o += NT) {
        double sum = 0.0;
#pragma unroll
        for (int g = 0; g < NW; ++g) sum += part[g * Wd + o];
        y[o] = sum;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512, 1) k(int n, int reps, int variant, long long* cyc, double* out) {
    double* Hs = smem_d + 8192; double* x = smem_d; double* y = smem_d + 512; double* buf = smem_d + 1024;   // buf: 7168 doubles
    for (int t = threadIdx.x; t < n * (n + 1) / 2; t += 512) Hs[t] = 1e-3 * ((t * 7) % 17) - 4e-3;
    for (int t = threadIdx.x; t < 512; t += 512) x[t] = (t < n) ? 1.0 + 0.01 * t : 0.0;
    __syncthreads();
    HView h{Hs, nullptr, 1000, buf};
    long long t1 = 0, t2 = 0;
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        long long t0 = clock64();
        if (variant == 0) symv_leaf<512>(h, n, x, y);
        else if (variant == 1) symv_v2<512>(h, n, x, y);
        else symv_v3<512>(h, n, x, y, buf);
        long long tm = clock64();
        if (variant == 0) syr_leaf<512>(h, n, y, 1e-9);
        else syr_v2<512>(h, n, y, 1e-9);
        t1 += tm - t0; t2 += clock64() - tm;
    }
    if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t1; cyc[2 * blockIdx.x + 1] = t2; }
    if (blockIdx.x == 0) {
        for (int t = threadIdx.x; t < n; t += 512) out[t] = y[t];
        for (int t = threadIdx.x; t < n * (n + 1) / 2; t += 512) out[512 + t] = Hs[t];
    }
}
int main() {
    long long* cyc; double* out; cudaMalloc(&cyc, 148 * 16); cudaMalloc(&out, (512 + 32768) * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    static double ref[512 + 32768], got[512 + 32768];
    for (int n : {64, 100, 140, 170, 200}) {
        for (int variant = 0; variant < 3; ++variant) {
            const int reps = 20;
            k<<<148, 512, 225 * 1024>>>(n, 2, variant, cyc, out);
            k<<<148, 512, 225 * 1024>>>(n, reps, variant, cyc, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long hc[296]; cudaMemcpy(hc, cyc, 296 * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(variant == 0 ? ref : got, out, (512 + n * (n + 1) / 2) * 8, cudaMemcpyDeviceToHost);
            double dy = 0, dh = 0;
            if (variant) {
                for (int t = 0; t < n; ++t) dy = fmax(dy, fabs(got[t] - ref[t]) / fmax(1e-300, fabs(ref[t])));
                for (int t = 0; t < n * (n + 1) / 2; ++t) dh = fmax(dh, fabs(got[512 + t] - ref[512 + t]));
            }
            double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += hc[2 * i]; b += hc[2 * i + 1]; }
            a /= 148.0 * reps; b /= 148.0 * reps;
            printf("n=%3d variant %d: symv %6.0f cyc   syr %6.0f cyc   max rel dy %.2e  max dH %.2e  %s\n", n, variant, a, b, dy, dh,
                   e ? cudaGetErrorString(e) : "");
        }
    }
    return 0;
}
