// ssqp_kernel.cuh — device code of the batched status-switching active-set QP solver (sm_100a, FP64).
//
// One CTA solves one QP from start to finish (Phase 1 simplex start + Phase 2 active-set loop);
// CTAs are persistent and pull QP indices from a global atomic queue (trip counts are heavy-tailed).
//
// What replaces what (reference = PharosAbad/StatusSwitchingQP.jl v1.0.2):
//   phase1()      <- initQP (src/SSQP.jl:461-560) + cDantzigLP (src/Simplex.jl:445-615); the reference
//                    re-inverts the basis (inv(lu(A[:,B]))) and recomputes Y=invB*A[:,F] on every pivot;
//                    here invB gets a product-form rank-1 update and reduced costs are a GEMV with
//                    pi = invB' c_B.  Basis rows are kept unsorted; ties resolve on the variable id,
//                    which is what the reference's sorted B + first-extremum findmin/findmax gives.
//   phase2()      <- solveQP(Q,S,x0) main loop (src/SSQP.jl:269-376).  The reference refactorises
//                    inv(cholesky(V[F,F])) and the Schur complement every trip (:322-331); here the
//                    inverse of the reduced KKT matrix  [V_FF AE'; AE 0]  (free variables + active rows)
//                    is kept explicitly — its blocks are exactly the reference's VQ, TC and -C — as a
//                    packed symmetric matrix, and a status switch is a bordered rank-1 add / remove
//                    update (north-star piece 2).  p and the multipliers are one symmetric GEMV
//                    (piece 3); the ratio test (aStep!, :61-134) and the dual sign test (KKTchk!,
//                    :136-188) are CTA arg-min reductions on (key, insertion-rank) pairs (piece 4).
//   free_k()      <- freeK! (src/SSQP.jl:35-59);  polish() <- polishSz! (src/SSQP.jl:10-32)
//
// Data layout (device, all FP64 column-major):
//   V     N x N            shared by all QPs (or one per QP), L2 resident (2 MB at N=500)
//   Ccol  M0 x N           [A;G], column k = constraint column of variable k (contiguous)
//   Crow  N x M0           its transpose: constraint row r contiguous over variables
//   cA    N                column norms of [A;G] (Simplex.jl:463-465), computed once per set_shared
//   per-QP q,d,u (N), b (M), g (J); outputs x (N), S (N+J) int32, status int64
//   per-CTA workspace in global memory (L2): packed lower-triangular-by-rows symmetric inverse
//   (row i at offset i(i+1)/2), aliased with the Phase-1 basis inverse invB (M0 x M0, column-major).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssqp {

constexpr int NT = 256;          // threads per CTA
constexpr int NWARP = NT / 32;

enum : int { S_IN = 0, S_DN = 1, S_UP = 2, S_OE = 3, S_EO = 4 };
constexpr int NSTATS = 16;
enum : int { ST_TRIPS = 0, ST_FALG, ST_MAXK, ST_MAXW, ST_LOOPS, ST_PIVOTS, ST_UPDATES, ST_REBUILDS, ST_MAXRES,
             ST_REFINES, ST_BYTES, ST_DEGEN, ST_CYC_P1, ST_CYC_VPASS, ST_CYC_CPASS, ST_CYC_SYMV };
// ST_CYC_*: SM cycles spent in Phase 1 / gradient passes / constraint passes / packed-inverse passes (symv+syr)

struct KParams {
    int N, M, J, M0, nmax;
    const double* V; long long strideV;
    const double* Ccol; const double* Crow; const double* cA;
    const double *q, *b, *g, *d, *u;
    const int* S0; const double* x0;
    double* x; int* S; long long* status; double* stats;
    double* work; long long wstride;
    unsigned long long* queue;
    long long nb;
    int max_iter; double tol, tolG, tolLP;
    int phase1_only;
};

__host__ __device__ inline int rup(int a, int m) { return (a + m - 1) / m * m; }

// shared-memory carve-up (same arithmetic on host and device)
struct SmemLayout {
    int Np, nmp, M0p;
    // double offsets
    int z, gr, pfull, rhs, sol, hv, colv, slack, cp, bg, lam, pi, pcol, qB, rvec, sig, buf, red;
    int ndbl;
    // int offsets (after doubles)
    int item, pos, Sst, Bv, supp, flist, evl, redi, misc;
    int nint;
    __host__ __device__ SmemLayout(int N, int M0, int J) {
        Np = rup(N, 4); nmp = rup(N + M0, 4); M0p = rup(M0 > 0 ? M0 : 1, 32);
        int o = 0;
        z = o; o += Np; gr = o; o += Np; pfull = o; o += Np;
        rhs = o; o += nmp; sol = o; o += nmp; hv = o; o += nmp; colv = o; o += nmp;
        slack = o; o += M0p; cp = o; o += M0p; bg = o; o += M0p; lam = o; o += M0p;
        pi = o; o += M0p; pcol = o; o += M0p; qB = o; o += M0p; rvec = o; o += M0p; sig = o; o += M0p;
        int bsz = NWARP * nmp;
        int need2 = NT / 32 * M0p;     // partial buffer of the constraint passes
        if (need2 > bsz) bsz = need2;
        buf = o; o += bsz;
        red = o; o += 2 * NWARP + 8;
        ndbl = o;
        int p = 0;
        item = p; p += nmp; pos = p; p += rup(N + M0, 4); Sst = p; p += rup(N + J + M0, 4);
        Bv = p; p += M0p; supp = p; p += Np; flist = p; p += Np; evl = p; p += rup(N + M0, 4);
        redi = p; p += 2 * NWARP + 8; misc = p; p += 32;
        nint = p;
    }
    __host__ __device__ size_t bytes() const { return (size_t)ndbl * 8 + (size_t)nint * 4; }
};

#ifdef __CUDACC__

// Julia isless on Float64: -0.0 < 0.0, NaN sorts last  (sort!(..., by=x->x.L), src/SSQP.jl:94,176)
__device__ __forceinline__ bool jl_isless(double a, double b) {
    if (a != a) return false;
    if (b != b) return true;
    if (a < b) return true;
    if (a == 0.0 && b == 0.0) return (__double_as_longlong(a) < 0) && !(__double_as_longlong(b) < 0);
    return false;
}
// (ka,ia) strictly precedes (kb,ib); id<0 means "no candidate"
__device__ __forceinline__ bool precedes(double ka, int ia, double kb, int ib) {
    if (ia < 0) return false;
    if (ib < 0) return true;
    if (jl_isless(ka, kb)) return true;
    if (jl_isless(kb, ka)) return false;
    return ia < ib;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct Ctx {
    const KParams* P;
    int N, M, J, M0, M0p;
    const double *V, *Ccol, *Crow, *cA, *q, *d, *u;
    double *z, *gr, *pfull, *rhs, *sol, *hv, *colv, *slack, *cp, *bg, *lam, *pi, *pcol, *qB, *rvec, *sig, *buf, *red;
    int *item, *pos, *Sst, *Bv, *supp, *flist, *evl, *redi, *misc;
    double* Kinv;        // packed symmetric inverse of the reduced KKT matrix (global workspace)
    int n;               // current order of the reduced KKT system (K + W)
    double bytes;        // streamed bytes (thread 0 only)
    long long cyc_v, cyc_c, cyc_k;   // cycle counters (thread 0 only)
};

// ---- block-wide deterministic reductions (all threads must call) ----------------------------------
static __device__ double block_sum(Ctx& c, double v) {
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
    if (l == 0) c.red[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NWARP; ++i) s += c.red[i];
    return s;
}
static __device__ double block_max(Ctx& c, double v) {
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    v = warp_max(v);
    __syncthreads();
    if (l == 0) c.red[w] = v;
    __syncthreads();
    double s = c.red[0];
#pragma unroll
    for (int i = 1; i < NWARP; ++i) s = fmax(s, c.red[i]);
    return s;
}
// arg-min under `precedes`; result broadcast to all threads
static __device__ void block_argmin(Ctx& c, double& key, int& id) {
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double k2 = __shfl_xor_sync(0xffffffffu, key, o);
        int i2 = __shfl_xor_sync(0xffffffffu, id, o);
        if (precedes(k2, i2, key, id)) { key = k2; id = i2; }
    }
    __syncthreads();
    if (l == 0) { c.red[w] = key; c.redi[w] = id; }
    __syncthreads();
    key = c.red[0]; id = c.redi[0];
#pragma unroll
    for (int i = 1; i < NWARP; ++i)
        if (precedes(c.red[i], c.redi[i], key, id)) { key = c.red[i]; id = c.redi[i]; }
}

// ordered compaction of {k in [0,cnt) : pred(k)} into out[]; returns the count (all threads)
template <class Pred>
static __device__ int block_compact(Ctx& c, int cnt, int* out, Pred pred) {
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    int base = 0;
    for (int s = 0; s < cnt; s += NT) {
        int k = s + threadIdx.x;
        bool p = (k < cnt) && pred(k);
        unsigned m = __ballot_sync(0xffffffffu, p);
        __syncthreads();
        if (l == 0) c.redi[w] = __popc(m);
        __syncthreads();
        int off = base, tot = base;
#pragma unroll
        for (int i = 0; i < NWARP; ++i) {
            int v = c.redi[i];
            if (i < w) off += v;
            tot += v;
        }
        if (p) out[off + __popc(m & ((1u << l) - 1u))] = k;
        base = tot;
    }
    __syncthreads();
    return base;
}

// ---- streaming passes -----------------------------------------------------------------------------
// out[r] = sum_t Ccol[r + list[t]*M0] * w[list[t]]   for r < M0   (constraint pass over a variable list)
static __device__ void cpass(Ctx& c, const int* list, int cnt, const double* w, double* out) {
    const long long t0_ = clock64();
    const int M0 = c.M0;
    if (M0 == 0) return;
    const int RW = c.M0p < NT ? c.M0p : NT;
    const int G = NT / RW;
    const int rl = threadIdx.x % RW, part = threadIdx.x / RW;
    if (part < G) {
        for (int r = rl; r < M0; r += RW) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int t = part;
            for (; t + 3 * G < cnt; t += 4 * G) {
                int k0 = list[t], k1 = list[t + G], k2 = list[t + 2 * G], k3 = list[t + 3 * G];
                double v0 = c.Ccol[r + (size_t)k0 * M0], v1 = c.Ccol[r + (size_t)k1 * M0];
                double v2 = c.Ccol[r + (size_t)k2 * M0], v3 = c.Ccol[r + (size_t)k3 * M0];
                a0 += v0 * w[k0]; a1 += v1 * w[k1]; a2 += v2 * w[k2]; a3 += v3 * w[k3];
            }
            for (; t < cnt; t += G) { int k0 = list[t]; a0 += c.Ccol[r + (size_t)k0 * M0] * w[k0]; }
            c.buf[part * c.M0p + r] = (a0 + a1) + (a2 + a3);
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < M0; r += NT) {
        double s = 0.0;
        for (int g = 0; g < G; ++g) s += c.buf[g * c.M0p + r];
        out[r] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) { c.bytes += 8.0 * M0 * cnt; c.cyc_c += clock64() - t0_; }
}

// gr[i] = q[i] + sum_t V[i + list[t]*N] * z[list[t]]    (gradient at z over the support of z)
static __device__ void vpass(Ctx& c, const int* list, int cnt) {
    const long long t0_ = clock64();
    const int N = c.N;
    for (int i = threadIdx.x; i < N; i += NT) {
        double a0 = c.q[i], a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int t = 0;
        for (; t + 7 < cnt; t += 8) {
            int k0 = list[t], k1 = list[t + 1], k2 = list[t + 2], k3 = list[t + 3];
            int k4 = list[t + 4], k5 = list[t + 5], k6 = list[t + 6], k7 = list[t + 7];
            double v0 = c.V[i + (size_t)k0 * N], v1 = c.V[i + (size_t)k1 * N];
            double v2 = c.V[i + (size_t)k2 * N], v3 = c.V[i + (size_t)k3 * N];
            double v4 = c.V[i + (size_t)k4 * N], v5 = c.V[i + (size_t)k5 * N];
            double v6 = c.V[i + (size_t)k6 * N], v7 = c.V[i + (size_t)k7 * N];
            a0 += v0 * c.z[k0]; a1 += v1 * c.z[k1]; a2 += v2 * c.z[k2]; a3 += v3 * c.z[k3];
            a0 += v4 * c.z[k4]; a1 += v5 * c.z[k5]; a2 += v6 * c.z[k6]; a3 += v7 * c.z[k7];
        }
        for (; t < cnt; ++t) { int k0 = list[t]; a0 += c.V[i + (size_t)k0 * N] * c.z[k0]; }
        c.gr[i] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (threadIdx.x == 0) { c.bytes += 8.0 * N * cnt; c.cyc_v += clock64() - t0_; }
}

__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

// y = S x for the packed symmetric S (order n) in global memory; each element is read once.
// Warp per row; per-lane column accumulators; deterministic cross-warp reduction through c.buf.
template <int CMAX>
static __device__ void symv(Ctx& c, const double* __restrict__ S, int n, const double* x, double* y) {
    const long long t0_ = clock64();
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int ld = rup(c.N + c.M0, 4);
    double cacc[CMAX];
#pragma unroll
    for (int t = 0; t < CMAX; ++t) cacc[t] = 0.0;
    for (int i = w; i < n; i += NWARP) {
        const double* row = S + tri(i);
        const double xi = x[i];
        double a[CMAX];
#pragma unroll
        for (int t = 0; t < CMAX; ++t) {
            int k = l + 32 * t;
            a[t] = (32 * t <= i && k <= i) ? row[k] : 0.0;
        }
        double racc = 0.0;
#pragma unroll
        for (int t = 0; t < CMAX; ++t) {
            int k = l + 32 * t;
            if (32 * t <= i) {
                if (k <= i) racc += a[t] * x[k];
                if (k < i) cacc[t] += a[t] * xi;
            }
        }
        racc = warp_sum(racc);
        if (l == 0) y[i] = racc;
    }
#pragma unroll
    for (int t = 0; t < CMAX; ++t) {
        int k = l + 32 * t;
        if (k < n) c.buf[w * ld + k] = cacc[t];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += NT) {
        double s = y[k];
#pragma unroll
        for (int g = 0; g < NWARP; ++g) s += c.buf[g * ld + k];
        y[k] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) { c.bytes += 4.0 * n * (n + 1); c.cyc_k += clock64() - t0_; }
}

// S += sigma * v v'   on the packed lower triangle (order n)
template <int CMAX>
static __device__ void syr(Ctx& c, double* __restrict__ S, int n, const double* v, double sigma) {
    const long long t0_ = clock64();
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = w; i < n; i += NWARP) {
        double* row = S + tri(i);
        const double ci = sigma * v[i];
        double a[CMAX];
#pragma unroll
        for (int t = 0; t < CMAX; ++t) {
            int k = l + 32 * t;
            a[t] = (32 * t <= i && k <= i) ? row[k] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < CMAX; ++t) {
            int k = l + 32 * t;
            if (32 * t <= i && k <= i) row[k] = a[t] + ci * v[k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { c.bytes += 8.0 * n * (n + 1); c.cyc_k += clock64() - t0_; }
}

// ---- reduced-KKT inverse maintenance ---------------------------------------------------------------
// entry (row item `a`, col item `b`) of the reduced KKT matrix; ids < N are variables, N + r constraint rows
__device__ __forceinline__ double kkt_entry(const Ctx& c, int a, int b) {
    const int N = c.N;
    if (a < N) return (b < N) ? c.V[a + (size_t)b * N] : c.Crow[a + (size_t)(b - N) * N];
    return (b < N) ? c.Crow[b + (size_t)(a - N) * N] : 0.0;
}

// Bordered add of item `it` (variable k, or N + row).  Returns 0 ok, 1 dependent/singular pivot (nothing changed).
template <int CMAX>
static __device__ int kinv_add(Ctx& c, int it) {
    const int n = c.n;
    const double diag = (it < c.N) ? c.V[it + (size_t)it * c.N] : 0.0;
    if (n == 0) {
        if (!(fabs(diag) > 0.0)) return 1;
        if (threadIdx.x == 0) { c.Kinv[0] = 1.0 / diag; c.item[0] = it; c.pos[it] = 0; }
        c.n = 1;
        __syncthreads();
        return 0;
    }
    for (int p = threadIdx.x; p < n; p += NT) c.colv[p] = kkt_entry(c, c.item[p], it);
    __syncthreads();
    symv<CMAX>(c, c.Kinv, n, c.colv, c.hv);
    double part = 0.0, apart = 0.0;
    for (int p = threadIdx.x; p < n; p += NT) { double t = c.colv[p] * c.hv[p]; part += t; apart += fabs(t); }
    const double dot = block_sum(c, part);
    const double sref = block_sum(c, apart) + fabs(diag);
    const double s = diag - dot;
    if (!(fabs(s) > 1e-12 * sref)) return 1;
    const double is = 1.0 / s;
    syr<CMAX>(c, c.Kinv, n, c.hv, is);
    double* row = c.Kinv + tri(n);
    for (int p = threadIdx.x; p < n; p += NT) row[p] = -c.hv[p] * is;
    if (threadIdx.x == 0) { row[n] = is; c.item[n] = it; c.pos[it] = n; }
    c.n = n + 1;
    __syncthreads();
    return 0;
}

// Remove item `it` from the system.  Returns 0 ok, 1 singular (nothing changed).
template <int CMAX>
static __device__ int kinv_remove(Ctx& c, int it) {
    const int n = c.n;
    const int j = c.pos[it];
    for (int p = threadIdx.x; p < n; p += NT)
        c.colv[p] = (p <= j) ? c.Kinv[tri(j) + p] : c.Kinv[tri(p) + j];
    __syncthreads();
    const double piv = c.colv[j];
    double apart = 0.0;
    for (int p = threadIdx.x; p < n; p += NT) apart = fmax(apart, fabs(c.colv[p]));
    const double cmax = block_max(c, apart);
    if (!(fabs(piv) > 1e-13 * cmax) || !(fabs(piv) > 0.0)) return 1;
    syr<CMAX>(c, c.Kinv, n, c.colv, -1.0 / piv);
    const int last = n - 1;
    if (j != last) {
        const double* lrow = c.Kinv + tri(last);
        for (int k = threadIdx.x; k < last; k += NT) {
            if (k < j) c.Kinv[tri(j) + k] = lrow[k];
            else if (k > j) c.Kinv[tri(k) + j] = lrow[k];
            else c.Kinv[tri(j) + j] = lrow[last];
        }
        __syncthreads();
        if (threadIdx.x == 0) { int li = c.item[last]; c.item[j] = li; c.pos[li] = j; }
    }
    if (threadIdx.x == 0) c.pos[it] = -1;
    c.n = last;
    __syncthreads();
    return 0;
}

// From-scratch build of the inverse for the current status vector: border in the free variables in
// ascending order (V_FF is positive definite -> every pivot > 0), then the equality rows, then the EO
// rows ascending; a row whose pivot vanishes is dependent on the earlier rows and is left out of the
// system (the role getRowsGJr plays in the reference, src/SSQP.jl:310-319).  Returns #dropped rows, or -1.
template <int CMAX>
static __device__ int kinv_rebuild(Ctx& c) {
    const int N = c.N, M = c.M, M0 = c.M0;
    for (int i = threadIdx.x; i < N + M0; i += NT) c.pos[i] = -1;
    c.n = 0;
    __syncthreads();
    for (int k = 0; k < N; ++k)
        if (c.Sst[k] == S_IN)
            if (kinv_add<CMAX>(c, k)) return -1;       // V_FF not positive definite (PosDefException)
    int dropped = 0;
    for (int r = 0; r < M0; ++r)
        if (r < M || c.Sst[N + r - M] == S_EO)
            dropped += kinv_add<CMAX>(c, N + r);
    return dropped;
}

#ifndef SSQP_NO_SOLVE_KERNEL
// ---- Phase 1: initQP + cDantzigLP ------------------------------------------------------------------
// returns 1 feasible, 0 infeasible, -1 numerical; fills c.z (x0) and c.Sst[0..N+J)
static __device__ int phase1(Ctx& c, double* stats) {
    const int N = c.N, M = c.M, J = c.J, M0 = c.M0;
    const int N0 = N + J, N1 = N0 + M0;
    const double tol = c.P->tolLP;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double* invB = c.Kinv;      // M0 x M0 column-major (aliases the Phase-2 workspace)
    int* S1 = c.Sst;            // N1 statuses: structurals, slacks, artificials

    for (int k = threadIdx.x; k < N1; k += NT) S1[k] = (k >= N0) ? S_IN : S_DN;
    for (int j = threadIdx.x; j < M0; j += NT) c.Bv[j] = N0 + j;
    for (int k = threadIdx.x; k < N; k += NT) c.z[k] = c.d[k];
    __syncthreads();
    if (M0 == 0) return 1;
    // q0 = A0*d0 ; sig ; qB = |q0 - b0|                                  (src/SSQP.jl:516-521)
    int cnt = block_compact(c, N, c.supp, [&](int k) { return c.z[k] != 0.0; });
    cpass(c, c.supp, cnt, c.z, c.rvec);
    for (int j = threadIdx.x; j < M0; j += NT) {
        double q0 = c.rvec[j];
        c.sig[j] = (c.bg[j] >= q0) ? 1.0 : -1.0;
        c.qB[j] = fabs(q0 - c.bg[j]);
    }
    for (int t = threadIdx.x; t < M0 * M0; t += NT) invB[t] = 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < M0; j += NT) invB[j + (size_t)j * M0] = c.sig[j];
    __syncthreads();

    long long loop = 0, pivots = 0;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    while (true) {
        // pi = invB' c_B : sum of the rows of invB whose basic variable is an artificial   (Simplex.jl:600)
        for (int i = w; i < M0; i += NWARP) {
            double s = 0.0;
            for (int j = l; j < M0; j += 32)
                if (c.Bv[j] >= N0) s += invB[j + (size_t)i * M0];
            s = warp_sum(s);
            if (l == 0) c.pi[i] = s;
        }
        __syncthreads();
        const bool bland = (loop + 1) > N1;            // loop += 1; if loop > N: Bland  (Simplex.jl:487-490)
        // pricing: h > tol candidates; largest-distance Dantzig  argmax(hp ./ cA)  (Simplex.jl:495)
        double bkey = 0.0; int bid = -1;
        for (int k = threadIdx.x; k < N1; k += NT) {
            const int st = S1[k];
            if (st == S_IN) continue;
            double h, ca = 1.0;
            if (k < N) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int i = 0;
                for (; i + 3 < M0; i += 4) {
                    double v0 = c.Crow[k + (size_t)i * N], v1 = c.Crow[k + (size_t)(i + 1) * N];
                    double v2 = c.Crow[k + (size_t)(i + 2) * N], v3 = c.Crow[k + (size_t)(i + 3) * N];
                    a0 += v0 * c.pi[i]; a1 += v1 * c.pi[i + 1]; a2 += v2 * c.pi[i + 2]; a3 += v3 * c.pi[i + 3];
                }
                for (; i < M0; ++i) a0 += c.Crow[k + (size_t)i * N] * c.pi[i];
                double rc = -((a0 + a1) + (a2 + a3));
                h = (st == S_DN) ? -rc : rc;
                ca = c.cA[k];
            } else if (k < N0) {
                double rc = -c.pi[M + (k - N)];
                h = (st == S_DN) ? -rc : rc;
            } else {
                double rc = 1.0 - c.sig[k - N0] * c.pi[k - N0];
                h = (st == S_DN) ? -rc : rc;
            }
            if (h > tol) {
                double key = bland ? 0.0 : -(h / ca);           // arg-max == arg-min of the negated score
                if (precedes(key, k, bkey, bid)) { bkey = key; bid = k; }
            }
        }
        if (threadIdx.x == 0) c.bytes += 8.0 * N * M0;
        block_argmin(c, bkey, bid);
        if (bid < 0) break;
        loop += 1;
        const int kin = bid;
        // p = invB * A1[:,kin]                                                            (Simplex.jl:497)
        for (int j = threadIdx.x; j < M0; j += NT) {
            double s = 0.0;
            if (kin < N) {
                const double* col = c.Ccol + (size_t)kin * M0;
                for (int i = 0; i < M0; ++i) s += invB[j + (size_t)i * M0] * col[i];
            } else if (kin < N0) {
                s = invB[j + (size_t)(M + kin - N) * M0];
            } else {
                s = c.sig[kin - N0] * invB[j + (size_t)(kin - N0) * M0];
            }
            c.pcol[j] = s;
        }
        __syncthreads();
        // ratio test (Simplex.jl:499-569): arg-min/arg-max over basis rows, ties -> lowest basic variable id
        const bool kd = (S1[kin] == S_DN);
        const double lo_k = (kin < N) ? c.d[kin] : 0.0;
        const double hi_k = (kin < N) ? c.u[kin] : INF;
        const bool fu = hi_k < INF;
        double rkey = 0.0; int rid = -1;
        for (int j = threadIdx.x; j < M0; j += NT) {
            const int i = c.Bv[j];
            const double pj = c.pcol[j];
            const double lo = (i < N) ? c.d[i] : 0.0;
            const double hi = (i < N) ? c.u[i] : INF;
            double gt; bool has = false;
            if (kd) {
                if (pj > tol) { gt = (c.qB[j] - lo) / pj; has = true; }
                else if (pj < -tol) { gt = (c.qB[j] - hi) / pj; has = true; }
            } else {
                if (pj > tol) { gt = (c.qB[j] - hi) / pj; has = true; }
                else if (pj < -tol) { gt = (c.qB[j] - lo) / pj; has = true; }
            }
            if (has) {
                double key = kd ? gt : -gt;
                if (precedes(key, i, rkey, rid)) { rkey = key; rid = i; }
            }
        }
        block_argmin(c, rkey, rid);
        int action;      // -1 flip to UP, -2 flip to DN, >=0 pivot on the row of basic variable rid
        if (kd) {
            if (rid < 0) {
                if (fu) action = -1; else return -1;                         // unbounded (cannot happen in Phase 1)
            } else {
                const double gl = rkey;
                if (fu) action = (gl >= hi_k - lo_k) ? -1 : 0;
                else { if (isinf(gl)) return -1; action = 0; }
            }
        } else {
            if (rid < 0) action = -2;
            else { const double gl = -rkey; action = (gl <= -(hi_k - lo_k)) ? -2 : 0; }
        }
        if (action == -1) { if (threadIdx.x == 0) S1[kin] = S_UP; }
        else if (action == -2) { if (threadIdx.x == 0) S1[kin] = S_DN; }
        else {
            // find the row of the leaving variable
            int lrow = -1;
            for (int j = threadIdx.x; j < M0; j += NT) if (c.Bv[j] == rid) lrow = j;
            if (lrow >= 0) c.misc[0] = lrow;
            __syncthreads();
            lrow = c.misc[0];
            const double pj = c.pcol[lrow];
            int Sl;
            if (kd) Sl = (pj > tol) ? S_DN : S_UP; else Sl = (pj > tol) ? S_UP : S_DN;
            // product-form update: row l /= p_l ; row j -= p_j * row l
            const double ipl = 1.0 / pj;
            // two-step to avoid the read/write race on row lrow: first stash the scaled pivot row
            for (int i = threadIdx.x; i < M0; i += NT) c.rvec[i] = invB[lrow + (size_t)i * M0] * ipl;
            __syncthreads();
            for (int t = threadIdx.x; t < M0 * M0; t += NT) {
                const int jj = t % M0, ii = t / M0;
                invB[t] = (jj == lrow) ? c.rvec[ii] : invB[t] - c.pcol[jj] * c.rvec[ii];
            }
            if (threadIdx.x == 0) {
                c.Bv[lrow] = kin; S1[kin] = S_IN; S1[rid] = Sl;
                c.bytes += 16.0 * M0 * M0;
            }
            pivots += 1;
        }
        __syncthreads();
        // q = invB * (b - sum_{nonbasic, x != 0} A1[:,k] x_k)    (fresh every loop, Simplex.jl:599)
        for (int k = threadIdx.x; k < N; k += NT) {
            const int st = S1[k];
            c.z[k] = (st == S_IN) ? 0.0 : ((st == S_UP) ? c.u[k] : c.d[k]);
        }
        __syncthreads();
        cnt = block_compact(c, N, c.supp, [&](int k) { return c.z[k] != 0.0; });
        cpass(c, c.supp, cnt, c.z, c.rvec);
        for (int j = threadIdx.x; j < M0; j += NT) c.rvec[j] = c.bg[j] - c.rvec[j];
        __syncthreads();
        for (int j = threadIdx.x; j < M0; j += NT) {
            double s = 0.0;
            for (int i = 0; i < M0; ++i) s += invB[j + (size_t)i * M0] * c.rvec[i];
            c.qB[j] = s;
        }
        __syncthreads();
    }
    // x[B] = q ; f = sum(artificials) ; status mapping                     (Simplex.jl:610, SSQP.jl:531-542)
    for (int k = threadIdx.x; k < N; k += NT) {
        const int st = S1[k];
        c.z[k] = (st == S_UP) ? c.u[k] : c.d[k];
    }
    __syncthreads();
    double fpart = 0.0;
    for (int j = threadIdx.x; j < M0; j += NT) {
        const int i = c.Bv[j];
        if (i < N) c.z[i] = c.qB[j];
        else if (i >= N0) fpart += c.qB[j];
    }
    const double f = block_sum(c, fpart);
    if (threadIdx.x == 0) { stats[ST_LOOPS] = (double)loop; stats[ST_PIVOTS] = (double)pivots; }
    __syncthreads();
    if (f > tol) return 0;
    for (int k = N + threadIdx.x; k < N0; k += NT) S1[k] = (S1[k] == S_IN) ? S_OE : S_EO;
    __syncthreads();
    return 1;
}

// ---- Phase 2 ---------------------------------------------------------------------------------------
template <int CMAX>
static __device__ long long phase2(Ctx& c, double* stats) {
    const int N = c.N, M = c.M, J = c.J, M0 = c.M0;
    const double tol = c.P->tol, tolG = c.P->tolG;
    const int maxIter = c.P->max_iter;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int* S = c.Sst;
    long long iter = 0;
    bool have_sys = false;
    int ndropped = 0;
    bool gr_valid = false;
    double falg = 0.0, maxres = 0.0;
    long long updates = 0, rebuilds = 0, degen = 0;
    int maxK = 0, maxW = 0;

    auto finish = [&](long long st) {
        if (threadIdx.x == 0) {
            stats[ST_TRIPS] = (double)(st > 0 ? st : iter);
            stats[ST_FALG] += falg; stats[ST_MAXK] = maxK; stats[ST_MAXW] = maxW;
            stats[ST_UPDATES] = (double)updates; stats[ST_REBUILDS] = (double)rebuilds;
            stats[ST_MAXRES] = maxres; stats[ST_DEGEN] = (double)degen;
        }
        return st;
    };

    while (true) {
        iter += 1;
        if (iter > maxIter) return finish(-iter);

        // K = |{S == IN}|
        int kpart = 0;
        for (int k = threadIdx.x; k < N; k += NT) kpart += (S[k] == S_IN);
        const int K = (int)(block_sum(c, (double)kpart) + 0.5);

        // gradient at z (fresh every time z changed): gr = V z + q
        if (!gr_valid) {
            int cnt = block_compact(c, N, c.supp, [&](int k) { return c.z[k] != 0.0; });
            vpass(c, c.supp, cnt);
            gr_valid = true;
        }

        if (K == 0) {   // freeK!  (src/SSQP.jl:35-59)
            falg += 2.0 * N * N;
            int any = 0;
            for (int k = threadIdx.x; k < N; k += NT) {
                const double p = c.gr[k];
                const int st = S[k];
                c.evl[k] = st;          // S0 = copy(S)
                if ((p >= -tol && st == S_UP) || (p <= tol && st == S_DN)) { S[k] = S_IN; any = 1; }
            }
            any = (block_sum(c, (double)any) > 0.0);
            if (!any) return finish(iter);
            double pm = 0.0;
            for (int k = threadIdx.x; k < N; k += NT) if (S[k] == S_IN) pm = fmax(pm, fabs(c.gr[k]));
            pm = block_max(c, pm);
            if (pm <= tol) {
                for (int k = threadIdx.x; k < N; k += NT) if (S[k] == S_IN) S[k] = c.evl[k];
                __syncthreads();
                return finish(iter);
            }
            have_sys = false;
            __syncthreads();
            continue;
        }

        if (!have_sys || ndropped > 0) {
            ndropped = kinv_rebuild<CMAX>(c);
            rebuilds += 1;
            if (ndropped < 0) return finish(-1);
            if (ndropped > 0) degen += 1;
            have_sys = true;
        }
        const int n = c.n;
        const int W = n - K;
        maxK = max(maxK, K); maxW = max(maxW, W);
        {
            int jo = 0;
            for (int j = threadIdx.x; j < J; j += NT) jo += (S[N + j] == S_OE);
            const double JO = block_sum(c, (double)jo);
            const double k = K, w = W, nn = N;
            falg += k * k * k / 3 + k * k * w + k * w * w + w * w * w / 3 + 2 * k * k + 4 * k * w + 2 * w * w +
                    2 * nn * nn + 2 * (nn - k) * w + 2 * JO * (nn + k);
        }

        // slack = [b;g] - [A;G] z   (bE and zo of the reference, src/SSQP.jl:295 and :79)
        {
            int cnt = block_compact(c, N, c.supp, [&](int k) { return c.z[k] != 0.0; });
            cpass(c, c.supp, cnt, c.z, c.slack);
            for (int r = threadIdx.x; r < M0; r += NT) c.slack[r] = c.bg[r] - c.slack[r];
            __syncthreads();
        }
        // reduced KKT solve:  [V_FF AE'; AE 0] [p; lam] = [-gr_F; slack_E]   ->  alpha = z_F + p
        for (int p = threadIdx.x; p < n; p += NT) {
            const int it = c.item[p];
            c.rhs[p] = (it < N) ? -c.gr[it] : c.slack[it - N];
        }
        __syncthreads();
        symv<CMAX>(c, c.Kinv, n, c.rhs, c.sol);
        double pm = 0.0;
        for (int p = threadIdx.x; p < n; p += NT) {
            const int it = c.item[p];
            if (it < N) { c.pfull[it] = c.sol[p]; pm = fmax(pm, fabs(c.sol[p])); }
        }
        for (int r = threadIdx.x; r < M0; r += NT) c.lam[r] = 0.0;
        __syncthreads();
        for (int p = threadIdx.x; p < n; p += NT) {
            const int it = c.item[p];
            if (it >= N) c.lam[it - N] = c.sol[p];
        }
        pm = block_max(c, pm);

        if (pm > tolG) {    // aStep!  (src/SSQP.jl:61-134)
            const int nf = block_compact(c, N, c.flist, [&](int k) { return S[k] == S_IN; });
            if (J > 0) cpass(c, c.flist, nf, c.pfull, c.cp);       // po = G[Og,F]*p (all rows computed)
            double bkey = 0.0; int bid = -1;
            for (int t = threadIdx.x; t < nf + J; t += NT) {
                double L; bool has = false; int id = -1;
                if (t < nf) {
                    const int j = c.flist[t];
                    const double tt = c.pfull[j], h = c.z[j];
                    const double dj = c.d[j], uj = c.u[j];
                    if (tt > tol && uj < INF) { L = (uj - h) / tt; has = true; id = j; }
                    else if (tt < -tol && dj > -INF) { L = (dj - h) / tt; has = true; id = j; }
                } else {
                    const int j = t - nf;
                    if (S[N + j] == S_OE) {
                        const double po = c.cp[M + j];
                        if (po > tol) { L = c.slack[M + j] / po; has = true; id = N + j; }
                    }
                }
                if (has && precedes(L, id, bkey, bid)) { bkey = L; bid = id; }
            }
            block_argmin(c, bkey, bid);
            const double L1 = (bid >= 0) ? bkey : 1.0;
            if (L1 < 1.0) {
                // collect every event with L - L1 <= tol (multi blocking), then step and switch statuses
                if (threadIdx.x == 0) c.misc[1] = 0;
                __syncthreads();
                for (int t = threadIdx.x; t < nf + J; t += NT) {
                    double L; bool has = false; int id = -1;
                    if (t < nf) {
                        const int j = c.flist[t];
                        const double tt = c.pfull[j], h = c.z[j];
                        const double dj = c.d[j], uj = c.u[j];
                        if (tt > tol && uj < INF) { L = (uj - h) / tt; has = true; id = j; }
                        else if (tt < -tol && dj > -INF) { L = (dj - h) / tt; has = true; id = -2 - j; }  // to DN
                    } else {
                        const int j = t - nf;
                        if (S[N + j] == S_OE) {
                            const double po = c.cp[M + j];
                            if (po > tol) { L = c.slack[M + j] / po; has = true; id = N + j; }
                        }
                    }
                    if (has && !(L - L1 > tol)) { int s = atomicAdd(&c.misc[1], 1); c.evl[s] = id; }
                }
                __syncthreads();
                const int nev = c.misc[1];
                if (threadIdx.x == 0) {       // deterministic order: ascending variable / row id
                    for (int a = 1; a < nev; ++a) {
                        int v = c.evl[a];
                        int kv = v < -1 ? -2 - v : v;
                        int b = a - 1;
                        while (b >= 0) {
                            int u2 = c.evl[b];
                            int ku = u2 < -1 ? -2 - u2 : u2;
                            if (ku <= kv) break;
                            c.evl[b + 1] = u2; --b;
                        }
                        c.evl[b + 1] = v;
                    }
                }
                for (int t = threadIdx.x; t < nf; t += NT) { const int j = c.flist[t]; c.z[j] += L1 * c.pfull[j]; }
                __syncthreads();
                for (int e = 0; e < nev; ++e) {
                    const int ev = c.evl[e];
                    int rc = 0;
                    if (ev < -1 || ev < N) {
                        const int k = ev < -1 ? -2 - ev : ev;
                        const int To = ev < -1 ? S_DN : S_UP;
                        if (threadIdx.x == 0) { S[k] = To; c.z[k] = (To == S_DN) ? c.d[k] : c.u[k]; }
                        __syncthreads();
                        if (c.pos[k] >= 0) rc = kinv_remove<CMAX>(c, k);
                    } else {
                        const int j = ev - N;
                        if (threadIdx.x == 0) S[N + j] = S_EO;
                        __syncthreads();
                        rc = kinv_add<CMAX>(c, N + M + j);
                    }
                    updates += 1;
                    if (rc) { ndropped = 1; }     // dependent working set: rebuild (with row purge) next trip
                }
                gr_valid = false;
                __syncthreads();
                continue;
            }
            // full step: z[F] = alpha
            for (int t = threadIdx.x; t < nf; t += NT) { const int j = c.flist[t]; c.z[j] += c.pfull[j]; }
            __syncthreads();
            int cnt = block_compact(c, N, c.supp, [&](int k) { return c.z[k] != 0.0; });
            vpass(c, c.supp, cnt);
        }
        // KKTchk!  (src/SSQP.jl:136-188): gamma = (V z + q)_B + AB' alphaL ; release the most negative
        const int nrow = block_compact(c, M0, c.evl, [&](int r) { return c.pos[N + r] >= 0; });
        double bkey = 0.0; int bid = -1;
        double res = 0.0;
        for (int k = threadIdx.x; k < N; k += NT) {
            double a0 = c.gr[k], a1 = 0.0;
            int t = 0;
            for (; t + 1 < nrow; t += 2) {
                const int r0 = c.evl[t], r1 = c.evl[t + 1];
                a0 += c.Crow[k + (size_t)r0 * N] * c.lam[r0];
                a1 += c.Crow[k + (size_t)r1 * N] * c.lam[r1];
            }
            if (t < nrow) { const int r0 = c.evl[t]; a0 += c.Crow[k + (size_t)r0 * N] * c.lam[r0]; }
            const double gam = a0 + a1;
            const int st = S[k];
            if (st == S_IN) res = fmax(res, fabs(gam));          // stationarity residual of the free set
            else if (st == S_UP && gam > tolG) { if (precedes(-gam, k, bkey, bid)) { bkey = -gam; bid = k; } }
            else if (st == S_DN && gam < -tolG) { if (precedes(gam, k, bkey, bid)) { bkey = gam; bid = k; } }
        }
        for (int j = threadIdx.x; j < J; j += NT) {
            if (S[N + j] == S_EO && c.pos[N + M + j] >= 0) {
                const double t = c.lam[M + j];
                if (t < -tolG && precedes(t, N + j, bkey, bid)) { bkey = t; bid = N + j; }
            }
        }
        if (threadIdx.x == 0) c.bytes += 8.0 * N * W;
        res = block_max(c, res);
        maxres = fmax(maxres, res);
        block_argmin(c, bkey, bid);
        if (bid >= 0) {
            int rc;
            if (bid < N) {
                if (threadIdx.x == 0) S[bid] = S_IN;
                __syncthreads();
                rc = kinv_add<CMAX>(c, bid);
            } else {
                if (threadIdx.x == 0) S[bid] = S_OE;
                __syncthreads();
                rc = kinv_remove<CMAX>(c, N + M + (bid - N));
            }
            updates += 1;
            if (rc) ndropped = 1;
            __syncthreads();
            continue;
        }
        // optimal: polishSz!  (src/SSQP.jl:10-32)
        for (int k = threadIdx.x; k < N; k += NT) {
            const int st = S[k];
            const double dk = c.d[k], uk = c.u[k];
            if (st == S_DN) c.z[k] = dk;
            else if (st == S_UP) c.z[k] = uk;
            else {
                if (fabs(c.z[k] - dk) < tol) { c.z[k] = dk; S[k] = S_DN; }
                else if (fabs(c.z[k] - uk) < tol) { c.z[k] = uk; S[k] = S_UP; }
            }
        }
        __syncthreads();
        if (J > 0) {
            int cnt = block_compact(c, N, c.supp, [&](int k) { return c.z[k] != 0.0; });
            cpass(c, c.supp, cnt, c.z, c.cp);
            for (int j = threadIdx.x; j < J; j += NT) S[N + j] = (fabs(c.bg[M + j] - c.cp[M + j]) < tol) ? S_EO : S_OE;
            __syncthreads();
        }
        return finish(iter);
    }
}

template <int CMAX>
__global__ void __launch_bounds__(NT, 2) ssqp_solve_kernel(const KParams P) {
    extern __shared__ double smem_d[];
    __shared__ long long s_qp;
    const SmemLayout L(P.N, P.M0, P.J);
    Ctx c;
    c.P = &P;
    c.N = P.N; c.M = P.M; c.J = P.J; c.M0 = P.M0; c.M0p = L.M0p;
    c.Ccol = P.Ccol; c.Crow = P.Crow; c.cA = P.cA;
    double* sd = smem_d;
    c.z = sd + L.z; c.gr = sd + L.gr; c.pfull = sd + L.pfull; c.rhs = sd + L.rhs; c.sol = sd + L.sol;
    c.hv = sd + L.hv; c.colv = sd + L.colv; c.slack = sd + L.slack; c.cp = sd + L.cp; c.bg = sd + L.bg;
    c.lam = sd + L.lam; c.pi = sd + L.pi; c.pcol = sd + L.pcol; c.qB = sd + L.qB; c.rvec = sd + L.rvec;
    c.sig = sd + L.sig; c.buf = sd + L.buf; c.red = sd + L.red;
    int* si = reinterpret_cast<int*>(sd + L.ndbl);
    c.item = si + L.item; c.pos = si + L.pos; c.Sst = si + L.Sst; c.Bv = si + L.Bv; c.supp = si + L.supp;
    c.flist = si + L.flist; c.evl = si + L.evl; c.redi = si + L.redi; c.misc = si + L.misc;
    c.Kinv = P.work + (size_t)blockIdx.x * P.wstride;

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_qp = (long long)atomicAdd(P.queue, 1ULL);
        __syncthreads();
        const long long qp = s_qp;
        if (qp >= P.nb) break;
        const int N = P.N, M = P.M, J = P.J, M0 = P.M0;
        c.V = P.V + (size_t)qp * P.strideV;
        c.q = P.q ? P.q + (size_t)qp * N : nullptr;
        c.d = P.d + (size_t)qp * N; c.u = P.u + (size_t)qp * N;
        c.n = 0; c.bytes = 0.0; c.cyc_v = c.cyc_c = c.cyc_k = 0;
        const long long tq0 = clock64();
        double* stats = P.stats + (size_t)qp * NSTATS;
        for (int t = threadIdx.x; t < NSTATS; t += NT) stats[t] = 0.0;
        for (int r = threadIdx.x; r < M0; r += NT) c.bg[r] = (r < M) ? P.b[(size_t)qp * M + r] : P.g[(size_t)qp * J + (r - M)];
        // finite lower bounds only (the reference's (-Inf,u] handling is defective, src/SSQP.jl:551-557)
        int bad = 0;
        for (int k = threadIdx.x; k < N; k += NT) { const double dk = c.d[k]; if (!(dk > -1e300) || dk != dk) bad = 1; }
        __syncthreads();
        bad = (block_sum(c, (double)bad) > 0.0);
        long long status;
        if (bad) {
            for (int k = threadIdx.x; k < N; k += NT) { c.z[k] = 0.0; c.Sst[k] = S_DN; }
            for (int j = threadIdx.x; j < J; j += NT) c.Sst[N + j] = S_OE;
            status = -1;
        } else if (P.S0 != nullptr && P.x0 != nullptr) {
            for (int k = threadIdx.x; k < N; k += NT) c.z[k] = P.x0[(size_t)qp * N + k];
            for (int k = threadIdx.x; k < N + J; k += NT) c.Sst[k] = P.S0[(size_t)qp * (N + J) + k];
            status = 1;
        } else {
            status = phase1(c, stats);
        }
        const long long tq1 = clock64();
        __syncthreads();
        if (status > 0 && !P.phase1_only) status = phase2<CMAX>(c, stats);
        __syncthreads();
        for (int k = threadIdx.x; k < N; k += NT) P.x[(size_t)qp * N + k] = c.z[k];
        for (int k = threadIdx.x; k < N + J; k += NT) P.S[(size_t)qp * (N + J) + k] = c.Sst[k];
        if (threadIdx.x == 0) {
            P.status[qp] = status; stats[ST_BYTES] = c.bytes;
            stats[ST_CYC_P1] = (double)(tq1 - tq0); stats[ST_CYC_VPASS] = (double)c.cyc_v;
            stats[ST_CYC_CPASS] = (double)c.cyc_c; stats[ST_CYC_SYMV] = (double)c.cyc_k;
            stats[ST_REFINES] = (double)(clock64() - tq0);      // total cycles of this QP (slot reused until refinement lands)
        }
    }
}

#endif  // SSQP_NO_SOLVE_KERNEL

#ifdef SSQP_NO_SOLVE_KERNEL   // helper kernels live in the C-ABI translation unit only
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

#endif  // helper kernels

#endif  // __CUDACC__
}  // namespace ssqp
