// ssqp_kernel.cuh — device code of the batched status-switching active-set QP solver (sm_100a, FP64).
//
// One CTA solves one QP from start to finish (Phase 1 simplex start + Phase 2 active-set loop);
// CTAs are persistent and pull QP indices from a global atomic queue (trip counts are heavy-tailed).
//
// What replaces what (reference = PharosAbad/StatusSwitchingQP.jl v1.0.2):
//   phase1()      <- initQP (src/SSQP.jl:461-560) + cDantzigLP (src/Simplex.jl:445-615).  The reference
//                    re-inverts the basis (inv(lu(A[:,B]))) and recomputes Y=invB*A[:,F], h and q on every pivot;
//                    here invB lives in shared memory and gets a product-form rank-1 update, the duals pi and the
//                    basic values follow each pivot in O(M0), and the reduced costs are one streaming GEMV over
//                    [A;G]'.  Ties resolve on the variable id, which is what the reference's sorted B +
//                    first-extremum findmin/findmax gives.
//   lp_solve()    <- SimplexLP (src/Simplex.jl:831-1034): the same pivot loop run twice (Phase-1 costs, then the
//                    LP's costs from the Phase-1 basis).
//   phase2()      <- solveQP(Q,S,x0) main loop (src/SSQP.jl:269-376).  The reference refactorises
//                    inv(cholesky(V[F,F])) and the Schur complement every trip (:322-331); here the
//                    inverse H of the reduced KKT matrix  [V_FF AE'; AE 0]  (free variables + active rows;
//                    its blocks are the reference's VQ, TC and -C) is kept as a packed symmetric matrix in
//                    SHARED MEMORY (rows beyond the capacity spill to an L2-resident global tail), and a
//                    status switch is a bordered rank-1 add / remove update (north-star piece 2) that also
//                    carries the solution (p, lambda) of the reduced system along in O(n).  The gradient is
//                    recomputed fresh (pass over V) at every KKT check; a refinement solve (fresh slacks, one
//                    symmetric GEMV with H on the fresh residual; piece 3) runs every 8th check and always before
//                    optimality is declared.  The ratio test (aStep!, :61-134) and the dual sign test (KKTchk!,
//                    :136-188) are CTA arg-min reductions on (key, insertion-rank) pairs (piece 4).
//                    A cycle watch recognises the exact release / block alternation on which the reference runs to
//                    maxIter and skips the remaining trips in pairs (same status, S and z).
//   simplex_loop_alt() <- stpEdgeLP / maxImprvLP (src/Simplex.jl:234-416, 641-813), the other two values of Settings.rule
//                    (general kernel flavour only).
//   xform_begin/end  <- the free-variable split and the (-Inf,u] negation of initQP / SimplexLP (src/SSQP.jl:484-509).
//   drive_out_artificials() <- SimplexLP's re-selection of the basis when an artificial stays basic (src/Simplex.jl:962-977).
//   chains (KParams::chain_len) <- the user loop solveQP(Q, S, x0) along a sweep over q: one CTA, z and S stay on chip.
//   purge_rows_gjr() <- getRowsGJr (src/utils.jl:49-86), only for degenerate working sets (more active rows than free
//                    variables, or a border pivot below the soft threshold PIV_SOFT: the reference's own test is the arbiter).
//   freeK!  (src/SSQP.jl:35-59) and polishSz! (src/SSQP.jl:10-32) are inlined in phase2().
// Round 2:
//   kinv_build_chol() <- the reference's per-trip factorisation (src/SSQP.jl:322-331: cholesky(V_FF), Schur complement) done ONCE
//                    per rebuild, in place on the packed inverse (used while the matrix fits in shared memory; beyond that the
//                    sequential bordering of kinv_rebuild streams the L2-resident tail better).
//   drift guard      refinement corrections above 16 tolG (z) / 1e-6 (multipliers), or 4096 updates, rebuild the inverse and
//                    redo the trip without advancing the trip counter (phase2: DRIFT_TOL, `redo`).
//   ccache_*         constraint columns of the free variables cached in the unused END of the inverse's shared memory, filled
//                    by TMA bulk copies (cp.async.bulk + mbarrier): the ratio test's pass runs from shared memory when the
//                    cache holds the whole free list (cpass_free).
//   reuse            a warm-started QP of a chain starts from the inverse its predecessor left behind (same V, A, G, S).
//
// Data layout (device, all FP64 column-major):
//   V     N x N            shared by all QPs (or one per QP), L2 resident (2 MB at N=500)
//   Ccol  M0 x N           [A;G], column k = constraint column of variable k (contiguous)
//   Crow  N x M0           its transpose: constraint row r contiguous over variables
//   cA    N                column norms of [A;G] (Simplex.jl:463-465), computed once per set_shared
//   per-QP q,d,u (N), b (M), g (J): staged into shared memory once per QP; outputs x (N), S (N+J) int32, status int64
//   H     packed lower-triangular-by-rows (row i at offset i(i+1)/2): rows < hrows in shared memory,
//         rows >= hrows in the CTA's global workspace.  Phase 1 aliases the same storage with invB
//         (M0 x M0, odd leading dimension -> conflict-free row and column access; in the workspace when too large).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssqp {

enum : int { S_IN = 0, S_DN = 1, S_UP = 2, S_OE = 3, S_EO = 4 };
constexpr int NSTATS = 56;
constexpr int TP_STAGE = 16384;      // outputs per group of the from-scratch factorisation's two-phase transforms (doubles of
                                     // per-CTA staging in the workspace, L2-resident)
enum : int { ST_TRIPS = 0, ST_FALG, ST_MAXK, ST_MAXW, ST_LOOPS, ST_PIVOTS, ST_UPDATES, ST_REBUILDS, ST_MAXRES,
             ST_CYCLES, ST_BYTES, ST_DEGEN, ST_CYC_P1, ST_CYC0 /* 13.. : the NCYC section timers below */,
             ST_DRIFT = 53 /* rebuilds forced by the drift guard (refinement correction above 16 tolG, or REBUILD_EVERY updates) */,
             ST_LAMERR = 54 /* largest relative correction a refinement made to the carried multipliers */,
             ST_NKKT = 55 /* sign tests (KKTchk!) run */ };
// section timers (SM cycles, thread 0): gradient pass, constraint passes, symmetric GEMV, rank-1 update, sign-test
// pass, Phase-1 pricing pass, Phase-1 basis-inverse work, ratio test, event application, sign test; then call counts
enum : int { CY_VPASS = 0, CY_CPASS, CY_SYMV, CY_SYR, CY_GAMMA, CY_P1PRICE, CY_REBUILD /* from-scratch builds of the inverse */, CY_RATIO, CY_EVENTS, CY_KKT,
             CY_NSYMV, CY_NSYR, CY_GLOAD, CY_GEPI,
             // exclusive timeline sections of a Phase-2 trip (TICK): every cycle of Phase 2 lands in exactly one of them
             T_TOP = 16, T_CPASS, T_RATIO, T_COLLECT, T_STEP, T_RM_GATHER, T_RM_CHECK, T_RM_SYR, T_RM_TAIL, T_AD_GATHER,
             T_AD_SYMV, T_AD_SUM, T_AD_SYR, T_AD_TAIL, T_COMPACT, T_VPASS, T_CPASSZ, T_RHS, T_FSYMV, T_APPLY, T_GAMMA,
             T_KKT, T_MISC, T_LAST, NCYC };
// Section timers (the CY_* statistics): clock reads by EVERY thread around each pass — on in the developer build, and in the
// product build only with -DSSQP_SECTION_TIMERS (a clock read is a scheduling fence: loads cannot be hoisted across it).
#if defined(SSQP_TIMELINE) || defined(SSQP_SECTION_TIMERS)
#define SSQP_CLK() clock64()
#else
#define SSQP_CLK() 0LL
#endif
#ifdef SSQP_TIMELINE      // developer build: the timeline costs ~12% (it perturbs the schedule), off by default
#define SSQP_TICK(c, slot) do { if (threadIdx.x == 0) { const long long t__ = clock64(); (c).cyc[slot] += t__ - (c).cyc[T_LAST]; (c).cyc[T_LAST] = t__; } } while (0)
#else
#define SSQP_TICK(c, slot) do { } while (0)
#endif

struct KParams {
    int N, M, J, M0;
    int hrows;               // rows of H kept in shared memory
    int hcap;                // doubles of shared memory reserved for H / invB
    const double* V; long long strideV;
    const double* Ccol; const double* Crow; const double* cA;
    const double *q, *b, *g, *d, *u;
    const int* S0; const double* x0;
    long long strideS0, strideX0;   // per-QP strides of the warm start (0: one start point shared by the whole batch)
    double* x; int* S; long long* status; double* stats;
    double* work; long long wstride;
    long long stage_off;     // offset, in the CTA's workspace, of the TP_STAGE doubles the from-scratch factorisation stages through
    unsigned long long* queue;
    long long nb;
    int max_iter; double tol, tolG, tolLP;
    int phase1_only;         // 1: stop after initQP
    int lp_mode;             // 1: SimplexLP (q = cost vector, V unused)
    int rule;                // pivot rule of the simplex (Settings.rule): 0 :Dantzig, 1 :stpEdgeLP, 2 :maxImprovement
    int chain_len;           // > 1: QPs [c*chain_len, (c+1)*chain_len) form a chain solved in order by one CTA, each warm-started
                             // from the previous one's (x, S) — solveQP(Q, S, x0), src/SSQP.jl:237, along a sweep over q
    int rebuild_mode;        // 0: from-scratch factorisation (kinv_build_chol); 1: sequential bordered updates (kinv_rebuild;
                             // SSQP_REBUILD=border, kept for A/B and as the reference for the tests)
    int debug_perturb;       // test knob (SSQP_DEBUG_PERTURB): every this many status switches of a QP the diagonal of its
                             // inverse is scaled by (1 + 1e-3) — the drift guard must notice and rebuild; 0: off
    int nfree_cap;           // most free variables (d = -Inf and u = +Inf) any QP of the batch has: Phase 1 splits each
                             // into two [0, Inf) columns (src/SSQP.jl:484-509) and needs that many extra status slots
};

__host__ __device__ inline int rup(int a, int m) { return (a + m - 1) / m * m; }
__host__ __device__ inline long long tri64(long long i) { return i * (i + 1) / 2; }

// shared-memory carve-up (same arithmetic on host and device)
struct SmemLayout {
    int Np, nmp, M0p, bufsz;
    int z, gr, qq, dd, uu, rhs, sol, hv, colv, slack, cp, bg, pi, pcol, qB, rvec, sig, buf, red, cyc, H;
    int ndbl;
    int item, pos, Sst, Bv, supp, flist, rlist, lpos, evl, redi, misc;
    int nint;
    __host__ __device__ SmemLayout(int N, int M0, int J, int NT, int hcap, int nfree = 0) {
        Np = rup(N, 4); nmp = rup(N + M0, 4); M0p = rup(M0 > 0 ? M0 : 1, 32);
        bufsz = 3 * Np > NT ? 3 * Np : NT;      // streaming GEMV: slices 1..3 of an N-row pass; symmetric GEMV: NT
        int o = 0;
        z = o; o += Np; gr = o; o += Np; qq = o; o += Np; dd = o; o += Np; uu = o; o += Np;
        rhs = o; o += nmp; sol = o; o += nmp; hv = o; o += nmp; colv = o; o += nmp;
        slack = o; o += M0p; cp = o; o += M0p; bg = o; o += M0p;
        pi = o; o += M0p; pcol = o; o += M0p; qB = o; o += M0p; rvec = o; o += M0p; sig = o; o += M0p;
        buf = o; o += bufsz;
        red = o; o += 4 * 32 + 8;
        cyc = o; o += 40;
        H = o; o += rup(hcap, 2);
        ndbl = o;
        int p = 0;
        item = p; p += nmp; pos = p; p += nmp; Sst = p; p += rup(N + J + nfree + M0, 4);
        Bv = p; p += M0p; supp = p; p += Np; flist = p; p += Np; rlist = p; p += M0p; lpos = p; p += nmp; evl = p; p += nmp;
        redi = p; p += 2 * 32 + 8; misc = p; p += 32;
        nint = p;
    }
    __host__ __device__ size_t bytes() const { return (size_t)ndbl * 8 + (size_t)nint * 4; }
};

static_assert(NCYC <= 40 && ST_CYC0 + NCYC <= ST_DRIFT && ST_NKKT < NSTATS, "stats layout");

#ifdef __CUDACC__

// Every translation unit that instantiates the kernel compiles ONE flavour of the device code, in its own inline
// namespace: two flavours of ssqp_solve_kernel<NT> with the same mangled name in one library are an ODR violation
// (the runtime then launches whichever it registered first — it did, nondeterministically).
#ifdef SSQP_ONLY_VW4
#define SSQP_NS vw4
#else
#define SSQP_NS any
#endif
inline namespace SSQP_NS {

extern __shared__ __align__(16) double smem_d[];      // the CTA's dynamic shared memory (SmemLayout); visible to every device function

// Julia isless on Float64 is a total order with -0.0 < 0.0 and NaN last (sort!(..., by=x->x.L), src/SSQP.jl:94,176).
// sortable() maps a double to an int64 whose signed order is exactly that order (branch-free compares).
__device__ __forceinline__ long long sortable(double x) {
    if (x != x) return 0x7fffffffffffffffLL;
    const long long b = __double_as_longlong(x);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double unsortable(long long k) {
    return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}
// arg-min candidate: (key, insertion rank); an empty candidate is (INT64_MAX, INT_MAX)
struct Cand {
    long long k; int id;
    __device__ __forceinline__ Cand() : k(0x7fffffffffffffffLL), id(0x7fffffff) {}
    __device__ __forceinline__ void offer(double key, int i) {
        const long long kk = sortable(key);
        if (kk < k || (kk == k && i < id)) { k = kk; id = i; }
    }
    __device__ __forceinline__ void merge(long long k2, int i2) { if (k2 < k || (k2 == k && i2 < id)) { k = k2; id = i2; } }
    __device__ __forceinline__ bool any() const { return id != 0x7fffffff; }
    __device__ __forceinline__ double key() const { return unsortable(k); }
};

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
__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }
// a / b for 0 <= a < 2^20, 1 <= b < 2^11 without the ~40-instruction integer division (thread layouts are recomputed at every
// pass): (a + 0.5) / b is at least 0.5 / b away from an integer, far more than the single-precision error of the quotient
__device__ __forceinline__ int fastdiv(int a, int b) { return (int)__fdividef((float)a + 0.5f, (float)b); }
// (the streaming loads of the L2-resident operands V and [A;G] bypass L1: VecLd below)

// Per-thread context (register resident; the solver is inlined into the kernel except for the two packed-inverse
// primitives symv_leaf / syr_leaf).  Measured alternatives: Ctx in shared memory with every function a real call
// (code 367 KB instead of 950 KB, but 30% slower: pointer reloads after every shared-memory store), and real calls
// for the streaming GEMV / reductions (generic instead of LDS accesses, 13% slower).
struct Ctx {
    const KParams* P;
    int N, M, J, M0, M0p, bufsz;
    const double *V, *Ccol, *Crow, *cA, *q, *d, *u;
    double *z, *gr, *pfull, *rhs, *sol, *hv, *colv, *slack, *cp, *bg, *lam, *pi, *pcol, *qB, *rvec, *sig, *buf, *red;
    int *item, *pos, *Sst, *Bv, *supp, *flist, *rlist, *lpos, *evl, *redi, *misc;
    int nf, nr;          // variables / rows currently in the reduced system (lengths of flist / rlist)
    int nfree;           // Phase 1: free variables of this QP (their ids, ascending, in c.flist)
    bool xform;          // Phase 1: some lower bound is -Inf -> c.gr holds the column signs (-1 for (-Inf,u] variables),
                         // c.d / c.u the bounds of the transformed LP (src/SSQP.jl:484-509)
    double* Hs;          // shared-memory part of the packed inverse (rows < R)
    double* Hgm;         // global tail, biased so that row i >= R starts at Hgm + tri(i)
    double* work;        // the CTA's global workspace (unbiased)
    int R;               // rows of H in shared memory
    int n;               // current order of the reduced KKT system (K + W)
    bool sol_valid;      // c.sol holds the solution of the current reduced system at the current z
    bool reusable;       // phase2 ended optimal with the inverse matching the final status vector (polishSz! changed nothing):
                         // the next QP of a chain — same V, A, G, b, g, d, u, only q differs — starts from this very inverse
    double lamerr;       // fresh_solve(refine): largest correction of a multiplier, relative to the largest multiplier
    int ncache;          // constraint columns of the free variables flist[0 .. ncache) are cached in shared memory (ccache_*)
    int cstate;          // bit 0: a bulk copy into the cache is in flight; bit 1: parity of the mbarrier phase it completes
    double bytes;        // streamed bytes (thread 0 only)
    long long* cyc;       // section timers / call counters (shared memory, thread 0 only)
    __device__ __forceinline__ double* hrow(int i) const { return (i < R ? Hs : Hgm) + tri(i); }
};

// ---- block-wide deterministic reductions (all threads must call) ----------------------------------
// Stage 1: warp shuffle tree; stage 2: every warp re-reduces the NW per-warp partials with a second shuffle
// tree (lane l holds partial l % NW), so the result is bit-identical in every thread and costs two barriers.
// (inlined; static shared scratch.  Real calls measured slower, see the note at struct Ctx.  Also measured and dropped:
// one barrier instead of two, with the per-warp partials alternating between two scratch sets — 0.4 % slower: the warps
// that arrive early wait at the next barrier instead.)
template <int NT>
static __device__ __forceinline__ double block_sum_leaf(double v) {
    constexpr int NW = NT / 32;
    __shared__ double red[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    v = warp_sum(v);
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = red[l & (NW - 1)];
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
struct Sum3 { double a, b, d; };
template <int NT>
static __device__ __forceinline__ Sum3 block_sum3_leaf(double a, double b, double d) {
    constexpr int NW = NT / 32;
    __shared__ double red[96];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    a = warp_sum(a); b = warp_sum(b); d = warp_sum(d);
    __syncthreads();
    if (l == 0) { red[w] = a; red[32 + w] = b; red[64 + w] = d; }
    __syncthreads();
    double s0 = red[l & (NW - 1)], s1 = red[32 + (l & (NW - 1))], s2 = red[64 + (l & (NW - 1))];
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    return Sum3{s0, s1, s2};
}
// max of NON-NEGATIVE doubles (every use in the solver is a max of |.|): integer redux on the bit pattern
__device__ __forceinline__ double warp_max_nonneg(double v) {
    const unsigned long long mb = (unsigned long long)__double_as_longlong(v);
    const unsigned xh = (unsigned)(mb >> 32), xl = (unsigned)mb;
    const unsigned mxh = __reduce_max_sync(0xffffffffu, xh);
    const unsigned mxl = __reduce_max_sync(0xffffffffu, xh == mxh ? xl : 0u);
    return __longlong_as_double((long long)(((unsigned long long)mxh << 32) | mxl));
}
template <int NT>
static __device__ __forceinline__ double block_max_leaf(double v) {
    constexpr int NW = NT / 32;
    __shared__ double red[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    v = warp_max_nonneg(v);
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    return warp_max_nonneg(red[l & (NW - 1)]);
}
// Warp-level arg-min of (sortable key, rank) and max of a NON-NEGATIVE double with the hardware integer reductions
// (redux.sync): a 64-bit key is reduced as (signed high word, unsigned low word), 2-3 dependent redux per value
// instead of five dependent 64-bit shuffle steps (a shuffle-tree arg-min measured 1.1k cycles per CTA reduction).
__device__ __forceinline__ void warp_argmin_max(long long& k, int& id, double& mx) {
    const int hi = (int)(k >> 32);
    const unsigned lo = (unsigned)k;
    const int mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const int mid = __reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? id : 0x7fffffff);
    k = ((long long)mhi << 32) | (long long)mlo; id = mid;
    const unsigned long long mb = (unsigned long long)__double_as_longlong(mx);      // mx >= 0: bit pattern is monotone
    const unsigned xh = (unsigned)(mb >> 32), xl = (unsigned)mb;
    const unsigned mxh = __reduce_max_sync(0xffffffffu, xh);
    const unsigned mxl = __reduce_max_sync(0xffffffffu, xh == mxh ? xl : 0u);
    mx = __longlong_as_double((long long)(((unsigned long long)mxh << 32) | mxl));
}
// arg-min of (key, rank) candidates fused with a max reduction (mx must be >= 0 or NaN-free); broadcast to all threads
struct CandMax { long long k; int id; double mx; };
template <int NT>
static __device__ __forceinline__ CandMax block_argmin_leaf(long long qk, int qid, double mx) {
    constexpr int NW = NT / 32;
    __shared__ long long redk[32];
    __shared__ double redm[32];
    __shared__ int redi[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    warp_argmin_max(qk, qid, mx);
    __syncthreads();
    if (l == 0) { redk[w] = qk; redi[w] = qid; redm[w] = mx; }
    __syncthreads();
    qk = redk[l & (NW - 1)]; qid = redi[l & (NW - 1)]; mx = redm[l & (NW - 1)];
    warp_argmin_max(qk, qid, mx);
    return CandMax{qk, qid, mx};
}
template <int NT> static __device__ __forceinline__ double block_sum(Ctx& c, double v) { return block_sum_leaf<NT>(v); }
template <int NT> static __device__ __forceinline__ double block_max(Ctx& c, double v) { return block_max_leaf<NT>(v); }
template <int NT> static __device__ __forceinline__ void block_sum3(Ctx& c, double& a, double& b, double& d) {
    const Sum3 r = block_sum3_leaf<NT>(a, b, d);
    a = r.a; b = r.b; d = r.d;
}
template <int NT> static __device__ __forceinline__ void block_argmin(Ctx& c, Cand& q) {
    const CandMax r = block_argmin_leaf<NT>(q.k, q.id, 0.0);
    q.k = r.k; q.id = r.id;
}
template <int NT> static __device__ __forceinline__ void block_argmin_max(Ctx& c, Cand& q, double& mx) {
    const CandMax r = block_argmin_leaf<NT>(q.k, q.id, mx);
    q.k = r.k; q.id = r.id; mx = r.mx;
}

// ordered compaction of {k in [0,cnt) : pred(k)} into out[]; returns the count (all threads)
template <int NT, class Pred>
static __device__ __forceinline__ int block_compact_impl(int* redi, int cnt, int* out, Pred pred) {
    constexpr int NW = NT / 32;
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    int base = 0;
    for (int s = 0; s < cnt; s += NT) {
        int k = s + threadIdx.x;
        bool p = (k < cnt) && pred(k);
        unsigned m = __ballot_sync(0xffffffffu, p);
        __syncthreads();
        if (l == 0) redi[w] = __popc(m);
        __syncthreads();
        int off = base, tot = base;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            int v = redi[i];
            if (i < w) off += v;
            tot += v;
        }
        if (p) out[off + __popc(m & ((1u << l) - 1u))] = k;
        base = tot;
    }
    __syncthreads();
    return base;
}

// ordered list of {k < cnt : x[k] != 0} (the support of z: the columns a gradient / slack pass has to read)
template <int NT, class Pred>
static __device__ __forceinline__ int block_compact(Ctx& c, int cnt, int* out, Pred pred) {
    return block_compact_impl<NT>(c.redi, cnt, out, pred);
}
template <int NT>
static __device__ __forceinline__ int compact_nonzero_leaf(int* redi, const double* x, int cnt, int* out) {
    return block_compact_impl<NT>(redi, cnt, out, [=](int k) { return x[k] != 0.0; });
}
template <int NT>
static __device__ __forceinline__ int compact_nonzero(Ctx& c, const double* x, int cnt, int* out) {
    return compact_nonzero_leaf<NT>(c.redi, x, cnt, out);
}

// ---- streaming passes over L2-resident column-major data -------------------------------------------
// out[r] = init[r] + sum_{t<cnt} col(t)[r] * wt(t)   for r < rows.   col(t): pointer to a contiguous column.
// Threads are laid out as (group of VW rows, slice of t); every thread keeps a batch of NB vector loads (VW*8
// bytes each: 256-bit LDG when rows % 4 == 0) in flight and the tail of the t range is predicated into the same
// batch (never a serial remainder).
// (ldp: predicated form — the destination registers keep their value (0) when pred == 0 — so that a batch of loads
// has no branches between them and the compiler can issue the index / weight loads first and the LDGs back to back)
#ifndef SSQP_LDVOL
#define SSQP_LDVOL volatile
#endif
template <int VW> struct VecLd;
template <> struct VecLd<4> {
    static __device__ __forceinline__ void ldp(const double* p, double* v, int pred) {
        asm SSQP_LDVOL ("{\n\t.reg .pred pp;\n\tsetp.ne.s32 pp, %5, 0;\n\t"
                     "@pp ld.global.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];\n\t}"
                     : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3]) : "l"(p), "r"(pred));
    }
    static __device__ __forceinline__ void ld(const double* p, double* v) {
        asm volatile("ld.global.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
                     : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    }
};
template <> struct VecLd<2> {
    static __device__ __forceinline__ void ldp(const double* p, double* v, int pred) {
        asm SSQP_LDVOL ("{\n\t.reg .pred pp;\n\tsetp.ne.s32 pp, %3, 0;\n\t"
                     "@pp ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];\n\t}"
                     : "+d"(v[0]), "+d"(v[1]) : "l"(p), "r"(pred));
    }
    static __device__ __forceinline__ void ld(const double* p, double* v) {
        asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
    }
};
template <> struct VecLd<1> {
    static __device__ __forceinline__ void ldp(const double* p, double* v, int pred) {
        asm SSQP_LDVOL ("{\n\t.reg .pred pp;\n\tsetp.ne.s32 pp, %2, 0;\n\t"
                     "@pp ld.global.L1::no_allocate.f64 %0, [%1];\n\t}"
                     : "+d"(v[0]) : "l"(p), "r"(pred));
    }
    static __device__ __forceinline__ void ld(const double* p, double* v) {
        asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v[0]) : "l"(p));
    }
};

// Shared-memory operands travel as OFFSETS into smem_d (doubles; `list` in ints) so that the non-inlined GEMV keeps
// them on LDS/STS instead of generic accesses.
struct GemvArgs {            // out[r] = init[r] + sum_{t<cnt} base[(list ? list[t] : t) * ld + r] * w[list ? list[t] : t]
    const double* base; long long ld;      // global (L2-resident) column-major operand
    int list_off;                          // int offset of the column list in shared memory, or -1 (identity)
    int w_off;                             // weights (shared memory)
    int cnt, rows;
    const double* init_g;                  // optional initial value in global memory ...
    int init_off;                          // ... or in shared memory (-1: none)
    int out_off;                           // result (shared memory)
    int buf_off, bufsz;                    // staging buffer for the slices' partial sums (shared memory, doubles)
    int cyc_off;                           // section timers (-1: none)
};
static __device__ __forceinline__ int soff(const double* p) { return (int)(p - smem_d); }
static __device__ __forceinline__ int ioff(const int* p) { return (int)(p - reinterpret_cast<const int*>(smem_d)); }

template <int NT, int VW, int NB>
static __device__ __forceinline__ void gemv_cols_vw(const GemvArgs& a) {
    const int rows = a.rows, cnt = a.cnt;
    const int G = rows / VW;                          // groups of VW consecutive rows (rows % VW == 0)
    // Thread (slice, row group): a warp reads 32 consecutive row groups of ONE column per load instruction (1 KB
    // contiguous with 256-bit loads; slices laid out inside a warp measured 25% slower).  Slice 0 accumulates into
    // `out`, slices 1.. into the staging buffer; they are combined in a fixed order (deterministic).
    int SL = G <= NT ? fastdiv(NT, G) : 1;
    if (SL > 16) SL = 16;
    if (SL > cnt) SL = cnt > 0 ? cnt : 1;
    while (SL > 1 && (SL - 1) * rows > a.bufsz) --SL;
    const int* list = a.list_off >= 0 ? reinterpret_cast<const int*>(smem_d) + a.list_off : nullptr;
    const double* wt = smem_d + a.w_off;
    const double* init = a.init_g ? a.init_g : (a.init_off >= 0 ? smem_d + a.init_off : nullptr);
    double* out = smem_d + a.out_off;
    double* buf = smem_d + a.buf_off;
    const double* base = a.base; const long long ld = a.ld;
    const long long tg0_ = SSQP_CLK();
    const int sl = G <= NT ? fastdiv(threadIdx.x, G) : 0;
    for (int g = G <= NT ? threadIdx.x - sl * G : threadIdx.x; g < G && sl < SL; g += NT) {      // one trip unless G > NT
        double acc[VW], acc2[VW];
#pragma unroll
        for (int q = 0; q < VW; ++q) { acc[q] = 0.0; acc2[q] = 0.0; }
        const double* bg = base + VW * g;
        for (int t0 = sl; t0 < cnt; t0 += NB * SL) {      // NB vector loads in flight, tail predicated, no branches
            double v[NB][VW], wv[NB];
            int kk[NB];
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                const int te = t0 + e * SL;
                const int tc = te < cnt ? te : cnt - 1;
                kk[e] = list ? list[tc] : tc;
            }
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                const double wk = wt[kk[e]];
                wv[e] = (t0 + e * SL < cnt) ? wk : 0.0;
#pragma unroll
                for (int q = 0; q < VW; ++q) v[e][q] = 0.0;
            }
#pragma unroll
            for (int e = 0; e < NB; ++e) VecLd<VW>::ldp(bg + (size_t)kk[e] * ld, v[e], (t0 + e * SL < cnt) ? 1 : 0);
#pragma unroll
            for (int e = 0; e < NB; ++e)
#pragma unroll
                for (int q = 0; q < VW; ++q) {
                    if (e & 1) acc2[q] += v[e][q] * wv[e]; else acc[q] += v[e][q] * wv[e];
                }
        }
        double* dst = (sl == 0) ? out + VW * g : buf + (size_t)(sl - 1) * rows + VW * g;
#pragma unroll
        for (int q = 0; q < VW; ++q) dst[q] = acc[q] + acc2[q];
    }
    const long long tg1_ = SSQP_CLK();
    __syncthreads();
    if (SL > 1 || init) {
        for (int r = threadIdx.x; r < rows; r += NT) {
            double sum = (init ? init[r] : 0.0) + out[r];
            for (int s2 = 1; s2 < SL; ++s2) sum += buf[(size_t)(s2 - 1) * rows + r];
            out[r] = sum;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && a.cyc_off >= 0) {
        long long* cyc = reinterpret_cast<long long*>(smem_d + a.cyc_off);
        cyc[CY_GLOAD] += tg1_ - tg0_; cyc[CY_GEPI] += SSQP_CLK() - tg1_;
    }
}

// Inlined at its call sites (a real call measured 20% slower even with LDS operands).  SSQP_ONLY_VW4 builds (problem
// sizes with N % 4 == 0 and (M+J) % 4 == 0) carry the 256-bit variant only: the kernel's code shrinks from 950 KB to
// 575 KB and runs 6% faster (the Phase-2 trip no longer overflows the instruction cache as badly).
#ifndef SSQP_NB4
#define SSQP_NB4 6      // 256-bit loads in flight per thread in the streaming passes (8 measured: see scripts/micro/README.md)
#endif
template <int NT>
static __device__ __forceinline__ void gemv_cols(const GemvArgs a) {
    if (a.rows <= 0) return;
#ifdef SSQP_ONLY_VW4
    gemv_cols_vw<NT, 4, SSQP_NB4>(a);
#else
    // vector width from the row count AND the alignment of the operand (a per-QP V handed in as a device pointer may sit
    // at any 8-byte offset: a 256-bit load from a misaligned base faults and leaves the context in a sticky error)
    const unsigned long long al = (unsigned long long)(uintptr_t)a.base | ((unsigned long long)a.ld << 3);
    if ((a.rows & 3) == 0 && (al & 31ULL) == 0) gemv_cols_vw<NT, 4, 6>(a);
    else if ((a.rows & 1) == 0 && (al & 15ULL) == 0) gemv_cols_vw<NT, 2, 8>(a);
    else gemv_cols_vw<NT, 1, 8>(a);
#endif
}

// out[r] = sum_t Ccol[r + list[t]*M0] * w[list[t]]   for r < M0   (constraint pass over a variable list)
template <int NT>
static __device__ void cpass(Ctx& c, const int* list, int cnt, const double* w, double* out) {
    const long long t0_ = SSQP_CLK();
    const int M0 = c.M0;
    if (M0 == 0) return;
    gemv_cols<NT>(GemvArgs{c.Ccol, M0, ioff(list), soff(w), cnt, M0, nullptr, -1, soff(out), soff(c.buf), c.bufsz, -1});
    if (threadIdx.x == 0) { c.bytes += 8.0 * M0 * cnt; c.cyc[CY_CPASS] += SSQP_CLK() - t0_; }
}

// gr[i] = q[i] + sum_t V[i + list[t]*N] * z[list[t]]    (gradient at z over the support of z)
template <int NT>
static __device__ void vpass(Ctx& c, const int* list, int cnt) {
    const long long t0_ = SSQP_CLK();
    const int N = c.N;
    gemv_cols<NT>(GemvArgs{c.V, N, ioff(list), soff(c.z), cnt, N, nullptr, soff(c.q), soff(c.gr), soff(c.buf), c.bufsz, soff(reinterpret_cast<double*>(c.cyc))});
    if (threadIdx.x == 0) { c.bytes += 8.0 * N * cnt; c.cyc[CY_VPASS] += SSQP_CLK() - t0_; }
}

// out[o] = sum_{m<nin} f(o, m)  for o < nout: threads laid out as (output, slice of m); slices are combined
// in a fixed order through c.buf.  For small dense operands that live in shared memory.
template <int NT, class F>
static __device__ __forceinline__ void small_reduce_leaf(double* buf, int nout, int nin, F f, double* out) {
    const int tid = threadIdx.x;
    const int Wd = rup(nout, 32);
    if (Wd <= NT) {
        int S = fastdiv(NT, Wd);
        if (S > nin) S = nin > 0 ? nin : 1;
        const int s = fastdiv(tid, Wd), o = tid - s * Wd;
        if (s < S) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (o < nout) {
                int m = s;
                for (; m + 3 * S < nin; m += 4 * S) {            // four loads in flight per thread
                    const double f0 = f(o, m), f1 = f(o, m + S), f2 = f(o, m + 2 * S), f3 = f(o, m + 3 * S);
                    a0 += f0; a1 += f1; a2 += f2; a3 += f3;
                }
                for (; m < nin; m += S) a0 += f(o, m);
            }
            buf[s * Wd + o] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        for (int o2 = tid; o2 < nout; o2 += NT) {
            double sum = 0.0;
            for (int g = 0; g < S; ++g) sum += buf[g * Wd + o2];
            out[o2] = sum;
        }
        __syncthreads();
    } else {
        for (int o = tid; o < nout; o += NT) {
            double a0 = 0.0;
            for (int m = 0; m < nin; ++m) a0 += f(o, m);
            out[o] = a0;
        }
        __syncthreads();
    }
}

// ---- packed symmetric inverse H: y = H x, H += sigma v v' ------------------------------------------
// The shared-memory block (rows < R) is read with symmetric indexing — thread (j, slice) walks a contiguous
// range [m0, m1) of "row j of the full matrix": H[tri(j)+m] for m <= j (triangular offsets of consecutive j fall
// in distinct banks) and H[tri(m)+j] for m > j (consecutive j -> consecutive words).  The range is cut at the
// warp's first / last row so that the two long segments are branch-free and only the <= 32 columns that cross
// the diagonal of the warp's rows take a per-element select.  Slices are combined in a fixed order.
// The global tail (rows >= R) is read twice, both times coalesced: warp-per-row for the row sums,
// thread-per-column for the column sums.
template <int NT, class F>
static __device__ __forceinline__ void small_reduce(Ctx& c, int nout, int nin, F f, double* out) {
    small_reduce_leaf<NT>(c.buf, nout, nin, f, out);
}

#ifdef SSQP_LEAF_INLINE
#define SSQP_LEAF __forceinline__
#else
#define SSQP_LEAF __noinline__
#endif
struct HView {            // what the packed-inverse kernels need (kept small: they are real calls, not inlined)
    double* Hs; double* Hgm; int R; double* buf;
};

template <int NT>
static __device__ SSQP_LEAF void symv_leaf(const HView h, int n, const double* x, double* y) {
    constexpr int NW = NT / 32;
    const int ns = n < h.R ? n : h.R;
    const double* Hs = h.Hs;
    const int tid = threadIdx.x;
    const int Wd = rup(ns, 32);
    if (Wd <= NT) {
        const int S = fastdiv(NT, Wd);
        const int chunk = rup(fastdiv(ns + S - 1, S), 2);
        const int s = fastdiv(tid, Wd), j = tid - s * Wd;
        if (s < S) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const int m0 = s * chunk;
            const int m1 = (m0 + chunk < ns) ? m0 + chunk : ns;
            if (j < ns && m0 < m1) {
                const int jlo = j & ~31;                            // first / last row of this warp
                const int jhi = (jlo + 31 < ns - 1) ? jlo + 31 : ns - 1;
                const int tj = tri(j);
                int m = m0;
                {   // segment A: m <= every row of the warp -> own row, contiguous
                    const int mA = (jlo < m1) ? jlo : m1;
                    const double* rowj = Hs + tj;
                    for (; m + 3 < mA; m += 4) {
                        a0 += rowj[m] * x[m]; a1 += rowj[m + 1] * x[m + 1];
                        a2 += rowj[m + 2] * x[m + 2]; a3 += rowj[m + 3] * x[m + 3];
                    }
                    for (; m < mA; ++m) a0 += rowj[m] * x[m];
                }
                int off = tri(m) + j;                               // H[m][j] for m > j
                {   // segment D: columns crossing the warp's diagonal block (one select per element)
                    const int mD = (jhi + 1 < m1) ? jhi + 1 : m1;
                    for (; m < mD; ++m) {
                        a1 += Hs[(m <= j) ? tj + m : off] * x[m];
                        off += m + 1;
                    }
                }
                {   // segment B: m > every row of the warp -> column j of rows m
                    for (; m + 3 < m1; m += 4) {
                        const int o1 = off + m + 1, o2 = o1 + m + 2, o3 = o2 + m + 3;
                        a0 += Hs[off] * x[m]; a1 += Hs[o1] * x[m + 1];
                        a2 += Hs[o2] * x[m + 2]; a3 += Hs[o3] * x[m + 3];
                        off = o3 + m + 4;
                    }
                    for (; m < m1; ++m) { a0 += Hs[off] * x[m]; off += m + 1; }
                }
            }
            h.buf[s * Wd + j] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        for (int o2 = tid; o2 < ns; o2 += NT) {
            double sum = 0.0;
            for (int g = 0; g < S; ++g) sum += h.buf[g * Wd + o2];
            y[o2] = sum;
        }
        __syncthreads();
    } else {
        for (int j = tid; j < ns; j += NT) {
            double a0 = 0.0;
            for (int m = 0; m < ns; ++m) a0 += ((m <= j) ? Hs[tri(j) + m] : Hs[tri(m) + j]) * x[m];
            y[j] = a0;
        }
        __syncthreads();
    }
    if (n > h.R) {
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        const double* Hg = h.Hgm;
        for (int i = h.R + w; i < n; i += NW) {
            const double* row = Hg + tri(i);
            double a0 = 0.0, a1 = 0.0;
            int k = l;
            for (; k + 96 <= i; k += 128) {
                double v0 = row[k], v1 = row[k + 32], v2 = row[k + 64], v3 = row[k + 96];
                a0 += v0 * x[k]; a1 += v1 * x[k + 32]; a0 += v2 * x[k + 64]; a1 += v3 * x[k + 96];
            }
            for (; k <= i; k += 32) a0 += row[k] * x[k];
            a0 = warp_sum(a0 + a1);
            if (l == 0) y[i] = a0;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += NT) {
            int i = (k + 1 > h.R) ? k + 1 : h.R;
            double a0 = 0.0, a1 = 0.0;
            for (; i + 7 < n; i += 8) {
                double v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = Hg[tri(i + e) + k];
#pragma unroll
                for (int e = 0; e < 8; e += 2) { a0 += v[e] * x[i + e]; a1 += v[e + 1] * x[i + e + 1]; }
            }
            for (; i < n; ++i) a0 += Hg[tri(i) + k] * x[i];
            y[k] += a0 + a1;
        }
        __syncthreads();
    }
}

template <int NT>
static __device__ __forceinline__ void symv(Ctx& c, int n, const double* x, double* y) {
    const long long t0_ = SSQP_CLK();
    symv_leaf<NT>(HView{c.Hs, c.Hgm, c.R, c.buf}, n, x, y);
    if (threadIdx.x == 0) {
        if (n > c.R) c.bytes += 16.0 * (tri(n) - tri(c.R));
        c.cyc[CY_SYMV] += SSQP_CLK() - t0_; c.cyc[CY_NSYMV] += 1;
    }
}

// H += sigma * v v'   on the packed lower triangle (order n); warp per row, 32 columns per step
template <int NT>
static __device__ SSQP_LEAF void syr_leaf(const HView h, int n, const double* __restrict__ v, double sigma) {
    constexpr int NW = NT / 32;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int ns = n < h.R ? n : h.R;
    double* __restrict__ Hs = h.Hs;
    // shared-memory rows, one 32-column chunk (column block q) at a time: the lane's v[k] is loaded once per column
    // block, and four independent rows are in flight per warp step (loads first, then FMAs, then stores)
    for (int q0 = 0; q0 < ns; q0 += 32) {
        const int k = q0 + l;
        const double vk = (k < ns) ? v[k] : 0.0;
        for (int i0 = q0 + w; i0 < ns; i0 += 4 * NW) {
            const int i1 = i0 + NW, i2 = i0 + 2 * NW, i3 = i0 + 3 * NW;
            const bool p0 = k <= i0, p1 = (i1 < ns) && (k <= i1), p2 = (i2 < ns) && (k <= i2), p3 = (i3 < ns) && (k <= i3);
            double* r0 = Hs + tri(i0) + k; double* r1 = Hs + tri(i1) + k; double* r2 = Hs + tri(i2) + k; double* r3 = Hs + tri(i3) + k;
            const double c0 = sigma * v[i0];
            const double c1 = (i1 < ns) ? sigma * v[i1] : 0.0;
            const double c2 = (i2 < ns) ? sigma * v[i2] : 0.0;
            const double c3 = (i3 < ns) ? sigma * v[i3] : 0.0;
            const double a0 = p0 ? *r0 : 0.0, a1 = p1 ? *r1 : 0.0, a2 = p2 ? *r2 : 0.0, a3 = p3 ? *r3 : 0.0;
            if (p0) *r0 = a0 + c0 * vk;
            if (p1) *r1 = a1 + c1 * vk;
            if (p2) *r2 = a2 + c2 * vk;
            if (p3) *r3 = a3 + c3 * vk;
        }
    }
    for (int i = ns + w; i < n; i += NW) {             // global tail rows
        double* row = h.Hgm + tri(i);
        const double ci = sigma * v[i];
        int k = l;
        for (; k + 96 <= i; k += 128) {
            double r0 = row[k], r1 = row[k + 32], r2 = row[k + 64], r3 = row[k + 96];
            row[k] = r0 + ci * v[k]; row[k + 32] = r1 + ci * v[k + 32];
            row[k + 64] = r2 + ci * v[k + 64]; row[k + 96] = r3 + ci * v[k + 96];
        }
        for (; k <= i; k += 32) row[k] += ci * v[k];
    }
    __syncthreads();
}

template <int NT>
static __device__ __forceinline__ void syr(Ctx& c, int n, const double* v, double sigma) {
    const long long t0_ = SSQP_CLK();
    syr_leaf<NT>(HView{c.Hs, c.Hgm, c.R, c.buf}, n, v, sigma);
    if (threadIdx.x == 0) {
        if (n > c.R) c.bytes += 16.0 * (tri(n) - tri(c.R));
        c.cyc[CY_SYR] += SSQP_CLK() - t0_; c.cyc[CY_NSYR] += 1;
    }
}

// ---- constraint-column cache (TMA bulk copies) ------------------------------------------------------
// The ratio test of every Phase-2 trip needs [A;G][:,F] * p (src/SSQP.jl:79-92).  From L2 that pass is pure latency
// (index -> weight -> address -> LDG -> reduce: 2.7-6.7k cycles for 80 KB).  The packed inverse rarely fills its
// shared-memory region (order n ~ 140 of a capacity of 194 rows at N=500), so the free END of that region caches the
// constraint columns of the free variables, one slot of M0 doubles per position of c.flist: slot t <-> variable flist[t],
// for t < c.ncache.  A slot is filled when a variable is released — a 1-D bulk copy global -> shared by the TMA engine
// (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP / SYNCS), issued by one thread and consumed trips later, so its L2
// latency is never exposed — moved when the free list swap-removes, and given up when the inverse grows into it.
// The bulk copy needs 16-byte granularity: M0 even (otherwise the cache stays empty and the pass streams from L2).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long* ccache_mbar() {          // the CTA's mbarrier (static shared memory)
    __shared__ __align__(8) unsigned long long mbar;
    return &mbar;
}
__device__ __forceinline__ void ccache_init() {          // once per CTA (thread 0), followed by a __syncthreads
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(ccache_mbar())));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool ccache_on(const Ctx& c) { return (c.M0 & 1) == 0 && c.M0 > 0; }
// (slots hang from the even floor of the capacity: with M0 even every slot is 16-byte aligned, for the bulk copies and for the
// 128-bit loads of the cached pass)
__device__ __forceinline__ double* ccache_slot(const Ctx& c, int t) { return c.Hs + ((c.P->hcap & ~1) - (t + 1) * c.M0); }
// slots that fit above a packed inverse of order n (rows beyond c.R live in global memory)
__device__ __forceinline__ int ccache_room(const Ctx& c, int n) {
    const int rows = n < c.R ? n : c.R;
    // floor((hcap - tri(rows)) / M0) without the integer division (every thread evaluates this once or twice a trip): the
    // single-precision quotient is within one of the exact one, and one slot fewer is always safe
    const int room = (int)(__fdividef((float)((c.P->hcap & ~1) - tri(rows)), (float)c.M0)) - 1;
    return room > 0 ? room : 0;
}
// all threads: wait for the bulk copy in flight (if any); afterwards the cache may be read through ordinary loads
__device__ __forceinline__ void ccache_wait(Ctx& c) {
    if (c.cstate & 1) {
        const unsigned mb = smem_u32(ccache_mbar()), par = (c.cstate >> 1) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tCCW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra CCD;\n\tbra CCW;\n\tCCD:\n\t}"
                     :: "r"(mb), "r"(par) : "memory");
        c.cstate = (c.cstate ^ 2) & ~1;
    }
}
// all threads: fill slots [t0, t1) with the columns of flist[t0 .. t1) — one mbarrier phase, copies issued by thread 0
template <int NT>
static __device__ __forceinline__ void ccache_fill(Ctx& c, int t0, int t1) {
    if (t1 <= t0) return;
    ccache_wait(c);
    __syncthreads();          // every ordinary access to the slots' memory (inverse rows, slot moves) is done
    if (threadIdx.x == 0) {
        const unsigned mb = smem_u32(ccache_mbar());
        const unsigned bytes = (unsigned)c.M0 * 8u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(bytes * (unsigned)(t1 - t0)) : "memory");
        for (int t = t0; t < t1; ++t) {
            const double* src = c.Ccol + (size_t)c.flist[t] * c.M0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(ccache_slot(c, t))), "l"(src), "r"(bytes), "r"(mb) : "memory");
        }
    }
    c.cstate |= 1;
}
// the free list is about to grow the inverse to order n1: give up the slots its rows will overwrite
__device__ __forceinline__ void ccache_shrink(Ctx& c, int n1) {
    if (c.ncache > 0) {
        const int room = ccache_room(c, n1);
        if (c.ncache > room) { ccache_wait(c); c.ncache = room; }
    }
}
// out[r] = sum_{t < nf} [A;G][r, flist[t]] * w[flist[t]]: cached columns from shared memory, the others from L2
template <int NT>
static __device__ void cpass_free(Ctx& c, const double* w, double* out) {
    const int M0 = c.M0, nc = c.ncache;
    if (M0 == 0) return;
    // The L2 pass costs its latency whatever its length, so the cache pays only when it holds EVERY column of the free
    // list (the top-up in phase2 works towards that); a partial cache is ignored for the pass.
    if (nc < c.nf) { cpass<NT>(c, c.flist, c.nf, w, out); return; }
    const long long t0_ = SSQP_CLK();
    ccache_wait(c);
    // threads = (pair of rows, slice of t): the slots are 16-byte aligned (M0 and hcap even), so a thread reads two rows of a
    // column with one 128-bit load; P = M0/2 row pairs x up to 16 slices (10 at M0 = 100: 8 columns per thread at nc = 80)
    const int P2 = M0 >> 1;
    if (P2 <= NT) {
        int SLc = fastdiv(NT, P2);
        if (SLc > 16) SLc = 16;
        while (SLc > 1 && (SLc - 1) * M0 > c.bufsz) --SLc;
        const int sl = fastdiv(threadIdx.x, P2), rp = threadIdx.x - sl * P2;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        const bool act = sl < SLc;
        if (act) {
            const double2* base = reinterpret_cast<const double2*>(c.Hs + ((c.P->hcap & ~1) - M0)) + rp;        // slot t at base - t*P2
            int t = sl;
            for (; t + 3 * SLc < nc; t += 4 * SLc) {
                const int k0 = c.flist[t], k1 = c.flist[t + SLc], k2 = c.flist[t + 2 * SLc], k3 = c.flist[t + 3 * SLc];
                const double w0 = w[k0], w1 = w[k1], w2 = w[k2], w3 = w[k3];
                const double2 v0 = base[-(t) * P2], v1 = base[-(t + SLc) * P2], v2 = base[-(t + 2 * SLc) * P2], v3 = base[-(t + 3 * SLc) * P2];
                a0 += v0.x * w0; a1 += v0.y * w0; b0 += v1.x * w1; b1 += v1.y * w1;
                a0 += v2.x * w2; a1 += v2.y * w2; b0 += v3.x * w3; b1 += v3.y * w3;
            }
            for (; t < nc; t += SLc) { const double wt = w[c.flist[t]]; const double2 v = base[-t * P2]; a0 += v.x * wt; a1 += v.y * wt; }
            a0 += b0; a1 += b1;
            if (sl > 0) *reinterpret_cast<double2*>(c.buf + (sl - 1) * M0 + 2 * rp) = make_double2(a0, a1);
        }
        __syncthreads();
        if (sl == 0) {
            for (int s2 = 1; s2 < SLc; ++s2) {
                const double2 v = *reinterpret_cast<const double2*>(c.buf + (s2 - 1) * M0 + 2 * rp);
                a0 += v.x; a1 += v.y;
            }
            *reinterpret_cast<double2*>(out + 2 * rp) = make_double2(a0, a1);
        }
    } else {
        for (int rr = threadIdx.x; rr < M0; rr += NT) {
            double sum = 0.0;
            const double* base = c.Hs + ((c.P->hcap & ~1) - M0) + rr;
            for (int t = 0; t < nc; ++t) sum += base[-t * M0] * w[c.flist[t]];
            out[rr] = sum;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) c.cyc[CY_CPASS] += SSQP_CLK() - t0_;
}

// ---- reduced-KKT inverse maintenance ---------------------------------------------------------------
// c.sol is indexed by ITEM ID (variable k at [k], constraint row r at [N + r]; 0 for items outside the system), so
// that c.pfull (= c.sol) is the direction p by variable and c.lam (= c.sol + N) the multipliers by row with no
// scatter step.  flist / rlist are the (unordered) lists of variables / rows in the system, kept incrementally.
static __device__ __forceinline__ void list_add(Ctx& c, int it) {        // thread 0 only
    if (it < c.N) { c.lpos[it] = c.nf; c.flist[c.nf] = it; }
    else { c.lpos[it] = c.nr; c.rlist[c.nr] = it - c.N; }
}
static __device__ __forceinline__ void list_remove(Ctx& c, int it) {     // thread 0 only
    const int q = c.lpos[it];
    if (it < c.N) { const int last = c.flist[c.nf - 1]; c.flist[q] = last; c.lpos[last] = q; }
    else { const int last = c.rlist[c.nr - 1]; c.rlist[q] = last; c.lpos[c.N + last] = q; }
}

// Bordered add of item `it` (variable k, or N + row); rnew = right-hand side entry of the new item
// (-gradient_k for a variable, slack_r for a row) used to carry c.sol along.
// Returns 0 ok, 1 dependent/singular pivot (nothing changed).
// Two thresholds on the border pivot s (relative to the sum of the magnitudes it is formed from):
//   PIV_SOFT  while the system is maintained incrementally: anything below it is SUSPECTED dependent and sends the working
//             set through the reference's own test — getRowsGJr([AE bE], tol), which the reference runs every trip
//             (src/SSQP.jl:310) — by way of a rebuild with the row purge.  Generous on purpose: the purge is the arbiter.
//   PIV_HARD  inside that rebuild, for the rows getRowsGJr kept: only a pivot that is roundoff (or of the wrong sign: the
//             reference's cholesky throws PosDefException, src/SSQP.jl:322,328) stops the solve with status -1.
constexpr double PIV_SOFT = 1e-8, PIV_HARD = 4e-15;
template <int NT>
static __device__ int kinv_add(Ctx& c, int it, double rnew, const bool hard = false) {
    const int n = c.n, N = c.N, M0 = c.M0;
    const double diag = (it < N) ? c.V[it + (size_t)it * N] : 0.0;
    ccache_shrink(c, n + 1);
    if (n == 0) {
        if (!(diag > 0.0)) return 1;
        if (threadIdx.x == 0) { c.hrow(0)[0] = 1.0 / diag; c.item[0] = it; c.pos[it] = 0; c.sol[it] = rnew / diag; list_add(c, it); }
        c.n = 1;
        if (it < N) c.nf += 1; else c.nr += 1;
        __syncthreads();
        if (it < N && ccache_on(c) && c.ncache == c.nf - 1 && ccache_room(c, c.n + 1) > c.ncache) { ccache_fill<NT>(c, c.ncache, c.ncache + 1); c.ncache += 1; }
        return 0;
    }
    if (it < N) {
        const double* vcol = c.V + (size_t)it * N;
        const double* ccol = c.Ccol + (size_t)it * M0;
        for (int p = threadIdx.x; p < n; p += NT) { const int a = c.item[p]; c.colv[p] = (a < N) ? vcol[a] : ccol[a - N]; }
    } else {
        const double* crow = c.Crow + (size_t)(it - N) * N;
        for (int p = threadIdx.x; p < n; p += NT) { const int a = c.item[p]; c.colv[p] = (a < N) ? crow[a] : 0.0; }
    }
    __syncthreads();
    SSQP_TICK(c, T_AD_GATHER);
    symv<NT>(c, n, c.colv, c.hv);
    SSQP_TICK(c, T_AD_SYMV);
    double part = 0.0, apart = 0.0, spart = 0.0;
    for (int p = threadIdx.x; p < n; p += NT) {
        const double cv = c.colv[p];
        const double t = cv * c.hv[p];
        part += t; apart += fabs(t);
        spart += cv * c.sol[c.item[p]];
    }
    block_sum3<NT>(c, part, apart, spart);
    const double s = diag - part;
    // (a VARIABLE is never purged by the reference: its pivot only has to be a positive number that is not roundoff —
    // cholesky(V[F,F]), src/SSQP.jl:322; the soft threshold is for ROWS, which getRowsGJr may drop)
    if (!(fabs(s) > ((hard || it < N) ? PIV_HARD : PIV_SOFT) * (apart + fabs(diag)))) return 1;        // dependent on the items already in the system
    if ((hard || it < N) && ((it < N) != (s > 0.0))) return 1;      // indefinite: V_FF (s > 0) / the Schur complement (s < 0)
    const double is = 1.0 / s;
    const double tnew = (rnew - spart) * is;
    SSQP_TICK(c, T_AD_SUM);
    syr<NT>(c, n, c.hv, is);
    SSQP_TICK(c, T_AD_SYR);
    double* row = c.hrow(n);
    for (int p = threadIdx.x; p < n; p += NT) {
        const double h = c.hv[p];
        row[p] = -h * is;
        c.sol[c.item[p]] -= h * tnew;
    }
    if (threadIdx.x == 0) { row[n] = is; c.item[n] = it; c.pos[it] = n; c.sol[it] = tnew; list_add(c, it); }
    c.n = n + 1;
    if (it < N) c.nf += 1; else c.nr += 1;
    __syncthreads();
    // a released variable takes the next slot of the constraint-column cache (bulk copy, consumed by later trips)
    if (it < N && ccache_on(c) && c.ncache == c.nf - 1 && ccache_room(c, c.n + 1) > c.ncache) { ccache_fill<NT>(c, c.ncache, c.ncache + 1); c.ncache += 1; }
    SSQP_TICK(c, T_AD_TAIL);
    return 0;
}

// Remove item `it` from the system (and from c.sol).  Returns 0 ok, 1 singular (nothing changed).
template <int NT>
static __device__ int kinv_remove(Ctx& c, int it) {
    const int n = c.n;
    const int j = c.pos[it];
    const int lq = c.lpos[it];          // position in the free list / row list (read before thread 0 edits the lists)
    // the pivot test |piv| > PIV_SOFT * max|column| rides on the barrier behind the gather (__syncthreads_or of the
    // per-element comparisons) instead of a block-wide max reduction of its own
    const double piv = c.hrow(j)[j];
    int small = !(fabs(piv) > 0.0);
    {
        const double* rowj = c.hrow(j);
        const double thr = fabs(piv);
        for (int p = threadIdx.x; p < n; p += NT) {
            const double v = (p <= j) ? rowj[p] : c.hrow(p)[j];
            c.colv[p] = v;
            small |= !(thr > PIV_SOFT * fabs(v));
        }
    }
    if (__syncthreads_or(small)) return 1;      // (what is left would be dependent: through the purge)
    SSQP_TICK(c, T_RM_GATHER);
    const double f = c.sol[it] / piv;
    SSQP_TICK(c, T_RM_CHECK);
    syr<NT>(c, n, c.colv, -1.0 / piv);
    SSQP_TICK(c, T_RM_SYR);
    for (int p = threadIdx.x; p < n; p += NT) if (p != j) c.sol[c.item[p]] -= c.colv[p] * f;
    const int last = n - 1;
    if (j != last) {                    // move the last item into slot j (item[j] is only rewritten by thread 0 below,
        const double* lrow = c.hrow(last);      // after its own sol update; every other thread skips p == j)
        double* rowj = c.hrow(j);
        for (int k = threadIdx.x; k < last; k += NT) {
            if (k < j) rowj[k] = lrow[k];
            else if (k > j) c.hrow(k)[j] = lrow[k];
            else rowj[j] = lrow[last];
        }
    }
    if (threadIdx.x == 0) {
        if (j != last) { const int li = c.item[last]; c.item[j] = li; c.pos[li] = j; }
        c.pos[it] = -1; c.sol[it] = 0.0;
        list_remove(c, it);
    }
    c.n = last;
    if (it < c.N) c.nf -= 1; else c.nr -= 1;
    __syncthreads();
    if (it < c.N && lq < c.ncache) {          // the constraint-column cache follows the swap-remove of the free list
        const int lastpos = c.nf;             // (the variable that was last now sits at position lq)
        if (lastpos < c.ncache) {             // it was cached: move its slot
            ccache_wait(c);
            if (lq != lastpos) {
                const double* src = ccache_slot(c, lastpos); double* dst = ccache_slot(c, lq);
                for (int r = threadIdx.x; r < c.M0; r += NT) dst[r] = src[r];
            }
            c.ncache = lastpos;
            __syncthreads();
        } else {
            c.ncache = lq;                    // it was not (the cache is partial, hence unused): keep the prefix before the hole
        }
    }
    SSQP_TICK(c, T_RM_TAIL);
    return 0;
}

// getRowsGJr (src/utils.jl:49-86) on X = [AE bE] (src/SSQP.jl:295,310), restated on the device: Gauss-Jordan with
// in-row column pivoting over ALL remaining columns (the bE column included), first maximum on ties, pivot
// threshold `tol`.  Only run for degenerate working sets (more active rows than free variables, or a dependent row
// met while bordering).  X lives in the CTA's global workspace (the inverse is rebuilt right after).
// keep[r] (r < M0) = 1 for the rows the reference keeps.  Returns the number of kept rows, or -1 when the kept
// rows outnumber the free variables.
// Core of getRowsGJr on X (nr x nc, column-major with leading dimension nr, global memory): in-row column pivoting over
// all remaining columns (tracked through c0, the data is not moved), first maximum on ties, pivot threshold `tol`.
// sel[rowid ? rowid[i] : i] = 1 for every selected row i (the caller clears sel).  prow: nc doubles, pcolm: nr doubles.
// Returns the number of selected rows.
template <int NT>
static __device__ int gjr_core(Ctx& c, double* X, int nr, int nc, double tol, int* c0, double* prow, double* pcolm,
                               int* sel, const int* rowid) {
    for (int t = threadIdx.x; t < nc; t += NT) c0[t] = t;
    __syncthreads();
    int i = 0, j = 0, kept = 0;
    while (i < nr && j < nc) {
        Cand best;
        for (int t = j + threadIdx.x; t < nc; t += NT) best.offer(-fabs(X[i + (size_t)nr * c0[t]]), t);
        block_argmin<NT>(c, best);
        const double m = -best.key();
        if (!(m > tol)) { i += 1; continue; }
        const int mt = best.id;
        if (threadIdx.x == 0) { sel[rowid ? rowid[i] : i] = 1; const int a = c0[mt]; c0[mt] = c0[j]; c0[j] = a; }
        __syncthreads();
        const int ncol = c0[j];
        const double d = X[i + (size_t)nr * ncol];
        __syncthreads();
        for (int t = j + threadIdx.x; t < nc; t += NT) {
            double* e = X + i + (size_t)nr * c0[t];
            const double v = *e / d;
            *e = v; prow[t] = v;
        }
        for (int k = threadIdx.x; k < nr; k += NT) pcolm[k] = X[k + (size_t)nr * ncol];
        __syncthreads();
        const int span = nc - j;
        for (long long t = threadIdx.x; t < (long long)nr * span; t += NT) {
            const int k = (int)(t % nr), tt = j + (int)(t / nr);
            if (k != i) X[k + (size_t)nr * c0[tt]] -= pcolm[k] * prow[tt];
        }
        __syncthreads();
        kept += 1; i += 1; j += 1;
    }
    return kept;
}

template <int NT>
static __device__ int purge_rows_gjr(Ctx& c, int* keep) {
    const int N = c.N, M = c.M, M0 = c.M0;
    const double tol = c.P->tol;
    int* S = c.Sst;
    const int nf = block_compact<NT>(c, N, c.flist, [&](int k) { return S[k] == S_IN; });
    const int nr = block_compact<NT>(c, M0, c.evl, [&](int r) { return r < M || S[N + r - M] == S_EO; });
    const int nbz = block_compact<NT>(c, N, c.supp, [&](int k) { return S[k] != S_IN && c.z[k] != 0.0; });
    cpass<NT>(c, c.supp, nbz, c.z, c.rvec);                       // AB * zB  (all rows)
    const int nc = nf + 1;
    double* X = c.work;
    int* c0 = c.item;
    double* prow = c.rhs;      // normalised pivot row (by column slot)
    double* pcolm = c.lam;     // pivot column before elimination
    for (int t = threadIdx.x; t < nr * nc; t += NT) {
        const int w = t % nr, ci = t / nr;
        const int r = c.evl[w];
        X[t] = (ci < nf) ? c.Crow[c.flist[ci] + (size_t)r * N] : (c.bg[r] - c.rvec[r]);
    }
    for (int r = threadIdx.x; r < M0; r += NT) keep[r] = 0;
    __syncthreads();
    const int kept = gjr_core<NT>(c, X, nr, nc, tol, c0, prow, pcolm, keep, c.evl);
    // more kept rows than free variables (a pivot was taken in the bE column): AE*inv(V_FF)*AE' is singular and the
    // reference's cholesky throws PosDefException (src/SSQP.jl:328)
    return kept > nf ? -1 : kept;
}

constexpr int NDROPX = 5;    // purged EO rows whose x-vectors (AE' \ GE[j,F]) are kept for KKTchk! (c.pi .. c.sig); any further
                             // purged EO row gets its multiplier from a solve on the fly (dropped_lda)

// a purged row j still takes part in KKTchk! through alphaL' * (AE' \ GE[j,F]) (src/SSQP.jl:156-160): with
// [V_FF AE'; AE 0] [hp; x] = [GE[j,F]; 0] and GE[j,F] in the row space of AE, x is that least-squares solution.
// (the ids of the first NDROPX purged EO rows are in c.misc[8..]; their x-vectors go to c.pi .. c.sig)
template <int NT>
static __device__ void dropped_xvectors(Ctx& c, int droppedEO) {
    const int N = c.N, M0 = c.M0;
    const int nx = droppedEO < NDROPX ? droppedEO : NDROPX;
    for (int dd = 0; dd < nx; ++dd) {
        const int r = c.misc[8 + dd];
        const int n = c.n;
        const double* crow = c.Crow + (size_t)r * N;
        for (int p = threadIdx.x; p < n; p += NT) { const int a = c.item[p]; c.colv[p] = (a < N) ? crow[a] : 0.0; }
        __syncthreads();
        symv<NT>(c, n, c.colv, c.hv);
        double* xd = c.pi + (size_t)dd * c.M0p;
        for (int q = threadIdx.x; q < M0; q += NT) xd[q] = 0.0;
        __syncthreads();
        for (int p = threadIdx.x; p < n; p += NT) { const int a = c.item[p]; if (a >= N) xd[a - N] = c.hv[p]; }
        __syncthreads();
    }
}

// From-scratch build of the inverse for the current status vector, carrying the solution of the reduced system
// along (needs c.gr and c.slack fresh at the current z): border in the free variables in ascending order (V_FF is
// positive definite -> every pivot > 0), then the equality rows, then the EO rows ascending.
//   use_gj = false: every active row is expected to be independent; returns -2 when one is not (caller retries
//                   with use_gj = true).
//   use_gj = true : rows are first purged the way the reference does it (getRowsGJr, src/SSQP.jl:310-319); a row the
//                   reference keeps but that is dependent makes its Schur complement singular (PosDefException)
//                   -> returns -1.
// Returns the number of purged rows.  c.Bv[r] = 1 for the rows kept (valid until the next rebuild); purged EQUALITY rows
// take no part in KKTchk! (iR = findall(ra .> M), src/SSQP.jl:152); c.misc[7] = number of purged EO rows, the ids of the
// first NDROPX of them in c.misc[8..] with their x-vectors in c.pi .. c.sig.
template <int NT>
static __device__ int kinv_rebuild(Ctx& c, bool use_gj) {
    const int N = c.N, M = c.M, M0 = c.M0;
    int* keep = c.Bv;
    if (use_gj && purge_rows_gjr<NT>(c, keep) < 0) return -1;
    for (int i = threadIdx.x; i < N + M0; i += NT) { c.pos[i] = -1; c.sol[i] = 0.0; }
    c.n = 0; c.nf = 0; c.nr = 0;
    ccache_wait(c); c.ncache = 0;
    c.sol_valid = true;
    __syncthreads();
    for (int k = 0; k < N; ++k)
        if (c.Sst[k] == S_IN)
            if (kinv_add<NT>(c, k, -c.gr[k], true)) return -1;       // V_FF not positive definite (PosDefException)
    int dropped = 0, droppedEO = 0;
    for (int r = 0; r < M0; ++r)
        if (r < M || c.Sst[N + r - M] == S_EO) {
            if (use_gj && !keep[r]) {
                if (r >= M) {
                    if (threadIdx.x == 0 && droppedEO < NDROPX) c.misc[8 + droppedEO] = r;
                    droppedEO += 1;
                }
                dropped += 1;
                continue;
            }
            if (kinv_add<NT>(c, N + r, c.slack[r], use_gj)) return use_gj ? -1 : -2;
        }
    if (!use_gj) for (int r = threadIdx.x; r < M0; r += NT) keep[r] = 1;
    if (threadIdx.x == 0) c.misc[7] = droppedEO;
    __syncthreads();
    dropped_xvectors<NT>(c, droppedEO);
    return dropped;
}

// ---- from-scratch factorisation (north-star piece 1) ----------------------------------------------------------------
// What the reference does on EVERY trip (src/SSQP.jl:322-331) — iV = inv(cholesky(V[F,F])), C = inv(cholesky(AE*iV*AE')),
// TC = iV*AE'*C, VQ = iV - TC*AE*iV — done here ONCE per rebuild, in place on the packed storage of the inverse
//     H = [VQ TC; TC' -C]      (items: the K free variables ascending, then the W kept rows ascending)
// instead of K + W sequential bordered updates (each a CTA-wide symmetric GEMV + rank-1 update with half a dozen barriers):
//   1  rows 0..K-1 <- V_FF (lower), rows K+r <- AE[r,:]          6  L2 = chol(Cinv), L2i = L2^-1 (in place, W-part)
//   2  L = chol(V_FF)      right-looking, warp per trailing row   7  U = L2i Y'          (rows last to first)
//   3  Li = L^-1           row by row                             8  P = U Li            (row by row)
//   4  Y' = AE Li'         (rows K.., first K entries)            9  VQ = Li'Li - P'P    (rows ascending)
//   5  Cinv = Y' Y         (W-part of rows K..)                  10  TC' = L2i' P;  11  -C = -L2i'L2i
// Every in-place transform is two-phase (all outputs of a group of rows into registers, barrier, write), the groups ordered so
// that nothing is overwritten before its last use (proto/chol_build_proto.py is the same sequence in numpy).  Unlike the
// bordered updates, whose pivots are differences of O(cond) numbers, the Cholesky pivots lose digits like cond, not cond^2:
// this is also the path the drift guard falls back to.
// tri_chol / tri_inv work on the m x m lower triangle whose element (i, j) is hrow(base + i)[base + j].
template <int NT, class FC, class FW>
static __device__ __forceinline__ void two_phase(double* stage, int total, FC compute, FW write) {
#pragma unroll 1
    for (int idx = threadIdx.x; idx < total; idx += NT) stage[idx] = compute(idx);
    __syncthreads();
#pragma unroll 1
    for (int idx = threadIdx.x; idx < total; idx += NT) write(idx, stage[idx]);
    __syncthreads();
}
// in-place Cholesky (lower); dref[i] = the original diagonal, a pivot <= thr * dref[i] ends it: returns the failing index + 1, 0 ok
template <int NT>
static __device__ int tri_chol(Ctx& c, int base, int m, const double* dref, double thr) {
    constexpr int NW = NT / 32;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int j = 0; j < m; ++j) {
        const double d = c.hrow(base + j)[base + j];       // (ordered after the previous step's trailing update by its barrier)
        if (!(d > thr * dref[j]) || !(d > 0.0)) return j + 1;
        const double sd = sqrt(d), isd = 1.0 / sd;
        for (int i = j + 1 + threadIdx.x; i < m; i += NT) {
            double* e = c.hrow(base + i) + base + j;
            const double v = *e * isd;
            *e = v; c.colv[i] = v;
        }
        if (threadIdx.x == 0) c.hv[j] = sd;               // the diagonal goes in at the end: other threads may still be reading d
        __syncthreads();
        for (int i = j + 1 + w; i < m; i += NW) {         // trailing update, warp per row
            const double ci = c.colv[i];
            double* row = c.hrow(base + i) + base;
            for (int k = j + 1 + l; k <= i; k += 32) row[k] -= ci * c.colv[k];
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < m; j += NT) c.hrow(base + j)[base + j] = c.hv[j];
    __syncthreads();
    return 0;
}
// in-place inverse of a lower-triangular matrix, row by row: Li[i,j] = -(sum_{k=j}^{i-1} L[i,k] Li[k,j]) / L[i,i]
template <int NT>
static __device__ void tri_inv(Ctx& c, int base, int m) {
    for (int i = 0; i < m; ++i) {
        double* rowi = c.hrow(base + i) + base;
        const double dii = rowi[i];
        for (int k = threadIdx.x; k < i; k += NT) c.colv[k] = rowi[k];
        __syncthreads();
        for (int j = threadIdx.x; j < i; j += NT) {
            double a0 = 0.0, a1 = 0.0;
            int k = j;
            for (; k + 1 < i; k += 2) { a0 += c.colv[k] * c.hrow(base + k)[base + j]; a1 += c.colv[k + 1] * c.hrow(base + k + 1)[base + j]; }
            if (k < i) a0 += c.colv[k] * c.hrow(base + k)[base + j];
            rowi[j] = -(a0 + a1) / dii;
        }
        if (threadIdx.x == 0) rowi[i] = 1.0 / dii;
        __syncthreads();
    }
}

template <int NT> static __device__ double fresh_solve(Ctx& c, bool refine);

// Same contract as kinv_rebuild (needs c.gr and c.slack fresh at the current z; returns the number of purged rows, -1
// PosDefException, -2 a row looks dependent and use_gj was false).
template <int NT>
static __device__ int kinv_build_chol(Ctx& c, bool use_gj) {
    const int N = c.N, M = c.M, M0 = c.M0;
    int* keep = c.Bv;
    int* S = c.Sst;
    if (use_gj && purge_rows_gjr<NT>(c, keep) < 0) return -1;
    if (!use_gj) { for (int r = threadIdx.x; r < M0; r += NT) keep[r] = 1; }
    for (int i = threadIdx.x; i < N + M0; i += NT) { c.pos[i] = -1; c.sol[i] = 0.0; }
    ccache_wait(c); c.ncache = 0;
    __syncthreads();
    const int K = block_compact<NT>(c, N, c.flist, [&](int k) { return S[k] == S_IN; });
    const int W = block_compact<NT>(c, M0, c.rlist, [&](int r) { return (r < M || S[N + r - M] == S_EO) && keep[r]; });
    const int nact = block_compact<NT>(c, M0, c.evl, [&](int r) { return (r < M || S[N + r - M] == S_EO); });
    const int droppedEO = block_compact<NT>(c, M0, c.evl, [&](int r) { return r >= M && S[N + r - M] == S_EO && !keep[r]; });
    const int dropped = nact - W;
    if (threadIdx.x == 0) {
        c.misc[7] = droppedEO;
        for (int t = 0; t < droppedEO && t < NDROPX; ++t) c.misc[8 + t] = c.evl[t];
    }
    const int n = K + W;
    double* stage = c.work + c.P->stage_off;
    for (int p = threadIdx.x; p < n; p += NT) {
        const int it = (p < K) ? c.flist[p] : N + c.rlist[p - K];
        c.item[p] = it; c.pos[it] = p; c.lpos[it] = (p < K) ? p : p - K;
    }
    c.n = n; c.nf = K; c.nr = W;
    __syncthreads();
    // 1. V_FF (lower) and the rows of AE; the diagonal of V_FF kept for the pivot test
    for (int idx = threadIdx.x; idx < K * K; idx += NT) {
        const int i = idx / K, j = idx - i * K;
        if (j <= i) c.hrow(i)[j] = c.V[c.flist[i] + (size_t)c.flist[j] * N];
    }
    for (int idx = threadIdx.x; idx < W * K; idx += NT) {
        const int r = idx / K, k = idx - r * K;
        c.hrow(K + r)[k] = c.Crow[c.flist[k] + (size_t)c.rlist[r] * N];
    }
    for (int i = threadIdx.x; i < K; i += NT) c.rhs[i] = c.V[c.flist[i] + (size_t)c.flist[i] * N];
    __syncthreads();
    // 2, 3
    if (tri_chol<NT>(c, 0, K, c.rhs, PIV_HARD)) return -1;                     // cholesky(V[F,F]) throws PosDefException
    tri_inv<NT>(c, 0, K);
    if (W > 0) {
        const int RG = (TP_STAGE) / (K > 0 ? K : 1);                        // rows K+r per two-phase group
        // 4. Y' = AE Li'   (own row only: any group order)
        for (int r0 = 0; r0 < W; r0 += RG) {
            const int rg = (W - r0 < RG) ? W - r0 : RG;
            two_phase<NT>(stage, rg * K,
                [&](int idx) { const int r = r0 + idx / K, i = idx % K; const double* li = c.hrow(i); const double* a = c.hrow(K + r);
                               double a0 = 0.0, a1 = 0.0; int k = 0;
                               
_Pragma("unroll 1")
                               for (; k + 1 <= i; k += 2) { a0 += li[k] * a[k]; a1 += li[k + 1] * a[k + 1]; }
                               if (k <= i) a0 += li[k] * a[k];
                               return a0 + a1; },
                [&](int idx, double v) { c.hrow(K + r0 + idx / K)[idx % K] = v; });
        }
        // 5. Cinv = Y' Y (lower) into the W-part; its diagonal kept for the pivot test
        for (int idx = threadIdx.x; idx < W * W; idx += NT) {
            const int r = idx / W, q = idx - r * W;
            if (q > r) continue;
            const double* yr = c.hrow(K + r); const double* yq = c.hrow(K + q);
            double a0 = 0.0, a1 = 0.0; int k = 0;
            
_Pragma("unroll 1")
                               for (; k + 1 < K; k += 2) { a0 += yr[k] * yq[k]; a1 += yr[k + 1] * yq[k + 1]; }
            if (k < K) a0 += yr[k] * yq[k];
            c.hrow(K + r)[K + q] = a0 + a1;
            if (q == r) c.rhs[r] = a0 + a1;
        }
        __syncthreads();
        // 6. a row the reference's getRowsGJr would have to look at shows up as a small pivot of chol(AE iV AE')
        if (tri_chol<NT>(c, K, W, c.rhs, use_gj ? PIV_HARD : PIV_SOFT)) return use_gj ? -1 : -2;
        tri_inv<NT>(c, K, W);
        // 7. U = L2i Y'   (row r reads rows q <= r: groups from the last rows to the first)
        for (int r1 = W; r1 > 0; r1 -= RG) {
            const int r0 = (r1 > RG) ? r1 - RG : 0, rg = r1 - r0;
            two_phase<NT>(stage, rg * K,
                [&](int idx) { const int r = r0 + idx / K, cc = idx % K; const double* l2 = c.hrow(K + r) + K;
                               double a0 = 0.0;
_Pragma("unroll 1")
                               for (int q = 0; q <= r; ++q) a0 += l2[q] * c.hrow(K + q)[cc];
                               return a0; },
                [&](int idx, double v) { c.hrow(K + r0 + idx / K)[idx % K] = v; });
        }
        // 8. P = U Li   (own row only)
        for (int r0 = 0; r0 < W; r0 += RG) {
            const int rg = (W - r0 < RG) ? W - r0 : RG;
            two_phase<NT>(stage, rg * K,
                [&](int idx) { const int r = r0 + idx / K, j = idx % K; const double* u = c.hrow(K + r);
                               double a0 = 0.0, a1 = 0.0; int i = j;
                               
_Pragma("unroll 1")
                               for (; i + 1 < K; i += 2) { a0 += u[i] * c.hrow(i)[j]; a1 += u[i + 1] * c.hrow(i + 1)[j]; }
                               if (i < K) a0 += u[i] * c.hrow(i)[j];
                               return a0 + a1; },
                [&](int idx, double v) { c.hrow(K + r0 + idx / K)[idx % K] = v; });
        }
    }
    // 9. VQ = Li'Li - P'P   (row i reads rows m >= i: groups of rows ascending; outputs (i, j <= i) laid out i1-wide)
    for (int i0 = 0; i0 < K;) {
        int i1 = i0 + 1;
        while (i1 < K && (i1 + 1 - i0) * (i1 + 1) <= TP_STAGE) ++i1;
        const int wd = i1;                                 // columns 0 .. i1-1 per row of the group
        two_phase<NT>(stage, (i1 - i0) * wd,
            [&](int idx) { const int i = i0 + idx / wd, j = idx % wd;
                           if (j > i) return 0.0;
                           double a0 = 0.0, a1 = 0.0;
_Pragma("unroll 1")
                           for (int m2 = i; m2 < K; ++m2) { const double* rm = c.hrow(m2); a0 += rm[i] * rm[j]; }
_Pragma("unroll 1")
                           for (int r = 0; r < W; ++r) { const double* pr = c.hrow(K + r); a1 += pr[i] * pr[j]; }
                           return a0 - a1; },
            [&](int idx, double v) { const int i = i0 + idx / wd, j = idx % wd; if (j <= i) c.hrow(i)[j] = v; });
        i0 = i1;
    }
    if (W > 0) {
        const int RG = (TP_STAGE) / (K > 0 ? K : 1);
        // 10. TC' = L2i' P   (row r reads rows q >= r: groups ascending)
        for (int r0 = 0; r0 < W; r0 += RG) {
            const int rg = (W - r0 < RG) ? W - r0 : RG;
            two_phase<NT>(stage, rg * K,
                [&](int idx) { const int r = r0 + idx / K, cc = idx % K;
                               double a0 = 0.0;
_Pragma("unroll 1")
                               for (int q = r; q < W; ++q) { const double* rq = c.hrow(K + q); a0 += rq[K + r] * rq[cc]; }
                               return a0; },
                [&](int idx, double v) { c.hrow(K + r0 + idx / K)[idx % K] = v; });
        }
        // 11. -C = -L2i'L2i   (W <= M0: one group as long as W*W <= TP_MAX*NT, else by rows ascending)
        for (int i0 = 0; i0 < W;) {
            int i1 = i0 + 1;
            while (i1 < W && (i1 + 1 - i0) * (i1 + 1) <= TP_STAGE) ++i1;
            const int wd = i1;
            two_phase<NT>(stage, (i1 - i0) * wd,
                [&](int idx) { const int i = i0 + idx / wd, j = idx % wd;
                               if (j > i) return 0.0;
                               double a0 = 0.0;
_Pragma("unroll 1")
                               for (int m2 = i; m2 < W; ++m2) { const double* rm = c.hrow(K + m2) + K; a0 += rm[i] * rm[j]; }
                               return -a0; },
                [&](int idx, double v) { const int i = i0 + idx / wd, j = idx % wd; if (j <= i) c.hrow(K + i)[K + j] = v; });
            i0 = i1;
        }
    }
    // the solution of the reduced system at the current point, then the multipliers' helpers of the purged rows
    fresh_solve<NT>(c, false);
    dropped_xvectors<NT>(c, droppedEO);
    return dropped;
}

// ---- Phase 1: initQP + cDantzigLP ------------------------------------------------------------------
// returns 1 feasible, 0 infeasible, -1 numerical; fills c.z (x0) and c.Sst[0..N+J)
// simplex_loop(mode): the pivot loop of cDantzigLP (src/Simplex.jl:486-607) from the basis in c.Bv / invB / c.qB / S1.
//   mode 0: Phase-1 costs (1 on the artificials, src/SSQP.jl:524 / Simplex.jl:917)         columns 0 .. N1-1
//   mode 1: the LP's own costs c.q on the structurals, 0 on the slacks (SimplexLP Phase 2)  columns 0 .. N0-1
// Returns 0 when no candidate is left (mode 1: 1 unique optimum / 2 some nonbasic reduced cost vanishes), 3 unbounded.
// invB (M0 x M0, column-major): in shared memory with an ODD leading dimension (rows and columns both conflict-free)
// when it fits next to the solver's vectors; otherwise in the CTA's global workspace (L2) with the leading dimension
// rounded up to 4, so that its columns take the 256-bit streaming path (config 5: M0 = 200 -> 321 KB).
__device__ __forceinline__ bool invb_in_smem(const Ctx& c) { return (long long)(c.M0 | 1) * c.M0 <= (long long)c.P->hcap; }
__device__ __forceinline__ int invb_ld(const Ctx& c) { return invb_in_smem(c) ? (c.M0 | 1) : rup(c.M0, 4); }
__device__ __forceinline__ double* invb_ptr(const Ctx& c) { return invb_in_smem(c) ? c.Hs : c.work; }

// Phase-1 start (src/SSQP.jl:511-526, the same construction as src/Simplex.jl:905-920): all-artificial basis
// invB = diag(+-1), x at the lower bounds, x_B = |A0 d0 - b0|
template <int NT>
static __device__ void simplex_init(Ctx& c) {
    const int N = c.N, J = c.J, M0 = c.M0;
    const int N0 = N + J + c.nfree, N1 = N0 + M0;      // columns: structurals, slacks, 2nd halves of the free variables, artificials
    const int ldB = invb_ld(c);
    double* invB = invb_ptr(c);
    int* S1 = c.Sst;
    for (int k = threadIdx.x; k < N1; k += NT) S1[k] = (k >= N0) ? S_IN : S_DN;
    for (int j = threadIdx.x; j < M0; j += NT) c.Bv[j] = N0 + j;
    // x at the lower bounds d0 of the transformed LP, held in the ORIGINAL variables: a (-Inf,u] column is negated and
    // starts at d0 = -u, i.e. x = u (the products A0[:,k]*d0[k] are the same bits); free variables start at 0
    for (int k = threadIdx.x; k < N; k += NT) c.z[k] = c.xform ? c.gr[k] * c.d[k] : c.d[k];
    __syncthreads();
    if (M0 == 0) return;
    // q0 = A0*d0 ; sig ; qB = |q0 - b0|                                  (src/SSQP.jl:516-521)
    int cnt = compact_nonzero<NT>(c, c.z, N, c.supp);
    cpass<NT>(c, c.supp, cnt, c.z, c.rvec);
    for (int j = threadIdx.x; j < M0; j += NT) {
        double q0 = c.rvec[j];
        c.sig[j] = (c.bg[j] >= q0) ? 1.0 : -1.0;
        c.qB[j] = fabs(q0 - c.bg[j]);
    }
    for (int t = threadIdx.x; t < ldB * M0; t += NT) invB[t] = 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < M0; j += NT) invB[j + (size_t)j * ldB] = c.sig[j];
    __syncthreads();

}

template <int NT>
static __device__ int simplex_loop(Ctx& c, const int mode, long long& loop, long long& pivots) {
    const int N = c.N, M = c.M, J = c.J, M0 = c.M0;
    const int NJ = N + J, N0 = NJ + c.nfree, N1 = N0 + M0;
    const int NC = (mode == 0) ? N1 : N0;          // columns of the LP being solved
    const bool xf = c.xform;
    const double* sgn = c.gr;   // xf: column signs of the structurals (-1: (-Inf,u] variable, column negated, src/SSQP.jl:506-509)
    const int* ivl = c.flist;   // ids of the free variables; column NJ + t is -A0[:, ivl[t]] with bounds [0, Inf)  (:493-503)
    const double tol = c.P->tolLP;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int ldB = invb_ld(c);
    double* invB = invb_ptr(c);
    const bool binv_global = !invb_in_smem(c);
    // (measured and dropped: the shared-memory case through a pointer the compiler knows is shared — LDS/STS instead of
    // generic LD/ST on every element of the per-pivot passes — 1.6 % slower)
    int* S1 = c.Sst;            // N1 statuses: structurals, slacks, artificials
    double* Api = c.pfull;      // [A;G]' pi over the structurals
    const double* cost = c.q;   // mode 1: structural costs
    // pi = invB' c_B : sum of the rows of invB whose basic variable is an artificial   (Simplex.jl:600).  Computed once;
    // a pivot then moves it by (reduced cost of the entering variable) x (new pivot row of invB), and x_B = q by
    // -(step) x (pivot column) — the textbook O(M0) updates of the quantities the reference recomputes from scratch.
    {
        const int* Bv = c.Bv;
        if (mode == 0)
            small_reduce<NT>(c, M0, M0, [=](int i, int j) { return (Bv[j] >= N0) ? invB[j + (size_t)i * ldB] : 0.0; }, c.pi);
        else
            small_reduce<NT>(c, M0, M0, [=](int i, int j) {
                const int bj = Bv[j];
                const double cb = (bj < N) ? (xf ? sgn[bj] * cost[bj] : cost[bj]) : (bj >= NJ && bj < N0) ? -cost[ivl[bj - NJ]] : 0.0;   // c0 (Simplex.jl:956-959)
                return (cb != 0.0) ? cb * invB[j + (size_t)i * ldB] : 0.0; }, c.pi);
    }
    int anyzero = 0;            // mode 1: some nonbasic reduced cost is (numerically) zero at the end -> status 2
    while (true) {
        const bool bland = (loop + 1) > NC;            // loop += 1; if loop > N: Bland  (Simplex.jl:487-490)
        // pricing: h > tol candidates; largest-distance Dantzig  argmax(hp ./ cA)  (Simplex.jl:495)
        Cand best;
        int zpart = 0;
        auto price = [&](int k, Cand& b) {
            const int st = S1[k];
            if (st == S_IN) return;
            double rc, ca = 1.0;
            if (k < N) { const double r0 = (mode == 0 ? 0.0 : cost[k]) - Api[k]; rc = xf ? sgn[k] * r0 : r0; ca = c.cA[k]; }
            else if (k < NJ) rc = -c.pi[M + (k - N)];
            else if (k < N0) { const int v = ivl[k - NJ]; rc = (mode == 0 ? 0.0 : -cost[v]) + Api[v]; ca = c.cA[v]; }
            else rc = 1.0 - c.sig[k - N0] * c.pi[k - N0];
            const double h = (st == S_DN) ? -rc : rc;
            if (fabs(h) < tol) zpart = 1;
            if (h > tol) b.offer(bland ? 0.0 : -(h / ca), k);           // arg-max == arg-min of the negated score
        };
        constexpr int CH = 256;
        if (bland && N >= 2 * CH) {
            // Bland's rule takes the candidate of LOWEST index: [A;G]' pi is formed CH structurals at a time and the pass
            // stops at the first chunk that holds a candidate (the tail of config 5's LPs spends 99 % of its loops here and
            // the full 1.6 MB pass per loop is what saturates L2); no candidate among the structurals -> every chunk was
            // formed, and the slacks / second halves / artificials are priced as usual.
            const long long tp_ = SSQP_CLK();
            bool found = false;
            int done_rows = 0;
            for (int k0 = 0; k0 < N && !found; k0 += CH) {
                const int rows = (N - k0 < CH) ? N - k0 : CH;
                gemv_cols<NT>(GemvArgs{c.Crow + k0, N, -1, soff(c.pi), M0, rows, nullptr, -1, soff(Api) + k0, soff(c.buf), c.bufsz, -1});
                Cand cb;
                for (int k = k0 + threadIdx.x; k < k0 + rows; k += NT) price(k, cb);
                block_argmin<NT>(c, cb);
                if (cb.any()) { best = cb; found = true; }
                done_rows = k0 + rows;
            }
            if (threadIdx.x == 0) { c.bytes += 8.0 * done_rows * M0; c.cyc[CY_P1PRICE] += SSQP_CLK() - tp_; }
            if (!found) {
                for (int k = N + threadIdx.x; k < NC; k += NT) price(k, best);
                block_argmin<NT>(c, best);
            }
        } else {
            // [A;G]' pi over the structurals: one streaming pass over Crow (N x M0, L2)
            const long long tp_ = SSQP_CLK();
            gemv_cols<NT>(GemvArgs{c.Crow, N, -1, soff(c.pi), M0, N, nullptr, -1, soff(Api), soff(c.buf), c.bufsz, -1});
            if (threadIdx.x == 0) { c.bytes += 8.0 * N * M0; c.cyc[CY_P1PRICE] += SSQP_CLK() - tp_; }
            SSQP_TICK(c, T_VPASS);       // (Phase-1 timeline: pricing pass)
            for (int k = threadIdx.x; k < NC; k += NT) price(k, best);
            block_argmin<NT>(c, best);
            SSQP_TICK(c, T_KKT);         // (Phase-1 timeline: pricing loop + arg-max)
        }
        if (!best.any()) {
            if (mode == 1) anyzero = (block_max<NT>(c, (double)zpart) > 0.0);       // ms = any(abs.(h) .< tol)  (Simplex.jl:612)
            break;
        }
        loop += 1;
        const int kin = best.id;
        const double rc_kin = (kin < N) ? (xf ? sgn[kin] : 1.0) * ((mode == 0 ? 0.0 : cost[kin]) - Api[kin])
                            : (kin < NJ) ? -c.pi[M + (kin - N)]
                            : (kin < N0) ? (mode == 0 ? 0.0 : -cost[ivl[kin - NJ]]) + Api[ivl[kin - NJ]]
                            : 1.0 - c.sig[kin - N0] * c.pi[kin - N0];
        // p = invB * A1[:,kin]                                                            (Simplex.jl:497)
        if (kin < N || (kin >= NJ && kin < N0)) {
            const double* col = c.Ccol + (size_t)(kin < N ? kin : ivl[kin - NJ]) * M0;
            if (!xf) for (int i = threadIdx.x; i < M0; i += NT) c.rvec[i] = col[i];
            else {
                const double sg = (kin < N) ? sgn[kin] : -1.0;
                for (int i = threadIdx.x; i < M0; i += NT) c.rvec[i] = sg * col[i];
            }
            __syncthreads();
            const double* rv = c.rvec;
            if (binv_global)         // invB in L2: its columns stream like any other column-major operand
                gemv_cols<NT>(GemvArgs{invB, ldB, -1, soff(c.rvec), M0, M0, nullptr, -1, soff(c.pcol), soff(c.buf), c.bufsz, -1});
            else
                small_reduce<NT>(c, M0, M0, [=](int j, int i) { return invB[j + (size_t)i * ldB] * rv[i]; }, c.pcol);
        } else {
            const int ci = (kin < NJ) ? (M + kin - N) : (kin - N0);
            const double sg = (kin < NJ) ? 1.0 : c.sig[kin - N0];
            for (int j = threadIdx.x; j < M0; j += NT) c.pcol[j] = sg * invB[j + (size_t)ci * ldB];
            __syncthreads();
        }
        SSQP_TICK(c, T_AD_SYMV);         // (Phase-1 timeline: p = invB * A1[:,kin])
        // ratio test (Simplex.jl:499-569): arg-min/arg-max over basis rows, ties -> lowest basic variable id
        const bool kd = (S1[kin] == S_DN);
        const double lo_k = (kin < N) ? c.d[kin] : 0.0;
        const double hi_k = (kin < N) ? c.u[kin] : INF;
        const bool fu = hi_k < INF;
        Cand rbest;
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
            if (has) rbest.offer(kd ? gt : -gt, i);
        }
        block_argmin<NT>(c, rbest);
        SSQP_TICK(c, T_RATIO);           // (Phase-1 timeline: ratio test)
        const int rid = rbest.any() ? rbest.id : -1;
        const double rkey = rbest.key();
        int action;      // -1 flip to UP, -2 flip to DN, >=0 pivot on the row of basic variable rid
        if (kd) {
            if (rid < 0) {
                if (fu) action = -1; else return 3;                          // unbounded  (Simplex.jl:523-526)
            } else {
                const double gl = rkey;
                if (fu) action = (gl >= hi_k - lo_k) ? -1 : 0;
                else { if (isinf(gl)) return 3; action = 0; }
            }
        } else {
            if (rid < 0) action = -2;
            else { const double gl = -rkey; action = (gl <= -(hi_k - lo_k)) ? -2 : 0; }
        }
        // x_B moves along -p by the step of the entering variable (a bound flip is a step of the full range)
        const double xold_k = kd ? lo_k : hi_k;
        const double gstep = (action == -1) ? (hi_k - lo_k) : (action == -2) ? -(hi_k - lo_k) : (kd ? rkey : -rkey);
        for (int j = threadIdx.x; j < M0; j += NT) c.qB[j] -= gstep * c.pcol[j];
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
            SSQP_TICK(c, T_STEP);        // (Phase-1 timeline: step of x_B, leaving row)
            // product-form update: row l /= p_l ; row j -= p_j * row l   (pivot row stashed first)
            const double ipl = 1.0 / pj;
            for (int i = threadIdx.x; i < M0; i += NT) c.rvec[i] = invB[lrow + (size_t)i * ldB] * ipl;
            __syncthreads();
            SSQP_TICK(c, T_RM_GATHER);   // (Phase-1 timeline: pivot row)
            if (binv_global && (M0 & 3) == 0) {
                // invB in L2: 256-bit read-modify-write of four rows x one column per step, four steps in flight
                const int G4 = M0 >> 2;
                const int total = G4 * M0;
                for (int idx0 = threadIdx.x; idx0 < total; idx0 += 4 * NT) {
                    double v[4][4]; int jg[4], ii[4]; bool ok[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int idx = idx0 + e * NT;
                        ok[e] = idx < total;
                        const int id2 = ok[e] ? idx : 0;
                        ii[e] = id2 / G4; jg[e] = id2 - ii[e] * G4;
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[e][q] = 0.0;
                        VecLd<4>::ldp(invB + 4 * jg[e] + (size_t)ii[e] * ldB, v[e], ok[e] ? 1 : 0);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (!ok[e]) continue;
                        const double rv2 = c.rvec[ii[e]];
                        double4 o;
                        const int j0 = 4 * jg[e];
                        o.x = (j0 == lrow) ? rv2 : v[e][0] - c.pcol[j0] * rv2;
                        o.y = (j0 + 1 == lrow) ? rv2 : v[e][1] - c.pcol[j0 + 1] * rv2;
                        o.z = (j0 + 2 == lrow) ? rv2 : v[e][2] - c.pcol[j0 + 2] * rv2;
                        o.w = (j0 + 3 == lrow) ? rv2 : v[e][3] - c.pcol[j0 + 3] * rv2;
                        *reinterpret_cast<double4*>(invB + j0 + (size_t)ii[e] * ldB) = o;
                    }
                }
            } else {
                const int Wd = c.M0p;
                const int cstep = Wd <= NT ? fastdiv(NT, Wd) : 1;
                if (Wd <= NT) {
                    const int i0 = fastdiv(threadIdx.x, Wd), jj = threadIdx.x - i0 * Wd;
                    if (jj < M0 && i0 < cstep) {
                        // four columns per step, loads first: written as `*e = *e - p * r[ii]` in a plain loop every store
                        // may alias the next r[ii] (two generic pointers) and the iterations serialise (4.8 k cycles per pivot
                        // instead of 2.8 k).
                        const double pjj = (jj == lrow) ? 0.0 : c.pcol[jj];
                        const bool piv = (jj == lrow);
                        const double* rv = smem_d + soff(c.rvec);
                        auto upd = [&](double* col) {
                            int ii = i0;
                            for (; ii + 3 * cstep < M0; ii += 4 * cstep) {
                                double* e0 = col + (size_t)ii * ldB; double* e1 = e0 + (size_t)cstep * ldB;
                                double* e2 = e1 + (size_t)cstep * ldB; double* e3 = e2 + (size_t)cstep * ldB;
                                const double r0 = rv[ii], r1 = rv[ii + cstep], r2 = rv[ii + 2 * cstep], r3 = rv[ii + 3 * cstep];
                                const double v0 = *e0, v1 = *e1, v2 = *e2, v3 = *e3;
                                *e0 = piv ? r0 : v0 - pjj * r0; *e1 = piv ? r1 : v1 - pjj * r1;
                                *e2 = piv ? r2 : v2 - pjj * r2; *e3 = piv ? r3 : v3 - pjj * r3;
                            }
                            for (; ii < M0; ii += cstep) {
                                double* e = col + (size_t)ii * ldB;
                                const double r = rv[ii];
                                *e = piv ? r : *e - pjj * r;
                            }
                        };
                        upd(invB + jj);      // (measured: a copy of the loop on a shared-memory-typed pointer — LDS/STS instead of
                                             //  generic LD/ST — is slower, 3.3 k vs 2.8 k cycles; the pass moves 160 KB through the
                                             //  128 B/clk shared-memory pipe: floor ~1.6 k)
                    }
                } else {
                    for (int t = threadIdx.x; t < M0 * M0; t += NT) {
                        const int jj = t % M0, ii = t / M0;
                        double* e = invB + jj + (size_t)ii * ldB;
                        *e = (jj == lrow) ? c.rvec[ii] : *e - c.pcol[jj] * c.rvec[ii];
                    }
                }
            }
            SSQP_TICK(c, T_RM_CHECK);    // (Phase-1 timeline: rank-1 update of invB, thread 0's share)
            for (int i = threadIdx.x; i < M0; i += NT) c.pi[i] += rc_kin * c.rvec[i];        // duals follow the pivot
            SSQP_TICK(c, T_RM_SYR);      // (Phase-1 timeline, developer build: basis-inverse update + duals)
            __syncthreads();       // (every thread has applied its qB update before slot lrow is overwritten)
            if (threadIdx.x == 0) { c.Bv[lrow] = kin; S1[kin] = S_IN; S1[rid] = Sl; c.qB[lrow] = xold_k + gstep; }
            pivots += 1;
        }
        __syncthreads();
        SSQP_TICK(c, T_RM_TAIL);         // (Phase-1 timeline: end of the loop)
    }
    return (mode == 1) ? (anyzero ? 2 : 1) : 0;
}

#ifndef SSQP_ONLY_VW4
// ---- the reference's other pivot rules (Settings.rule, src/types.jl:397): stpEdgeLP (src/Simplex.jl:234-416, the
// second definition — the one in effect) and maxImprvLP (src/Simplex.jl:641-813).  Both look at Y[:,k] = invB*A1[:,k]
// of EVERY candidate column on every loop (a norm for the steepest edge, a full ratio test for the greatest
// improvement), so a loop costs one basis-inverse GEMV per candidate instead of one in total; the CTA walks the
// candidates in ascending order and keeps the first best, which is the reference's argmax / findmax.  Same data and
// pivot update as simplex_loop; compiled into the general kernel flavour only (the host selects it when rule != 0).
// rule 1 = :stpEdgeLP, 2 = :maxImprovement.  Returns like simplex_loop (0 / 1 / 2 / 3), -1 at the loop cap.
template <int NT>
static __device__ int simplex_loop_alt(Ctx& c, const int mode, const int rule, long long& loop, long long& pivots) {
    const int N = c.N, M = c.M, J = c.J, M0 = c.M0;
    const int NJ = N + J, N0 = NJ + c.nfree, N1 = N0 + M0;
    const int NC = (mode == 0) ? N1 : N0;
    const bool xf = c.xform;
    const double* sgn = c.gr;
    const int* ivl = c.flist;
    const double tol = c.P->tolLP;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const int ldB = invb_ld(c);
    double* invB = invb_ptr(c);      // (in the CTA's global workspace when it does not fit in shared memory: generic accesses, slower)
    int* S1 = c.Sst;
    double* Api = c.pfull;
    const double* cost = c.q;
    double* hval = c.hv;         // signed reduced cost h of every column (0 for basic ones): NC doubles (hv + colv)
    int* cand = c.lpos;          // candidates h > tol, ascending: NC ints (lpos + evl)
    {
        const int* Bv = c.Bv;
        if (mode == 0)
            small_reduce<NT>(c, M0, M0, [=](int i, int j) { return (Bv[j] >= N0) ? invB[j + (size_t)i * ldB] : 0.0; }, c.pi);
        else
            small_reduce<NT>(c, M0, M0, [=](int i, int j) {
                const int bj = Bv[j];
                const double cb = (bj < N) ? (xf ? sgn[bj] * cost[bj] : cost[bj]) : (bj >= NJ && bj < N0) ? -cost[ivl[bj - NJ]] : 0.0;
                return (cb != 0.0) ? cb * invB[j + (size_t)i * ldB] : 0.0; }, c.pi);
    }
    // p = invB * A1[:,k] into c.pcol
    auto column_p = [&](int k) {
        if (k < N || (k >= NJ && k < N0)) {
            const double* col = c.Ccol + (size_t)(k < N ? k : ivl[k - NJ]) * M0;
            const double sg = (k < N) ? (xf ? sgn[k] : 1.0) : -1.0;
            for (int i = threadIdx.x; i < M0; i += NT) c.rvec[i] = sg * col[i];
            __syncthreads();
            const double* rv = c.rvec;
            small_reduce<NT>(c, M0, M0, [=](int j, int i) { return invB[j + (size_t)i * ldB] * rv[i]; }, c.pcol);
        } else {
            const int ci = (k < NJ) ? (M + k - N) : (k - N0);
            const double sg = (k < NJ) ? 1.0 : c.sig[k - N0];
            __syncthreads();
            for (int j = threadIdx.x; j < M0; j += NT) c.pcol[j] = sg * invB[j + (size_t)ci * ldB];
            __syncthreads();
        }
    };
    // ratio test of c.pcol for entering variable k.  variant 0: cDantzigLP / stpEdgeLP form, 1: maxImprvLP form (:682-751).
    // action: -1 flip to UP, -2 flip to DN, 0 pivot (rid = leaving variable), 3 unbounded; gl = signed step of x_k.
    auto ratio = [&](int k, int variant, int& rid, double& gl) -> int {
        const bool kd = (S1[k] == S_DN);
        const double lo_k = (k < N) ? c.d[k] : 0.0, hi_k = (k < N) ? c.u[k] : INF;
        const bool fu = hi_k < INF;
        Cand rb;
        for (int j = threadIdx.x; j < M0; j += NT) {
            const int i = c.Bv[j];
            const double pj = c.pcol[j];
            const double lo = (i < N) ? c.d[i] : 0.0, hi = (i < N) ? c.u[i] : INF;
            double gt; bool has = false;
            if (kd) {
                if (pj > tol) { gt = (c.qB[j] - lo) / pj; has = true; }
                else if (pj < -tol) { gt = (c.qB[j] - hi) / pj; has = true; }
            } else {
                if (pj > tol && (variant == 0 || hi < INF)) { gt = (c.qB[j] - hi) / pj; has = true; }
                else if (pj < -tol) { gt = (c.qB[j] - lo) / pj; has = true; }
            }
            if (has) rb.offer(kd ? gt : -gt, i);
        }
        block_argmin<NT>(c, rb);
        rid = rb.any() ? rb.id : -1;
        const double key = rb.key();
        if (kd) {
            if (rid < 0) { if (fu) { gl = hi_k - lo_k; return -1; } return 3; }
            if (fu) { if (key >= hi_k - lo_k) { gl = hi_k - lo_k; return -1; } }
            else if (isinf(key)) return 3;
            gl = key;
            return 0;
        }
        if (rid < 0) { gl = -(hi_k - lo_k); return -2; }
        if (-key <= -(hi_k - lo_k)) { gl = -(hi_k - lo_k); return -2; }
        gl = -key;
        return 0;
    };
    int anyzero = 0;
    bool Edge = true;
    while (true) {
        gemv_cols<NT>(GemvArgs{c.Crow, N, -1, soff(c.pi), M0, N, nullptr, -1, soff(Api), soff(c.buf), c.bufsz, -1});
        int zpart = 0;
        for (int k = threadIdx.x; k < NC; k += NT) {
            const int st = S1[k];
            double h = 0.0;
            if (st != S_IN) {
                double rc;
                if (k < N) { const double r0 = (mode == 0 ? 0.0 : cost[k]) - Api[k]; rc = xf ? sgn[k] * r0 : r0; }
                else if (k < NJ) rc = -c.pi[M + (k - N)];
                else if (k < N0) { const int v = ivl[k - NJ]; rc = (mode == 0 ? 0.0 : -cost[v]) + Api[v]; }
                else rc = 1.0 - c.sig[k - N0] * c.pi[k - N0];
                h = (st == S_DN) ? -rc : rc;
                if (fabs(h) < tol) zpart = 1;
            }
            hval[k] = h;
        }
        __syncthreads();
        const int nH = block_compact<NT>(c, NC, cand, [&](int k) { return hval[k] > tol; });
        if (nH == 0) {
            if (mode == 1) anyzero = (block_max<NT>(c, (double)zpart) > 0.0);
            break;
        }
        loop += 1;
        if (loop > 100000LL + 50LL * NC) return -1;      // neither rule has an anti-cycling safeguard in the reference: do not hang the device
        int kin = cand[0], action = 0, rid = -1;
        double gl = 0.0;
        if (rule == 1) {
            if (Edge) {       // se = hp.^2 ./ (1 .+ sum(Y[:,ih].^2)) ; argmax keeps the first maximum  (:278-280)
                double best = -1.0;
                for (int t = 0; t < nH; ++t) {
                    const int k = cand[t];
                    column_p(k);
                    double part = 0.0;
                    for (int j = threadIdx.x; j < M0; j += NT) part += c.pcol[j] * c.pcol[j];
                    const double y = block_sum<NT>(c, part) + 1.0;
                    const double se = hval[k] * hval[k] / y;
                    if (t == 0 || se > best) { best = se; kin = k; }
                }
            }
            column_p(kin);
            action = ratio(kin, 0, rid, gl);
            if (action == 3) return 3;
            if (action == 0 && Edge && fabs(gl) < tol) { Edge = false; continue; }     // zero step: first candidate once (:377-380)
        } else {
            double best = -1.0; int bact = 0, brid = -1; double bgl = 0.0;
            for (int t = 0; t < nH; ++t) {       // a ratio test per candidate; k = argmax |h .* g|  (:676-757)
                const int k = cand[t];
                column_p(k);
                int r1; double g1;
                const int a1 = ratio(k, 1, r1, g1);
                if (a1 == 3) return 3;
                const double sc = fabs(hval[k] * g1);
                if (t == 0 || sc > best) { best = sc; kin = k; bact = a1; brid = r1; bgl = g1; }
            }
            column_p(kin);
            action = bact; rid = brid; gl = bgl;
        }
        // step of the entering variable, then flip or pivot (same updates as simplex_loop)
        const bool kd = (S1[kin] == S_DN);
        const double lo_k = (kin < N) ? c.d[kin] : 0.0, hi_k = (kin < N) ? c.u[kin] : INF;
        const double xold_k = kd ? lo_k : hi_k;
        double rc_kin;
        {
            const double hk = hval[kin];
            rc_kin = kd ? -hk : hk;
        }
        for (int j = threadIdx.x; j < M0; j += NT) c.qB[j] -= gl * c.pcol[j];
        if (action == -1) { if (threadIdx.x == 0) S1[kin] = S_UP; }
        else if (action == -2) { if (threadIdx.x == 0) S1[kin] = S_DN; }
        else {
            int lrow = -1;
            for (int j = threadIdx.x; j < M0; j += NT) if (c.Bv[j] == rid) lrow = j;
            if (lrow >= 0) c.misc[0] = lrow;
            __syncthreads();
            lrow = c.misc[0];
            const double pj = c.pcol[lrow];
            int Sl;
            if (kd) Sl = (pj > tol) ? S_DN : S_UP; else Sl = (pj > tol) ? S_UP : S_DN;
            const double ipl = 1.0 / pj;
            for (int i = threadIdx.x; i < M0; i += NT) c.rvec[i] = invB[lrow + (size_t)i * ldB] * ipl;
            __syncthreads();
            for (int t = threadIdx.x; t < M0 * M0; t += NT) {
                const int jj = t % M0, ii = t / M0;
                double* e = invB + jj + (size_t)ii * ldB;
                *e = (jj == lrow) ? c.rvec[ii] : *e - c.pcol[jj] * c.rvec[ii];
            }
            for (int i = threadIdx.x; i < M0; i += NT) c.pi[i] += rc_kin * c.rvec[i];
            __syncthreads();
            if (threadIdx.x == 0) { c.Bv[lrow] = kin; S1[kin] = S_IN; S1[rid] = Sl; c.qB[lrow] = xold_k + gl; }
            pivots += 1;
            Edge = true;
        }
        __syncthreads();
    }
    return (mode == 1) ? (anyzero ? 2 : 1) : 0;
}
#endif  // !SSQP_ONLY_VW4

// the pivot loop Settings.rule asks for
template <int NT>
static __device__ __forceinline__ int simplex_run(Ctx& c, const int mode, long long& loop, long long& pivots) {
#ifndef SSQP_ONLY_VW4
    if (c.P->rule != 0) return simplex_loop_alt<NT>(c, mode, c.P->rule, loop, pivots);
#endif
    return simplex_loop<NT>(c, mode, loop, pivots);
}

// x from the statuses and x_B (Simplex.jl:610); returns f = sum of the basic artificials
template <int NT>
static __device__ double simplex_assemble(Ctx& c) {
    const int N = c.N, J = c.J, M0 = c.M0;
    const int NJ = N + J, N0 = NJ + c.nfree;
    int* S1 = c.Sst;
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
        else if (i >= NJ) c.z[c.flist[i - NJ]] -= c.qB[j];      // x[iv] .-= x0[N+J+1:N+J+n]  (src/SSQP.jl:540-543; the 1st half is nonbasic)
    }
    return block_sum<NT>(c, fpart);
}

// ---- Phase 1 of solveQP: initQP (src/SSQP.jl:461-560) ------------------------------------------------
// returns 1 feasible, 0 infeasible, -1 numerical; fills c.z (x0) and c.Sst[0..N+J)
// Bounds of the transformed LP (src/SSQP.jl:493-509 = src/Simplex.jl:879-887), staged over c.d / c.u while the simplex
// runs: free variables -> [0, Inf) (ids, ascending, in c.flist), (-Inf,u] -> [-u, Inf) with column sign -1 in c.gr.
template <int NT>
static __device__ void xform_begin(Ctx& c) {
    if (!c.xform) return;
    const int N = c.N;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double* ds = const_cast<double*>(c.d); double* us = const_cast<double*>(c.u);
    block_compact<NT>(c, N, c.flist, [&](int k) { return ds[k] == -INF && us[k] == INF; });      // iv: c.nfree entries
    for (int k = threadIdx.x; k < N; k += NT) {
        const double dk = ds[k], uk = us[k];
        const bool fd = (dk == -INF), fv = fd && (uk == INF);
        c.gr[k] = (fd && !fv) ? -1.0 : 1.0;
        if (fv) ds[k] = 0.0;
        else if (fd) { ds[k] = -uk; us[k] = INF; }
    }
    __syncthreads();
}
template <int NT>
static __device__ void xform_end(Ctx& c, const double* dg, const double* ug) {      // the QP's own bounds back into c.d / c.u
    if (!c.xform) return;
    double* ds = const_cast<double*>(c.d); double* us = const_cast<double*>(c.u);
    for (int k = threadIdx.x; k < c.N; k += NT) { ds[k] = dg[k]; us[k] = ug[k]; }
    __syncthreads();
}

// Variables without a lower bound (src/SSQP.jl:484-509): a free variable is split into two [0, Inf) columns (the 2nd
// half is column N+J+t = -A0[:, iv[t]]); a (-Inf, u] variable is replaced by its negative on [-u, Inf) (column negated).
// After the simplex the halves are recombined, the free variables become IN and the negated ones are flipped back
// (:540-558).  DEVIATION, documented in DESIGN.md: the reference's status flip for the negated variables is a no-op
// comparison (`S[k] == UP`, :552-557) that leaves them DN at x = u — and Phase 2 then reads d = -Inf; here a negated
// variable that ends Phase 1 at its (transformed) lower bound becomes UP, which is what the loop was written for.
// dg / ug: this QP's d and u in global memory (c.d / c.u carry the transformed bounds while the simplex runs).
template <int NT>
static __device__ int phase1(Ctx& c, double* stats, const double* dg, const double* ug) {
    const int N = c.N, J = c.J, M0 = c.M0;
    const int NJ = N + J;
    const double tol = c.P->tolLP;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int* S1 = c.Sst;
    xform_begin<NT>(c);
    auto restore_bounds = [&]() { xform_end<NT>(c, dg, ug); };
    simplex_init<NT>(c);
    if (M0 == 0) {        // no rows: every variable stays at its start value (free ones at 0, IN)
        if (c.xform) {
            for (int k = threadIdx.x; k < N; k += NT) {
                if (dg[k] == -INF) S1[k] = (ug[k] == INF) ? S_IN : S_UP;
            }
            __syncthreads();
        }
        restore_bounds();
        return 1;
    }
    long long loop = 0, pivots = 0;
#ifdef SSQP_TIMELINE
    if (threadIdx.x == 0) c.cyc[T_LAST] = clock64();
#endif
    { const int r1 = simplex_run<NT>(c, 0, loop, pivots); if (r1 == 3 || r1 < 0) { restore_bounds(); return -1; } }      // (unbounded cannot happen in Phase 1)
    const double f = simplex_assemble<NT>(c);
    if (threadIdx.x == 0) { stats[ST_LOOPS] = (double)loop; stats[ST_PIVOTS] = (double)pivots; }
    __syncthreads();
    if (f > tol) { restore_bounds(); return 0; }
    for (int k = N + threadIdx.x; k < NJ; k += NT) S1[k] = (S1[k] == S_IN) ? S_OE : S_EO;
    if (c.xform) {
        for (int t = threadIdx.x; t < c.nfree; t += NT) S1[c.flist[t]] = S_IN;
        for (int k = threadIdx.x; k < N; k += NT)
            if (c.gr[k] < 0.0) { c.z[k] = -c.z[k]; if (S1[k] == S_DN) S1[k] = S_UP; }
    }
    __syncthreads();
    restore_bounds();
    return 1;
}

// In-place inverse of the M0 x M0 matrix a (column-major, leading dimension ld) by Gauss-Jordan with row pivoting
// (the reference calls inv(lu(A0[:,B])), src/Simplex.jl:974).  perm: M0 ints, prow / pcolm: M0 doubles.  Returns 0, or
// -1 when a pivot vanishes (SingularException).
template <int NT>
static __device__ int invert_in_place(Ctx& c, double* a, int ld, int* perm, double* prow, double* pcolm) {
    const int M0 = c.M0;
    for (int p = 0; p < M0; ++p) {
        Cand best;
        for (int r = p + threadIdx.x; r < M0; r += NT) best.offer(-fabs(a[r + (size_t)p * ld]), r);
        block_argmin<NT>(c, best);
        const double mag = -best.key();
        if (!(mag > 0.0) || !(mag < __longlong_as_double(0x7ff0000000000000LL))) return -1;
        const int pr = best.id;
        if (threadIdx.x == 0) perm[p] = pr;
        if (pr != p)
            for (int j = threadIdx.x; j < M0; j += NT) {
                const double t = a[p + (size_t)j * ld]; a[p + (size_t)j * ld] = a[pr + (size_t)j * ld]; a[pr + (size_t)j * ld] = t;
            }
        __syncthreads();
        const double piv = a[p + (size_t)p * ld];
        __syncthreads();
        for (int j = threadIdx.x; j < M0; j += NT) {
            const double v = ((j == p) ? 1.0 : a[p + (size_t)j * ld]) / piv;
            a[p + (size_t)j * ld] = v; prow[j] = v;
        }
        for (int r = threadIdx.x; r < M0; r += NT) pcolm[r] = (r == p) ? 0.0 : a[r + (size_t)p * ld];
        __syncthreads();
        for (int t = threadIdx.x; t < M0 * M0; t += NT) {
            const int r = t % M0, j = t / M0;
            if (r != p) {
                const double base = (j == p) ? 0.0 : a[r + (size_t)j * ld];
                a[r + (size_t)j * ld] = base - pcolm[r] * prow[j];
            }
        }
        __syncthreads();
    }
    for (int p = M0 - 1; p >= 0; --p) {            // the row swaps, undone on the columns of the inverse in reverse order
        const int pr = perm[p];
        if (pr != p)
            for (int r = threadIdx.x; r < M0; r += NT) {
                const double t = a[r + (size_t)p * ld]; a[r + (size_t)p * ld] = a[r + (size_t)pr * ld]; a[r + (size_t)pr * ld] = t;
            }
        __syncthreads();
    }
    return 0;
}

// Artificial variables still basic after Phase 1 of SimplexLP (at level zero): the reference re-selects the basis
// (src/Simplex.jl:962-977) — ic = [basic non-artificial columns; the other columns, ascending], the independent ones
// picked by getRowsGJr(A0[:, ic]', tol) — takes a fresh inverse of A0[:, B] and reads x_B off the current point.
// Restated literally: X = A0[:, ic]' (N0 x M0) in the CTA's global workspace, then invert_in_place.
// Returns 0, or -1 when no full basis comes out / the basis is singular.
template <int NT>
static __device__ int drive_out_artificials(Ctx& c) {
    const int N = c.N, M = c.M, J = c.J, M0 = c.M0;
    const int NJ = N + J, N0 = NJ + c.nfree;
    const double tol = c.P->tolLP;
    const bool xf = c.xform;
    const double* sgn = c.gr;
    const int* ivl = c.flist;
    int* S1 = c.Sst;
    int* ic = c.lpos;            // N0 ints  (lpos + evl are contiguous: 2 (N + M0) ints)
    int* sel = c.item;           // N0 ints  (item + pos)
    double* xv = c.rhs;          // N0 doubles (rhs + sol)
    double* pcolm = c.hv;        // N0 doubles (hv + colv)
    const int ldB = invb_ld(c);
    double* invB = invb_ptr(c);
    double* X = c.work + (invb_in_smem(c) ? 0 : (size_t)ldB * M0);
    // the current point over the N0 columns of the transformed LP
    for (int k = threadIdx.x; k < N0; k += NT) xv[k] = (k < N) ? ((S1[k] == S_UP) ? c.u[k] : c.d[k]) : 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < M0; j += NT) { const int i = c.Bv[j]; if (i < N0) xv[i] = c.qB[j]; }
    const int nbas = block_compact<NT>(c, N0, ic, [&](int k) { return S1[k] == S_IN; });
    block_compact<NT>(c, N0, ic + nbas, [&](int k) { return S1[k] != S_IN; });
    for (long long t = threadIdx.x; t < (long long)N0 * M0; t += NT) {
        const int w = (int)(t % N0), ci = (int)(t / N0);
        const int k = ic[w];
        double v;
        if (k < N) { v = c.Ccol[ci + (size_t)k * M0]; if (xf) v *= sgn[k]; }
        else if (k < NJ) v = (ci == M + (k - N)) ? 1.0 : 0.0;
        else v = -c.Ccol[ci + (size_t)ivl[k - NJ] * M0];
        X[t] = v;
    }
    for (int w = threadIdx.x; w < N0; w += NT) sel[w] = 0;
    __syncthreads();
    const int kept = gjr_core<NT>(c, X, N0, M0, tol, c.rlist, c.pcol, pcolm, sel, nullptr);
    if (kept != M0) return -1;                       // inv(lu(A0[:,B])) of a non-square basis throws
    int miss = 0;
    for (int w = threadIdx.x; w < nbas; w += NT) miss |= (sel[w] == 0);       // a basic column judged dependent: not restated
    if (block_max<NT>(c, (double)miss) > 0.0) return -1;
    for (int w = nbas + threadIdx.x; w < N0; w += NT) if (sel[w]) S1[ic[w]] = S_IN;      // S[iA] .= IN
    __syncthreads();
    const int nB = block_compact<NT>(c, N0, c.Bv, [&](int k) { return S1[k] == S_IN; });     // B = sort(ic[ra])
    if (nB != M0) return -1;
    for (int j = threadIdx.x; j < M0; j += NT) c.qB[j] = xv[c.Bv[j]];                         // q = x[B]
    for (int t = threadIdx.x; t < M0 * M0; t += NT) {
        const int ci = t % M0, j = t / M0;
        const int k = c.Bv[j];
        double v;
        if (k < N) { v = c.Ccol[ci + (size_t)k * M0]; if (xf) v *= sgn[k]; }
        else if (k < NJ) v = (ci == M + (k - N)) ? 1.0 : 0.0;
        else v = -c.Ccol[ci + (size_t)ivl[k - NJ] * M0];
        invB[ci + (size_t)j * ldB] = v;
    }
    __syncthreads();
    return invert_in_place<NT>(c, invB, ldB, c.rlist, c.pcol, c.rvec);
}

// ---- SimplexLP (src/Simplex.jl:831-1034): Phase 1 on the slack form with artificials (free variables split, (-Inf,u]
// ones negated), the drive-out of artificials that stay basic (:962-977), then Phase 2 with the LP's costs from that
// basis and the free-variable epilogue.  Returns the reference's status: 1 unique optimum, 2 infinitely many optima,
// 3 unbounded, 0 infeasible, -1 numerical.  [A 0; G I] must have full row rank: the reference's rank purge (:889-902)
// runs on the host side of the C ABI (solver.py / the Julia glue), before the rows reach the device.
template <int NT>
static __device__ int lp_solve(Ctx& c, double* stats, const double* dg, const double* ug) {
    const int N = c.N, M = c.M, J = c.J, M0 = c.M0;
    const int NJ = N + J, N0 = NJ + c.nfree;
    const double tol = c.P->tolLP;
    int* S1 = c.Sst;
    if (c.xform && M0 == 0) return -1;       // (no rows and no lower bound: not on the device path)
    xform_begin<NT>(c);
    simplex_init<NT>(c);
    long long loop = 0, pivots = 0;
    int status = 1;
    if (M0 > 0) {
        { const int r1 = simplex_run<NT>(c, 0, loop, pivots); if (r1 == 3 || r1 < 0) { xform_end<NT>(c, dg, ug); return -1; } }
        const double f = simplex_assemble<NT>(c);
        if (fabs(f) > tol) {                                              // feasible region is empty  (:923-927)
            if (threadIdx.x == 0) { stats[ST_LOOPS] = (double)loop; stats[ST_PIVOTS] = (double)pivots; }
            __syncthreads();
            xform_end<NT>(c, dg, ug);
            return 0;                                                     // (S is returned as it stands, like the reference)
        }
        int art = 0;
        for (int j = threadIdx.x; j < M0; j += NT) art |= (c.Bv[j] >= N0);
        if (block_max<NT>(c, (double)art) > 0.0 && drive_out_artificials<NT>(c) != 0) { xform_end<NT>(c, dg, ug); return -1; }
        long long loop2 = 0;                 // every cDantzigLP call counts its own loops (the Bland switch depends on it)
        status = simplex_run<NT>(c, 1, loop2, pivots);
        loop += loop2;
    } else {
        // no rows at all: every variable moves to the bound its cost prefers (cDantzigLP with M = 0 flips them one by one)
        int unb = 0, zero = 0;
        for (int k = threadIdx.x; k < N; k += NT) {
            const double ck = c.q[k];
            if (ck < -tol) { if (c.u[k] < __longlong_as_double(0x7ff0000000000000LL)) S1[k] = S_UP; else unb = 1; }
            if (fabs(ck) < tol) zero = 1;
        }
        const double flags = block_max<NT>(c, (double)(2 * unb + zero));
        status = (flags >= 2.0) ? 3 : (flags >= 1.0 ? 2 : 1);
    }
    simplex_assemble<NT>(c);
    if (threadIdx.x == 0) { stats[ST_LOOPS] = (double)loop; stats[ST_PIVOTS] = (double)pivots; }
    for (int k = N + threadIdx.x; k < NJ; k += NT) S1[k] = (S1[k] == S_IN) ? S_OE : S_EO;
    __syncthreads();
    if (c.nfree > 0) {
        // Free variables (src/Simplex.jl:996-1021): a basic 2nd half moves to the variable itself (IN), and the status is
        // recomputed as 2 when some reduced cost of a nonbasic column among the first N+J vanishes, else 1 — also after
        // an unbounded Phase 2, as in the reference.  c.pi and Api = [A;G]' pi (c.pfull) are those of the final basis.
        for (int t = threadIdx.x; t < c.nfree; t += NT) if (S1[NJ + t] == S_IN) S1[c.flist[t]] = S_IN;
        __syncthreads();
        int zpart = 0;
        for (int k = threadIdx.x; k < NJ; k += NT) {
            if (S1[k] == S_IN || S1[k] == S_OE) continue;                  // basic (slacks: IN was renamed OE above)
            const double h = (k < N) ? c.gr[k] * (c.q[k] - c.pfull[k]) : -c.pi[M + (k - N)];
            if (fabs(h) < tol) zpart = 1;
        }
        status = (block_max<NT>(c, (double)zpart) > 0.0) ? 2 : 1;
    }
    if (c.xform) {       // (-Inf,u] variables back to the caller's sign (:1023-1032)
        for (int k = threadIdx.x; k < N; k += NT)
            if (c.gr[k] < 0.0) { c.z[k] = -c.z[k]; if (S1[k] == S_DN) S1[k] = S_UP; }
        __syncthreads();
    }
    xform_end<NT>(c, dg, ug);
    return status;
}

// ---- Phase 2 ---------------------------------------------------------------------------------------
// gradient at z over the support of z (and, with `slack_too`, the slacks [b;g] - [A;G] z): fresh values
template <int NT>
static __device__ void fresh_grad(Ctx& c, bool need_gr, bool slack_too) {
    const int N = c.N, M0 = c.M0;
    SSQP_TICK(c, T_MISC);
    const int cnt = compact_nonzero<NT>(c, c.z, N, c.supp);
    SSQP_TICK(c, T_COMPACT);
    if (need_gr) vpass<NT>(c, c.supp, cnt);
    SSQP_TICK(c, T_VPASS);
    if (slack_too && M0 > 0) {
        cpass<NT>(c, c.supp, cnt, c.z, c.slack);         // (bE / zo of the reference, src/SSQP.jl:295 and :79)
        for (int r = threadIdx.x; r < M0; r += NT) c.slack[r] = c.bg[r] - c.slack[r];
        __syncthreads();
    }
    SSQP_TICK(c, T_CPASSZ);
}

// fresh solve of the reduced KKT system at the current z:  [V_FF AE'; AE 0] [p; lam] = [-gr_F; slack_E]
// (c.gr and c.slack must be fresh).  refine = false: c.sol <- (p, lam).  refine = true: z_F += p (one step of
// iterative refinement of the point the updated inverse produced), c.sol <- (0, lam); returns max |p|.
template <int NT>
static __device__ double fresh_solve(Ctx& c, bool refine) {
    const int N = c.N, n = c.n;
    for (int p = threadIdx.x; p < n; p += NT) {
        const int it = c.item[p];
        c.rhs[p] = (it < N) ? -c.gr[it] : c.slack[it - N];
    }
    __syncthreads();
    SSQP_TICK(c, T_RHS);
    symv<NT>(c, n, c.rhs, c.colv);
    SSQP_TICK(c, T_FSYMV);
    double pm = 0.0, dl = 0.0, lm = 0.0;
    for (int p = threadIdx.x; p < n; p += NT) {
        const int it = c.item[p];
        const double v = c.colv[p];
        if (refine && it < N) { c.z[it] += v; c.sol[it] = 0.0; pm = fmax(pm, fabs(v)); }
        else {
            if (refine) { dl = fmax(dl, fabs(v - c.sol[it])); lm = fmax(lm, fabs(v)); }      // carried multiplier vs fresh one
            c.sol[it] = v;
        }
    }
    c.sol_valid = true;
    if (refine) {
        pm = block_max<NT>(c, pm);
        dl = block_max<NT>(c, dl); lm = block_max<NT>(c, lm);
        c.lamerr = (lm > 0.0) ? dl / lm : 0.0;          // relative error the updates had left in the multipliers
        SSQP_TICK(c, T_APPLY);
        return pm;
    }
    __syncthreads();
    SSQP_TICK(c, T_APPLY);
    return 0.0;
}

template <int NT>
static __device__ long long phase2(Ctx& c, double* stats, const bool reuse = false) {
    const int N = c.N, M = c.M, J = c.J, M0 = c.M0;
    const double tol = c.P->tol, tolG = c.P->tolG;
    const int maxIter = c.P->max_iter;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    constexpr int REFINE_EVERY = 8;      // KKT checks between two refinement solves (the optimal one always refines)
    int* S = c.Sst;
    long long iter = 0;
    // reuse: warm start inside a chain (ssqp_solve_sweep) from a QP that left its inverse behind for exactly this status
    // vector — the reduced KKT matrix depends on (V, A, G, S) only, so the new QP needs a fresh right-hand side (one
    // gradient pass, one symmetric GEMV), not a rebuild: a warm-started QP of a sweep costs its 1-2 trips and nothing else.
    bool have_sys = reuse;
    int ndropped = 0;
    bool gr_fresh = false;
    double falg = 0.0, maxres = 0.0;
    long long updates = 0, rebuilds = 0, degen = 0, nkkt = 0;
    // Drift guard.  The reference refactorises every trip (src/SSQP.jl:322-331); here the inverse lives through ~1000 rank-1
    // updates per QP.  A refinement solve measures the damage: its correction pm = max |H r| is what the updated inverse got
    // wrong at the current point (3e-12 on the benchmark's QPs).  Above DRIFT_TOL = 16 tolG the inverse is rebuilt from
    // scratch and the trip is redone on exact data (`redo`: the trip counter does not advance); REBUILD_EVERY status switches
    // without a rebuild force one as well.
    const double DRIFT_TOL = 16.0 * tolG;
    constexpr double DRIFT_LAM = 1e-6;       // ... or a multiplier was off by more than this, relative (healthy: see ST_LAMERR)
    double maxlam = 0.0;
    constexpr long long REBUILD_EVERY = 4096;
    long long drift_rebuilds = 0, last_rebuild_at = 0;
    bool redo = false;
    int maxK = 0, maxW = 0;
    // Cycle watch.  At a degenerate vertex the reference's method can release a variable (KKTchk!) and block it again with a
    // zero-length step (aStep!) for ever; solveQP then runs to maxIter (src/SSQP.jl:271-274).  A period is [step trip with
    // ONE event and L1 == 0][sign-test trip that releases one item], z untouched.  Nothing but (S, z) and the check counter
    // modulo REFINE_EVERY carries over from one period to the next, so once the same period has been seen CYC_PERIODS times
    // in a row (every residue of the counter included) the alternation is exact for ever and the remaining trips are
    // skipped in pairs: same status -(maxIter+1), same S, same z as running them (a from-scratch rebuild each, 3.6 s for
    // the config-4 QP that does this; tests: test_qp_on_which_the_reference_cycles_until_max_iter).
    // The watch's state lives in shared memory (c.misc[24..30], thread 0 only; held in registers it cost 1.4 % through spills).
    constexpr int CYC_PERIODS = 2 * REFINE_EVERY;
    enum { CY_STEP = 24, CY_PREV_STEP, CY_PREV_KKT, CY_COUNT, CY_NSTEP, CY_ZMOD, CY_SKIP };
    if (threadIdx.x == 0) {
        c.misc[CY_STEP] = -3; c.misc[CY_PREV_STEP] = -3; c.misc[CY_PREV_KKT] = -3; c.misc[CY_COUNT] = 0; c.misc[CY_NSTEP] = 0;
        c.misc[CY_ZMOD] = 0; c.misc[CY_SKIP] = 0;
    }
    c.sol_valid = false;
    c.reusable = false;
    if (!reuse) { c.nf = c.nr = 0; ccache_wait(c); c.ncache = 0; }
    if (threadIdx.x == 0) { c.misc[1] = 0; c.cyc[T_LAST] = clock64(); }

    auto finish = [&](long long st) {
        if (threadIdx.x == 0) {
            stats[ST_TRIPS] = (double)(st > 0 ? st : iter);
            stats[ST_FALG] += falg; stats[ST_MAXK] = maxK; stats[ST_MAXW] = maxW;
            stats[ST_UPDATES] = (double)updates; stats[ST_REBUILDS] = (double)rebuilds;
            stats[ST_MAXRES] = maxres; stats[ST_DEGEN] = (double)degen; stats[ST_DRIFT] = (double)drift_rebuilds;
            stats[ST_LAMERR] = maxlam; stats[ST_NKKT] = (double)nkkt;
        }
        return st;
    };

    // K = |{S == IN}|, JO = |{S == OE}|: counted once, then tracked through the status switches
    int K, JO;
    {
        int kpart = 0, jpart = 0;
        for (int k = threadIdx.x; k < N; k += NT) kpart += (S[k] == S_IN);
        for (int j = threadIdx.x; j < J; j += NT) jpart += (S[N + j] == S_OE);
        const double cnts = block_sum<NT>(c, (double)kpart + 1048576.0 * (double)jpart);
        JO = (int)(cnts / 1048576.0);
        K = (int)(cnts - 1048576.0 * JO + 0.5);
    }

    while (true) {
        if (!redo) {
            iter += 1;
            if (iter > maxIter) return finish(-iter);
        }
        const bool redoing = redo;
        redo = false;
        if (have_sys && updates - last_rebuild_at >= REBUILD_EVERY) { have_sys = false; c.sol_valid = false; drift_rebuilds += 1; }

        if (K == 0) {   // freeK!  (src/SSQP.jl:35-59)
            if (threadIdx.x == 0) { c.misc[CY_COUNT] = 0; c.misc[CY_STEP] = -3; c.misc[CY_NSTEP] = 0; c.misc[CY_ZMOD] = 1; }
            if (!gr_fresh) { fresh_grad<NT>(c, true, false); gr_fresh = true; }
            if (threadIdx.x == 0) falg += 2.0 * N * N;
            int cntin = 0;
            for (int k = threadIdx.x; k < N; k += NT) {
                const double p = c.gr[k];
                const int st = S[k];
                c.evl[k] = st;          // S0 = copy(S)
                if ((p >= -tol && st == S_UP) || (p <= tol && st == S_DN)) { S[k] = S_IN; cntin += 1; }
            }
            const int nin = (int)(block_sum<NT>(c, (double)cntin) + 0.5);
            if (nin == 0) return finish(iter);
            double pm = 0.0;
            for (int k = threadIdx.x; k < N; k += NT) if (S[k] == S_IN) pm = fmax(pm, fabs(c.gr[k]));
            pm = block_max<NT>(c, pm);
            if (pm <= tol) {
                for (int k = threadIdx.x; k < N; k += NT) if (S[k] == S_IN) S[k] = c.evl[k];
                __syncthreads();
                return finish(iter);
            }
            K = nin;
            have_sys = false;
            c.sol_valid = false;
            __syncthreads();
            continue;
        }

        bool fresh_now = false;
        const int W0 = M + (J - JO);                    // rows of [A; G_E] before the redundancy purge
        if (!have_sys || ndropped > 0 || W0 > K) {
            // fresh gradient and slacks at z, then border everything in; the solution comes along
            fresh_grad<NT>(c, !gr_fresh, true); gr_fresh = true;
            int rc;
            const long long tb_ = clock64();          // (once per rebuild: kept in the product build, it is the A/B figure of the builders)
            // (measured and dropped: the two builders as real calls on a copy of the context, to keep their code out of the
            // trip's instruction stream — 13 % slower: the copy pins the context's fields to the stack)
            // The factorisation works on whole rows and columns of the packed matrix: it is the faster builder (x1.6-1.8)
            // as long as the matrix fits in shared memory; once rows live in the L2-resident tail its column walks (one
            // 8-byte load per sector) cost more than the bordered updates, which stream the tail row by row.  Measured along
            // a frontier sweep (scripts/gpu_sweep_diag.py, cycles per rebuild, factorisation / bordering): K+W = 88: 0.47 M /
            // 0.85 M; 177: 1.85 M / 1.95 M; 200: 2.4 M / 2.3 M; 262: 6.9 M / 4.4 M; 310: 13.6 M / 7.0 M — except right after a
            // freeK! mass release, where W is the equality rows only (K ~ 300, W = 1: 0.34 M / 0.60 M).
            const bool chol = (c.P->rebuild_mode != 1) && (K + W0 <= c.R || (W0 <= 8 && K + W0 <= c.R + 128));
            rc = -2;
            for (int tryno = 0; tryno < 2 && rc == -2; ++tryno) {       // (one inlined copy of each builder)
                const bool gj = tryno == 1 || ndropped > 0 || W0 > K;
                rc = chol ? kinv_build_chol<NT>(c, gj) : kinv_rebuild<NT>(c, gj);
            }
            if (threadIdx.x == 0) c.cyc[CY_REBUILD] += clock64() - tb_;
            rebuilds += 1;
            last_rebuild_at = updates;
            if (rc < 0) return finish(-1);
            ndropped = rc;
            if (ndropped > 0) degen += 1;
            have_sys = true;
            fresh_now = true;
        }
        const int n = c.n;
        const int W = n - K;
        maxK = max(maxK, K); maxW = max(maxW, W);
        if (!redoing && threadIdx.x == 0) {        // (F_alg accounting of SURVEY 8d: thread 0's statistic)
            const double k = K, w = W, nn = N;
            const double third = 1.0 / 3.0;       // (a statistic: no division on the trip's critical path)
            falg += ((k * third + w + 2.0) * k + (w + 4.0) * w) * k + ((w * third + 2.0) * w + 2.0 * (nn - k)) * w +
                    2.0 * nn * nn + 2.0 * (double)JO * (nn + k);
        }
        if (!c.sol_valid) {
            fresh_grad<NT>(c, !gr_fresh, true); gr_fresh = true;
            fresh_solve<NT>(c, false);
            fresh_now = true;
        }

        // ---- aStep!  (src/SSQP.jl:61-134): one pass gives max|p| and the ratio test ------------------------------
        bool stepped = false;
        SSQP_TICK(c, T_TOP);
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (J > 0) {       // po = G[Og,F]*p (all rows computed), cached columns first topped up (at most 8 bulk copies a trip)
                if (attempt == 0 && ccache_on(c) && c.ncache < c.nf && ccache_room(c, c.n + 1) >= c.nf) {
                    // (only when the whole free list fits above the inverse: a cache that cannot become complete is never used)
                    int t1 = c.nf;
                    if (t1 > c.ncache + 8) t1 = c.ncache + 8;
                    ccache_fill<NT>(c, c.ncache, t1); c.ncache = t1;
                }
                cpass_free<NT>(c, c.sol, c.cp);
            }
            SSQP_TICK(c, T_CPASS);
            const long long tr_ = SSQP_CLK();
            Cand best;
            double pm = 0.0;
            // (the thread's own first candidates — variable threadIdx.x, row threadIdx.x — stay in registers for the event
            // collection below: at N, J <= NT that is every candidate, and the second pass over S/sol/u/d/z is not needed)
            double myLv = 0.0, myLr = 0.0, mytt = 0.0;
            int myidv = 0;
            bool hasv = false, hasr = false, isF = false;
            for (int k = threadIdx.x; k < N; k += NT) {
                if (S[k] != S_IN) continue;
                const double tt = c.sol[k];
                pm = fmax(pm, fabs(tt));
                double L = 0.0; int id = 0; bool has = false;
                if (tt > tol) { const double uk = c.u[k]; if (uk < INF) { L = (uk - c.z[k]) / tt; id = k; has = true; } }
                else if (tt < -tol) { const double dk = c.d[k]; if (dk > -INF) { L = (dk - c.z[k]) / tt; id = -2 - k; has = true; } }
                if (has) best.offer(L, k);
                if (k == threadIdx.x) { myLv = L; myidv = id; hasv = has; mytt = tt; isF = true; }
            }
            for (int j = threadIdx.x; j < J; j += NT) {
                if (S[N + j] != S_OE) continue;
                const double po = c.cp[M + j];
                if (po > tol) {
                    const double L = c.slack[M + j] / po;
                    best.offer(L, N + j);
                    if (j == threadIdx.x) { myLr = L; hasr = true; }
                }
            }
            block_argmin_max<NT>(c, best, pm);
            SSQP_TICK(c, T_RATIO);
            if (!(pm > tolG)) {
                if (threadIdx.x == 0) c.cyc[CY_RATIO] += SSQP_CLK() - tr_;
                if (fresh_now) break;               // the direction vanishes: go to the sign test, z unchanged
                // A direction that is clearly zero — roundoff of the updates, far below tolG: the vertex-to-vertex trips of the
                // return-seeking QPs, where K = W and z_F is pinned by the active rows — goes to the sign test as it is (the
                // gradient is recomputed there, optimality is only certified after a refinement, the drift guard watches the
                // inverse).  Only a norm within 64x of the threshold is confirmed with fresh data first.
                if (!(pm > tolG * 0.015625)) break;
                fresh_grad<NT>(c, !gr_fresh, true); gr_fresh = true;
                fresh_solve<NT>(c, false);
                fresh_now = true;
                continue;
            }
            const double L1 = best.any() ? best.key() : 1.0;
            if (L1 < 1.0) {
                // collect every event with L - L1 <= tol (multi blocking), then step and switch statuses
                if (hasv && !(myLv - L1 > tol)) { const int s = atomicAdd(&c.misc[1], 1); c.evl[s] = myidv; }
                if (hasr && !(myLr - L1 > tol)) { const int s = atomicAdd(&c.misc[1], 1); c.evl[s] = N + threadIdx.x; }
                for (int k = threadIdx.x + NT; k < N; k += NT) {        // (only when N > NT)
                    if (S[k] != S_IN) continue;
                    const double tt = c.sol[k];
                    double L; int id = 0; bool has = false;
                    if (tt > tol) { const double uk = c.u[k]; if (uk < INF) { L = (uk - c.z[k]) / tt; id = k; has = true; } }
                    else if (tt < -tol) { const double dk = c.d[k]; if (dk > -INF) { L = (dk - c.z[k]) / tt; id = -2 - k; has = true; } }
                    if (has && !(L - L1 > tol)) { const int s = atomicAdd(&c.misc[1], 1); c.evl[s] = id; }
                }
                for (int j = threadIdx.x + NT; j < J; j += NT) {        // (only when J > NT)
                    if (S[N + j] != S_OE) continue;
                    const double po = c.cp[M + j];
                    if (po > tol && !(c.slack[M + j] / po - L1 > tol)) { const int s = atomicAdd(&c.misc[1], 1); c.evl[s] = N + j; }
                }
                // step: z_F += L1 p; the solution of the same system at the new point is p' = (1 - L1) p, lam' = lam
                // (own elements only: no barrier needed between the collection and the step)
                {
                    const double sc = 1.0 - L1;
                    if (isF) { c.z[threadIdx.x] += L1 * mytt; c.sol[threadIdx.x] = sc * mytt; }
                    for (int k = threadIdx.x + NT; k < N; k += NT)
                        if (S[k] == S_IN) { const double tt = c.sol[k]; c.z[k] += L1 * tt; c.sol[k] = sc * tt; }
                    if (J > 0) for (int r = threadIdx.x; r < M0; r += NT) c.slack[r] -= L1 * c.cp[r];
                }
                __syncthreads();
                SSQP_TICK(c, T_COLLECT);
                const int nev = c.misc[1];
                if (threadIdx.x == 0) {
                    c.misc[CY_NSTEP] += 1;
                    c.misc[CY_STEP] = (nev == 1 && L1 == 0.0) ? c.evl[0] : -3;
                    if (L1 != 0.0) c.misc[CY_ZMOD] = 1;
                }
                if (threadIdx.x == 0) {
                    for (int a = 1; a < nev; ++a) {          // deterministic order: ascending variable / row id
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
                gr_fresh = false;
                __syncthreads();
                const long long te_ = SSQP_CLK();
                if (threadIdx.x == 0) { c.misc[1] = 0; c.cyc[CY_RATIO] += te_ - tr_; }
                SSQP_TICK(c, T_STEP);
                for (int e = 0; e < nev; ++e) {
                    const int ev = c.evl[e];
                    int rc = 0;
                    if (ev < -1 || ev < N) {
                        const int k = ev < -1 ? -2 - ev : ev;
                        const int To = ev < -1 ? S_DN : S_UP;
                        if (threadIdx.x == 0) { S[k] = To; c.z[k] = (To == S_DN) ? c.d[k] : c.u[k]; }
                        K -= 1;       // (no barrier: the update reads neither S nor z, and ends with one)
                        if (c.pos[k] >= 0) rc = kinv_remove<NT>(c, k);
                    } else {
                        const int j = ev - N;
                        if (threadIdx.x == 0) S[N + j] = S_EO;
                        JO -= 1;
                        rc = kinv_add<NT>(c, N + M + j, c.slack[M + j]);
                    }
                    updates += 1;
                    if (rc) { ndropped = 1; c.sol_valid = false; }     // dependent working set: rebuild (with row purge) next trip
                }
                __syncthreads();
                if (c.P->debug_perturb > 0 && updates / c.P->debug_perturb != (updates - nev) / c.P->debug_perturb) {
                    for (int p = threadIdx.x; p < c.n; p += NT) c.hrow(p)[p] *= 1.0 + 1e-3;      // test knob: damage the inverse (every debug_perturb updates)
                    __syncthreads();
                }
                if (threadIdx.x == 0) c.cyc[CY_EVENTS] += SSQP_CLK() - te_;
                stepped = true;      // (marks "continue the outer loop")
                break;
            }
            // full step: z[F] = alpha; at alpha the direction vanishes
            if (threadIdx.x == 0) c.misc[CY_ZMOD] = 1;
            for (int k = threadIdx.x; k < N; k += NT)
                if (S[k] == S_IN) { c.z[k] += c.sol[k]; c.sol[k] = 0.0; }
            if (J > 0) for (int r = threadIdx.x; r < M0; r += NT) c.slack[r] -= c.cp[r];
            gr_fresh = false;
            fresh_now = false;
            __syncthreads();
            if (threadIdx.x == 0) c.cyc[CY_RATIO] += SSQP_CLK() - tr_;
            SSQP_TICK(c, T_STEP);
            break;
        }
        if (stepped) continue;

        // ---- KKTchk!  (src/SSQP.jl:136-188): gamma = (V z + q)_B + AB' alphaL ; release the most negative ---------
        // The gradient is always fresh here; the multipliers come from the updated solution, except every
        // REFINE_EVERY-th check and before optimality is declared, when a fresh solve refines z_F and lam.
        bool refined = fresh_now;
        if (!fresh_now) {
            const bool do_refine = (nkkt % REFINE_EVERY) == 0;
            // (measured and dropped: the gradient pass over V[:, supp z] and the multiplier pass over [A;G]' fused into one
            // two-segment streaming pass — 5.7 % slower: two accumulator sets and the per-load segment select cost more
            // than the second pass's ramp-up)
            fresh_grad<NT>(c, true, do_refine); gr_fresh = true;
            if (do_refine) {
                const double pm = fresh_solve<NT>(c, true);
                maxres = fmax(maxres, pm); maxlam = fmax(maxlam, c.lamerr); refined = true;
                if (threadIdx.x == 0) c.misc[CY_ZMOD] = 1;
                if ((pm > DRIFT_TOL || c.lamerr > DRIFT_LAM) && !redoing) {       // the inverse has drifted: rebuild, redo this trip on exact data
                    have_sys = false; c.sol_valid = false; gr_fresh = false; drift_rebuilds += 1; redo = true;
                    __syncthreads();
                    continue;
                }
            }
        }
        nkkt += 1;
        int bid = -1;
        for (int pass = 0; pass < 2; ++pass) {
            {
                const long long tg_ = SSQP_CLK();
                gemv_cols<NT>(GemvArgs{c.Crow, N, ioff(c.rlist), soff(c.lam), c.nr, N, nullptr, soff(c.gr), soff(c.hv), soff(c.buf), c.bufsz, -1});
                if (threadIdx.x == 0) { c.bytes += 8.0 * N * c.nr; c.cyc[CY_GAMMA] += SSQP_CLK() - tg_; }
                SSQP_TICK(c, T_GAMMA);
            }
            const long long tk_ = SSQP_CLK();
            Cand best;
            for (int k = threadIdx.x; k < N; k += NT) {
                const double gam = c.hv[k];
                const int st = S[k];
                if (st == S_UP && gam > tolG) best.offer(-gam, k);
                else if (st == S_DN && gam < -tolG) best.offer(gam, k);
            }
            for (int j = threadIdx.x; j < J; j += NT) {
                if (S[N + j] == S_EO && c.pos[N + M + j] >= 0) {
                    const double t = c.lam[M + j];
                    if (t < -tolG) best.offer(t, N + j);
                }
            }
            if (ndropped > 0) {       // purged EO rows: Lda[j] = alphaL' * (AE' \ GE[j,F])  (src/SSQP.jl:156-160)
                const int nEO = c.misc[7];
                const int nx = nEO < NDROPX ? nEO : NDROPX;
                for (int dd = 0; dd < nx; ++dd) {
                    const int r = c.misc[8 + dd];
                    const double* xd = c.pi + (size_t)dd * c.M0p;
                    double part = 0.0;
                    for (int q = threadIdx.x; q < M0; q += NT) part += c.lam[q] * xd[q];
                    const double Lda = block_sum<NT>(c, part);
                    if (Lda < -tolG && threadIdx.x == 0) best.offer(Lda, N + (r - M));
                }
                if (nEO > NDROPX) {   // more purged EO rows than slots: one solve each, on the fly
                    const int rlast = c.misc[8 + NDROPX - 1];
                    for (int r = rlast + 1; r < M0; ++r) {
                        if (S[N + r - M] != S_EO || c.Bv[r]) continue;
                        const int n2 = c.n;
                        const double* crow = c.Crow + (size_t)r * N;
                        for (int p = threadIdx.x; p < n2; p += NT) { const int a = c.item[p]; c.colv[p] = (a < N) ? crow[a] : 0.0; }
                        __syncthreads();
                        symv<NT>(c, n2, c.colv, c.rhs);
                        double part = 0.0;
                        for (int p = threadIdx.x; p < n2; p += NT) { const int a = c.item[p]; if (a >= N) part += c.lam[a - N] * c.rhs[p]; }
                        const double Lda = block_sum<NT>(c, part);
                        if (Lda < -tolG && threadIdx.x == 0) best.offer(Lda, N + (r - M));
                    }
                }
            }
            block_argmin<NT>(c, best);
            bid = best.any() ? best.id : -1;
            if (threadIdx.x == 0) c.cyc[CY_KKT] += SSQP_CLK() - tk_;
            SSQP_TICK(c, T_KKT);
            if (bid >= 0 || refined) break;
            // optimality must be certified on refined values: fresh slacks, fresh solve, then test once more
            fresh_grad<NT>(c, false, true);
            const double pm = fresh_solve<NT>(c, true);
            maxres = fmax(maxres, pm); maxlam = fmax(maxlam, c.lamerr);
            refined = true;
            if (threadIdx.x == 0) c.misc[CY_ZMOD] = 1;
            if ((pm > DRIFT_TOL || c.lamerr > DRIFT_LAM) && !redoing) { redo = true; break; }       // optimality is never certified through a drifted inverse
        }
        if (redo) {
            have_sys = false; c.sol_valid = false; gr_fresh = false; drift_rebuilds += 1;
            __syncthreads();
            continue;
        }
        if (bid >= 0) {
            const long long te_ = SSQP_CLK();
            int rc;
            if (bid < N) {
                if (threadIdx.x == 0) S[bid] = S_IN;
                K += 1;
                rc = kinv_add<NT>(c, bid, -c.gr[bid]);
            } else {
                if (threadIdx.x == 0) S[bid] = S_OE;
                JO += 1;
                rc = (c.pos[N + M + (bid - N)] >= 0) ? kinv_remove<NT>(c, N + M + (bid - N)) : 0;
            }
            updates += 1;
            if (rc) { ndropped = 1; c.sol_valid = false; }
            __syncthreads();
            if (threadIdx.x == 0) c.cyc[CY_EVENTS] += SSQP_CLK() - te_;
            // cycle watch (see the declaration): one more identical period?
            if (threadIdx.x == 0) {
                const bool clean = (c.misc[CY_ZMOD] == 0) && (c.misc[CY_NSTEP] == 1);
                const int st = c.misc[CY_STEP];
                if (clean && st != -3 && st == c.misc[CY_PREV_STEP] && bid == c.misc[CY_PREV_KKT]) c.misc[CY_COUNT] += 1;
                else c.misc[CY_COUNT] = 0;
                c.misc[CY_PREV_STEP] = clean ? st : -3;
                c.misc[CY_PREV_KKT] = bid;
                c.misc[CY_STEP] = -3; c.misc[CY_NSTEP] = 0; c.misc[CY_ZMOD] = 0;
                c.misc[CY_SKIP] = (c.misc[CY_COUNT] >= CYC_PERIODS) ? 1 : 0;
            }
            __syncthreads();
            if (c.misc[CY_SKIP] && (long long)maxIter - iter > 2)
                iter = (long long)maxIter - (((long long)maxIter - iter) & 1LL);      // skip whole periods; the outcome is the same
            continue;
        }
        // optimal: polishSz!  (src/SSQP.jl:10-32)
        int changed = 0;        // polishSz! relabelled something: the inverse no longer matches S (not reusable by the next QP of a chain)
        for (int k = threadIdx.x; k < N; k += NT) {
            const int st = S[k];
            const double dk = c.d[k], uk = c.u[k];
            if (st == S_DN) c.z[k] = dk;
            else if (st == S_UP) c.z[k] = uk;
            else {
                if (fabs(c.z[k] - dk) < tol) { c.z[k] = dk; S[k] = S_DN; changed = 1; }
                else if (fabs(c.z[k] - uk) < tol) { c.z[k] = uk; S[k] = S_UP; changed = 1; }
            }
        }
        changed = __syncthreads_or(changed);
        if (J > 0) {
            int cnt = compact_nonzero<NT>(c, c.z, N, c.supp);
            cpass<NT>(c, c.supp, cnt, c.z, c.cp);
            int ch2 = 0;
            for (int j = threadIdx.x; j < J; j += NT) {
                const int ns = (fabs(c.bg[M + j] - c.cp[M + j]) < tol) ? S_EO : S_OE;
                ch2 |= (ns != S[N + j]);
                S[N + j] = ns;
            }
            changed |= __syncthreads_or(ch2);
        }
        c.reusable = have_sys && ndropped == 0 && !changed;
        return finish(iter);
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, (NT >= 512 ? 1 : NT >= 256 ? 2 : 4)) ssqp_solve_kernel(const __grid_constant__ KParams P) {
    __shared__ long long s_qp;
    Ctx c;
    int L_qq, L_dd, L_uu;
    {
        const SmemLayout L(P.N, P.M0, P.J, NT, P.hcap, P.nfree_cap);
        c.P = &P;
        c.N = P.N; c.M = P.M; c.J = P.J; c.M0 = P.M0; c.M0p = L.M0p; c.bufsz = L.bufsz;
        c.Ccol = P.Ccol; c.Crow = P.Crow; c.cA = P.cA;
        double* sd = smem_d;
        c.z = sd + L.z; c.gr = sd + L.gr; c.rhs = sd + L.rhs; c.sol = sd + L.sol;
        c.pfull = c.sol; c.lam = c.sol + P.N;          // views of the id-indexed solution (direction by variable, multipliers by row)
        c.hv = sd + L.hv; c.colv = sd + L.colv; c.slack = sd + L.slack; c.cp = sd + L.cp; c.bg = sd + L.bg;
        c.pi = sd + L.pi; c.pcol = sd + L.pcol; c.qB = sd + L.qB; c.rvec = sd + L.rvec;
        c.sig = sd + L.sig; c.buf = sd + L.buf; c.red = sd + L.red; c.Hs = sd + L.H;
        c.cyc = reinterpret_cast<long long*>(sd + L.cyc);
        int* si = reinterpret_cast<int*>(sd + L.ndbl);
        c.item = si + L.item; c.pos = si + L.pos; c.Sst = si + L.Sst; c.Bv = si + L.Bv; c.supp = si + L.supp;
        c.flist = si + L.flist; c.rlist = si + L.rlist; c.lpos = si + L.lpos; c.evl = si + L.evl;
        c.redi = si + L.redi; c.misc = si + L.misc;
        c.work = P.work + (size_t)blockIdx.x * P.wstride;
        c.R = P.hrows;
        c.Hgm = c.work - tri64(P.hrows);
        L_qq = L.qq; L_dd = L.dd; L_uu = L.uu;
    }

    c.ncache = 0; c.cstate = 0; c.reusable = false;
    if (threadIdx.x == 0) ccache_init();
    __syncthreads();
    const int chain = P.chain_len > 1 ? P.chain_len : 1;
    bool carry = false;          // c.z / c.Sst hold the optimal point of the previous QP of the chain
    for (long long pulled = 0, qp = 0; ; ++pulled) {
        if (pulled % chain == 0) {          // next unit of work: one QP, or one chain of QPs
            __syncthreads();
            if (threadIdx.x == 0) s_qp = (long long)atomicAdd(P.queue, 1ULL);
            __syncthreads();
            // The queue is walked from the END of the batch: sweeps are usually ordered by increasing L / mu, which is also
            // increasing cost (trip counts grow ~3x along config 4), and handing out the expensive QPs first keeps the SMs
            // busy to the end of the launch.  Any order is valid — every QP writes only its own outputs.
            qp = ((P.nb + chain - 1) / chain - 1 - s_qp) * chain;
            carry = false;
        } else {
            qp += 1;
        }
        if (qp < 0 || qp >= P.nb) break;
        const int N = P.N, M = P.M, J = P.J, M0 = P.M0;
        c.V = P.V + (size_t)qp * P.strideV;
        {   // per-QP vectors q, d, u: one coalesced read into shared memory (ratio tests and the gradient pass use them
            // every trip; from global memory each use cost an exposed L2 round trip)
            double* qs = smem_d + L_qq; double* ds = smem_d + L_dd; double* us = smem_d + L_uu;
            for (int k = threadIdx.x; k < N; k += NT) {
                qs[k] = P.q ? P.q[(size_t)qp * N + k] : 0.0;
                ds[k] = P.d[(size_t)qp * N + k]; us[k] = P.u[(size_t)qp * N + k];
            }
            c.q = qs; c.d = ds; c.u = us;
        }
        // (inside a chain the previous QP may have left its inverse for this QP: same V, A, G, b, g, d, u and status vector)
        const bool reuse = carry && c.reusable && P.strideV == 0;
        c.bytes = 0.0; c.sol_valid = false; c.nfree = 0; c.xform = false;
        if (!reuse) {
            c.n = 0; c.nf = 0; c.nr = 0; c.reusable = false;
            ccache_wait(c); c.ncache = 0;      // (no bulk copy of the previous QP may still be landing in the inverse's storage)
        }
        if (threadIdx.x == 0) for (int t = 0; t < NCYC; ++t) c.cyc[t] = 0;
        const long long tq0 = clock64();
        double* stats = P.stats + (size_t)qp * NSTATS;
        for (int t = threadIdx.x; t < NSTATS; t += NT) stats[t] = 0.0;
        for (int r = threadIdx.x; r < M0; r += NT) c.bg[r] = (r < M) ? P.b[(size_t)qp * M + r] : P.g[(size_t)qp * J + (r - M)];
        // variables without a lower bound: free ones (u = +Inf) are split in Phase 1, (-Inf,u] ones negated (xform_begin)
        int bad;
        {
            const double INF = __longlong_as_double(0x7ff0000000000000LL);
            int nbad = 0, nneg = 0, nfv = 0;
            for (int k = threadIdx.x; k < N; k += NT) {
                const double dk = c.d[k], uk = c.u[k];
                if (dk != dk || uk != uk) nbad = 1;
                else if (dk == -INF) { if (uk == INF) nfv += 1; else nneg += 1; }
            }
            __syncthreads();
            const double packed = block_sum<NT>(c, 4398046511104.0 * nbad + 2097152.0 * nneg + (double)nfv);     // 2^42, 2^21
            const long long pk = (long long)(packed + 0.5);
            const int tfv = (int)(pk & 2097151LL), tneg = (int)((pk >> 21) & 2097151LL);
            bad = (pk >> 42) != 0 || tfv > P.nfree_cap;
            c.xform = !bad && (tfv + tneg > 0);
            c.nfree = c.xform ? tfv : 0;
        }
        long long status;
        if (bad) {
            for (int k = threadIdx.x; k < N; k += NT) { c.z[k] = 0.0; c.Sst[k] = S_DN; }
            for (int j = threadIdx.x; j < J; j += NT) c.Sst[N + j] = S_OE;
            status = -1;
        } else if (carry) {
            status = 1;                     // warm start from the previous QP of the chain: x and S are already in place
        } else if (P.S0 != nullptr && P.x0 != nullptr) {
            for (int k = threadIdx.x; k < N; k += NT) c.z[k] = P.x0[(size_t)qp * P.strideX0 + k];
            for (int k = threadIdx.x; k < N + J; k += NT) c.Sst[k] = P.S0[(size_t)qp * P.strideS0 + k];
            status = 1;
        } else if (P.lp_mode) {
            status = lp_solve<NT>(c, stats, P.d + (size_t)qp * N, P.u + (size_t)qp * N);
        } else {
            status = phase1<NT>(c, stats, P.d + (size_t)qp * N, P.u + (size_t)qp * N);
        }
        const long long tq1 = clock64();
        __syncthreads();
        if (status > 0 && !P.phase1_only && !P.lp_mode) status = phase2<NT>(c, stats, reuse);
        carry = (chain > 1) && status > 0 && !P.phase1_only && !P.lp_mode;
        __syncthreads();
        for (int k = threadIdx.x; k < N; k += NT) P.x[(size_t)qp * N + k] = c.z[k];
        for (int k = threadIdx.x; k < N + J; k += NT) P.S[(size_t)qp * (N + J) + k] = c.Sst[k];
        if (threadIdx.x == 0) {
            P.status[qp] = status; stats[ST_BYTES] = c.bytes;
            stats[ST_CYC_P1] = (double)(tq1 - tq0);
            for (int t = 0; t < NCYC; ++t) stats[ST_CYC0 + t] = (double)c.cyc[t];
            stats[ST_CYCLES] = (double)(clock64() - tq0);
        }
    }
}

}  // inline namespace SSQP_NS
#endif  // __CUDACC__
}  // namespace ssqp
