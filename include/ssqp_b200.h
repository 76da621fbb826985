/* ssqp_b200.h — C ABI of libssqp_b200.so: batched status-switching active-set QP on NVIDIA B200.
 *
 * This is the drop-in boundary for the `solveQP` hot path of PharosAbad/StatusSwitchingQP.jl
 * (reference v1.0.2).  The reference has no FFI layer; the seam is the Julia function boundary
 *
 *     solveQP(Q::QP{Float64}; settings, settingsLP) -> (z, S, status)        src/SSQP.jl:224-234
 *     solveQP(Q::QP{Float64}, S, x0; settings)      -> (z, S, status)        src/SSQP.jl:237-377
 *
 * and each entry point below is what a Julia `ccall` (or Python ctypes) binds in order to solve a
 * whole batch of QPs in one call (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - plain C, host pointers unless the name says `_device`; all matrices column-major (Julia layout);
 *     per-QP arrays are stored one QP after another (QP i's vector starts at base + i*len).
 *   - every function returns 0 on success and a negative ssqp_error otherwise; nothing throws or
 *     aborts.  Per-QP numerical outcomes are reported in status[i] with the reference's meaning:
 *        >0  optimal, value = number of Phase-2 trips (iter)          src/SSQP.jl:281,374
 *         0  infeasible (Phase 1)                                      src/SSQP.jl:534-537
 *        -1  numerical error (Julia would throw PosDef/SingularException)
 *        -(max_iter+1)  iteration cap                                  src/SSQP.jl:272-274
 *   - status vector S: int32 codes IN=0, DN=1, UP=2, OE=3, EO=4       src/types.jl:17-23
 *   - there is NO CPU fallback: without a usable CUDA device ssqp_create fails with SSQP_ERR_CUDA.
 *   - a ctx is not thread-safe (one ctx per host thread); calls are blocking unless noted.
 */
#ifndef SSQP_B200_H
#define SSQP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssqp_ctx ssqp_ctx;

/* mirrors struct Settings{Float64}, src/types.jl:390-408 (defaults 7777, 2^-26, 2^-33, :column, :Dantzig) */
typedef struct {
    int32_t max_iter;
    double  tol;
    double  tolG;
    int32_t rule;   /* pivot rule of the simplex (src/types.jl:397): 0 = :Dantzig, 1 = :stpEdgeLP, 2 = :maxImprovement.  initQP
                       reads settingsLP.rule, SimplexLP settings.rule (src/SSQP.jl:474-482, src/Simplex.jl:853-858) */
    int32_t pivot;  /* read nowhere live in the reference (src/SSQP.jl:258-259); kept for layout parity */
} ssqp_settings;

typedef enum {
    SSQP_OK = 0,
    SSQP_ERR_ARG = -1,          /* bad argument (NULL, negative size, NaN bounds, ...) */
    SSQP_ERR_CUDA = -2,         /* CUDA runtime error / no device; see ssqp_last_error */
    SSQP_ERR_UNSUPPORTED = -3,  /* input outside the device path (problem too large for shared memory) */
    SSQP_ERR_STATE = -4         /* call order (solve before set_shared, ...) */
} ssqp_error;

/* per-QP statistics written by the kernel (SSQP_NSTATS doubles per QP), see ssqp_get_stats */
enum {
    SSQP_STAT_TRIPS = 0,       /* Phase-2 trips (== status when optimal) */
    SSQP_STAT_FALG = 1,        /* algorithmic FLOPs, SURVEY.md 8(d) F_alg, from the QP's own K_t, W_t */
    SSQP_STAT_MAXK = 2,
    SSQP_STAT_MAXW = 3,
    SSQP_STAT_LP_LOOPS = 4,    /* Phase-1 simplex loops */
    SSQP_STAT_LP_PIVOTS = 5,
    SSQP_STAT_UPDATES = 6,     /* rank-1 add/remove updates applied to the reduced-KKT inverse */
    SSQP_STAT_REBUILDS = 7,    /* from-scratch rebuilds of the reduced-KKT inverse */
    SSQP_STAT_MAXRES = 8,      /* largest iterative-refinement correction |dz| seen (health of the updated inverse) */
    SSQP_STAT_CYCLES = 9,      /* SM cycles spent on this QP */
    SSQP_STAT_BYTES = 10,      /* bytes streamed from L2/HBM by the QP's passes (V, [A;G], inverse tail) */
    SSQP_STAT_DEGEN = 11,      /* rebuilds that purged dependent rows (getRowsGJr path) */
    SSQP_STAT_CYC_PHASE1 = 12, /* SM cycles of Phase 1 */
    SSQP_STAT_CYC_SECTION0 = 13, /* 13..22: cycles in gradient pass, constraint passes, symmetric GEMV, rank-1 update,
                                    sign-test pass, Phase-1 pricing pass, from-scratch builds of the inverse, ratio test, event
                                    application, sign test; 23, 24: symmetric-GEMV / rank-1-update call counts */
    SSQP_STAT_DRIFT = 53,      /* rebuilds forced by the drift guard (refinement correction above 16 tolG, or 4096 updates) */
    SSQP_STAT_LAMERR = 54,     /* largest relative correction a refinement made to the carried multipliers */
    SSQP_STAT_NKKT = 55,       /* sign tests (KKTchk!) run */
    SSQP_NSTATS = 56           /* 29..51: exclusive per-section timeline of Phase 2 (developer diagnostics, see scripts/gpu_check.py) */
};

void ssqp_default_settings(ssqp_settings* s);                       /* src/types.jl:401-408 */

/* Create a context on the given CUDA devices (device_ids == NULL -> device 0..n_devices-1). */
int ssqp_create(ssqp_ctx** ctx, const int32_t* device_ids, int32_t n_devices);
int ssqp_destroy(ssqp_ctx* ctx);

/* Upload the data shared by every QP of the following batches and replicate it on every device:
 * V (N x N, may be NULL when every batch passes V_per_qp), A (M x N), G (J x N).   (QP fields, src/types.jl:214-227) */
int ssqp_set_shared(ssqp_ctx* ctx, int32_t N, int32_t M, int32_t J,
                    const double* V, const double* A, const double* G);

/* Solve nb QPs (a loop of solveQP calls, src/SSQP.jl:224-234).  Shards by QP index over the ctx's devices
 * (no collective).  V_per_qp: NULL -> shared V; else N*N*nb.  S0/x0: NULL -> cold start through Phase 1
 * (initQP, src/SSQP.jl:461-560); else warm start solveQP(Q,S,x0) (src/SSQP.jl:237).
 * Bounds: d may be -Inf and u +Inf.  Phase 1 splits a free variable into two columns and negates a (-Inf,u] one
 * (src/SSQP.jl:484-509, 540-558); a negated variable that ends Phase 1 at its bound is returned UP — the reference's
 * own flip (:552-557) is a no-op comparison that leaves it DN with d = -Inf (deviations).
 * Outputs: x (N*nb), S ((N+J)*nb), status (nb). */
int ssqp_solve_batch(ssqp_ctx* ctx, int64_t nb,
                     const double* V_per_qp,
                     const double* q, const double* b, const double* g,
                     const double* d, const double* u,
                     const int32_t* S0, const double* x0,
                     const ssqp_settings* settings, const ssqp_settings* settingsLP,
                     double* x, int32_t* S, int64_t* status);

/* Sweep over the linear term (the frontier sweeps QP(P, q, L) of src/types.jl:303-319): nb QPs in chains of chain_len
 * consecutive QPs that share b, g, d, u (checked; only q — or V_per_qp — varies along a chain).  One CTA solves a chain in
 * order: its first QP starts cold (initQP; once for the whole batch when b, g, d, u are shared by all of it) and every next
 * one is the reference's warm start solveQP(Q, S, x0) (src/SSQP.jl:237) from the previous QP's result — the loop
 *     x, S, st = solveQP(Qs[1]);  for Q in Qs[2:end]  x, S, st = solveQP(Q, S, x)  end
 * with chains running in parallel.  status[i] is that call's own iteration count.  nb % chain_len == 0; chains are
 * sharded over the ctx's devices whole. */
int ssqp_solve_sweep(ssqp_ctx* ctx, int64_t nb, int64_t chain_len,
                     const double* V_per_qp,
                     const double* q, const double* b, const double* g,
                     const double* d, const double* u,
                     const ssqp_settings* settings, const ssqp_settings* settingsLP,
                     double* x, int32_t* S, int64_t* status);

/* Same, but every pointer is a DEVICE pointer on the ctx's first device and the call only enqueues the
 * kernels on `stream` (a cudaStream_t passed as void*; NULL = the ctx's own stream) and returns without
 * synchronising: the timed region of bench.py's HBM-resident `value`.  Single device.  ONE launch in flight per ctx: the
 * kernel's workspace, work queue and statistics belong to the ctx, so a second call — on any stream — must wait for the
 * first (use one ctx per concurrent stream).  V_per_qp may sit at any 8-byte aligned address (the 256-bit streaming loads
 * are used when it is 32-byte aligned).  The bounds are not scanned on the host here (that would synchronise): Phase 1
 * sizes its status array for the number of FREE variables (d = -Inf and u = +Inf) per QP announced with
 * ssqp_set_free_var_capacity (default 0); a QP with more gets status -1.  Returns SSQP_ERR_ARG for NULL b (M > 0), NULL g
 * (J > 0) or only one of S0 / x0. */
int ssqp_solve_batch_device(ssqp_ctx* ctx, int64_t nb,
                            const double* V_per_qp,
                            const double* q, const double* b, const double* g,
                            const double* d, const double* u,
                            const int32_t* S0, const double* x0,
                            const ssqp_settings* settings, const ssqp_settings* settingsLP,
                            double* x, int32_t* S, int64_t* status, void* stream);

/* Batch of LPs  min c'x  s.t. Ax=b, Gx<=g, d<=x<=u  sharing A and G (ssqp_set_shared with V = NULL): the reference's
 * two-phase SimplexLP (src/Simplex.jl:831-1034; Phase 1 and Phase 2 are both cDantzigLP, :445-615) with the default
 * Dantzig rule.  c is N*nb.  status[i]: 1 unique optimum, 2 infinitely many optima, 3 unbounded, 0 infeasible,
 * -1 numerical.  Free and (-Inf,u] variables (:861-887, 996-1032) and the drive-out of artificial variables that stay
 * basic after Phase 1 (:962-977) run on the device.  [A 0; G I] must have full row rank: the reference's redundancy purge
 * (rank + getRowsGJr, :889-902) belongs to the caller's side of this ABI (solver.py / julia/SSQPB200.jl do it); with
 * dependent rows an artificial variable cannot be driven out and the LP gets -1. */
int ssqp_solve_lp_batch(ssqp_ctx* ctx, int64_t nb,
                        const double* c, const double* b, const double* g,
                        const double* d, const double* u,
                        const ssqp_settings* settings,
                        double* x, int32_t* S, int64_t* status);

/* Phase 1 only (initQP, src/SSQP.jl:461-560): x0 (N*nb), S ((N+J)*nb), status (1 feasible / 0 infeasible / -1). */
int ssqp_init_batch(ssqp_ctx* ctx, int64_t nb,
                    const double* b, const double* g, const double* d, const double* u,
                    const ssqp_settings* settingsLP,
                    double* x0, int32_t* S, int64_t* status);

/* Device-pointer entry only: the largest number of free variables (d = -Inf and u = +Inf; initQP splits each into two
 * columns, src/SSQP.jl:484-509) any QP of the following ssqp_solve_batch_device calls may have.  The host-pointer entries
 * scan the bounds themselves. */
int ssqp_set_free_var_capacity(ssqp_ctx* ctx, int32_t max_free_vars_per_qp);

/* Copy the per-QP statistics (SSQP_NSTATS doubles per QP) of the last host-pointer batch to `stats`. */
int ssqp_get_stats(ssqp_ctx* ctx, int64_t nb, double* stats);
/* Device-pointer variant: stats of the last ssqp_solve_batch_device call (device buffer owned by ctx). */
int ssqp_get_stats_device(ssqp_ctx* ctx, int64_t nb, double* stats_host);

/* Kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
int64_t ssqp_launch_count(const ssqp_ctx* ctx);
/* Elapsed device time (ms, CUDA events on the launch stream) of the solve kernel in the last call, max over devices. */
double ssqp_last_kernel_ms(const ssqp_ctx* ctx);

/* FP64 FMA peak microbenchmark on the ctx's first device: returns achieved TFLOP/s (roofline denominator). */
double ssqp_measure_fp64_peak(ssqp_ctx* ctx);
/* L2->SM streaming read bandwidth microbenchmark (GB/s) over a `mbytes`-MB buffer re-read `reps` times. */
double ssqp_measure_read_bw(ssqp_ctx* ctx, int32_t mbytes, int32_t reps);

const char* ssqp_last_error(const ssqp_ctx* ctx);
/* Human-readable launch configuration of the last solve on the first device (CTA width, inverse rows kept in
 * shared memory, bytes of shared memory, CTAs per SM, grid) — diagnostics for bench.py / profiles. */
const char* ssqp_last_launch_config(const ssqp_ctx* ctx);
int32_t ssqp_device_count(void);   /* number of visible CUDA devices (0 when none / no driver) */
const char* ssqp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SSQP_B200_H */
