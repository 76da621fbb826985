// ssqp_capi.cu — host runtime + C ABI of libssqp_b200.so (see include/ssqp_b200.h).
//
// Host side of the drop-in boundary for the reference's solveQP (src/SSQP.jl:224-377): device contexts,
// replication of the shared V/A/G, sharding of a batch by QP index over the ctx's devices (interleaved,
// i mod G, because trip counts are heavy-tailed), H2D/D2H staging and kernel launches.  No collectives:
// QPs are independent.  There is no CPU fallback — every entry point needs a CUDA device.
#include "../../include/ssqp_b200.h"
#include "ssqp_kernel.cuh"           // KParams / SmemLayout (the kernel itself is instantiated in ssqp_inst.cu)
#include "ssqp_helpers.cuh"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace ssqp;

// one translation unit per CTA width (csrc/ssqp_inst.cu compiled with -DSSQP_NT=256|512)
typedef void (*ssqp_kernel_fn)(const KParams);
ssqp_kernel_fn ssqp_kernel_ptr_512_any();
ssqp_kernel_fn ssqp_kernel_ptr_256_any();
ssqp_kernel_fn ssqp_kernel_ptr_512_vw4();      // N % 4 == 0 && (M+J) % 4 == 0: 256-bit streaming loads only
ssqp_kernel_fn ssqp_kernel_ptr_256_vw4();
ssqp_kernel_fn ssqp_kernel_ptr_128_any();      // small problems: four 128-thread CTAs per SM hide each other's latency chains
ssqp_kernel_fn ssqp_kernel_ptr_128_vw4();

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Device {
    int id = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf V, A, G, Ccol, Crow, cA;          // shared problem data
    DevBuf work, queue, stats;               // kernel workspace
    DevBuf q, b, g, d, u, Vq, S0, x0, x, S, status;   // staging of the device's shard
    int grid = 0;                            // CTAs of the last sizing
    long long wstride = 0;
    double last_ms = 0.0;
    std::string err;
    std::string last_cfg;                    // launch configuration of the last solve (diagnostics)
};

}  // namespace

struct ssqp_ctx {
    std::vector<Device> dev;
    int N = 0, M = 0, J = 0;
    bool have_shared = false, have_V = false;
    std::atomic<int64_t> launches{0};        // (the per-device host threads of a multi-device batch all count here)
    int64_t last_nb = 0;
    std::string err;
    std::vector<int64_t> shard_cnt;          // per-device QP counts of the last host batch
    bool bcast_start = false;                // the warm start of the batch being launched is one shared point (stride 0)
    int nfree_cap = 0;                       // most free variables (d = -Inf, u = +Inf) of any QP in the batch being launched
    int chain_len = 1;                       // chain length of the batch being launched (ssqp_solve_sweep), 1 = independent QPs
    int free_cap_device = 0;                 // ssqp_set_free_var_capacity: free variables per QP the device-pointer entry sizes for
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            errs = buf_;                                                                           \
            return SSQP_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

int check_settings(const ssqp_settings* s, const ssqp_settings* slp, std::string& errs) {
    // Settings.rule (src/types.jl:397): 0 :Dantzig, 1 :stpEdgeLP, 2 :maxImprovement — the rule of the simplex (initQP reads
    // settingsLP.rule, SimplexLP settings.rule); solveQP's own settings.rule is not read by the reference's Phase 2
    if (s && (s->rule < 0 || s->rule > 2)) { errs = "settings.rule must be 0 (:Dantzig), 1 (:stpEdgeLP) or 2 (:maxImprovement)"; return SSQP_ERR_ARG; }
    if (slp && (slp->rule < 0 || slp->rule > 2)) { errs = "settingsLP.rule must be 0 (:Dantzig), 1 (:stpEdgeLP) or 2 (:maxImprovement)"; return SSQP_ERR_ARG; }
    return SSQP_OK;
}

// enqueue the solve of nb QPs whose per-QP arrays are device pointers on D: size shared memory (packed inverse
// rows kept on chip), grid and workspace, then launch
int launch_solve(ssqp_ctx* ctx, Device& D, int64_t nb, const double* Vq, const double* q, const double* b,
                  const double* g, const double* d, const double* u, const int32_t* S0, const double* x0,
                  const ssqp_settings& st, const ssqp_settings& stlp, double* x, int32_t* S, int64_t* status,
                  cudaStream_t stream, int phase1_only /* 0 solveQP, 1 initQP only, 2 SimplexLP */, std::string& errs) {
    const int N = ctx->N, M = ctx->M, J = ctx->J, M0 = M + J;
    // CTA width: 512 threads (one CTA per SM, the inverse fills shared memory) for the N ~ 500 shapes; 256 for mid sizes; 128
    // for small problems, whose stages are pure latency and whose inverse is small enough for 3-4 CTAs per SM (config 2,
    // N = 100: 297 k QPs/s with 3 x 128 threads against 233 k with 2 x 256; at N = 200 the inverse fills the SM and 256 wins)
    int NTv = (N + M0 >= 320) ? 512 : (N + M0 >= 160) ? 256 : 128;
    if (stlp.rule != 0 && NTv < 256) NTv = 256;      // (the steepest-edge scores are CTA-wide sums: keep the summation order the
                                                     //  path-exact tests of that rule were pinned with)
    if (const char* e = getenv("SSQP_NT")) { int t = atoi(e); if (t == 128 || t == 256 || t == 512) NTv = t; }
    bool vw4 = (N % 4 == 0) && (M0 % 4 == 0) && M0 > 0 && (!Vq || ((uintptr_t)Vq % 32 == 0));
    if (const char* e = getenv("SSQP_FLAVOUR")) { if (!strcmp(e, "any")) vw4 = false; }     // test knob: force the general flavour
    if (stlp.rule != 0) vw4 = false;         // the other pivot rules live in the general flavour only
    ssqp_kernel_fn fn = vw4 ? ((NTv == 512) ? ssqp_kernel_ptr_512_vw4() : (NTv == 256) ? ssqp_kernel_ptr_256_vw4() : ssqp_kernel_ptr_128_vw4())
                            : ((NTv == 512) ? ssqp_kernel_ptr_512_any() : (NTv == 256) ? ssqp_kernel_ptr_256_any() : ssqp_kernel_ptr_128_any());
    const long long nmax = N + M0;
    const long long full = nmax * (nmax + 1) / 2;
    const long long ldB = M0 | 1, invB = ldB * M0;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, fn));
    const size_t SMEM_MAX = 227 * 1024 - ((fa.sharedSizeBytes + 15) / 16) * 16;   // opt-in limit per CTA minus the kernel's static bytes
    const int nfree = ctx->nfree_cap;
    const size_t base = SmemLayout(N, M0, J, NTv, 0, nfree).bytes();
    if (base + 8 * 64 > SMEM_MAX) { errs = "problem too large for the device path (shared memory)"; return SSQP_ERR_UNSUPPORTED; }
    long long hcap, hrows;
    long long want = full > invB ? full : invB;
    if (base + 8 * (size_t)want <= SMEM_MAX) { hcap = want; hrows = nmax; }
    else {
        hcap = (long long)((SMEM_MAX - base) / 8) & ~1LL;
        hrows = (long long)((std::sqrt(8.0 * (double)hcap + 1.0) - 1.0) / 2.0);
        while (hrows * (hrows + 1) / 2 > hcap) --hrows;
        if (hrows > nmax) hrows = nmax;
    }
    // Mid-size problems (256-thread CTAs) whose full inverse would fill the SM: two CTAs per SM with at least half of the
    // rows on chip beat one CTA with all of them (N=200, J=30: 51 k QPs/s with 128 of 231 rows on chip and 2 CTAs/SM against
    // 38 k; the two solves hide each other's latency chains, the spilled tail streams from L2).  Not at N ~ 500: the
    // vectors alone take 79 KB there and the inverse would keep 92 of its ~150 live rows (measured slower).
    if (NTv == 256 && base + 8 * (size_t)want > SMEM_MAX / 2) {
        const size_t half = (size_t)228 * 1024 / 2 - 1024 - ((fa.sharedSizeBytes + 15) / 16) * 16;      // two CTAs, 1 KB reserved each
        if (half > base + 8 * (size_t)invB) {
            long long hc2 = (long long)((half - base) / 8) & ~1LL;
            long long hr2 = (long long)((std::sqrt(8.0 * (double)hc2 + 1.0) - 1.0) / 2.0);
            while (hr2 * (hr2 + 1) / 2 > hc2) --hr2;
            if (hr2 < nmax && 2 * hr2 >= nmax && hc2 >= invB) { hcap = hc2; hrows = hr2; }
        }
    }
    if (const char* e = getenv("SSQP_HROWS")) {          // experiment knob: keep fewer rows on chip (more CTAs per SM)
        long long r = atoll(e);
        if (r >= 1 && r < hrows) { hrows = r; hcap = r * (r + 1) / 2; }
    }
    if (hcap < 2) hcap = 2;
    const size_t smem = SmemLayout(N, M0, J, NTv, (int)hcap, nfree).bytes();
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, NTv, smem));
    if (occ < 1) { errs = "kernel does not fit on an SM"; return SSQP_ERR_CUDA; }
    const int chain = ctx->chain_len > 1 ? ctx->chain_len : 1;
    int64_t grid = (int64_t)D.sms * occ;
    if (grid > (nb + chain - 1) / chain) grid = (nb + chain - 1) / chain;
    if (grid < 1) grid = 1;
    long long w = full - hrows * (hrows + 1) / 2;
    const long long invBg = (long long)((M0 + 3) / 4 * 4) * M0;      // invB in the workspace: leading dimension rounded up to 4
    if (invB > hcap && invBg > w) w = invBg;
    const long long gj = (long long)M0 * (N + 1);        // [AE bE] of the degenerate-row purge (getRowsGJr)
    if (gj > w) w = gj;
    if (phase1_only == 2) {                              // SimplexLP: A0[:, ic]' of the drive-out of basic artificials, behind invB
        const long long dx = (invB > hcap ? invBg : 0) + (long long)(N + J + nfree) * M0;
        if (dx > w) w = dx;
    }
    w = (w + 31) / 16 * 16;
    const long long stage_off = w;                       // staging of the from-scratch factorisation's in-place transforms
    if (phase1_only == 0) w += TP_STAGE;
    D.wstride = w;
    D.grid = (int)grid;
    CK(D.work.ensure((size_t)w * grid * sizeof(double)));
    CK(D.queue.ensure(sizeof(unsigned long long)));
    CK(D.stats.ensure((size_t)nb * NSTATS * sizeof(double)));
    KParams P;
    memset(&P, 0, sizeof P);
    P.N = N; P.M = M; P.J = J; P.M0 = M0; P.hrows = (int)hrows; P.hcap = (int)hcap;
    if (Vq) { P.V = Vq; P.strideV = (long long)N * N; }
    else { P.V = D.V.as<double>(); P.strideV = 0; }
    P.Ccol = D.Ccol.as<double>(); P.Crow = D.Crow.as<double>(); P.cA = D.cA.as<double>();
    P.q = q; P.b = b; P.g = g; P.d = d; P.u = u;
    P.S0 = S0; P.x0 = x0;
    P.strideS0 = ctx->bcast_start ? 0 : (long long)(N + J);
    P.strideX0 = ctx->bcast_start ? 0 : (long long)N;
    P.x = x; P.S = S; P.status = (long long*)status; P.stats = D.stats.as<double>();
    P.work = D.work.as<double>(); P.wstride = D.wstride; P.stage_off = stage_off;
    P.queue = D.queue.as<unsigned long long>();
    P.nb = nb;
    P.max_iter = st.max_iter; P.tol = st.tol; P.tolG = st.tolG; P.tolLP = stlp.tol;
    P.phase1_only = (phase1_only == 1);
    P.lp_mode = (phase1_only == 2);
    P.nfree_cap = nfree;
    P.chain_len = chain;
    P.rule = stlp.rule;
    if (const char* e = getenv("SSQP_DEBUG_PERTURB")) P.debug_perturb = atoi(e);      // test knob of the drift guard
    if (const char* e = getenv("SSQP_REBUILD")) P.rebuild_mode = !strcmp(e, "border") ? 1 : 0;      // A/B knob: sequential bordering
    CK(cudaMemsetAsync(D.queue.p, 0, sizeof(unsigned long long), stream));
    CK(cudaEventRecord(D.ev0, stream));
    fn<<<D.grid, NTv, smem, stream>>>(P);
    CK(cudaGetLastError());
    CK(cudaEventRecord(D.ev1, stream));
    ctx->launches += 1;
    D.last_cfg = "NT=" + std::to_string(NTv) + (vw4 ? " vw4" : " any") + " hrows=" + std::to_string(hrows) + " smem=" + std::to_string(smem) +
                 " occ=" + std::to_string(occ) + " grid=" + std::to_string(grid);
    return SSQP_OK;
}

}  // namespace

extern "C" {

void ssqp_default_settings(ssqp_settings* s) {
    s->max_iter = 7777;
    s->tol = std::ldexp(1.0, -26);
    s->tolG = std::ldexp(1.0, -33);
    s->rule = 0;
    s->pivot = 0;
}

int32_t ssqp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* ssqp_version(void) { return "ssqp_b200 0.1.0 (sm_100a)"; }

static thread_local std::string g_create_err;
int ssqp_destroy(ssqp_ctx* ctx);

int ssqp_create(ssqp_ctx** out, const int32_t* device_ids, int32_t n_devices) {
    if (!out || n_devices < 1) return SSQP_ERR_ARG;
    *out = nullptr;
    int have = ssqp_device_count();
    if (have < 1) return SSQP_ERR_CUDA;      // no CPU fallback
    ssqp_ctx* ctx = new ssqp_ctx();
    std::string& errs = ctx->err;
    ctx->dev.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        Device& D = ctx->dev[i];
        D.id = device_ids ? device_ids[i] : i;
        if (D.id < 0 || D.id >= have) { ssqp_destroy(ctx); return SSQP_ERR_ARG; }
        cudaError_t e = cudaSetDevice(D.id);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&D.ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&D.ev1);
        cudaDeviceProp prop;
        if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, D.id);
        if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); cudaGetLastError(); ssqp_destroy(ctx); return SSQP_ERR_CUDA; }      // (frees what the earlier devices got)
        D.sms = prop.multiProcessorCount;
    }
    (void)errs;
    *out = ctx;
    return SSQP_OK;
}

int ssqp_destroy(ssqp_ctx* ctx) {
    if (!ctx) return SSQP_ERR_ARG;
    for (Device& D : ctx->dev) {
        if (!D.stream && !D.ev0 && !D.ev1) continue;      // (a device ssqp_create never reached)
        cudaSetDevice(D.id);
        if (D.stream) cudaStreamSynchronize(D.stream);
        for (DevBuf* b : {&D.V, &D.A, &D.G, &D.Ccol, &D.Crow, &D.cA, &D.work, &D.queue, &D.stats, &D.q, &D.b, &D.g, &D.d,
                          &D.u, &D.Vq, &D.S0, &D.x0, &D.x, &D.S, &D.status})
            b->release();
        if (D.ev0) cudaEventDestroy(D.ev0);
        if (D.ev1) cudaEventDestroy(D.ev1);
        if (D.stream) cudaStreamDestroy(D.stream);
    }
    delete ctx;
    return SSQP_OK;
}

const char* ssqp_last_launch_config(const ssqp_ctx* ctx) { return (ctx && !ctx->dev.empty()) ? ctx->dev[0].last_cfg.c_str() : ""; }
const char* ssqp_last_error(const ssqp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int64_t ssqp_launch_count(const ssqp_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
double ssqp_last_kernel_ms(const ssqp_ctx* ctx) {
    double m = 0.0;
    if (ctx) for (const Device& D : ctx->dev) m = D.last_ms > m ? D.last_ms : m;
    return m;
}

int ssqp_set_shared(ssqp_ctx* ctx, int32_t N, int32_t M, int32_t J, const double* V, const double* A, const double* G) {
    if (!ctx) return SSQP_ERR_ARG;
    std::string& errs = ctx->err;
    if (N < 1 || M < 0 || J < 0 || (M > 0 && !A) || (J > 0 && !G)) { errs = "set_shared: bad sizes or NULL A/G"; return SSQP_ERR_ARG; }
    const int M0 = M + J;
    for (Device& D : ctx->dev) {
        CK(cudaSetDevice(D.id));
        if (V) {
            CK(D.V.ensure((size_t)N * N * 8));
            CK(cudaMemcpyAsync(D.V.p, V, (size_t)N * N * 8, cudaMemcpyHostToDevice, D.stream));
        }
        CK(D.A.ensure((size_t)(M > 0 ? M : 1) * N * 8));
        CK(D.G.ensure((size_t)(J > 0 ? J : 1) * N * 8));
        CK(D.Ccol.ensure((size_t)(M0 > 0 ? M0 : 1) * N * 8));
        CK(D.Crow.ensure((size_t)(M0 > 0 ? M0 : 1) * N * 8));
        CK(D.cA.ensure((size_t)N * 8));
        if (M > 0) CK(cudaMemcpyAsync(D.A.p, A, (size_t)M * N * 8, cudaMemcpyHostToDevice, D.stream));
        if (J > 0) CK(cudaMemcpyAsync(D.G.p, G, (size_t)J * N * 8, cudaMemcpyHostToDevice, D.stream));
        if (M0 > 0) {
            ssqp_stack_kernel<<<64, 256, 0, D.stream>>>(N, M, J, D.A.as<double>(), D.G.as<double>(), D.Ccol.as<double>());
            CK(cudaGetLastError());
        }
        ssqp_prep_kernel<<<(N + 127) / 128, 128, 0, D.stream>>>(N, M0, D.Ccol.as<double>(), D.Crow.as<double>(), D.cA.as<double>());
        CK(cudaGetLastError());
        ctx->launches += (M0 > 0) ? 2 : 1;
        CK(cudaStreamSynchronize(D.stream));
    }
    ctx->N = N; ctx->M = M; ctx->J = J;
    ctx->have_shared = true;
    ctx->have_V = (V != nullptr);
    return SSQP_OK;
}

static int solve_host(ssqp_ctx* ctx, int64_t nb, const double* Vq, const double* q, const double* b, const double* g,
                      const double* d, const double* u, const int32_t* S0, const double* x0, const ssqp_settings* settings,
                      const ssqp_settings* settingsLP, double* x, int32_t* S, int64_t* status, int phase1_only,
                      int64_t chain_len = 1) {
    if (!ctx) return SSQP_ERR_ARG;
    std::string& errs = ctx->err;
    const int64_t U = chain_len > 1 ? chain_len : 1;      // QPs per unit of work (a chain is solved in order by one CTA)
    if (U > 1) {
        if (phase1_only != 0 || nb % U != 0) { errs = "solve_sweep: nb must be a multiple of chain_len"; return SSQP_ERR_ARG; }
        // a warm start from the neighbour's optimum must be feasible: every QP of a chain has the same b, g, d, u
        const int N_ = ctx->N, M_ = ctx->M, J_ = ctx->J;
        auto same_in_chain = [&](const double* a, size_t len) {
            if (!a || len == 0) return true;
            for (int64_t c0 = 0; c0 < nb; c0 += U)
                for (int64_t t = 1; t < U; ++t)
                    if (memcmp(a + (size_t)c0 * len, a + (size_t)(c0 + t) * len, len * sizeof(double)) != 0) return false;
            return true;
        };
        if (!same_in_chain(b, (size_t)M_) || !same_in_chain(g, (size_t)J_) || !same_in_chain(d, (size_t)N_) || !same_in_chain(u, (size_t)N_)) {
            errs = "solve_sweep: the QPs of a chain must share b, g, d and u (only q may vary along a chain)"; return SSQP_ERR_ARG;
        }
    }
    if (!ctx->have_shared) { errs = "solve before ssqp_set_shared"; return SSQP_ERR_STATE; }
    const int N = ctx->N, M = ctx->M, J = ctx->J;
    if (nb < 0 || !d || !u || !x || !S || !status || (!phase1_only && !q) || (M > 0 && !b) || (J > 0 && !g)) {
        errs = "solve_batch: NULL argument"; return SSQP_ERR_ARG;
    }
    if (!Vq && !ctx->have_V && phase1_only == 0) { errs = "no V: pass V to ssqp_set_shared or V_per_qp"; return SSQP_ERR_STATE; }
    if ((S0 == nullptr) != (x0 == nullptr)) { errs = "warm start needs both S0 and x0"; return SSQP_ERR_ARG; }
    ssqp_settings st, stlp;
    ssqp_default_settings(&st);
    if (settings) st = *settings;
    stlp = settingsLP ? *settingsLP : st;
    int rc = check_settings(&st, &stlp, errs);
    if (rc) return rc;
    // Variables without a lower bound (src/SSQP.jl:484-509, 540-558): Phase 1 splits a free variable into two columns and
    // negates a (-Inf,u] one; the kernel needs the largest number of free variables of any QP to size its status array.
    // (SimplexLP has the same branches: src/Simplex.jl:861-887, 996-1032.)
    int nfree_cap = 0;
    {
        // (O(nb*N) reads of the caller's d and u — 0.5 GB for the named 65 536 x 500 batch: split over host threads so that
        // the scan stays out of the end-to-end time)
        const int64_t work_items = nb * (int64_t)N;
        int nth = work_items > (1 << 20) ? (int)std::thread::hardware_concurrency() : 1;
        if (nth > 16) nth = 16;
        if (nth < 1) nth = 1;
        std::vector<int> caps(nth, 0), bad(nth, 0);
        auto scan = [&](int t) {
            const int64_t i0 = nb * t / nth, i1 = nb * (t + 1) / nth;
            int cap = 0, nan = 0;
            for (int64_t i = i0; i < i1; ++i) {
                int nfv = 0;
                const double* di = d + (size_t)i * N; const double* ui = u + (size_t)i * N;
                for (int k = 0; k < N; ++k) {
                    const double dk = di[k], uk = ui[k];
                    if (dk != dk || uk != uk) nan = 1;
                    if (dk == -INFINITY && uk == INFINITY) nfv += 1;
                }
                if (nfv > cap) cap = nfv;
            }
            caps[t] = cap; bad[t] = nan;
        };
        if (nth == 1) scan(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nth; ++t) th.emplace_back(scan, t);
            for (auto& t : th) t.join();
        }
        for (int t = 0; t < nth; ++t) {
            if (bad[t]) { errs = "d / u must not be NaN"; return SSQP_ERR_ARG; }
            if (caps[t] > nfree_cap) nfree_cap = caps[t];
        }
    }
    // Phase-1 de-duplication: initQP depends on (A, G, b, g, d, u) only (src/SSQP.jl:461-530).  When those
    // are bit-identical for every QP of the batch (a frontier sweep over q = -L*E, src/types.jl:303-319), Phase 1 runs once
    // and every QP starts Phase 2 from that point — exactly the point its own Phase 1 would have produced.
    std::vector<double> x0_shared;
    std::vector<int32_t> S0_shared;
    bool bcast = false;
    if (phase1_only == 0 && !S0 && nb >= 2 && !getenv("SSQP_NO_DEDUP")) {
        auto same = [&](const double* a, size_t len) {
            if (!a || len == 0) return true;
            for (int64_t i = 1; i < nb; ++i)
                if (memcmp(a, a + (size_t)i * len, len * sizeof(double)) != 0) return false;
            return true;
        };
        if (same(g, (size_t)J) && same(b, (size_t)M) && same(u, (size_t)N) && same(d, (size_t)N)) {
            x0_shared.resize(N); S0_shared.resize(N + J);
            int64_t st1 = 0;
            rc = solve_host(ctx, 1, nullptr, nullptr, b, g, d, u, nullptr, nullptr, &stlp, &stlp, x0_shared.data(),
                            S0_shared.data(), &st1, 1);
            if (rc) return rc;
            if (st1 == 1) { S0 = S0_shared.data(); x0 = x0_shared.data(); bcast = true; }
        }
    }
    const int G_ = (int)ctx->dev.size();
    ctx->shard_cnt.assign(G_, 0);
    ctx->last_nb = nb;
    if (nb == 0) return SSQP_OK;

    ctx->bcast_start = bcast;          // launch parameters of this batch: set once, before the per-device threads start
    ctx->nfree_cap = nfree_cap;
    ctx->chain_len = (int)U;
    std::vector<int> rcs(G_, 0);
    std::vector<std::string> es(G_);
    auto work = [&](int gi) {
        Device& D = ctx->dev[gi];
        std::string& errs = es[gi];
        auto body = [&]() -> int {
            const int64_t units = nb / U;
            const int64_t ucnt = (units - gi + G_ - 1) / G_;  // units (QPs, or chains of U QPs) gi, gi+G, gi+2G, ...
            const int64_t cnt = ucnt * U;
            ctx->shard_cnt[gi] = cnt;      // (one slot per device thread)
            if (cnt <= 0) return SSQP_OK;
            CK(cudaSetDevice(D.id));
            auto h2d = [&](DevBuf& B, const void* src, size_t len1) -> int {    // interleaved gather of the shard
                const size_t len = len1 * (size_t)U;
                const int64_t cnt = ucnt;
                if (!src || len == 0) return SSQP_OK;
                CK(B.ensure(len * cnt));
                CK(cudaMemcpy2DAsync(B.p, len, (const char*)src + (size_t)gi * len, len * G_, len, cnt, cudaMemcpyHostToDevice, D.stream));
                return SSQP_OK;
            };
            int r;
            if ((r = h2d(D.q, q, (size_t)N * 8))) return r;
            if ((r = h2d(D.b, b, (size_t)M * 8))) return r;
            if ((r = h2d(D.g, g, (size_t)J * 8))) return r;
            if ((r = h2d(D.d, d, (size_t)N * 8))) return r;
            if ((r = h2d(D.u, u, (size_t)N * 8))) return r;
            if ((r = h2d(D.Vq, Vq, (size_t)N * N * 8))) return r;
            if (bcast) {       // one shared start point: a single row, read with stride 0
                CK(D.S0.ensure((size_t)(N + J) * 4)); CK(D.x0.ensure((size_t)N * 8));
                CK(cudaMemcpyAsync(D.S0.p, S0, (size_t)(N + J) * 4, cudaMemcpyHostToDevice, D.stream));
                CK(cudaMemcpyAsync(D.x0.p, x0, (size_t)N * 8, cudaMemcpyHostToDevice, D.stream));
            } else {
                if ((r = h2d(D.S0, S0, (size_t)(N + J) * 4))) return r;
                if ((r = h2d(D.x0, x0, (size_t)N * 8))) return r;
            }
            CK(D.x.ensure((size_t)N * 8 * cnt));
            CK(D.S.ensure((size_t)(N + J) * 4 * cnt));
            CK(D.status.ensure((size_t)8 * cnt));
            r = launch_solve(ctx, D, cnt, Vq ? D.Vq.as<double>() : nullptr, D.q.as<double>(), D.b.as<double>(),
                             D.g.as<double>(), D.d.as<double>(), D.u.as<double>(), S0 ? D.S0.as<int32_t>() : nullptr,
                             x0 ? D.x0.as<double>() : nullptr, st, stlp, D.x.as<double>(), D.S.as<int32_t>(),
                             D.status.as<int64_t>(), D.stream, phase1_only, errs);
            if (r) return r;
            {
                const size_t lx = (size_t)N * 8 * U, ls = (size_t)(N + J) * 4 * U, lt = (size_t)8 * U;
                CK(cudaMemcpy2DAsync((char*)x + (size_t)gi * lx, lx * G_, D.x.p, lx, lx, ucnt, cudaMemcpyDeviceToHost, D.stream));
                CK(cudaMemcpy2DAsync((char*)S + (size_t)gi * ls, ls * G_, D.S.p, ls, ls, ucnt, cudaMemcpyDeviceToHost, D.stream));
                CK(cudaMemcpy2DAsync((char*)status + (size_t)gi * lt, lt * G_, D.status.p, lt, lt, ucnt, cudaMemcpyDeviceToHost, D.stream));
            }
            CK(cudaStreamSynchronize(D.stream));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, D.ev0, D.ev1));
            D.last_ms = ms;
            return SSQP_OK;
        };
        rcs[gi] = body();
    };
    if (G_ == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int gi = 0; gi < G_; ++gi) th.emplace_back(work, gi);
        for (auto& t : th) t.join();
    }
    for (int gi = 0; gi < G_; ++gi)
        if (rcs[gi]) { errs = es[gi]; return rcs[gi]; }
    return SSQP_OK;
}

int ssqp_solve_batch(ssqp_ctx* ctx, int64_t nb, const double* V_per_qp, const double* q, const double* b, const double* g,
                     const double* d, const double* u, const int32_t* S0, const double* x0, const ssqp_settings* settings,
                     const ssqp_settings* settingsLP, double* x, int32_t* S, int64_t* status) {
    return solve_host(ctx, nb, V_per_qp, q, b, g, d, u, S0, x0, settings, settingsLP, x, S, status, 0);
}

int ssqp_solve_sweep(ssqp_ctx* ctx, int64_t nb, int64_t chain_len, const double* V_per_qp, const double* q, const double* b,
                     const double* g, const double* d, const double* u, const ssqp_settings* settings,
                     const ssqp_settings* settingsLP, double* x, int32_t* S, int64_t* status) {
    return solve_host(ctx, nb, V_per_qp, q, b, g, d, u, nullptr, nullptr, settings, settingsLP, x, S, status, 0, chain_len);
}

int ssqp_solve_lp_batch(ssqp_ctx* ctx, int64_t nb, const double* c, const double* b, const double* g, const double* d,
                        const double* u, const ssqp_settings* settings, double* x, int32_t* S, int64_t* status) {
    if (!c) { if (ctx) ctx->err = "solve_lp_batch: NULL cost vector"; return SSQP_ERR_ARG; }
    return solve_host(ctx, nb, nullptr, c, b, g, d, u, nullptr, nullptr, settings, settings, x, S, status, 2);
}

int ssqp_init_batch(ssqp_ctx* ctx, int64_t nb, const double* b, const double* g, const double* d, const double* u,
                    const ssqp_settings* settingsLP, double* x0, int32_t* S, int64_t* status) {
    return solve_host(ctx, nb, nullptr, nullptr, b, g, d, u, nullptr, nullptr, settingsLP, settingsLP, x0, S, status, 1);
}

int ssqp_solve_batch_device(ssqp_ctx* ctx, int64_t nb, const double* V_per_qp, const double* q, const double* b,
                            const double* g, const double* d, const double* u, const int32_t* S0, const double* x0,
                            const ssqp_settings* settings, const ssqp_settings* settingsLP, double* x, int32_t* S,
                            int64_t* status, void* stream) {
    if (!ctx) return SSQP_ERR_ARG;
    std::string& errs = ctx->err;
    if (!ctx->have_shared) { errs = "solve before ssqp_set_shared"; return SSQP_ERR_STATE; }
    if (nb < 0 || !q || !d || !u || !x || !S || !status || (ctx->M > 0 && !b) || (ctx->J > 0 && !g)) {
        errs = "solve_batch_device: NULL argument"; return SSQP_ERR_ARG;
    }
    if ((S0 == nullptr) != (x0 == nullptr)) { errs = "warm start needs both S0 and x0"; return SSQP_ERR_ARG; }
    if (V_per_qp && ((uintptr_t)V_per_qp & 7)) { errs = "V_per_qp must be 8-byte aligned"; return SSQP_ERR_ARG; }
    if (!V_per_qp && !ctx->have_V) { errs = "no V: pass V to ssqp_set_shared or V_per_qp"; return SSQP_ERR_STATE; }
    ssqp_settings st, stlp;
    ssqp_default_settings(&st);
    if (settings) st = *settings;
    stlp = settingsLP ? *settingsLP : st;
    int rc = check_settings(&st, &stlp, errs);
    if (rc) return rc;
    Device& D = ctx->dev[0];
    CK(cudaSetDevice(D.id));
    ctx->last_nb = nb;
    cudaStream_t s = stream ? (cudaStream_t)stream : D.stream;
    ctx->bcast_start = false;
    ctx->chain_len = 1;
    // device-pointer entry: the bounds are not scanned on the host (that would need a synchronisation); the caller states
    // how many free variables (d = -Inf and u = +Inf) a QP may have with ssqp_set_free_var_capacity (default 0); a QP with
    // more gets status -1
    ctx->nfree_cap = ctx->free_cap_device;
    return launch_solve(ctx, D, nb, V_per_qp, q, b, g, d, u, S0, x0, st, stlp, x, S, status, s, 0, errs);
}

int ssqp_set_free_var_capacity(ssqp_ctx* ctx, int32_t max_free_vars_per_qp) {
    if (!ctx || max_free_vars_per_qp < 0) return SSQP_ERR_ARG;
    ctx->free_cap_device = max_free_vars_per_qp;
    return SSQP_OK;
}

int ssqp_get_stats(ssqp_ctx* ctx, int64_t nb, double* stats) {
    if (!ctx || !stats) return SSQP_ERR_ARG;
    std::string& errs = ctx->err;
    if (nb != ctx->last_nb) { errs = "get_stats: nb differs from the last batch"; return SSQP_ERR_ARG; }
    const int G_ = (int)ctx->dev.size();
    for (int gi = 0; gi < G_; ++gi) {
        Device& D = ctx->dev[gi];
        const int64_t U = ctx->chain_len > 1 ? ctx->chain_len : 1;       // units of the last batch (chains stay together)
        const int64_t ucnt = (nb / U - gi + G_ - 1) / G_;
        if (ucnt <= 0) continue;
        CK(cudaSetDevice(D.id));
        const size_t lu = (size_t)NSTATS * 8 * U;
        CK(cudaMemcpy2D((char*)stats + (size_t)gi * lu, lu * G_, D.stats.p, lu, lu, ucnt, cudaMemcpyDeviceToHost));
    }
    return SSQP_OK;
}

int ssqp_get_stats_device(ssqp_ctx* ctx, int64_t nb, double* stats_host) {
    if (!ctx || !stats_host) return SSQP_ERR_ARG;
    std::string& errs = ctx->err;
    Device& D = ctx->dev[0];
    if (nb != ctx->last_nb || D.stats.cap < (size_t)nb * NSTATS * 8) { errs = "get_stats_device: nb differs from the last batch"; return SSQP_ERR_ARG; }
    CK(cudaSetDevice(D.id));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(stats_host, D.stats.p, (size_t)nb * NSTATS * 8, cudaMemcpyDeviceToHost));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, D.ev0, D.ev1) == cudaSuccess) D.last_ms = ms; else cudaGetLastError();
    return SSQP_OK;
}

double ssqp_measure_fp64_peak(ssqp_ctx* ctx) {
    if (!ctx) return -1.0;
    Device& D = ctx->dev[0];
    if (cudaSetDevice(D.id) != cudaSuccess) return -1.0;
    const int blocks = D.sms * 8, threads = 256, iters = 20000;
    double* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) return -1.0;
    ssqp_dfma_kernel<<<blocks, threads, 0, D.stream>>>(out, 100);
    cudaEventRecord(D.ev0, D.stream);
    ssqp_dfma_kernel<<<blocks, threads, 0, D.stream>>>(out, iters);
    cudaEventRecord(D.ev1, D.stream);
    cudaStreamSynchronize(D.stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, D.ev0, D.ev1);
    cudaFree(out);
    ctx->launches += 2;
    const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
    return flops / (ms * 1e-3) / 1e12;
}

double ssqp_measure_read_bw(ssqp_ctx* ctx, int32_t mbytes, int32_t reps) {
    if (!ctx || mbytes < 1 || reps < 1) return -1.0;
    Device& D = ctx->dev[0];
    if (cudaSetDevice(D.id) != cudaSuccess) return -1.0;
    const size_t bytes = (size_t)mbytes << 20;
    double2* in = nullptr; double* out = nullptr;
    const int blocks = D.sms * 8, threads = 256;
    if (cudaMalloc(&in, bytes) != cudaSuccess) return -1.0;
    if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) { cudaFree(in); return -1.0; }
    cudaMemsetAsync(in, 0, bytes, D.stream);
    ssqp_readbw_kernel<<<blocks, threads, 0, D.stream>>>(in, (long long)(bytes / 16), 1, out);
    cudaEventRecord(D.ev0, D.stream);
    ssqp_readbw_kernel<<<blocks, threads, 0, D.stream>>>(in, (long long)(bytes / 16), reps, out);
    cudaEventRecord(D.ev1, D.stream);
    cudaStreamSynchronize(D.stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, D.ev0, D.ev1);
    cudaFree(in); cudaFree(out);
    ctx->launches += 2;
    return (double)bytes * reps / (ms * 1e-3) / 1e9;
}

}  // extern "C"
