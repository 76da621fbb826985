// ssqp_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A from-scratch C++ restatement, in *reference form*, of the algorithm of
// PharosAbad/StatusSwitchingQP.jl v1.0.2 for the `solveQP` hot path:
//   src/SSQP.jl:10-32    polishSz!      -> polish_sz
//   src/SSQP.jl:35-59    freeK!         -> free_k
//   src/SSQP.jl:61-134   aStep!         -> a_step
//   src/SSQP.jl:136-188  KKTchk!        -> kkt_chk
//   src/SSQP.jl:224-234  solveQP(Q)     -> ssqp_oracle_solve
//   src/SSQP.jl:237-377  solveQP(Q,S,x0)-> solve_phase2
//   src/SSQP.jl:461-560  initQP         -> init_qp
//   src/Simplex.jl:445-615 cDantzigLP   -> c_dantzig_lp
//   src/Simplex.jl:831-1034 SimplexLP   -> simplex_lp (free and (-Inf,u] variables included)
//   src/utils.jl:49-86   getRowsGJr     -> get_rows_gjr
// "Reference form" = refactorise every trip with explicit inverses
// (inv(cholesky(.)), inv(lu(.)) on every simplex pivot), i.e. the reference's
// own operation count.  Julia's stdlib LinearAlgebra (OpenBLAS/LAPACK dpotrf/
// dpotri/dgetrf/dgetri/dgemm, unpinned version) is replaced by the small dense
// routines below, so floating-point roundoff differs from Julia's at the
// 1e-16 level; decisions are thresholded at tol=2^-26 / tolG=2^-33.
//
// PARITY STATUS: pinned only by the reference's two known-answer tests
// (test/runtests.jl:22-32 Status[UP,IN,IN] through solveQP; test/runtests.jl:7-19 status 3 through the
// on-path cDantzigLP on the slack form of that LP), by optimality certificates recomputed independently
// in numpy/scipy (tests/test_oracle.py) and by the committed golden vectors (tests/golden/).  The reference
// itself (Julia) cannot be executed in this environment: x/iteration-count
// parity is "parity unpinned" beyond those KATs.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (libssqp_b200.so) never
// links or calls it.
//
// Exceptions in Julia (PosDefException / SingularException) propagate out of
// solveQP; here they are mapped to status = -1 (numerical error).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef std::vector<double> vec;
typedef std::vector<int> ivec;
const double INF = std::numeric_limits<double>::infinity();

enum : int32_t { IN = 0, DN = 1, UP = 2, OE = 3, EO = 4 };   // src/types.jl:17-23

// column-major dense matrix (Julia layout)
struct Mat {
    int r = 0, c = 0;
    vec a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    inline double& operator()(int i, int j) { return a[(size_t)i + (size_t)j * r]; }
    inline double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * r]; }
};

struct NumErr {};   // stands for PosDefException / SingularException

// ---- LAPACK form -------------------------------------------------------------------------------------------
// Julia's stdlib LinearAlgebra runs the dense algebra of this path on OpenBLAS/LAPACK: inv(cholesky(X)) = dpotrf('U')
// + dpotri (src/SSQP.jl:322,328), inv(lu(X)) = dgetrf + dgetri (src/Simplex.jl:590), every `*` = dgemm / dgemv.  When the
// Python loader hands over the Fortran-ABI entry points of scipy's bundled OpenBLAS (ssqp_oracle_set_lapack; taken from
// scipy.linalg.cython_lapack / cython_blas.__pyx_capi__, LP64 ints, BLAS threads pinned to 1), the routines below call the
// same library family and the same routines as the reference; otherwise (or after ssqp_oracle_use_lapack(0)) the scalar
// loops run.  Both forms stay available: the tests compare them (decisions must not depend on LAPACK-level roundoff).
typedef void (*potrf_fn)(char*, int*, double*, int*, int*);
typedef void (*getrf_fn)(int*, int*, double*, int*, int*, int*);
typedef void (*getri_fn)(int*, double*, int*, int*, double*, int*, int*);
typedef void (*gemm_fn)(char*, char*, int*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
typedef void (*gemv_fn)(char*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
struct Lapack {
    potrf_fn potrf = nullptr, potri = nullptr;
    getrf_fn getrf = nullptr;
    getri_fn getri = nullptr;
    gemm_fn gemm = nullptr;
    gemv_fn gemv = nullptr;
    int have = 0, on = 0;
} g_la;
inline bool la_on() { return g_la.on != 0; }
inline void la_gemm(char ta, char tb, int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    double one = 1.0, zero = 0.0;
    if (m == 0 || n == 0) return;
    if (k == 0) { for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) C[i + (size_t)j * ldc] = 0.0; return; }
    g_la.gemm(&ta, &tb, &m, &n, &k, &one, const_cast<double*>(A), &lda, const_cast<double*>(B), &ldb, &zero, C, &ldc);
}
inline void la_gemv(char t, int m, int n, const double* A, int lda, const double* x, double* y) {
    double one = 1.0, zero = 0.0;
    int inc = 1;
    g_la.gemv(&t, &m, &n, &one, const_cast<double*>(A), &lda, const_cast<double*>(x), &inc, &zero, y, &inc);
}

// C = A * B
Mat matmul(const Mat& A, const Mat& B) {
    Mat C(A.r, B.c);
    if (la_on()) { la_gemm('N', 'N', A.r, B.c, A.c, A.a.data(), A.r > 0 ? A.r : 1, B.a.data(), B.r > 0 ? B.r : 1, C.a.data(), C.r > 0 ? C.r : 1); return C; }
    for (int j = 0; j < B.c; ++j)
        for (int k = 0; k < A.c; ++k) {
            double b = B(k, j);
            if (b == 0.0) continue;
            const double* ap = &A.a[(size_t)k * A.r];
            double* cp = &C.a[(size_t)j * C.r];
            for (int i = 0; i < A.r; ++i) cp[i] += ap[i] * b;
        }
    return C;
}
// C = A * B'
Mat matmul_nt(const Mat& A, const Mat& B) {
    Mat C(A.r, B.r);
    if (la_on()) { la_gemm('N', 'T', A.r, B.r, A.c, A.a.data(), A.r > 0 ? A.r : 1, B.a.data(), B.r > 0 ? B.r : 1, C.a.data(), C.r > 0 ? C.r : 1); return C; }
    for (int k = 0; k < A.c; ++k)
        for (int j = 0; j < B.r; ++j) {
            double b = B(j, k);
            if (b == 0.0) continue;
            const double* ap = &A.a[(size_t)k * A.r];
            double* cp = &C.a[(size_t)j * C.r];
            for (int i = 0; i < A.r; ++i) cp[i] += ap[i] * b;
        }
    return C;
}
// y = A * x
vec matvec(const Mat& A, const vec& x) {
    vec y(A.r, 0.0);
    if (la_on() && A.r > 0 && A.c > 0) { la_gemv('N', A.r, A.c, A.a.data(), A.r, x.data(), y.data()); return y; }
    for (int k = 0; k < A.c; ++k) {
        double b = x[k];
        if (b == 0.0) continue;
        const double* ap = &A.a[(size_t)k * A.r];
        for (int i = 0; i < A.r; ++i) y[i] += ap[i] * b;
    }
    return y;
}
// y = A' * x
vec matvec_t(const Mat& A, const vec& x) {
    vec y(A.c, 0.0);
    if (la_on() && A.r > 0 && A.c > 0) { la_gemv('T', A.r, A.c, A.a.data(), A.r, x.data(), y.data()); return y; }
    for (int k = 0; k < A.c; ++k) {
        const double* ap = &A.a[(size_t)k * A.r];
        double s = 0.0;
        for (int i = 0; i < A.r; ++i) s += ap[i] * x[i];
        y[k] = s;
    }
    return y;
}

// inv(cholesky(X)) for symmetric X  (dpotrf + dpotri in the reference, src/SSQP.jl:322,328)
Mat inv_cholesky(const Mat& X) {
    int n = X.r;
    if (la_on() && n > 0) {      // cholesky(X) reads the upper triangle (dpotrf 'U'); inv(::Cholesky) = dpotri + copytri!
        Mat R = X;
        char U = 'U';
        int info = 0, lda = n;
        g_la.potrf(&U, &n, R.a.data(), &lda, &info);
        if (info != 0) throw NumErr();                 // PosDefException
        g_la.potri(&U, &n, R.a.data(), &lda, &info);
        if (info != 0) throw NumErr();                 // SingularException
        for (int j = 0; j < n; ++j)
            for (int i = j + 1; i < n; ++i) R(i, j) = R(j, i);
        return R;
    }
    Mat L(n, n);
    // lower Cholesky, column by column
    for (int j = 0; j < n; ++j) {
        double d = X(j, j);
        for (int k = 0; k < j; ++k) d -= L(j, k) * L(j, k);
        if (!(d > 0.0)) throw NumErr();
        d = std::sqrt(d);
        L(j, j) = d;
        for (int i = j + 1; i < n; ++i) {
            double s = X(i, j);
            for (int k = 0; k < j; ++k) s -= L(i, k) * L(j, k);
            L(i, j) = s / d;
        }
    }
    // Li = inv(L) (lower)
    Mat Li(n, n);
    for (int j = 0; j < n; ++j) {
        Li(j, j) = 1.0 / L(j, j);
        for (int i = j + 1; i < n; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s -= L(i, k) * Li(k, j);
            Li(i, j) = s / L(i, i);
        }
    }
    // inv(X) = Li' * Li
    Mat R(n, n);
    for (int j = 0; j < n; ++j)
        for (int i = j; i < n; ++i) {
            double s = 0.0;
            for (int k = i; k < n; ++k) s += Li(k, i) * Li(k, j);
            R(i, j) = s;
            R(j, i) = s;
        }
    return R;
}

// inv(lu(X)) with partial pivoting (dgetrf + dgetri in the reference, src/Simplex.jl:590)
Mat inv_lu(const Mat& X) {
    int n = X.r;
    if (la_on() && n > 0) {      // lu(X) = dgetrf (SingularException when a pivot is exactly zero), inv(::LU) = dgetri
        Mat R = X;
        std::vector<int> ipiv(n);
        int info = 0, lda = n;
        g_la.getrf(&n, &n, R.a.data(), &lda, ipiv.data(), &info);
        if (info != 0) throw NumErr();
        int lwork = 64 * n;
        vec work((size_t)lwork);
        g_la.getri(&n, R.a.data(), &lda, ipiv.data(), work.data(), &lwork, &info);
        if (info != 0) throw NumErr();
        return R;
    }
    Mat A = X;
    ivec piv(n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double m = std::fabs(A(k, k));
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(A(i, k)) > m) { m = std::fabs(A(i, k)); p = i; }
        piv[k] = p;
        if (m == 0.0) throw NumErr();   // SingularException
        if (p != k)
            for (int j = 0; j < n; ++j) std::swap(A(k, j), A(p, j));
        double d = A(k, k);
        for (int i = k + 1; i < n; ++i) A(i, k) /= d;
        for (int j = k + 1; j < n; ++j) {
            double t = A(k, j);
            if (t == 0.0) continue;
            double* cj = &A.a[(size_t)j * n];
            const double* ck = &A.a[(size_t)k * n];
            for (int i = k + 1; i < n; ++i) cj[i] -= ck[i] * t;
        }
    }
    // solve A * Xinv = P * I column by column
    Mat R(n, n);
    vec y(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) y[i] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < n; ++k) std::swap(y[k], y[piv[k]]);
        for (int k = 0; k < n; ++k) {           // forward, unit lower (column sweep)
            double t = y[k];
            if (t == 0.0) continue;
            const double* ck = &A.a[(size_t)k * n];
            for (int i = k + 1; i < n; ++i) y[i] -= ck[i] * t;
        }
        for (int k = n - 1; k >= 0; --k) {      // backward, upper (column sweep)
            y[k] /= A(k, k);
            double t = y[k];
            if (t == 0.0) continue;
            const double* ck = &A.a[(size_t)k * n];
            for (int i = 0; i < k; ++i) y[i] -= ck[i] * t;
        }
        for (int i = 0; i < n; ++i) R(i, c) = y[i];
    }
    return R;
}

// Minimum-norm least-squares solve  x = X \ y  for a tall/any K x W matrix X via
// column-pivoted Householder QR (Julia `\` on a rectangular matrix, src/SSQP.jl:158).
vec lstsq(const Mat& Xin, const vec& yin) {
    int m = Xin.r, n = Xin.c;
    Mat A = Xin;
    vec y = yin;
    ivec perm(n);
    for (int j = 0; j < n; ++j) perm[j] = j;
    int rk = 0;
    int kmax = std::min(m, n);
    vec cn(n);
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += A(i, j) * A(i, j);
        cn[j] = s;
    }
    double tol0 = 0.0;
    for (int k = 0; k < kmax; ++k) {
        int p = k;
        for (int j = k; j < n; ++j) {
            double s = 0;
            for (int i = k; i < m; ++i) s += A(i, j) * A(i, j);
            cn[j] = s;
            if (cn[j] > cn[p]) p = j;
        }
        if (k == 0) tol0 = std::sqrt(cn[p]) * 2.220446049250313e-16 * std::max(m, n);
        if (std::sqrt(cn[p]) <= tol0) break;
        if (p != k) {
            for (int i = 0; i < m; ++i) std::swap(A(i, k), A(i, p));
            std::swap(perm[k], perm[p]);
            std::swap(cn[k], cn[p]);         // (round-2 fix: the norm travels with its column; without it the reflector was
                                             // built from the wrong norm and the purged-row multipliers of KKTchk! were wrong —
                                             // found by the full-shard parity check, tests/test_oracle.py::test_lstsq_matches_numpy)
        }
        double nrm = std::sqrt(cn[k]);
        double alpha = A(k, k) > 0 ? -nrm : nrm;
        vec v(m - k);
        for (int i = k; i < m; ++i) v[i - k] = A(i, k);
        v[0] -= alpha;
        double vn = 0;
        for (double t : v) vn += t * t;
        if (vn > 0) {
            for (int j = k; j < n; ++j) {
                double s = 0;
                for (int i = k; i < m; ++i) s += v[i - k] * A(i, j);
                s = 2 * s / vn;
                for (int i = k; i < m; ++i) A(i, j) -= s * v[i - k];
            }
            double s = 0;
            for (int i = k; i < m; ++i) s += v[i - k] * y[i];
            s = 2 * s / vn;
            for (int i = k; i < m; ++i) y[i] -= s * v[i - k];
        }
        rk = k + 1;
    }
    // basic solution on the leading rk columns (full-rank case == least squares solution)
    vec z(n, 0.0);
    for (int k = rk - 1; k >= 0; --k) {
        double s = y[k];
        for (int j = k + 1; j < rk; ++j) s -= A(k, j) * z[j];
        z[k] = s / A(k, k);
    }
    vec x(n, 0.0);
    for (int j = 0; j < n; ++j) x[perm[j]] = z[j];
    return x;
}

// ---------------------------------------------------------------------------------------
// getRowsGJr  (src/utils.jl:49-86): Gauss-Jordan with in-row column pivoting; returns the
// independent rows (0-based) and l1.
// ---------------------------------------------------------------------------------------
void get_rows_gjr(const Mat& X, double tol, ivec& rows, int& l1) {
    Mat A = X;
    int nr = A.r, nc = A.c;
    rows.clear();
    ivec c0(nc);
    for (int k = 0; k < nc; ++k) c0[k] = k;
    l1 = 0;
    int i = 0, j = 0;
    while (i < nr && j < nc) {
        double m = -1.0;
        int mj = j;
        for (int k = j; k < nc; ++k) {            // findmax: first maximal element
            double t = std::fabs(A(i, c0[k]));
            if (t > m) { m = t; mj = k; }
        }
        if (m <= tol) {
            i += 1;
        } else {
            rows.push_back(i);
            std::swap(c0[mj], c0[j]);
            int n = c0[j];
            double d = A(i, n);
            for (int k = j; k < nc; ++k) A(i, c0[k]) /= d;
            for (int k = 0; k < nr; ++k) {
                if (k != i) {
                    double dk = A(k, n);
                    for (int l = j; l < nc; ++l) A(k, c0[l]) -= dk * A(i, c0[l]);
                }
            }
            l1 = j + 1;
            i += 1;
            j += 1;
        }
    }
}

// ---------------------------------------------------------------------------------------
// cDantzigLP  (src/Simplex.jl:445-615).  B (sorted, 0-based) and S are mutated.
// returns status 1/2/3; x (length N) and invB updated.
// counters: loops, pivots (for F_alg accounting)
// ---------------------------------------------------------------------------------------
struct LPStats { int64_t loops = 0, pivots = 0, flips = 0; };

int c_dantzig_lp(const vec& c, const Mat& A, const vec& b, const vec& d, const vec& u,
                 ivec& B, std::vector<int32_t>& S, Mat& invB, vec q, double tol, vec& x,
                 LPStats* st) {
    int N = (int)c.size();
    int M = (int)b.size();
    std::vector<char> F(N, 1);
    for (int j = 0; j < M; ++j) F[B[j]] = 0;
    vec gt(M, 0.0);
    ivec ip(M, 0);
    std::vector<int32_t> Sb(M, DN);

    vec ud(N), du(N);
    std::vector<char> fu(N);
    for (int k = 0; k < N; ++k) { ud[k] = u[k] - d[k]; du[k] = -ud[k]; fu[k] = u[k] < INF; }

    vec cA(N, 0.0);
    x = d;
    for (int k = 0; k < N; ++k) {
        double s = 0.0;
        for (int i = 0; i < M; ++i) s += A(i, k) * A(i, k);
        cA[k] = std::sqrt(s);
    }
    for (int k = 0; k < N; ++k)
        if (S[k] == UP) x[k] = u[k];

    // helpers working on the current F/B
    ivec iF;           // findall(F)
    Mat Y;             // invB * A[:,F]
    vec h;             // signed reduced costs over F
    ivec iH;           // candidate variable ids
    vec hp;            // candidate values

    auto build_iF = [&]() {
        iF.clear();
        for (int k = 0; k < N; ++k) if (F[k]) iF.push_back(k);
    };
    auto compute_Y = [&]() {
        Mat AF(M, (int)iF.size());
        for (size_t t = 0; t < iF.size(); ++t)
            std::memcpy(&AF.a[t * M], &A.a[(size_t)iF[t] * M], sizeof(double) * M);
        Y = matmul(invB, AF);
    };
    auto compute_h = [&]() {
        vec cB(M);
        for (int j = 0; j < M; ++j) cB[j] = c[B[j]];
        vec yc = matvec_t(Y, cB);
        h.resize(iF.size());
        iH.clear();
        hp.clear();
        for (size_t t = 0; t < iF.size(); ++t) {
            double v = c[iF[t]] - yc[t];
            if (S[iF[t]] == DN) v = -v;
            h[t] = v;
            if (v > tol) { iH.push_back(iF[t]); hp.push_back(v); }
        }
    };

    build_iF();
    compute_Y();
    compute_h();

    int nH = (int)iH.size();
    bool Bland = false;
    int64_t loop = 0;
    while (nH > 0) {
        loop += 1;
        if (st) st->loops += 1;
        if (loop > N) Bland = true;

        int k0 = 0;
        if (!Bland) {                     // argmax(hp ./ cA[iH]): first maximum
            double best = hp[0] / cA[iH[0]];
            for (int t = 1; t < nH; ++t) {
                double v = hp[t] / cA[iH[t]];
                if (v > best) { best = v; k0 = t; }
            }
        }
        int k = iH[k0];
        vec p(M, 0.0);
        for (int jj = 0; jj < M; ++jj) {
            double a = A(jj, k);
            if (a == 0.0) continue;
            for (int i = 0; i < M; ++i) p[i] += invB(i, jj) * a;
        }
        bool kd = (S[k] == DN);
        int m = 0;
        int l = 0;           // 1-based row, or -1 / -2 for flips (reference convention)
        int32_t Sl = DN;
        if (kd) {
            for (int j = 0; j < M; ++j) {
                int i = B[j];
                if (p[j] > tol) {
                    gt[m] = (q[j] - d[i]) / p[j]; ip[m] = j; Sb[m] = DN; m += 1;
                } else if (p[j] < -tol) {
                    gt[m] = (q[j] - u[i]) / p[j]; ip[m] = j; Sb[m] = UP; m += 1;
                }
            }
            if (m == 0) {
                if (fu[k]) {
                    l = -1;
                } else {
                    for (int j = 0; j < M; ++j) x[B[j]] = q[j];
                    return 3;
                }
            } else {
                double gl = gt[0]; int li = 0;           // findmin: first minimum
                for (int t = 1; t < m; ++t) if (gt[t] < gl) { gl = gt[t]; li = t; }
                if (fu[k]) {
                    if (gl >= ud[k]) {
                        l = -1;
                    } else {
                        Sl = Sb[li]; l = ip[li] + 1;
                    }
                } else {
                    if (std::isinf(gl)) {
                        for (int j = 0; j < M; ++j) x[B[j]] = q[j];
                        return 3;
                    }
                    Sl = Sb[li]; l = ip[li] + 1;
                }
            }
        } else {
            for (int j = 0; j < M; ++j) {
                int i = B[j];
                if (p[j] > tol) {
                    gt[m] = (q[j] - u[i]) / p[j]; ip[m] = j; Sb[m] = UP; m += 1;
                } else if (p[j] < -tol) {
                    gt[m] = (q[j] - d[i]) / p[j]; ip[m] = j; Sb[m] = DN; m += 1;
                }
            }
            if (m == 0) {
                l = -2;
            } else {
                double gl = gt[0]; int li = 0;           // findmax: first maximum
                for (int t = 1; t < m; ++t) if (gt[t] > gl) { gl = gt[t]; li = t; }
                if (gl <= du[k]) {
                    l = -2;
                } else {
                    Sl = Sb[li]; l = ip[li] + 1;
                }
            }
        }

        if (l == -1) {
            S[k] = UP; x[k] = u[k];
            if (st) st->flips += 1;
        } else if (l == -2) {
            S[k] = DN; x[k] = d[k];
            if (st) st->flips += 1;
        } else if (l > 0) {
            int mrow = l - 1;
            int lv = B[mrow];
            F[k] = 0; F[lv] = 1;
            B[mrow] = k;
            std::sort(B.begin(), B.end());
            Mat AB(M, M);
            for (int j = 0; j < M; ++j)
                std::memcpy(&AB.a[(size_t)j * M], &A.a[(size_t)B[j] * M], sizeof(double) * M);
            invB = inv_lu(AB);
            S[k] = IN;
            S[lv] = Sl;
            x[lv] = (Sl == DN) ? d[lv] : u[lv];
            build_iF();
            compute_Y();
            if (st) st->pivots += 1;
        }

        // q = invB*b - Y*x[F]
        {
            vec xF(iF.size());
            for (size_t t = 0; t < iF.size(); ++t) xF[t] = x[iF[t]];
            vec ib = matvec(invB, b);
            vec yx = matvec(Y, xF);
            for (int j = 0; j < M; ++j) q[j] = ib[j] - yx[j];
        }
        compute_h();
        nH = (int)iH.size();
    }

    for (int j = 0; j < M; ++j) x[B[j]] = q[j];
    bool ms = false;
    for (double v : h) if (std::fabs(v) < tol) { ms = true; break; }
    return ms ? 2 : 1;
}

// ---------------------------------------------------------------------------------------
// stpEdgeLP (src/Simplex.jl:234-416, the second definition — the one in effect) and maxImprvLP (src/Simplex.jl:641-813):
// the reference's other two pivot rules, same calling convention as cDantzigLP.  rule 1 = :stpEdgeLP, 2 = :maxImprovement.
//   stpEdge : entering variable = argmax h^2 / (1 + |Y[:,k]|^2) over the candidates; after a pivot with a zero step the
//             rule falls back to the first candidate once (`Edge = false; continue`), Edge returns with the next pivot.
//   maxImprv: a full ratio test for EVERY candidate, entering variable = argmax |h .* g| ; no anti-cycling safeguard.
// ---------------------------------------------------------------------------------------
int c_alt_rule_lp(int rule, const vec& c, const Mat& A, const vec& b, const vec& d, const vec& u,
                  ivec& B, std::vector<int32_t>& S, Mat& invB, vec q, double tol, vec& x, LPStats* st) {
    int N = (int)c.size();
    int M = (int)b.size();
    std::vector<char> F(N, 1);
    for (int j = 0; j < M; ++j) F[B[j]] = 0;
    vec gt(M, 0.0);
    ivec ip(M, 0);
    std::vector<int32_t> Sb(M, DN);
    vec ud(N), du(N);
    std::vector<char> fu(N);
    for (int k = 0; k < N; ++k) { ud[k] = u[k] - d[k]; du[k] = -ud[k]; fu[k] = u[k] < INF; }
    x = d;
    for (int k = 0; k < N; ++k) if (S[k] == UP) x[k] = u[k];

    ivec iF; Mat Y; vec h; ivec iH; vec hp; ivec ih;      // ih: positions of the candidates inside F
    auto build_iF = [&]() { iF.clear(); for (int k = 0; k < N; ++k) if (F[k]) iF.push_back(k); };
    auto compute_Y = [&]() {
        Mat AF(M, (int)iF.size());
        for (size_t t = 0; t < iF.size(); ++t) std::memcpy(&AF.a[t * M], &A.a[(size_t)iF[t] * M], sizeof(double) * M);
        Y = matmul(invB, AF);
    };
    auto compute_h = [&]() {
        vec cB(M);
        for (int j = 0; j < M; ++j) cB[j] = c[B[j]];
        vec yc = matvec_t(Y, cB);
        h.resize(iF.size()); iH.clear(); hp.clear(); ih.clear();
        for (size_t t = 0; t < iF.size(); ++t) {
            double v = c[iF[t]] - yc[t];
            if (S[iF[t]] == DN) v = -v;
            h[t] = v;
            if (v > tol) { iH.push_back(iF[t]); hp.push_back(v); ih.push_back((int)t); }
        }
    };
    // ratio test of column p for entering variable k.  variant 0: cDantzigLP / stpEdgeLP form; 1: maxImprvLP form.
    // returns l (-1 / -2 flips, >0 1-based row, 0 = unbounded), step gl, leaving state Sl
    auto ratio = [&](int k, const double* p, int variant, double& gl, int32_t& Sl) -> int {
        bool kd = (S[k] == DN);
        int m = 0;
        Sl = DN; gl = 0.0;
        if (kd) {
            for (int j = 0; j < M; ++j) {
                int i = B[j];
                if (p[j] > tol) { gt[m] = (q[j] - d[i]) / p[j]; ip[m] = j; Sb[m] = DN; m += 1; }
                else if (p[j] < -tol) { gt[m] = (q[j] - u[i]) / p[j]; ip[m] = j; Sb[m] = UP; m += 1; }
            }
            if (m == 0) {
                if (fu[k]) { gl = ud[k]; return -1; }
                return 0;
            }
            int li = 0; gl = gt[0];
            for (int t = 1; t < m; ++t) if (gt[t] < gl) { gl = gt[t]; li = t; }
            if (fu[k]) {
                if (gl >= ud[k]) { gl = ud[k]; return -1; }
            } else if (std::isinf(gl)) return 0;
            Sl = Sb[li];
            return ip[li] + 1;
        }
        for (int j = 0; j < M; ++j) {
            int i = B[j];
            if (p[j] > tol && (variant == 0 || u[i] < INF)) { gt[m] = (q[j] - u[i]) / p[j]; ip[m] = j; Sb[m] = UP; m += 1; }
            else if (p[j] < -tol) { gt[m] = (q[j] - d[i]) / p[j]; ip[m] = j; Sb[m] = DN; m += 1; }
        }
        if (m == 0) { gl = du[k]; return -2; }
        int li = 0; gl = gt[0];
        for (int t = 1; t < m; ++t) if (gt[t] > gl) { gl = gt[t]; li = t; }
        if (gl <= du[k]) { gl = du[k]; return -2; }
        Sl = Sb[li];
        return ip[li] + 1;
    };

    build_iF(); compute_Y(); compute_h();
    int nH = (int)iH.size();
    bool Edge = true;
    while (nH > 0) {
        if (st) st->loops += 1;
        int k0 = 0, l; double gl; int32_t Sl;
        if (rule == 1) {
            if (Edge) {                                   // se = hp.^2 ./ (sum(Y[:,ih].^2, dims=1) .+ 1); argmax: first maximum
                double best = -1.0;
                for (int t = 0; t < nH; ++t) {
                    double y = 0.0;
                    for (int i = 0; i < M; ++i) { double v = Y(i, ih[t]); y += v * v; }
                    double se = hp[t] * hp[t] / (y + 1.0);
                    if (t == 0 || se > best) { best = se; k0 = t; }
                }
            }
            int k = iH[k0];
            vec p = matvec(invB, vec(&A.a[(size_t)k * M], &A.a[(size_t)k * M] + M));      // p = invB * A[:, k]
            l = ratio(k, p.data(), 0, gl, Sl);
            if (l == 0) { for (int j = 0; j < M; ++j) x[B[j]] = q[j]; return 3; }
            if (l > 0 && Edge && std::fabs(gl) < tol) { Edge = false; continue; }           // zero step: Bland's choice once
        } else {
            vec g(nH); ivec ig(nH); std::vector<int32_t> vS(nH);
            for (int t = 0; t < nH; ++t) {
                int lt = ratio(iH[t], &Y.a[(size_t)ih[t] * M], 1, g[t], vS[t]);
                if (lt == 0) { for (int j = 0; j < M; ++j) x[B[j]] = q[j]; return 3; }
                ig[t] = lt;
            }
            double best = -1.0;
            for (int t = 0; t < nH; ++t) { double v = std::fabs(hp[t] * g[t]); if (t == 0 || v > best) { best = v; k0 = t; } }
            l = ig[k0]; Sl = vS[k0]; gl = g[k0];
        }
        int k = iH[k0];
        if (l == -1) { S[k] = UP; x[k] = u[k]; if (st) st->flips += 1; }
        else if (l == -2) { S[k] = DN; x[k] = d[k]; if (st) st->flips += 1; }
        else {
            if (rule == 1) Edge = true;
            int mrow = l - 1;
            int lv = B[mrow];
            F[k] = 0; F[lv] = 1;
            B[mrow] = k;
            std::sort(B.begin(), B.end());
            Mat AB(M, M);
            for (int j = 0; j < M; ++j) std::memcpy(&AB.a[(size_t)j * M], &A.a[(size_t)B[j] * M], sizeof(double) * M);
            invB = inv_lu(AB);
            S[k] = IN; S[lv] = Sl;
            x[lv] = (Sl == DN) ? d[lv] : u[lv];
            build_iF(); compute_Y();
            if (st) st->pivots += 1;
        }
        {
            vec xF(iF.size());
            for (size_t t = 0; t < iF.size(); ++t) xF[t] = x[iF[t]];
            vec ib = matvec(invB, b);
            vec yx = matvec(Y, xF);
            for (int j = 0; j < M; ++j) q[j] = ib[j] - yx[j];
        }
        compute_h();
        nH = (int)iH.size();
    }
    for (int j = 0; j < M; ++j) x[B[j]] = q[j];
    bool ms = false;
    for (double v : h) if (std::fabs(v) < tol) { ms = true; break; }
    return ms ? 2 : 1;
}

// the pivot rule initQP / SimplexLP dispatch on (settingsLP.rule / settings.rule; 0 :Dantzig, 1 :stpEdgeLP, 2 :maxImprovement)
static int g_rule = 0;
int solve_lp_rule(const vec& c, const Mat& A, const vec& b, const vec& d, const vec& u, ivec& B, std::vector<int32_t>& S,
                  Mat& invB, vec q, double tol, vec& x, LPStats* st) {
    if (g_rule == 0) return c_dantzig_lp(c, A, b, d, u, B, S, invB, q, tol, x, st);
    return c_alt_rule_lp(g_rule, c, A, b, d, u, B, S, invB, q, tol, x, st);
}

// ---------------------------------------------------------------------------------------
struct QPView {
    int N, M, J;
    const double *V, *A, *G, *q, *b, *g, *d, *u;   // column-major: V NxN, A MxN, G JxN
};

// initQP (src/SSQP.jl:461-560) ; rule fixed to Dantzig (default; src/types.jl:405)
// g_fix_flip = 0 (default): literal restatement, including the reference's no-op status flip of the (-Inf,u] variables
// (`S[k] == UP` is a comparison, and k runs over 1:m instead of id, :552-557), which leaves such a variable DN at x = u.
// g_fix_flip = 1: what that loop was written for — a negated variable at its transformed lower bound becomes UP.  The
// device path implements the fixed form; tests that use (-Inf,u] variables switch it on.
static int g_fix_flip = 0;
int init_qp(const QPView& Q, double tol, vec& x, std::vector<int32_t>& S, LPStats* st) {
    int N = Q.N, M = Q.M, J = Q.J;
    ivec iv, id;
    for (int k = 0; k < N; ++k) {
        bool fu = Q.u[k] == INF, fd = Q.d[k] == -INF;
        if (fu && fd) iv.push_back(k);
        else if (fd) id.push_back(k);
    }
    int n = (int)iv.size();
    int M0 = M + J, N0 = N + J + n;
    Mat A0(M0, N0);
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < M; ++i) A0(i, k) = Q.A[i + (size_t)k * M];
        for (int i = 0; i < J; ++i) A0(M + i, k) = Q.G[i + (size_t)k * J];
    }
    for (int i = 0; i < J; ++i) A0(M + i, N + i) = 1.0;
    for (int t = 0; t < n; ++t)
        for (int i = 0; i < M0; ++i) A0(i, N + J + t) = -A0(i, iv[t]);
    vec b0(M0), d0(N0, 0.0), u0(N0, INF);
    for (int i = 0; i < M; ++i) b0[i] = Q.b[i];
    for (int i = 0; i < J; ++i) b0[M + i] = Q.g[i];
    for (int k = 0; k < N; ++k) { d0[k] = Q.d[k]; u0[k] = Q.u[k]; }
    for (int k : iv) d0[k] = 0.0;
    for (int k : id) {
        d0[k] = -u0[k];
        u0[k] = INF;
        for (int i = 0; i < M0; ++i) A0(i, k) = -A0(i, k);
    }
    int N1 = M0 + N0;
    std::vector<int32_t> S1(N1, DN);
    ivec B(M0);
    for (int j = 0; j < M0; ++j) { B[j] = N0 + j; S1[B[j]] = IN; }
    Mat invB(M0, M0);
    vec qv = matvec(A0, d0);
    for (int j = 0; j < M0; ++j) invB(j, j) = (b0[j] >= qv[j]) ? 1.0 : -1.0;
    for (int j = 0; j < M0; ++j) qv[j] = std::fabs(qv[j] - b0[j]);
    vec c1(N1, 0.0);
    for (int j = 0; j < M0; ++j) c1[N0 + j] = 1.0;
    Mat A1(M0, N1);
    std::memcpy(A1.a.data(), A0.a.data(), sizeof(double) * (size_t)M0 * N0);
    for (int j = 0; j < M0; ++j) A1(j, N0 + j) = invB(j, j);
    vec d1(N1, 0.0), u1(N1, INF);
    for (int k = 0; k < N0; ++k) { d1[k] = d0[k]; u1[k] = u0[k]; }

    vec x0;
    solve_lp_rule(c1, A1, b0, d1, u1, B, S1, invB, qv, tol, x0, st);

    x.assign(x0.begin(), x0.begin() + N);
    S.assign(S1.begin(), S1.begin() + N + J);
    double f = 0.0;
    for (int k = N0; k < N1; ++k) f += x0[k];
    if (f > tol) return 0;
    for (int k = N; k < N + J; ++k) S[k] = (S[k] == IN) ? OE : EO;
    if (n > 0) {
        for (int t = 0; t < n; ++t) { x[iv[t]] -= x0[N + J + t]; S[iv[t]] = IN; }
    }
    if (!id.empty()) {
        for (int k : id) x[k] = -x[k];
        // src/SSQP.jl:552-557: the status flip loop is a no-op comparison in the reference
        if (g_fix_flip)
            for (int k : id) if (S[k] == DN) S[k] = UP;
    }
    return 1;
}

// SimplexLP(P::LP) (src/Simplex.jl:831-1034), rule = Dantzig, min = true, Phase1 = false, including the free-variable
// split and the (-Inf,u] negation (:861-887) with their epilogue (:996-1032: the halves are recombined, basic 2nd halves
// move to the variable itself, the status 1/2 is recomputed from the reduced costs of the first N+J columns — also after
// an unbounded Phase 2, whose 3 is overwritten, as in the reference).  `rank(A0)` (Julia: SVD) is replaced by the number
// of pivots a full-pivoting Gauss-Jordan finds above 1e-10*max|A0|.  Returns the reference's status 1 / 2 / 3 / 0 / -1;
// x (N), S (N+J).
int simplex_lp(int N, int M, int J, const double* c, const double* A, const double* G, const double* b, const double* g,
               const double* d, const double* u, double tol, vec& x, std::vector<int32_t>& S, LPStats* st) {
    ivec iv, id;                                                            // :861-867
    for (int k = 0; k < N; ++k) {
        bool fu = u[k] == INF, fd = d[k] == -INF;
        if (fu && fd) iv.push_back(k);
        else if (fd) id.push_back(k);
    }
    int n = (int)iv.size();
    int nj = N + J, M0 = M + J, N0 = nj + n;
    Mat A0(M0, N0);
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < M; ++i) A0(i, k) = A[i + (size_t)k * M];
        for (int i = 0; i < J; ++i) A0(M + i, k) = G[i + (size_t)k * J];
    }
    for (int i = 0; i < J; ++i) A0(M + i, N + i) = 1.0;
    for (int t = 0; t < n; ++t)
        for (int i = 0; i < M0; ++i) A0(i, nj + t) = -A0(i, iv[t]);
    vec b0(M0), d0(N0, 0.0), u0(N0, INF);
    for (int i = 0; i < M; ++i) b0[i] = b[i];
    for (int i = 0; i < J; ++i) b0[M + i] = g[i];
    for (int k = 0; k < N; ++k) { d0[k] = d[k]; u0[k] = u[k]; }
    for (int k : iv) d0[k] = 0.0;                                           // :879-887
    for (int k : id) {
        d0[k] = -u0[k];
        u0[k] = INF;
        for (int i = 0; i < M0; ++i) A0(i, k) = -A0(i, k);
    }
    x.assign(N, 0.0); S.assign(nj, DN);
    // purge redundancy (:889-902)
    {
        Mat T = A0;
        double amax = 0.0;
        for (double v : T.a) amax = std::max(amax, std::fabs(v));
        int m0 = 0;
        std::vector<char> ru(M0, 0), cu(N0, 0);
        for (int step = 0; step < M0; ++step) {
            double best = 1e-10 * amax; int bi = -1, bj = -1;
            for (int j = 0; j < N0; ++j) if (!cu[j]) for (int i = 0; i < M0; ++i) if (!ru[i] && std::fabs(T(i, j)) > best) { best = std::fabs(T(i, j)); bi = i; bj = j; }
            if (bi < 0) break;
            ru[bi] = 1; cu[bj] = 1; m0++;
            for (int i = 0; i < M0; ++i) if (!ru[i]) {
                double f = T(i, bj) / T(bi, bj);
                if (f != 0.0) for (int j = 0; j < N0; ++j) if (!cu[j]) T(i, j) -= f * T(bi, j);
            }
        }
        if (m0 < M0) {
            Mat X(M0, N0 + 1);
            std::memcpy(X.a.data(), A0.a.data(), sizeof(double) * (size_t)M0 * N0);
            for (int i = 0; i < M0; ++i) X(i, N0) = b0[i];
            ivec ra; int la;
            get_rows_gjr(X, tol, ra, la);
            if ((int)ra.size() != la) return 0;
            if (m0 != la) return -1;
            Mat A2(m0, N0); vec b2(m0);
            for (int r = 0; r < m0; ++r) { for (int j = 0; j < N0; ++j) A2(r, j) = A0(ra[r], j); b2[r] = b0[ra[r]]; }
            A0 = A2; b0 = b2; M0 = m0;
        }
    }
    int N1 = M0 + N0;
    std::vector<int32_t> S1(N1, DN);
    ivec B(M0);
    for (int j = 0; j < M0; ++j) { B[j] = N0 + j; S1[B[j]] = IN; }
    Mat invB(M0, M0);
    vec qv = matvec(A0, d0);
    for (int j = 0; j < M0; ++j) invB(j, j) = (b0[j] >= qv[j]) ? 1.0 : -1.0;
    for (int j = 0; j < M0; ++j) qv[j] = std::fabs(qv[j] - b0[j]);
    vec c1(N1, 0.0);
    for (int j = 0; j < M0; ++j) c1[N0 + j] = 1.0;
    Mat A1(M0, N1);
    std::memcpy(A1.a.data(), A0.a.data(), sizeof(double) * (size_t)M0 * N0);
    for (int j = 0; j < M0; ++j) A1(j, N0 + j) = invB(j, j);
    vec d1(N1, 0.0), u1(N1, INF);
    for (int k = 0; k < N0; ++k) { d1[k] = d0[k]; u1[k] = u0[k]; }
    vec x1;
    solve_lp_rule(c1, A1, b0, d1, u1, B, S1, invB, qv, tol, x1, st);         // Phase 1 (:921)
    double f = 0.0;
    for (int k = N0; k < N1; ++k) f += x1[k];
    if (std::fabs(f) > tol) {                                               // :923-927 (S returned as it stands)
        x.assign(x1.begin(), x1.begin() + N);
        S.assign(S1.begin(), S1.begin() + nj);
        return 0;
    }
    // Phase 2 (:955-987)
    vec q(M0);
    for (int j = 0; j < M0; ++j) q[j] = x1[B[j]];
    vec c0(N0, 0.0);
    for (int k = 0; k < N; ++k) c0[k] = c[k];
    for (int k : id) c0[k] = -c0[k];                                        // :958-959
    for (int t = 0; t < n; ++t) c0[nj + t] = -c0[iv[t]];
    ivec iB;
    for (int j = 0; j < M0; ++j) if (B[j] < N0) iB.push_back(B[j]);
    if ((int)iB.size() < M0) {                                              // artificial variables in the basis: drive them out
        std::vector<char> F(N0, 1);
        for (int k : iB) F[k] = 0;
        ivec ic = iB;
        for (int k = 0; k < N0; ++k) if (F[k]) ic.push_back(k);
        Mat Xt((int)ic.size(), M0);                                         // A0[:, ic]'
        for (size_t r = 0; r < ic.size(); ++r) for (int i = 0; i < M0; ++i) Xt((int)r, i) = A0(i, ic[r]);
        ivec ra; int la;
        get_rows_gjr(Xt, tol, ra, la);
        ivec Bn;
        for (int r : ra) Bn.push_back(ic[r]);
        std::sort(Bn.begin(), Bn.end());
        if ((int)Bn.size() != M0) return -1;                                // inv(lu(A0[:,B])) would throw on a non-square basis
        for (int k : Bn) if (std::find(iB.begin(), iB.end(), k) == iB.end()) S1[k] = IN;
        B = Bn;
        Mat AB(M0, M0);
        for (int j = 0; j < M0; ++j) for (int i = 0; i < M0; ++i) AB(i, j) = A0(i, B[j]);
        try { invB = inv_lu(AB); } catch (NumErr&) { return -1; }
        for (int j = 0; j < M0; ++j) q[j] = x1[B[j]];
    }
    std::vector<int32_t> S0(S1.begin(), S1.begin() + N0);
    vec x0;
    int iH;
    try { iH = solve_lp_rule(c0, A0, b0, d0, u0, B, S0, invB, q, tol, x0, st); } catch (NumErr&) { return -1; }
    x.assign(x0.begin(), x0.begin() + N);
    S.assign(S0.begin(), S0.begin() + nj);
    for (int k = N; k < nj; ++k) S[k] = (S[k] == IN) ? OE : EO;
    if (n > 0) {                                                            // free variables (:996-1021)
        for (int t = 0; t < n; ++t) x[iv[t]] -= x0[nj + t];
        for (int j = 0; j < M0; ++j) {
            int t = B[j];
            if (t >= nj) { B[j] = iv[t - nj]; S[B[j]] = IN; }               // move IN to part 1
        }
        std::vector<char> F(nj, 1);
        for (int j = 0; j < M0; ++j) F[B[j]] = 0;
        Mat AB(M0, M0);
        for (int j = 0; j < M0; ++j) for (int i = 0; i < M0; ++i) AB(i, j) = A0(i, B[j]);
        Mat iBm;
        try { iBm = inv_lu(AB); } catch (NumErr&) { return -1; }
        vec cB(M0);
        for (int j = 0; j < M0; ++j) cB[j] = c0[B[j]];
        vec pi = matvec_t(iBm, cB);                                         // Y' c0[B] = A0[:,F]' (invB' c0[B])
        bool ms = false;
        for (int k = 0; k < nj; ++k) if (F[k]) {
            double hk = c0[k];
            for (int i = 0; i < M0; ++i) hk -= A0(i, k) * pi[i];
            if (std::fabs(hk) < tol) ms = true;
        }
        iH = ms ? 2 : 1;
    }
    if (!id.empty()) {                                                      // flip back (:1023-1032)
        for (int k : id) { x[k] = -x[k]; if (S[k] == DN) S[k] = UP; }
    }
    return iH;
}

// polishSz! (src/SSQP.jl:10-32)
void polish_sz(std::vector<int32_t>& S, vec& z, const QPView& Q, double tol) {
    int N = Q.N, J = Q.J;
    for (int k = 0; k < N; ++k) {
        if (S[k] == DN) z[k] = Q.d[k];
        else if (S[k] == UP) z[k] = Q.u[k];
        else {
            if (std::fabs(z[k] - Q.d[k]) < tol) { z[k] = Q.d[k]; S[k] = DN; }
            else if (std::fabs(z[k] - Q.u[k]) < tol) { z[k] = Q.u[k]; S[k] = UP; }
        }
    }
    for (int j = 0; j < J; ++j) {
        double s = 0.0;
        for (int k = 0; k < N; ++k) s += z[k] * Q.G[j + (size_t)k * J];
        S[N + j] = (std::fabs(Q.g[j] - s) < tol) ? EO : OE;
    }
}

// freeK! (src/SSQP.jl:35-59)
int free_k(std::vector<int32_t>& S, const vec& z, const QPView& Q, double tol) {
    int N = Q.N;
    vec p(N);
    for (int i = 0; i < N; ++i) p[i] = Q.q[i];
    for (int k = 0; k < N; ++k) {
        double zk = z[k];
        if (zk == 0.0) continue;
        const double* vc = Q.V + (size_t)k * N;
        for (int i = 0; i < N; ++i) p[i] += vc[i] * zk;
    }
    std::vector<int32_t> S0(S.begin(), S.end());
    bool t = true;
    for (int k = 0; k < N; ++k) {
        if ((p[k] >= -tol && S[k] == UP) || (p[k] <= tol && S[k] == DN)) { S[k] = IN; t = false; }
    }
    if (t) return 1;
    double nrm = 0.0;
    int cnt = 0;
    for (int k = 0; k < N; ++k)
        if (S[k] == IN) { cnt++; nrm = std::max(nrm, std::fabs(p[k])); }
    if (cnt > 0 && nrm <= tol) {
        for (int k = 0; k < N; ++k) if (S[k] == IN) S[k] = S0[k];
        return 1;
    }
    return -1;
}

struct Event { int32_t From, To; int id; double L; };   // src/types.jl:39-44 (id 0-based here)

// Julia isless on Float64: -0.0 < 0.0, NaN last
inline bool jl_isless(double a, double b) {
    if (std::isnan(a)) return false;
    if (std::isnan(b)) return true;
    if (a < b) return true;
    if (a == 0.0 && b == 0.0) return std::signbit(a) && !std::signbit(b);
    return false;
}

struct Trace {          // optional per-trip log
    int32_t* buf = nullptr;     // rows of 4: K, W, kind(0 freeK,1 step,2 release,3 optimal), count
    int64_t cap = 0, n = 0;
    void add(int K, int W, int kind, int cnt) {
        if (buf && n < cap) { buf[4 * n] = K; buf[4 * n + 1] = W; buf[4 * n + 2] = kind; buf[4 * n + 3] = cnt; }
        n++;
    }
};

// aStep! (src/SSQP.jl:61-134)
int a_step(const vec& p, vec& z, std::vector<int32_t>& S, const ivec& iF, const ivec& iOg,
           const vec& alpha, const QPView& Q, double tol, int* nblocked) {
    int N = Q.N, J = Q.J;
    std::vector<Event> Lo;
    for (size_t k = 0; k < alpha.size(); ++k) {
        int j = iF[k];
        double t = p[k], h = z[j];
        double dL = (Q.d[j] - h) / t, uL = (Q.u[j] - h) / t;
        if (t > tol && Q.u[j] < INF) Lo.push_back({IN, UP, j, uL});
        else if (t < -tol && Q.d[j] > -INF) Lo.push_back({IN, DN, j, dL});
    }
    if (J > 0) {
        for (size_t k = 0; k < iOg.size(); ++k) {
            int j = iOg[k];
            double gz = 0.0;                                  // G[Og,:]*z
            for (int i = 0; i < N; ++i) gz += Q.G[j + (size_t)i * J] * z[i];
            double zo = Q.g[j] - gz;
            double po = 0.0;                                  // G[Og,F]*p
            for (size_t t = 0; t < iF.size(); ++t) po += Q.G[j + (size_t)iF[t] * J] * p[t];
            if (po > tol) Lo.push_back({OE, EO, j, zo / po});
        }
    }
    double L1 = 1.0;
    if (!Lo.empty()) {
        std::stable_sort(Lo.begin(), Lo.end(), [](const Event& a, const Event& b) { return jl_isless(a.L, b.L); });
        L1 = Lo[0].L;
    }
    *nblocked = 0;
    if (L1 < 1.0) {
        for (size_t k = 0; k < iF.size(); ++k) z[iF[k]] += L1 * p[k];
        for (size_t i = 0; i < Lo.size(); ++i) {
            const Event& Lt = Lo[i];
            if (Lt.L - L1 > tol) break;
            int k = Lt.id;
            if (Lt.To == EO) k += N;
            S[k] = Lt.To;
            if (k < N) z[k] = (Lt.To == DN) ? Q.d[k] : Q.u[k];
            (*nblocked)++;
        }
        return -1;
    } else {
        for (size_t k = 0; k < iF.size(); ++k) z[iF[k]] = alpha[k];
        return 1;
    }
}

// KKTchk! (src/SSQP.jl:136-188)
int kkt_chk(std::vector<int32_t>& S, const ivec& iF, const ivec& iB, const ivec& iEg, const vec& gamma,
            const vec& alphaL, const Mat& AE, const QPView& Q, const ivec& idAE, const ivec& ra,
            double tolG) {
    int N = Q.N, M = Q.M, J = Q.J;
    std::vector<Event> Li;
    for (size_t k = 0; k < gamma.size(); ++k) {
        int j = iB[k];
        double t = gamma[k];
        if (S[j] == UP && t > tolG) Li.push_back({UP, IN, j, -t});
        else if (S[j] == DN && t < -tolG) Li.push_back({DN, IN, j, t});
    }
    int JE = (int)iEg.size();
    if (JE > 0) {
        ivec iE(JE, -1);
        for (size_t r = 0; r < ra.size(); ++r)
            if (ra[r] >= M) iE[idAE[ra[r]]] = (int)r;      // ra .> M  (1-based) == ra >= M (0-based)
        for (int j = 0; j < JE; ++j) {
            int k = iE[j];
            double t;
            if (k < 0) {
                // x = AE' \ GE[j,F] ; Lda = alphaL' * x       (src/SSQP.jl:158-159)
                Mat AEt(AE.c, AE.r);
                for (int a = 0; a < AE.r; ++a)
                    for (int c = 0; c < AE.c; ++c) AEt(c, a) = AE(a, c);
                vec rhs(iF.size());
                for (size_t c = 0; c < iF.size(); ++c) rhs[c] = Q.G[iEg[j] + (size_t)iF[c] * J];
                vec xs = lstsq(AEt, rhs);
                t = 0.0;
                for (size_t a = 0; a < xs.size(); ++a) t += alphaL[a] * xs[a];
            } else {
                t = alphaL[k];
            }
            if (t < -tolG) Li.push_back({EO, OE, iEg[j], t});
        }
    }
    if (!Li.empty()) {
        size_t best = 0;                 // stable sort, take the first == first minimal key
        for (size_t i = 1; i < Li.size(); ++i)
            if (jl_isless(Li[i].L, Li[best].L)) best = i;
        int k = Li[best].id;
        if (Li[best].To == OE) k += N;
        S[k] = Li[best].To;
        return -1;
    }
    return 1;
}

struct QPStats {
    int64_t trips = 0;
    double falg = 0.0;      // SURVEY 8d F_alg
    double fref = 0.0;      // reference-form flops
    int32_t maxK = 0, maxW = 0;
};

// solveQP(Q, S, x0) main loop (src/SSQP.jl:237-377)
int64_t solve_phase2(const QPView& Q, std::vector<int32_t>& S, vec& z, int maxIter, double tol, double tolG,
                     Trace* tr, QPStats* qs) {
    int N = Q.N, M = Q.M, J = Q.J;
    int64_t iter = 0;
    while (true) {
        iter += 1;
        if (iter > maxIter) return -iter;

        ivec iF, iB;
        for (int k = 0; k < N; ++k) (S[k] == IN ? iF : iB).push_back(k);
        int K = (int)iF.size();
        if (K == 0) {
            int st = free_k(S, z, Q, tol);
            if (tr) tr->add(0, 0, 0, 0);
            if (qs) { qs->trips++; qs->falg += 2.0 * N * N; qs->fref += 2.0 * N * N; }
            if (st > 0) return iter;
            continue;
        }
        ivec iEg, iOg;
        for (int j = 0; j < J; ++j) {
            if (S[N + j] == EO) iEg.push_back(j);
            else if (S[N + j] == OE) iOg.push_back(j);
        }
        int JE = (int)iEg.size();
        int W0 = M + JE;
        int NB = N - K;
        Mat AE(W0, K), AB(W0, NB);
        for (int c = 0; c < K; ++c) {
            int k = iF[c];
            for (int i = 0; i < M; ++i) AE(i, c) = Q.A[i + (size_t)k * M];
            for (int i = 0; i < JE; ++i) AE(M + i, c) = Q.G[iEg[i] + (size_t)k * J];
        }
        for (int c = 0; c < NB; ++c) {
            int k = iB[c];
            for (int i = 0; i < M; ++i) AB(i, c) = Q.A[i + (size_t)k * M];
            for (int i = 0; i < JE; ++i) AB(M + i, c) = Q.G[iEg[i] + (size_t)k * J];
        }
        ivec idAE(W0);
        for (int i = 0; i < M; ++i) idAE[i] = i;
        for (int i = 0; i < JE; ++i) idAE[M + i] = i;
        vec zB(NB);
        for (int c = 0; c < NB; ++c) zB[c] = z[iB[c]];
        vec bE(W0);
        {
            vec t = matvec(AB, zB);
            for (int i = 0; i < M; ++i) bE[i] = Q.b[i] - t[i];
            for (int i = 0; i < JE; ++i) bE[M + i] = Q.g[iEg[i]] - t[M + i];
        }
        // ra, la = getRowsGJr([AE bE], tol)
        ivec ra;
        int la;
        {
            Mat X(W0, K + 1);
            std::memcpy(X.a.data(), AE.a.data(), sizeof(double) * (size_t)W0 * K);
            for (int i = 0; i < W0; ++i) X(i, K) = bE[i];
            get_rows_gjr(X, tol, ra, la);
        }
        int W = (int)ra.size();
        if (W < W0) {
            if (W != la) return -1;      // unreachable by construction (l1 == length(rows))
            Mat AE2(W, K), AB2(W, NB);
            vec bE2(W);
            for (int r = 0; r < W; ++r) {
                for (int c = 0; c < K; ++c) AE2(r, c) = AE(ra[r], c);
                for (int c = 0; c < NB; ++c) AB2(r, c) = AB(ra[r], c);
                bE2[r] = bE[ra[r]];
            }
            AE = AE2; AB = AB2; bE = bE2;
        }
        if (qs) {
            qs->trips++;
            qs->maxK = std::max(qs->maxK, K);
            qs->maxW = std::max(qs->maxW, W);
            double k = K, w = W, n = N, jo = (double)iOg.size();
            qs->falg += k * k * k / 3 + k * k * w + k * w * w + w * w * w / 3 + 2 * k * k + 4 * k * w + 2 * w * w +
                        2 * n * n + 2 * (n - k) * w + 2 * jo * (n + k);
            qs->fref += k * k * k + 4 * k * k * w + 4 * k * w * w + w * w * w + 2.0 * W0 * W0 * (k + 1) + 2 * k * k +
                        2 * n * n + 2 * (n - k) * w + 2 * jo * (n + k);
        }

        Mat VFF(K, K);
        for (int c = 0; c < K; ++c)
            for (int r = 0; r < K; ++r) VFF(r, c) = Q.V[iF[r] + (size_t)iF[c] * N];
        Mat iV, C;
        try {
            iV = inv_cholesky(VFF);
        } catch (NumErr&) { return -1; }
        Mat VBF(NB, K);
        for (int c = 0; c < K; ++c)
            for (int r = 0; r < NB; ++r) VBF(r, c) = Q.V[iB[r] + (size_t)iF[c] * N];
        vec cvec = matvec_t(VBF, zB);
        for (int c = 0; c < K; ++c) cvec[c] += Q.q[iF[c]];
        Mat mT = matmul_nt(iV, AE);           // K x W
        C = matmul(AE, mT);                   // W x W
        for (int j = 0; j < W; ++j)
            for (int i = 0; i < j; ++i) {
                double s = (C(i, j) + C(j, i)) / 2;
                C(i, j) = s; C(j, i) = s;
            }
        try {
            C = inv_cholesky(C);
        } catch (NumErr&) { return -1; }
        Mat TC = matmul(mT, C);               // K x W
        Mat VQ = matmul_nt(mT, TC);           // K x K : mT * TC'
        for (size_t t = 0; t < VQ.a.size(); ++t) VQ.a[t] = iV.a[t] - VQ.a[t];
        vec alpha = matvec(TC, bE);
        {
            vec t = matvec(VQ, cvec);
            for (int c = 0; c < K; ++c) alpha[c] -= t[c];
        }
        vec p(K);
        double pn = 0.0;
        for (int c = 0; c < K; ++c) { p[c] = alpha[c] - z[iF[c]]; pn = std::max(pn, std::fabs(p[c])); }

        if (pn > tolG) {
            int nb = 0;
            int st = a_step(p, z, S, iF, iOg, alpha, Q, tol, &nb);
            if (st < 0) {
                if (tr) tr->add(K, W, 1, nb);
                continue;
            }
        }
        // alphaL = -(TC'*c + C*bE)
        vec alphaL = matvec_t(TC, cvec);
        {
            vec t = matvec(C, bE);
            for (int i = 0; i < W; ++i) alphaL[i] = -(alphaL[i] + t[i]);
        }
        // gamma = VBF*alpha + V[B,B]*zB + q[B] + AB'*alphaL
        vec gamma = matvec(VBF, alpha);
        if (la_on()) {       // the reference's own evaluation order: ((VBF*alpha + V[B,B]*zB) + q[B]) + AB'*alphaL, each product a dgemv
            Mat VBB(NB, NB);
            for (int c = 0; c < NB; ++c) {
                const double* vc = Q.V + (size_t)iB[c] * N;
                for (int r = 0; r < NB; ++r) VBB(r, c) = vc[iB[r]];
            }
            vec t2 = matvec(VBB, zB);
            vec t = matvec_t(AB, alphaL);
            for (int r = 0; r < NB; ++r) gamma[r] = ((gamma[r] + t2[r]) + Q.q[iB[r]]) + t[r];
        } else {
            for (int c = 0; c < NB; ++c) {
                double zc = zB[c];
                if (zc == 0.0) continue;
                const double* vc = Q.V + (size_t)iB[c] * N;
                for (int r = 0; r < NB; ++r) gamma[r] += vc[iB[r]] * zc;
            }
            vec t = matvec_t(AB, alphaL);
            for (int r = 0; r < NB; ++r) gamma[r] += Q.q[iB[r]] + t[r];
        }
        int st = kkt_chk(S, iF, iB, iEg, gamma, alphaL, AE, Q, idAE, ra, tolG);
        if (st > 0) {
            polish_sz(S, z, Q, tol);
            if (tr) tr->add(K, W, 3, 0);
            return iter;
        }
        if (tr) tr->add(K, W, 2, 1);
    }
}

}  // namespace

extern "C" {

struct ssqp_oracle_settings {
    int32_t max_iter;
    double tol;
    double tolG;
};

// One QP.  Returns the reference's `status` (iter>0 | 0 infeasible | -1 numerical | -(maxIter+1)).
// mc is the QP constructor's validity code (src/types.jl:240-284); mc<=0 -> early return
// (src/SSQP.jl:226-228): x=0, S[0:N]=DN, status=-1 (S has length N there; S[N:N+J] is left untouched).
// stats (nullable, 8 doubles): trips, F_alg, F_ref, maxK, maxW, lp_loops, lp_pivots, lp_flips
int64_t ssqp_oracle_solve(int32_t N, int32_t M, int32_t J, const double* V, const double* A, const double* G,
                          const double* q, const double* b, const double* g, const double* d, const double* u,
                          int32_t mc, const ssqp_oracle_settings* set, const ssqp_oracle_settings* setLP,
                          const int32_t* S0, const double* x0,   // nullable warm start (src/SSQP.jl:237)
                          double* x, int32_t* S, int32_t* trace, int64_t trace_cap, int64_t* trace_n,
                          double* stats) {
    QPView Q{N, M, J, V, A, G, q, b, g, d, u};
    if (trace_n) *trace_n = 0;
    if (stats) for (int i = 0; i < 8; ++i) stats[i] = 0.0;
    if (mc <= 0) {
        for (int k = 0; k < N; ++k) { x[k] = 0.0; S[k] = DN; }
        return -1;
    }
    vec z;
    std::vector<int32_t> Sv;
    LPStats lps;
    if (S0 && x0) {
        z.assign(x0, x0 + N);
        Sv.assign(S0, S0 + N + J);
    } else {
        int st;
        try {
            st = init_qp(Q, setLP->tol, z, Sv, &lps);
        } catch (NumErr&) { st = -1; }
        if (st <= 0) {
            for (int k = 0; k < N; ++k) x[k] = z.size() == (size_t)N ? z[k] : 0.0;
            for (int k = 0; k < N + J; ++k) S[k] = Sv.size() == (size_t)(N + J) ? Sv[k] : DN;
            if (stats) { stats[5] = (double)lps.loops; stats[6] = (double)lps.pivots; stats[7] = (double)lps.flips; }
            return st;
        }
    }
    Trace tr;
    tr.buf = trace; tr.cap = trace_cap;
    QPStats qs;
    int64_t status = solve_phase2(Q, Sv, z, set->max_iter, set->tol, set->tolG, &tr, &qs);
    for (int k = 0; k < N; ++k) x[k] = z[k];
    for (int k = 0; k < N + J; ++k) S[k] = Sv[k];
    if (trace_n) *trace_n = tr.n;
    if (stats) {
        stats[0] = (double)qs.trips; stats[1] = qs.falg; stats[2] = qs.fref; stats[3] = qs.maxK; stats[4] = qs.maxW;
        stats[5] = (double)lps.loops; stats[6] = (double)lps.pivots; stats[7] = (double)lps.flips;
        // phase-1 F_alg / F_ref
        double M0 = M + J, N1 = N + J + M0;
        stats[1] += 4 * M0 * N1 * lps.loops + 2 * M0 * M0 * lps.pivots;
        stats[2] += 4 * M0 * N1 * lps.loops + (2 * M0 * M0 * M0 + 2 * M0 * M0 * (N1 - M0)) * lps.pivots;
    }
    return status;
}

// Phase-1 only (initQP): returns 1 feasible / 0 infeasible / -1 numerical; x (N), S (N+J)
int64_t ssqp_oracle_init(int32_t N, int32_t M, int32_t J, const double* A, const double* G, const double* b,
                         const double* g, const double* d, const double* u, double tol, double* x, int32_t* S,
                         double* stats) {
    QPView Q{N, M, J, nullptr, A, G, nullptr, b, g, d, u};
    vec z;
    std::vector<int32_t> Sv;
    LPStats lps;
    int st;
    try {
        st = init_qp(Q, tol, z, Sv, &lps);
    } catch (NumErr&) { return -1; }
    for (int k = 0; k < N; ++k) x[k] = z[k];
    for (int k = 0; k < N + J; ++k) S[k] = Sv[k];
    if (stats) { stats[0] = (double)lps.loops; stats[1] = (double)lps.pivots; stats[2] = (double)lps.flips; }
    return st;
}

// Batch over nb QPs, OpenMP over the batch (one QP per thread, no nested parallelism) — the
// "Julia Threads.@threads loop over solveQP" stand-in used as the CPU baseline.
// Strides (in doubles) of 0 mean "shared by every QP".
int32_t ssqp_oracle_solve_batch(int32_t N, int32_t M, int32_t J, int64_t nb, const double* V, int64_t sV,
                                const double* A, const double* G, const double* q, int64_t sq, const double* b,
                                int64_t sb, const double* g, int64_t sg, const double* d, int64_t sd,
                                const double* u, int64_t su, const ssqp_oracle_settings* set,
                                const ssqp_oracle_settings* setLP, double* x, int32_t* S, int64_t* status,
                                double* stats /* nullable, 8*nb */, int32_t nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < nb; ++i) {
        status[i] = ssqp_oracle_solve(N, M, J, V + i * sV, A, G, q + i * sq, b + i * sb, g + i * sg, d + i * sd,
                                      u + i * su, 1, set, setLP, nullptr, nullptr, x + i * N, S + i * (N + J),
                                      nullptr, 0, nullptr, stats ? stats + 8 * i : nullptr);
    }
    return used;
}

// Phase 1 (initQP) of a batch, OpenMP over the batch: x0 (N*nb), S (N+J)*nb, status nb, stats 3*nb (loops, pivots, flips; nullable)
int32_t ssqp_oracle_init_batch(int32_t N, int32_t M, int32_t J, int64_t nb, const double* A, const double* G, const double* b,
                               int64_t sb, const double* g, int64_t sg, const double* d, int64_t sd, const double* u, int64_t su,
                               double tol, double* x, int32_t* S, int64_t* status, double* stats, int32_t nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < nb; ++i)
        status[i] = ssqp_oracle_init(N, M, J, A, G, b + i * sb, g + i * sg, d + i * sd, u + i * su, tol, x + i * N,
                                     S + i * (N + J), stats ? stats + 3 * i : nullptr);
    return used;
}

// lstsq (Julia's `\` on a rectangular matrix, src/SSQP.jl:158) exposed for unit tests: X is m x n column-major, y length m, x length n out
void ssqp_oracle_lstsq(int32_t m, int32_t n, const double* X, const double* y, double* x) {
    Mat Xm(m, n);
    std::memcpy(Xm.a.data(), X, sizeof(double) * (size_t)m * n);
    vec yv(y, y + m);
    vec r = lstsq(Xm, yv);
    for (int i = 0; i < n; ++i) x[i] = r[i];
}

// getRowsGJr exposed for unit tests: X is nr x nc column-major; rows (nr ints, 0-based) out; returns count; *l1 out
int32_t ssqp_oracle_get_rows_gjr(int32_t nr, int32_t nc, const double* X, double tol, int32_t* rows, int32_t* l1) {
    Mat Xm(nr, nc);
    std::memcpy(Xm.a.data(), X, sizeof(double) * (size_t)nr * nc);
    ivec r;
    int l;
    get_rows_gjr(Xm, tol, r, l);
    for (size_t i = 0; i < r.size(); ++i) rows[i] = r[i];
    *l1 = l;
    return (int32_t)r.size();
}

// cDantzigLP (src/Simplex.jl:445-615) exposed for unit tests: bounded revised simplex from a given basis.
// A is M x N column-major; B (M ints, sorted ascending, 0-based) and S (N Status codes) are in/out; invB (M x M
// column-major) and q (= x_B, M) describe the starting basis.  Returns status 1 (unique) / 2 (many) / 3 (unbounded),
// or -1 on a numerical error (Julia would throw SingularException).
int32_t ssqp_oracle_dantzig_lp(int32_t N, int32_t M, const double* c, const double* A, const double* b,
                               const double* d, const double* u, int32_t* B, int32_t* S, const double* invB,
                               const double* q, double tol, double* x) {
    vec cv(c, c + N), bv(b, b + M), dv(d, d + N), uv(u, u + N), qv(q, q + M), xv;
    Mat Am(M, N), iB(M, M);
    std::memcpy(Am.a.data(), A, sizeof(double) * (size_t)M * N);
    std::memcpy(iB.a.data(), invB, sizeof(double) * (size_t)M * M);
    ivec Bv(B, B + M);
    std::vector<int32_t> Sv(S, S + N);
    int st;
    try {
        st = c_dantzig_lp(cv, Am, bv, dv, uv, Bv, Sv, iB, qv, tol, xv, nullptr);
    } catch (NumErr&) { return -1; }
    for (int j = 0; j < M; ++j) B[j] = Bv[j];
    for (int k = 0; k < N; ++k) { S[k] = Sv[k]; x[k] = xv[k]; }
    return st;
}

// SimplexLP (src/Simplex.jl:831-1034) for one LP; returns status 1/2/3/0/-1 (or -99: free / (-Inf,u] variables, not restated)
int32_t ssqp_oracle_simplex_lp(int32_t N, int32_t M, int32_t J, const double* c, const double* A, const double* G,
                               const double* b, const double* g, const double* d, const double* u, double tol,
                               double* x, int32_t* S, double* stats /* nullable: loops, pivots, flips */) {
    vec xv; std::vector<int32_t> Sv; LPStats lps;
    int st;
    try { st = simplex_lp(N, M, J, c, A, G, b, g, d, u, tol, xv, Sv, &lps); } catch (NumErr&) { st = -1; }
    for (int k = 0; k < N && k < (int)xv.size(); ++k) x[k] = xv[k];
    for (int k = 0; k < N + J && k < (int)Sv.size(); ++k) S[k] = Sv[k];
    if (stats) { stats[0] = (double)lps.loops; stats[1] = (double)lps.pivots; stats[2] = (double)lps.flips; }
    return st;
}

// 1: the (-Inf,u] variables that end initQP at their bound become UP (what src/SSQP.jl:552-557 was written for); 0: literal
// LAPACK form: ptrs = {dpotrf, dpotri, dgetrf, dgetri, dgemm, dgemv} (Fortran ABI, 32-bit ints); NULL table -> scalar form
void ssqp_oracle_set_lapack(void** ptrs) {
    if (!ptrs) { g_la.have = 0; g_la.on = 0; return; }
    g_la.potrf = (potrf_fn)ptrs[0]; g_la.potri = (potrf_fn)ptrs[1]; g_la.getrf = (getrf_fn)ptrs[2];
    g_la.getri = (getri_fn)ptrs[3]; g_la.gemm = (gemm_fn)ptrs[4]; g_la.gemv = (gemv_fn)ptrs[5];
    g_la.have = 1; g_la.on = 1;
}
int32_t ssqp_oracle_use_lapack(int32_t on) { g_la.on = (on && g_la.have) ? 1 : 0; return g_la.on; }
int32_t ssqp_oracle_lapack_form() { return g_la.on; }

void ssqp_oracle_set_fix_flip(int32_t on) { g_fix_flip = on ? 1 : 0; }
// pivot rule of initQP / SimplexLP: 0 :Dantzig (default), 1 :stpEdgeLP, 2 :maxImprovement  (Settings.rule, src/types.jl:397)
void ssqp_oracle_set_rule(int32_t rule) { g_rule = (rule == 1 || rule == 2) ? rule : 0; }

int32_t ssqp_oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
