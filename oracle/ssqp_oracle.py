"""ctypes loader for the CPU oracle (oracle/ssqp_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

IN, DN, UP, OE, EO = 0, 1, 2, 3, 4


class Settings(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("tol", C.c_double), ("tolG", C.c_double)]


def default_settings(max_iter=7777, tol=2.0 ** -26, tolG=2.0 ** -33):
    return Settings(max_iter, tol, tolG)


def build(force=False):
    so = os.path.join(_HERE, "libssqp_oracle.so")
    src = os.path.join(_HERE, "ssqp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libssqp_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        lp = C.POINTER(C.c_int64)
        sp = C.POINTER(Settings)
        _LIB.ssqp_oracle_solve.restype = C.c_int64
        _LIB.ssqp_oracle_solve.argtypes = [C.c_int32] * 3 + [dp] * 8 + [C.c_int32, sp, sp, ip, dp, dp, ip, ip,
                                                                     C.c_int64, lp, dp]
        _LIB.ssqp_oracle_init.restype = C.c_int64
        _LIB.ssqp_oracle_init.argtypes = [C.c_int32] * 3 + [dp] * 6 + [C.c_double, dp, ip, dp]
        _LIB.ssqp_oracle_solve_batch.restype = C.c_int32
        _LIB.ssqp_oracle_solve_batch.argtypes = [C.c_int32] * 3 + [C.c_int64, dp, C.c_int64, dp, dp] + \
            [dp, C.c_int64] * 5 + [sp, sp, dp, ip, lp, dp, C.c_int32]
        _LIB.ssqp_oracle_get_rows_gjr.restype = C.c_int32
        _LIB.ssqp_oracle_get_rows_gjr.argtypes = [C.c_int32, C.c_int32, dp, C.c_double, ip, ip]
        _LIB.ssqp_oracle_max_threads.restype = C.c_int32
        _LIB.ssqp_oracle_simplex_lp.restype = C.c_int32
        _LIB.ssqp_oracle_simplex_lp.argtypes = [C.c_int32] * 3 + [dp] * 7 + [C.c_double, dp, ip, dp]
        _LIB.ssqp_oracle_dantzig_lp.restype = C.c_int32
        _LIB.ssqp_oracle_dantzig_lp.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, dp, dp, ip, ip, dp, dp, C.c_double, dp]
        _LIB.ssqp_oracle_init_batch.restype = C.c_int32
        _LIB.ssqp_oracle_init_batch.argtypes = [C.c_int32] * 3 + [C.c_int64, dp, dp] + [dp, C.c_int64] * 4 + \
            [C.c_double, dp, ip, lp, dp, C.c_int32]
        _LIB.ssqp_oracle_set_lapack.argtypes = [C.POINTER(C.c_void_p)]
        _LIB.ssqp_oracle_use_lapack.restype = C.c_int32
        _LIB.ssqp_oracle_use_lapack.argtypes = [C.c_int32]
        _LIB.ssqp_oracle_lapack_form.restype = C.c_int32
    return _LIB


_LAPACK_KEEP = []          # keeps scipy's modules (and so its OpenBLAS) alive


def _capsule_ptr(mod, name):
    cap = mod.__pyx_capi__[name]
    api = C.pythonapi
    api.PyCapsule_GetName.restype = C.c_char_p
    api.PyCapsule_GetName.argtypes = [C.py_object]
    api.PyCapsule_GetPointer.restype = C.c_void_p
    api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    return api.PyCapsule_GetPointer(cap, api.PyCapsule_GetName(cap))


def use_lapack(on=True):
    """LAPACK form of the oracle: the dense algebra runs on scipy's bundled OpenBLAS through the very routines Julia's
    LinearAlgebra calls on this path (dpotrf+dpotri, dgetrf+dgetri, dgemm, dgemv) with BLAS threads pinned to 1 (the
    batch is threaded over QPs).  on=False: the scalar loops.  Returns the form now in effect ("lapack" / "scalar")."""
    L = lib()
    if on and not _LAPACK_KEEP:
        from scipy.linalg import cython_lapack as cl, cython_blas as cb
        try:
            import threadpoolctl
            _LAPACK_KEEP.append(threadpoolctl.threadpool_limits(limits=1, user_api="blas"))
        except Exception:                                   # pragma: no cover
            os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        ptrs = (C.c_void_p * 6)(_capsule_ptr(cl, "dpotrf"), _capsule_ptr(cl, "dpotri"), _capsule_ptr(cl, "dgetrf"),
                                _capsule_ptr(cl, "dgetri"), _capsule_ptr(cb, "dgemm"), _capsule_ptr(cb, "dgemv"))
        L.ssqp_oracle_set_lapack(ptrs)
        _LAPACK_KEEP.extend([cl, cb])
    return "lapack" if L.ssqp_oracle_use_lapack(1 if on else 0) else "scalar"


def form():
    return "lapack" if lib().ssqp_oracle_lapack_form() else "scalar"


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def solve_qp(V, A, G, q, b, g, d, u, settings=None, settingsLP=None, mc=1, S0=None, x0=None, trace=False):
    """solveQP(Q) / solveQP(Q,S,x0)  (src/SSQP.jl:224-377).  Returns dict(x,S,status,stats[,trace])."""
    L = lib()
    V = _f(V); N = V.shape[0]
    A = _f(np.reshape(A, (-1, N))); G = _f(np.reshape(G, (-1, N)))
    M, J = A.shape[0], G.shape[0]
    q, b, g, d, u = (np.ascontiguousarray(t, dtype=np.float64).ravel() for t in (q, b, g, d, u))
    st = settings or default_settings()
    stl = settingsLP or st
    x = np.zeros(N); S = np.full(N + J, -1, dtype=np.int32)
    stats = np.zeros(8)
    cap = 4 * (st.max_iter + 2) if trace else 0
    tr = np.zeros((max(cap, 1), 4), dtype=np.int32)
    tn = C.c_int64(0)
    s0p = _ip(np.ascontiguousarray(S0, dtype=np.int32)) if S0 is not None else None
    x0p = _dp(np.ascontiguousarray(x0, dtype=np.float64)) if x0 is not None else None
    status = L.ssqp_oracle_solve(N, M, J, _dp(V), _dp(A), _dp(G), _dp(q), _dp(b), _dp(g), _dp(d), _dp(u), mc,
                                 C.byref(st), C.byref(stl), s0p, x0p, _dp(x), _ip(S),
                                 _ip(tr) if trace else None, cap, C.byref(tn), _dp(stats))
    out = dict(x=x, S=S, status=int(status), stats=stats)
    if trace:
        out["trace"] = tr[:tn.value].copy()
    return out


def init_qp(A, G, b, g, d, u, tol=2.0 ** -26):
    """initQP (src/SSQP.jl:461-560): returns (x0, S, status, stats[loops,pivots,flips])."""
    L = lib()
    d = np.ascontiguousarray(d, dtype=np.float64).ravel(); N = d.size
    A = _f(np.reshape(A, (-1, N))); G = _f(np.reshape(G, (-1, N)))
    M, J = A.shape[0], G.shape[0]
    b, g, u = (np.ascontiguousarray(t, dtype=np.float64).ravel() for t in (b, g, u))
    x = np.zeros(N); S = np.zeros(N + J, dtype=np.int32); stats = np.zeros(3)
    st = L.ssqp_oracle_init(N, M, J, _dp(A), _dp(G), _dp(b), _dp(g), _dp(d), _dp(u), tol, _dp(x), _ip(S), _dp(stats))
    return x, S, int(st), stats


def init_batch(A, G, b, g, d, u, tol=2.0 ** -26, nthreads=0):
    """initQP over a batch (OpenMP, one QP per thread).  b: (nb,M) or (M,), g: (nb,J) or (J,), d,u: (nb,N) or (N,).
    Returns dict(x (nb,N), S (nb,N+J), status (nb,), stats (nb,3) [loops, pivots, flips])."""
    L = lib()
    d = np.ascontiguousarray(d, dtype=np.float64); N = d.shape[-1]
    A = _f(np.reshape(A, (-1, N))); G = _f(np.reshape(G, (-1, N)))
    M, J = A.shape[0], G.shape[0]
    arrs = {"b": np.ascontiguousarray(b, dtype=np.float64), "g": np.ascontiguousarray(g, dtype=np.float64),
            "d": d, "u": np.ascontiguousarray(u, dtype=np.float64)}
    nb = max(v.shape[0] if v.ndim == 2 else 1 for v in arrs.values())
    st = {k: (v.shape[1] if v.ndim == 2 else 0) for k, v in arrs.items()}
    x = np.zeros((nb, N)); S = np.zeros((nb, N + J), dtype=np.int32); status = np.zeros(nb, dtype=np.int64); stats = np.zeros((nb, 3))
    L.ssqp_oracle_init_batch(N, M, J, nb, _dp(A), _dp(G), _dp(arrs["b"]), st["b"], _dp(arrs["g"]), st["g"], _dp(arrs["d"]), st["d"],
                             _dp(arrs["u"]), st["u"], tol, _dp(x), _ip(S), status.ctypes.data_as(C.POINTER(C.c_int64)), _dp(stats), nthreads)
    return dict(x=x, S=S, status=status, stats=stats)


def solve_batch(V, A, G, q, b, g, d, u, settings=None, settingsLP=None, nthreads=0, want_stats=False):
    """Loop of solveQP over a batch, OpenMP-threaded (one QP per thread).

    V: (N,N) shared or (nb,N,N) per-QP (each symmetric); q,d,u: (nb,N) or (N,); b: (nb,M) or (M,); g: (nb,J) or (J,).
    Returns dict(x (nb,N), S (nb,N+J), status (nb,), threads, stats)."""
    L = lib()
    V = np.ascontiguousarray(V, dtype=np.float64)
    N = V.shape[-1]
    A = _f(np.reshape(A, (-1, N))); G = _f(np.reshape(G, (-1, N)))
    M, J = A.shape[0], G.shape[0]
    arrs = {}
    nb = None
    for name, t, w in (("q", q, N), ("b", b, M), ("g", g, J), ("d", d, N), ("u", u, N)):
        t = np.ascontiguousarray(t, dtype=np.float64)
        if t.ndim == 2:
            nb = t.shape[0] if nb is None else nb
            assert t.shape == (nb, w), (name, t.shape)
        arrs[name] = t
    if V.ndim == 3:
        nb = V.shape[0] if nb is None else nb
    assert nb is not None, "at least one per-QP array is needed to define the batch size"
    strides = {k: (v.shape[1] if v.ndim == 2 else 0) for k, v in arrs.items()}
    sV = N * N if V.ndim == 3 else 0
    st = settings or default_settings()
    stl = settingsLP or st
    x = np.zeros((nb, N)); S = np.zeros((nb, N + J), dtype=np.int32); status = np.zeros(nb, dtype=np.int64)
    stats = np.zeros((nb, 8)) if want_stats else None
    used = L.ssqp_oracle_solve_batch(N, M, J, nb, _dp(V), sV, _dp(A), _dp(G),
                                     _dp(arrs["q"]), strides["q"], _dp(arrs["b"]), strides["b"],
                                     _dp(arrs["g"]), strides["g"], _dp(arrs["d"]), strides["d"],
                                     _dp(arrs["u"]), strides["u"], C.byref(st), C.byref(stl),
                                     _dp(x), _ip(S), status.ctypes.data_as(C.POINTER(C.c_int64)),
                                     _dp(stats) if want_stats else None, nthreads)
    return dict(x=x, S=S, status=status, threads=int(used), stats=stats)


def set_fix_flip(on):
    """initQP's status flip of the (-Inf,u] variables (src/SSQP.jl:552-557) is a no-op comparison in the reference.
    on=True: the intended flip (DN -> UP), which is what the device path implements; on=False (default): literal."""
    lib().ssqp_oracle_set_fix_flip(1 if on else 0)


def set_rule(rule):
    """Pivot rule used by initQP / SimplexLP (Settings.rule, src/types.jl:397): "Dantzig" (cDantzigLP), "stpEdgeLP"
    (src/Simplex.jl:234-416) or "maxImprovement" (maxImprvLP, src/Simplex.jl:641-813)."""
    lib().ssqp_oracle_set_rule({"Dantzig": 0, "stpEdgeLP": 1, "maxImprovement": 2}[rule])


def get_rows_gjr(X, tol=2.0 ** -33):
    L = lib()
    X = _f(X)
    nr, nc = X.shape
    rows = np.zeros(max(nr, 1), dtype=np.int32)
    l1 = C.c_int32(0)
    n = L.ssqp_oracle_get_rows_gjr(nr, nc, _dp(X), tol, _ip(rows), C.byref(l1))
    return rows[:n].copy(), int(l1.value)


def dantzig_lp(c, A, b, d, u, B, S, invB=None, q=None, tol=2.0 ** -26):
    """cDantzigLP (src/Simplex.jl:445-615) from the basis B (0-based, sorted).  Returns (status, x, B, S)."""
    L = lib()
    A = _f(A); M, N = A.shape
    c, b, d, u = (np.ascontiguousarray(t, dtype=np.float64).ravel() for t in (c, b, d, u))
    B = np.ascontiguousarray(B, dtype=np.int32).copy(); S = np.ascontiguousarray(S, dtype=np.int32).copy()
    invB = _f(np.eye(M) if invB is None else invB)
    q = np.ascontiguousarray(b if q is None else q, dtype=np.float64).ravel()
    x = np.zeros(N)
    st = L.ssqp_oracle_dantzig_lp(N, M, _dp(c), _dp(A), _dp(b), _dp(d), _dp(u), _ip(B), _ip(S), _dp(invB), _dp(q), tol, _dp(x))
    return int(st), x, B, S


def simplex_lp(c, A, G, b, g, d, u, tol=2.0 ** -26):
    """SimplexLP(P::LP) (src/Simplex.jl:831-1034).  Returns dict(x, S, status, stats[loops,pivots,flips])."""
    L = lib()
    c = np.ascontiguousarray(c, dtype=np.float64).ravel(); N = c.size
    A = _f(np.reshape(A, (-1, N))); G = _f(np.reshape(G, (-1, N)))
    M, J = A.shape[0], G.shape[0]
    b, g, d, u = (np.ascontiguousarray(t, dtype=np.float64).ravel() for t in (b, g, d, u))
    x = np.zeros(N); S = np.full(N + J, DN, dtype=np.int32); stats = np.zeros(3)
    st = L.ssqp_oracle_simplex_lp(N, M, J, _dp(c), _dp(A), _dp(G), _dp(b), _dp(g), _dp(d), _dp(u), tol, _dp(x), _ip(S), _dp(stats))
    return dict(x=x, S=S, status=int(st), stats=stats)
