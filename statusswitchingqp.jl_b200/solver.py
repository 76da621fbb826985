"""solveQP on the GPU — mirror of the reference's solver entry points (src/SSQP.jl:213-377).

    solveQP(Q)                 -> (z, S, status)                one QP        (src/SSQP.jl:224)
    solveQP(Q, S, x0)          -> (z, S, status)                warm start    (src/SSQP.jl:237)
    solveQP([Q1, Q2, ...])     -> list of (z, S, status)        a loop of solveQP calls as ONE device batch
    solveQP_batch(V, A, G, q, b, g, d, u) -> (X, S, status)     array form of the same batch

Every solve goes through the C ABI (libssqp_b200.so).  There is no CPU fallback.
"""
import numpy as np
from . import capi
from .types import QP, LP, Settings, Status, DN

_ctx = None
_ctx_key = None


def context(devices=None):
    """Process-wide Context (created lazily).  devices=None -> device 0."""
    global _ctx, _ctx_key
    key = tuple(devices) if devices is not None else (0,)
    if _ctx is None or _ctx_key != key:
        if _ctx is not None:
            _ctx.close()
        _ctx = capi.Context(list(key))
        _ctx_key = key
    return _ctx


def _settings(s):
    return (s or Settings()).to_c()


def solveQP_batch(V, A, G, q, b, g, d, u, S0=None, x0=None, settings=None, settingsLP=None, ctx=None,
                  return_stats=False):
    """Batch of QPs sharing A and G (and V when V.ndim == 2; V.ndim == 3 -> one V per QP).
    q,d,u: (nb,N); b: (nb,M); g: (nb,J).  Returns X (nb,N), S (nb,N+J) int32 Status codes, status (nb,) int64."""
    ctx = ctx or context()
    V = np.asarray(V, dtype=np.float64)
    shared_V = V if V.ndim == 2 else None
    ctx.set_shared(shared_V, A, G)
    st = _settings(settings)
    stlp = _settings(settingsLP) if settingsLP is not None else st
    out = ctx.solve_batch(q, b, g, d, u, V_per_qp=(None if shared_V is not None else V), S0=S0, x0=x0,
                          settings=st, settingsLP=stlp)
    if return_stats:
        return out + (ctx.stats(out[0].shape[0]),)
    return out


def solveQP_sweep(V, A, G, q, b, g, d, u, chain_len, settings=None, settingsLP=None, ctx=None, return_stats=False):
    """Warm-started sweep over the linear term: the batch is cut into chains of `chain_len` consecutive QPs
    that share b, g, d, u; inside a chain QP t+1 is solveQP(Q[t+1], S[t], x[t]) (src/SSQP.jl:237), i.e. the loop a user
    of the reference writes along a frontier (QP(P, q, L), src/types.jl:303-319); the chains run in parallel on the device.
    Returns X, S, status like solveQP_batch; status[i] is the iteration count of that (warm-started) call."""
    ctx = ctx or context()
    V = np.asarray(V, dtype=np.float64)
    shared_V = V if V.ndim == 2 else None
    ctx.set_shared(shared_V, A, G)
    st = _settings(settings)
    stlp = _settings(settingsLP) if settingsLP is not None else st
    out = ctx.solve_sweep(q, b, g, d, u, chain_len, V_per_qp=(None if shared_V is not None else V), settings=st, settingsLP=stlp)
    if return_stats:
        return out + (ctx.stats(out[0].shape[0]),)
    return out


def solveQP(Q, S=None, x0=None, settings=None, settingsLP=None, ctx=None):
    """Drop-in for the reference's solveQP.  `Q` may be a QP or a sequence of QPs of equal shape that share
    A and G (the frontier-sweep constructors QP.with_L / QP.with_mu produce exactly that)."""
    if isinstance(Q, QP):
        if Q.mc <= 0:                                                   # src/SSQP.jl:226-228
            return np.zeros(Q.N), np.full(Q.N, int(DN), dtype=np.int32), -1
        X, Sv, status = solveQP_batch(Q.V, Q.A, Q.G, Q.q[None], Q.b[None], Q.g[None], Q.d[None], Q.u[None],
                                      S0=None if S is None else np.asarray(S, dtype=np.int32)[None],
                                      x0=None if x0 is None else np.asarray(x0, dtype=np.float64)[None],
                                      settings=settings, settingsLP=settingsLP, ctx=ctx)
        if S is not None:                       # the reference mutates the caller's S in place
            np.asarray(S)[...] = Sv[0]
        return X[0], Sv[0], int(status[0])
    Qs = list(Q)
    if not Qs:
        return []
    P0 = Qs[0]
    good = [i for i, P in enumerate(Qs) if P.mc > 0]
    res = [None] * len(Qs)
    for i, P in enumerate(Qs):
        if P.mc <= 0:
            res[i] = (np.zeros(P.N), np.full(P.N, int(DN), dtype=np.int32), -1)
        if (P.N, P.M, P.J) != (P0.N, P0.M, P0.J) or not (P.A is P0.A or np.array_equal(P.A, P0.A)) \
                or not (P.G is P0.G or np.array_equal(P.G, P0.G)):
            raise ValueError("a device batch must share N, M, J, A and G; split the list by shape")
    if good:
        same_V = all(Qs[i].V is Qs[good[0]].V for i in good)
        V = Qs[good[0]].V if same_V else np.stack([Qs[i].V for i in good])
        stack = lambda name: np.stack([getattr(Qs[i], name) for i in good])
        X, Sv, status = solveQP_batch(V, P0.A, P0.G, stack("q"), stack("b"), stack("g"), stack("d"), stack("u"),
                                      settings=settings, settingsLP=settingsLP, ctx=ctx)
        for t, i in enumerate(good):
            res[i] = (X[t], Sv[t], int(status[t]))
    return res


def initQP_batch(A, G, b, g, d, u, settingsLP=None, ctx=None):
    """Phase 1 only (initQP, src/SSQP.jl:461-560) for a batch: returns x0 (nb,N), S (nb,N+J), status (nb,)."""
    ctx = ctx or context()
    ctx.set_shared(None, A, G)
    return ctx.init_batch(b, g, d, u, settingsLP=_settings(settingsLP))


def getRowsGJr(X, tol=2.0 ** -33):
    """getRowsGJr (src/utils.jl:49-86): rows of X that Gauss-Jordan elimination with in-row column pivoting finds
    independent, and l1 = the last pivot position.  Host-side helper of SimplexLP's redundancy purge (0-based rows)."""
    A = np.array(X, dtype=np.float64)
    nr, nc = A.shape
    rows = []
    c0 = np.arange(nc)
    l1 = 0
    i = j = 0
    while i < nr and j < nc:
        cols = c0[j:]
        mj = int(np.argmax(np.abs(A[i, cols])))          # first maximum, like findmax
        if not (abs(A[i, cols[mj]]) > tol):
            i += 1
            continue
        rows.append(i)
        c0[j + mj], c0[j] = c0[j], c0[j + mj]
        cols = c0[j:]
        n = c0[j]
        A[i, cols] /= A[i, n]
        f = A[:, n].copy()
        f[i] = 0.0
        A[:, cols] -= np.outer(f, A[i, cols])
        l1 = j + 1
        i += 1
        j += 1
    return rows, l1


def _lp_row_purge(A, G, b, g, d, u, tol):
    """SimplexLP's `purge redundancy` step (src/Simplex.jl:869-902) for one LP: the slack form A0 = [A 0 -A[:,iv]; G I -G[:,iv]]
    with the (-Inf,u] columns negated, m0 = rank(A0); when m0 < M+J the rows getRowsGJr([A0 b0], tol) keeps.
    Returns (rows or None when nothing is dropped, status or None): status 0 infeasible / -1 numerical end the solve."""
    M, N = A.shape
    J = G.shape[0]
    M0 = M + J
    if M0 == 0:
        return None, None
    iv = np.flatnonzero(np.isinf(u) & (u > 0) & np.isinf(d) & (d < 0))
    idn = np.flatnonzero(np.isinf(d) & (d < 0) & ~(np.isinf(u) & (u > 0)))
    A0 = np.block([[A, np.zeros((M, J)), -A[:, iv]], [G, np.eye(J), -G[:, iv]]])
    A0[:, idn] = -A0[:, idn]
    sv = np.linalg.svd(A0, compute_uv=False)
    m0 = int((sv > min(A0.shape) * np.finfo(np.float64).eps * sv[0]).sum()) if sv.size and sv[0] > 0 else 0   # rank(A0)
    if m0 >= M0:
        return None, None
    ra, la = getRowsGJr(np.hstack([A0, np.concatenate([b, g])[:, None]]), tol)
    if len(ra) != la:
        return None, 0
    if m0 != la:
        return None, -1
    return ra, None


def SimplexLP_batch(A, G, c, b, g, d, u, settings=None, ctx=None):
    """Batch of LPs sharing A and G: the reference's two-phase SimplexLP (src/Simplex.jl:831-1034, Dantzig rule).
    c,d,u: (nb,N); b: (nb,M); g: (nb,J).  Returns X (nb,N), S (nb,N+J), status (nb,) with the reference's codes:
    1 unique optimum, 2 infinitely many optima, 3 unbounded, 0 infeasible, -1 numerical / not on the device path."""
    ctx = ctx or context()
    c, b, g, d, u = (np.ascontiguousarray(t, dtype=np.float64) for t in (c, b, g, d, u))
    N = c.shape[1]
    A = np.asarray(A, dtype=np.float64).reshape(-1, N)
    G = np.asarray(G, dtype=np.float64).reshape(-1, N)
    M, J = A.shape[0], G.shape[0]
    nb = c.shape[0]
    st = settings or Settings()
    # The device needs [A 0; G I] of full row rank.  rank(A0) depends on A, G and on which variables are free, so it is
    # checked once per distinct bound pattern; only rank-deficient inputs take the reference's purge (:889-902), per LP.
    groups = {}
    early = {}
    pattern_rank_ok = {}
    for i in range(nb):
        key = (np.isinf(d[i]) & (d[i] < 0)).tobytes() + (np.isinf(u[i]) & (u[i] > 0)).tobytes() if M > 0 else b""
        if key not in pattern_rank_ok:
            rows, stat = _lp_row_purge(A, G, b[i], g[i], d[i], u[i], st.tol)
            pattern_rank_ok[key] = rows is None and stat is None
        if pattern_rank_ok[key]:
            groups.setdefault(None, []).append(i)
            continue
        rows, stat = _lp_row_purge(A, G, b[i], g[i], d[i], u[i], st.tol)
        if stat is not None:
            early[i] = stat
        elif rows is not None and any(r >= M for r in set(range(M + J)) - set(rows)):
            early[i] = -1                        # an inequality row judged redundant: not on the device path
        else:
            groups.setdefault(None if rows is None else tuple(r for r in rows if r < M), []).append(i)
    X = np.zeros((nb, N))
    Sv = np.full((nb, N + J), int(DN), dtype=np.int32)
    status = np.zeros(nb, dtype=np.int64)
    for i, stat in early.items():               # (zeros(N), fill(DN, N), status) like the reference's early returns
        status[i] = stat
    for rows, idx in groups.items():
        idx = np.asarray(idx)
        Ar, br = (A, b[idx]) if rows is None else (A[list(rows)], b[idx][:, list(rows)])
        ctx.set_shared(None, Ar, G)
        Xg, Sg, sg = ctx.solve_lp_batch(c[idx], np.ascontiguousarray(br), g[idx], d[idx], u[idx], settings=_settings(settings))
        X[idx], Sv[idx], status[idx] = Xg, Sg, sg
    return X, Sv, status


def SimplexLP(P, *arrays, settings=None, min=True, ctx=None):
    """Drop-in for SimplexLP(P::LP; settings, min) (src/Simplex.jl:831).  `P` may be an LP or a sequence of LPs sharing A
    and G.  min=False maximises (the reference negates the cost vector, :981-983).
    Array form, SimplexLP(c, A, b, d, u; settings, min) (src/Simplex.jl:1036-1196): the equality-constrained LP
    min c'x s.t. Ax = b, d <= x <= u — the same two-phase algorithm without the slack block; returns (x, S, status)
    with S of length N."""
    sgn = 1.0 if min else -1.0
    if arrays:
        if len(arrays) != 4:
            raise TypeError("SimplexLP(c, A, b, d, u): expected the five arrays of the reference's array form")
        A, b, d, u = arrays
        return SimplexLP(LP(P, A, b, d=d, u=u), settings=settings, min=min, ctx=ctx)
    if isinstance(P, LP):
        if P.mc <= 0:                                                   # src/Simplex.jl:848-850
            return np.zeros(P.N), np.full(P.N, int(DN), dtype=np.int32), -1
        X, Sv, status = SimplexLP_batch(P.A, P.G, sgn * P.c[None], P.b[None], P.g[None], P.d[None], P.u[None], settings=settings, ctx=ctx)
        return X[0], Sv[0], int(status[0])
    Ps = list(P)
    if not Ps:
        return []
    P0 = Ps[0]
    for Q in Ps:
        if (Q.N, Q.M, Q.J) != (P0.N, P0.M, P0.J) or not np.array_equal(Q.A, P0.A) or not np.array_equal(Q.G, P0.G):
            raise ValueError("a device batch must share N, M, J, A and G; split the list by shape")
    res = [None] * len(Ps)
    good = [i for i, Q in enumerate(Ps) if Q.mc > 0]
    for i, Q in enumerate(Ps):
        if Q.mc <= 0:
            res[i] = (np.zeros(Q.N), np.full(Q.N, int(DN), dtype=np.int32), -1)
    if good:
        stack = lambda name: np.stack([getattr(Ps[i], name) for i in good])
        X, Sv, status = SimplexLP_batch(P0.A, P0.G, sgn * stack("c"), stack("b"), stack("g"), stack("d"), stack("u"), settings=settings, ctx=ctx)
        for t, i in enumerate(good):
            res[i] = (X[t], Sv[t], int(status[t]))
    return res
