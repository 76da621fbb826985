"""Synthetic portfolio-QP workloads of the BASELINE.json configs.

Pure numpy (PCG64 via numpy.random.default_rng) so that the oracle, the CUDA path and the bench see
bit-identical inputs.  Shapes follow examples/SSQPspeed.jl:53-75 of the reference (sum-to-one
equality, box bounds) with the additions BASELINE.json names (general inequalities G x <= g).
"""
import numpy as np


def factor_model(N, seed, nf=10):
    """Factor-model asset returns -> (V, E): sample covariance (symmetrised) and sample mean."""
    rng = np.random.default_rng(seed)
    T = 2 * N + 10
    Bf = rng.normal(0.0, 0.3, (N, nf))
    f = rng.normal(0.0, 0.01, (T, nf))
    idio = rng.normal(0.0, 0.02, (T, N)) * rng.uniform(0.5, 1.5, N)
    mu = rng.normal(5e-4, 5e-4, N)
    R = mu + f @ Bf.T + idio
    V = np.cov(R, rowvar=False)
    V = (V + V.T) / 2
    E = R.mean(axis=0)
    return np.ascontiguousarray(V), np.ascontiguousarray(E)


def ineq_rows(N, J, rng, density=0.2):
    G = rng.uniform(0.0, 1.0, (J, N)) * (rng.uniform(0.0, 1.0, (J, N)) < density)
    g = (G.sum(axis=1) / N) * rng.uniform(0.9, 1.3, J)
    return G, g


def config1(N=300, seed=0, L=0.1):
    """Single mean-variance QP as in examples/SSQPspeed.jl (N~300, 1'x=1, 0<=x<=3/32)."""
    V, E = factor_model(N, seed)
    return dict(V=V, A=np.ones((1, N)), G=np.zeros((0, N)), q=(-L * E)[None, :], b=np.ones((1, 1)),
                g=np.zeros((1, 0)), d=np.zeros((1, N)), u=np.full((1, N), 3.0 / 32), E=E)


def config2(nb=4096, N=100, shared_V=True, seed=1):
    """Batch of random portfolio QPs N=100 (1 equality, box bounds u=0.1)."""
    if shared_V:
        V, E = factor_model(N, seed)
        Ls = np.logspace(-3, np.log10(3.0), nb)
        q = -Ls[:, None] * E[None, :]
    else:
        Vs, qs = [], []
        rng = np.random.default_rng(seed)
        for i in range(nb):
            Vi, Ei = factor_model(N, seed + 1 + i)
            Vs.append(Vi)
            qs.append(-rng.uniform(0.001, 3.0) * Ei)
        V = np.stack(Vs)
        q = np.stack(qs)
    return dict(V=V, A=np.ones((1, N)), G=np.zeros((0, N)), q=q, b=np.ones((nb, 1)), g=np.zeros((nb, 0)),
                d=np.zeros((nb, N)), u=np.full((nb, N), 0.1))


# config 3's target returns: evenly spaced between the return of the minimum-variance portfolio and 0.98 x the
# largest feasible return.  Both ends are properties of the seed-2 problem, computed once with the CPU oracle
# (solveQP with q = 0 and without the return row -> E'x = MU_MINVAR; SimplexLP max E'x -> MU_MAX, equal to HiGHS' value to
# 1e-17) by tests/golden/make_golden.py --mu-range, and pinned here so that the workload does not need the oracle.
CONFIG3_MU_MINVAR = 0.0005125008117583505
CONFIG3_MU_MAX = 0.002311775114739044


def config3(nb=1024, N=500, J=50, seed=2, mu_lo=None, mu_hi=None):
    """Efficient-frontier sweep: target-return QPs sharing V, A=[1';E'], b_i=[1;mu_i], q=0, J ineqs.

    mu range (default problem: N=500, J=50, seed=2): SURVEY 8d's — from the minimum-variance portfolio's return to
    0.98 x the maximum feasible return.  Other shapes fall back to a quantile band of E that is feasible for the G drawn."""
    V, E = factor_model(N, seed)
    rng = np.random.default_rng(seed + 1000)
    G, g = ineq_rows(N, J, rng)
    A = np.vstack([np.ones((1, N)), E[None, :]])
    named = (N, J, seed) == (500, 50, 2)
    if mu_lo is None:
        mu_lo = CONFIG3_MU_MINVAR if named else float(np.quantile(E, 0.45))
    if mu_hi is None:
        mu_hi = 0.98 * CONFIG3_MU_MAX if named else float(np.quantile(E, 0.70))
    mus = np.linspace(mu_lo, mu_hi, nb)
    b = np.stack([np.ones(nb), mus], axis=1)
    return dict(V=V, A=A, G=G, q=np.zeros((nb, N)), b=b, g=np.tile(g, (nb, 1)), d=np.zeros((nb, N)),
                u=np.full((nb, N), 0.05), E=E)


def config4(nb=65536, N=500, J=99, seed=3, start=0, total=None, index=None):
    """Portfolio QPs N=500, M=1, J=99 sharing V/A/G; per-QP q and g.  The batch is defined over `total`
    global QP indices (L log-spaced over the FULL batch, per-QP RNG streams keyed by the global index), so any
    shard — `start:start+nb`, or an explicit `index` array (e.g. rank::world for interleaved sharding) — is
    reproducible on its own."""
    if index is None:
        total = nb if total is None else total
        index = np.arange(start, start + nb)
    else:
        index = np.asarray(index, dtype=np.int64)
        total = int(index.max()) + 1 if total is None else total
    nb = index.size
    V, E = factor_model(N, seed)
    rng = np.random.default_rng(seed + 1000)
    G, g = ineq_rows(N, J, rng)
    Ls = np.logspace(-3, np.log10(3.0), total)[index]
    q = np.empty((nb, N))
    gi = np.empty((nb, J))
    sE = E.std()
    for i in range(nb):
        r = np.random.default_rng([seed, 7, int(index[i])])
        q[i] = -Ls[i] * (E + 0.1 * sE * r.standard_normal(N))
        gi[i] = g * r.uniform(0.95, 1.05, J)
    return dict(V=V, A=np.ones((1, N)), G=G, q=q, b=np.ones((nb, 1)), g=gi, d=np.zeros((nb, N)),
                u=np.full((nb, N), 0.05), E=E, index=index)


def kat_3asset():
    """The reference's own QP known-answer test (test/runtests.jl:22-32): expect S == [UP, IN, IN]."""
    V = np.array([[1 / 100, 1 / 80, 1 / 100], [1 / 80, 1 / 16, 1 / 40], [1 / 100, 1 / 40, 1 / 25]])
    return dict(V=V, A=np.ones((1, 3)), G=np.zeros((0, 3)), q=np.zeros((1, 3)), b=np.ones((1, 1)),
                g=np.zeros((1, 0)), d=np.zeros((1, 3)), u=np.array([[0.7, np.inf, 0.7]]))


def config5(nb=16384, N=1000, M=20, J=180, seed=4, index=None, total=None):
    """Batched LPs (BASELINE config 5): N=1000, M=20 equalities, J=180 inequalities sharing A and G (same sparsity
    recipe as the QP configs), x* ~ U(0,0.5), b = A x*, g = G x* + U(0,0.1), d=0, u=1, per-LP costs c_i ~ N(0,1)."""
    if index is None:
        index = np.arange(nb)
    index = np.asarray(index, dtype=np.int64)
    nb = index.size
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.0, 1.0, (M, N)) * (rng.uniform(0.0, 1.0, (M, N)) < 0.2)
    G = rng.uniform(0.0, 1.0, (J, N)) * (rng.uniform(0.0, 1.0, (J, N)) < 0.2)
    xs = rng.uniform(0.0, 0.5, N)
    b = A @ xs
    g = G @ xs + rng.uniform(0.0, 0.1, J)
    c = np.empty((nb, N))
    for i in range(nb):
        c[i] = np.random.default_rng([seed, 11, int(index[i])]).standard_normal(N)
    return dict(A=A, G=G, c=c, b=np.tile(b, (nb, 1)), g=np.tile(g, (nb, 1)), d=np.zeros((nb, N)), u=np.ones((nb, N)), index=index)


def kat_lp_unbounded():
    """The reference's own LP known-answer test (test/runtests.jl:7-19): SimplexLP -> status 3."""
    return dict(c=np.array([[-3.0, -2.0]]), A=np.zeros((0, 2)), b=np.zeros((1, 0)), G=np.array([[-1.0, 3.0], [1.0, -5.0]]),
                g=np.array([[12.0, 5.0]]), d=np.zeros((1, 2)), u=np.full((1, 2), np.inf))


def general_bounds(nb=6, N=40, M=3, J=12, seed=11):
    """Strictly convex QPs whose variables mix all four bound kinds — box [d,u], free (-Inf,+Inf), upper-only (-Inf,u]
    and lower-only [d,+Inf) — i.e. the branches of initQP that split / negate columns (src/SSQP.jl:484-509, 540-558).
    Feasible by construction (a hidden point inside the bounds satisfies Ax=b, Gx<=g)."""
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((N, N))
    V = B @ B.T / N + 0.1 * np.eye(N)
    V = (V + V.T) / 2
    A = rng.standard_normal((M, N))
    G = rng.standard_normal((J, N))
    xs = rng.uniform(-1, 1, (nb, N))
    b = xs @ A.T
    g = xs @ G.T + rng.uniform(0.0, 0.5, (nb, J))
    d = np.full((nb, N), -1.5)
    u = np.full((nb, N), 1.5)
    kind = rng.integers(0, 4, (nb, N))
    d[kind == 1] = -np.inf
    u[kind == 1] = np.inf
    d[kind == 2] = -np.inf
    u[kind == 3] = np.inf
    q = rng.standard_normal((nb, N))
    return dict(V=V, A=A, G=G, q=q, b=b, g=g, d=d, u=u, kind=kind)


def general_bounds_lp(nb=6, N=30, M=4, J=14, seed=3, bounded=True):
    """LPs whose variables mix box, free, upper-only and lower-only bounds (SimplexLP's split / negation branches,
    src/Simplex.jl:861-887, 996-1032).  Feasible by construction; `bounded`: the cost is a dual-feasible combination of
    the rows (c = A'mu - G'lam, lam > 0), so the optimum is finite whatever the bounds are."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, N))
    G = rng.standard_normal((J, N))
    xs = rng.uniform(-1, 1, (nb, N))
    b = xs @ A.T
    g = xs @ G.T + rng.uniform(0.0, 0.5, (nb, J))
    d = np.full((nb, N), -1.5)
    u = np.full((nb, N), 1.5)
    kind = rng.integers(0, 4, (nb, N))
    d[kind == 1] = -np.inf
    u[kind == 1] = np.inf
    d[kind == 2] = -np.inf
    u[kind == 3] = np.inf
    if bounded:
        c = rng.standard_normal((nb, M)) @ A - rng.uniform(0.1, 1.0, (nb, J)) @ G
    else:
        c = rng.standard_normal((nb, N))
    return dict(A=A, G=G, c=c, b=b, g=g, d=d, u=u, kind=kind)


def degenerate_lps(kind, nb=4, N=24, M=4, J=8, seed=7):
    """Feasible LPs that exercise SimplexLP's rarely taken branches (src/Simplex.jl):
    "zero_row": an equality row with non-positive coefficients and rhs 0 at d = 0 — its artificial variable stays basic at
                level zero after Phase 1 and has to be driven out (:962-977);
    "dup_row":  equality row 3 = 2*row 0 - row 1 with a consistent rhs — rank([A 0; G I]) < M+J, the redundancy purge drops a row (:889-902);
    "dup_row_inconsistent": the same with rhs of every second LP shifted — the purge reports a numerical error (-1)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, N))
    G = rng.standard_normal((J, N))
    xs = rng.uniform(0.1, 1, (nb, N))
    d = np.zeros((nb, N))
    u = np.full((nb, N), 2.0)
    if kind == "zero_row":
        A[1] = 0.0
        A[1, :5] = -rng.uniform(0.5, 1.5, 5)
        xs[:, :5] = 0.0
    if kind in ("dup_row", "dup_row_inconsistent"):
        A[3] = 2.0 * A[0] - A[1]
    b = xs @ A.T
    if kind == "dup_row_inconsistent":
        b[1::2, 3] += 1.0
    g = xs @ G.T + rng.uniform(0.0, 0.5, (nb, J))
    c = rng.standard_normal((nb, N))
    return dict(A=A, G=G, c=c, b=b, g=g, d=d, u=u)
