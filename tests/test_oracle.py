"""CPU tests of the oracle (oracle/ssqp_oracle.cpp): the reference's own known-answer tests, an independent
restatement of getRowsGJr, optimality certificates recomputed in numpy/scipy, and the committed golden vectors."""
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def S():
    import ssqp_b200
    return ssqp_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ssqp_oracle
    ssqp_oracle.build()
    return ssqp_oracle


def test_reference_kat_qp(S, O):
    """test/runtests.jl:22-32: solveQP(QP(V; u=[0.7,Inf,0.7])) -> Status[UP, IN, IN]."""
    k = S.workloads.kat_3asset()
    r = O.solve_qp(k["V"], k["A"], k["G"], k["q"][0], k["b"][0], k["g"][0], k["d"][0], k["u"][0])
    assert list(r["S"]) == [O.UP, O.IN, O.IN]
    assert r["status"] == 2
    np.testing.assert_allclose(r["x"], [0.7, 11 / 210, 52 / 210], rtol=1e-13)     # hand-derived optimum


def test_reference_kat_lp_unbounded(O):
    """test/runtests.jl:7-19: min -3x1-2x2, -x1+3x2<=12, x1-5x2<=5, x>=0 -> status 3.  The on-path function is
    cDantzigLP (src/Simplex.jl:445-615); it is run on the slack form from the all-slack basis."""
    c = np.array([-3.0, -2, 0, 0]); A = np.array([[-1.0, 3, 1, 0], [1, -5, 0, 1]]); b = np.array([12.0, 5])
    st, x, B, Sv = O.dantzig_lp(c, A, b, np.zeros(4), np.full(4, np.inf), B=[2, 3], S=[O.DN, O.DN, O.IN, O.IN])
    assert st == 3


def gjr_python(X, tol):
    """Independent, literal transcription of getRowsGJr's control flow (src/utils.jl:49-86) in numpy."""
    A = np.array(X, dtype=float)
    nr, nc = A.shape
    rows, c0, l1 = [], list(range(nc)), 0
    i = j = 0
    while i < nr and j < nc:
        seg = np.abs(A[i, c0[j:]])
        mj = int(np.argmax(seg)); m = seg[mj]; mj += j
        if m <= tol:
            i += 1
            continue
        rows.append(i)
        c0[mj], c0[j] = c0[j], c0[mj]
        n = c0[j]
        cols = c0[j:]
        A[i, cols] = A[i, cols] / A[i, n]
        for k in range(nr):
            if k != i:
                dk = A[k, n]
                A[k, cols] = A[k, cols] - dk * A[i, cols]
        l1 = j + 1
        i += 1; j += 1
    return rows, l1


def test_get_rows_gjr_matches_literal_transcription(O):
    rng = np.random.default_rng(0)
    for trial in range(40):
        nr, nc = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        X = rng.normal(size=(nr, nc))
        if nr > 2 and trial % 2 == 0:
            X[-1] = X[0] * 2 - X[1]                      # a dependent row
        if trial % 5 == 0:
            X[:, -1] = 0.0
        rows, l1 = O.get_rows_gjr(X, tol=2.0 ** -26)
        rp, lp = gjr_python(X, 2.0 ** -26)
        assert list(rows) == rp and l1 == lp
    rows, l1 = O.get_rows_gjr(np.array([[1.0, 1, 1, 1], [2, 2, 2, 2], [0, 1, 0, 0.5]]))
    assert list(rows) == [0, 2]


def kkt_certificate(c, x, Sv, tolG=1e-8):
    """Feasibility, complementarity with the status vector, stationarity and multiplier signs, recomputed in numpy."""
    V, A, G, q, b, g, d, u = (c[k] for k in ("V", "A", "G", "q", "b", "g", "d", "u"))
    N = x.size
    Sz, Se = Sv[:N], Sv[N:]
    assert np.abs(A @ x - b).max(initial=0) < 1e-9
    assert (G @ x <= g + 1e-9).all() and (x >= d - 1e-12).all() and (x <= u + 1e-12).all()
    assert np.all(x[Sz == 1] == d[Sz == 1]) and np.all(x[Sz == 2] == u[Sz == 2])
    F = Sz == 0; E = Se == 4
    gr = V @ x + q
    AE = np.vstack([A, G[E]])
    lam = np.linalg.lstsq(AE[:, F].T, -gr[F], rcond=None)[0]
    assert np.abs(gr[F] + AE[:, F].T @ lam).max(initial=0) < tolG
    gam = gr + AE.T @ lam
    assert (gam[Sz == 1] >= -tolG).all() and (gam[Sz == 2] <= tolG).all()
    assert (lam[A.shape[0]:] >= -tolG).all()


def test_oracle_solutions_satisfy_kkt(S, O):
    rng = np.random.default_rng(7)
    for trial in range(12):
        N, M, J = int(rng.integers(4, 30)), int(rng.integers(0, 3)), int(rng.integers(0, 6))
        B = rng.normal(size=(N, N)); V = B @ B.T / N + 0.05 * np.eye(N)
        x0 = rng.uniform(0.1, 0.9, N)
        A = rng.normal(size=(M, N)); G = rng.normal(size=(J, N))
        c = dict(V=V, A=A, G=G, q=rng.normal(size=N), b=A @ x0, g=G @ x0 + rng.uniform(0, 0.5, J), d=np.zeros(N),
                 u=np.ones(N))
        r = O.solve_qp(V, A, G, c["q"], c["b"], c["g"], c["d"], c["u"])
        assert r["status"] > 0
        kkt_certificate(c, r["x"], r["S"])


def test_oracle_objective_matches_scipy(S, O):
    from scipy.optimize import minimize, LinearConstraint, Bounds
    k = S.workloads.config2(nb=3, N=25)
    for i in range(3):
        V, q = k["V"], k["q"][i]
        r = O.solve_qp(V, k["A"], k["G"], q, k["b"][i], k["g"][i], k["d"][i], k["u"][i])
        f = lambda x: 0.5 * x @ V @ x + q @ x
        res = minimize(f, np.full(25, 1 / 25), jac=lambda x: V @ x + q, hess=lambda x: V, method="trust-constr",
                       constraints=[LinearConstraint(k["A"], k["b"][i], k["b"][i])], bounds=Bounds(k["d"][i], k["u"][i]),
                       options=dict(gtol=1e-12, xtol=1e-14, maxiter=3000))
        # trust-constr is an interior-point method (barrier parameter ~4e-10): the active-set optimum can only be lower
        assert f(r["x"]) <= res.fun + 1e-12
        assert abs(f(r["x"]) - res.fun) < 2e-3 * abs(res.fun)


def test_oracle_matches_golden_vectors(S, O):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from this same oracle: a regression pin."""
    import make_golden
    n = 0
    for name, c in make_golden.cases():
        gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
        r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
        assert np.array_equal(r["status"], gold["status"]), name
        assert np.array_equal(r["S"], gold["S"].astype(np.int32)), name
        ok = gold["status"] > 0
        assert np.abs(r["x"] - gold["x"])[ok].max(initial=0) < 1e-12, name
        n += 1
    assert n >= 7


def test_degenerate_golden_is_the_posdef_exception_case(S):
    """QP 280 of the 296-QP config-4 batch: a multi-blocking step leaves more active rows than free variables, the
    reference's Schur complement is singular (PosDefException, src/SSQP.jl:328) -> status -1."""
    gold = np.load(os.path.join(HERE, "golden", "config4_degenerate_qp280_of_296.npz"))
    assert gold["status"].tolist()[1] == -1 and gold["status"][0] > 0 and gold["status"][2] > 0


def test_reference_kat_lp_through_simplexlp(S, O):
    """test/runtests.jl:7-19 end to end: SimplexLP(LP(c, A, b; d, u, G, g)) -> status == 3."""
    k = S.workloads.kat_lp_unbounded()
    r = O.simplex_lp(k["c"][0], k["A"], k["G"], k["b"][0], k["g"][0], k["d"][0], k["u"][0])
    assert r["status"] == 3


def test_oracle_simplexlp_matches_highs(O):
    from scipy.optimize import linprog
    rng = np.random.default_rng(0)
    for t in range(8):
        N, M, J = 10, 2, 5
        A = rng.normal(size=(M, N)); x0 = rng.uniform(0.1, 0.9, N); b = A @ x0
        G = rng.normal(size=(J, N)); g = G @ x0 + rng.uniform(0, 0.5, J)
        c = rng.normal(size=N)
        r = O.simplex_lp(c, A, G, b, g, np.zeros(N), np.ones(N))
        lp = linprog(c, A_ub=G, b_ub=g, A_eq=A, b_eq=b, bounds=[(0, 1)] * N, method="highs")
        assert r["status"] in (1, 2)
        assert abs(c @ r["x"] - lp.fun) < 1e-10
    # infeasible: sum(x) = 5 with 0 <= x <= 1, N = 3
    r = O.simplex_lp(np.ones(3), np.ones((1, 3)), np.zeros((0, 3)), [5.0], [], np.zeros(3), np.ones(3))
    assert r["status"] == 0


def test_oracle_general_bounds_match_scipy(S, O):
    """Free and (-Inf,u] variables (src/SSQP.jl:484-509, 540-558).  Literal restatement: the reference's no-op status flip
    (:552-557) leaves a negated variable DN with d = -Inf and polishSz! then writes x = -Inf; with the flip repaired
    (set_fix_flip) the optimum agrees with an independent SLSQP solve."""
    from scipy.optimize import minimize
    c = S.workloads.general_bounds(nb=3, N=24, M=2, J=8, seed=21)
    V, A, G = c["V"], c["A"], c["G"]
    literal_finite = []
    for i in range(3):
        r = O.solve_qp(V, A, G, c["q"][i], c["b"][i], c["g"][i], c["d"][i], c["u"][i])
        literal_finite.append(bool(np.isfinite(r["x"]).all()))
    assert not all(literal_finite)
    O.set_fix_flip(True)
    try:
        for i in range(3):
            r = O.solve_qp(V, A, G, c["q"][i], c["b"][i], c["g"][i], c["d"][i], c["u"][i])
            assert r["status"] > 0 and np.isfinite(r["x"]).all()
            x = r["x"]
            assert np.abs(A @ x - c["b"][i]).max() < 1e-9 and (G @ x - c["g"][i]).max() < 1e-9
            assert (x - c["u"][i]).max() <= 0 and (c["d"][i] - x).max() <= 0
            cons = [dict(type="eq", fun=lambda y: A @ y - c["b"][i]), dict(type="ineq", fun=lambda y: c["g"][i] - G @ y)]
            res = minimize(lambda y: 0.5 * y @ V @ y + c["q"][i] @ y, np.zeros(24), jac=lambda y: V @ y + c["q"][i],
                           bounds=list(zip(c["d"][i], c["u"][i])), constraints=cons, method="SLSQP",
                           options=dict(maxiter=500, ftol=1e-14))
            f = 0.5 * x @ V @ x + c["q"][i] @ x
            assert abs(f - res.fun) <= 1e-7 * max(1.0, abs(f))
            Sx = r["S"][:24]
            assert (Sx[c["kind"][i] == 1] == O.IN).all() and not (Sx[c["kind"][i] == 2] == O.DN).any()
    finally:
        O.set_fix_flip(False)


def test_oracle_simplexlp_general_bounds_match_highs(S, O):
    """SimplexLP with free and (-Inf,u] variables (src/Simplex.jl:861-887, 996-1032) against HiGHS on bounded LPs."""
    from scipy.optimize import linprog
    w = S.workloads.general_bounds_lp(nb=5, N=30, M=4, J=14, seed=3, bounded=True)
    for i in range(5):
        r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
        res = linprog(w["c"][i], A_ub=w["G"], b_ub=w["g"][i], A_eq=w["A"], b_eq=w["b"][i],
                      bounds=list(zip(w["d"][i], w["u"][i])), method="highs")
        assert r["status"] in (1, 2) and res.status == 0
        x = r["x"]
        assert abs(w["c"][i] @ x - res.fun) <= 1e-8 * max(1.0, abs(res.fun))
        assert np.abs(w["A"] @ x - w["b"][i]).max() < 1e-8 and (w["G"] @ x - w["g"][i]).max() < 1e-8
        assert (x - w["u"][i]).max() <= 1e-12 and (w["d"][i] - x).max() <= 1e-12
        assert not (r["S"][:30][w["kind"][i] == 2] == O.DN).any()


def test_oracle_other_pivot_rules_match_highs(S, O):
    """stpEdgeLP / maxImprvLP (src/Simplex.jl:234-416, 641-813) restated in the oracle: same optimum as HiGHS through
    SimplexLP, and the same QP solution as the Dantzig rule through initQP + solveQP."""
    from scipy.optimize import linprog
    w = S.workloads.general_bounds_lp(nb=4, N=30, M=4, J=14, seed=3, bounded=True)
    c = S.workloads.config4(nb=2, N=60, J=12)
    ref = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
    try:
        for rule in ("stpEdgeLP", "maxImprovement"):
            O.set_rule(rule)
            for i in range(4):
                r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
                res = linprog(w["c"][i], A_ub=w["G"], b_ub=w["g"][i], A_eq=w["A"], b_eq=w["b"][i],
                              bounds=list(zip(w["d"][i], w["u"][i])), method="highs")
                assert r["status"] in (1, 2) and abs(w["c"][i] @ r["x"] - res.fun) <= 1e-8 * max(1.0, abs(res.fun))
            r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
            assert (r["status"] > 0).all() and np.array_equal(r["S"], ref["S"])
            assert np.abs(r["x"] - ref["x"]).max() <= 1e-9 * np.abs(ref["x"]).max()
    finally:
        O.set_rule("Dantzig")


def test_oracle_reproduces_the_cycling_qp(S, O):
    """QP 25306 of the 32 768-QP config-4 batch: the reference's method alternates between K = 42 and K = 41 at a degenerate
    vertex until maxIter (src/SSQP.jl:271-274).  The literal restatement must do the same: status -(maxIter+1), and a trace
    whose tail is the exact two-trip alternation the device's cycle watch relies on."""
    c = S.workloads.config4(index=np.array([25306]), total=32768)
    r = O.solve_qp(c["V"], c["A"], c["G"], c["q"][0], c["b"][0], c["g"][0], c["d"][0], c["u"][0], trace=True)
    assert r["status"] == -7778
    tail = r["trace"][-200:]
    assert np.array_equal(tail[::2], np.tile(tail[0], (100, 1))) and np.array_equal(tail[1::2], np.tile(tail[1], (100, 1)))
    assert sorted({int(tail[0][0]), int(tail[1][0])}) == [41, 42]


def test_lstsq_matches_numpy(O):
    """The least-squares solve behind the purged-row multipliers of KKTchk! (x = AE' \\ GE[j,F], src/SSQP.jl:158).  Round 2's
    full-shard parity check caught a bug here (column norms not swapped with their columns in the pivoted QR): two QPs of
    the 8 192 took another release at a degenerate vertex than the device — and the device was the one following the
    reference's formula."""
    import ctypes as C
    L = O.lib()
    dp = C.POINTER(C.c_double)
    L.ssqp_oracle_lstsq.argtypes = [C.c_int32, C.c_int32, dp, dp, dp]
    rng = np.random.default_rng(0)
    for m, n in ((73, 73), (10, 4), (5, 5), (30, 30), (40, 12), (7, 1)):
        X = np.asfortranarray(rng.uniform(0, 1, (m, n)) * (rng.uniform(0, 1, (m, n)) < 0.6))
        y = rng.standard_normal(m)
        x = np.zeros(n)
        L.ssqp_oracle_lstsq(m, n, X.ctypes.data_as(dp), y.ctypes.data_as(dp), x.ctypes.data_as(dp))
        xr = np.linalg.lstsq(X, y, rcond=None)[0]
        assert np.abs(x - xr).max() <= 1e-10 * max(1.0, np.abs(xr).max()), (m, n)


def test_oracle_forms_agree_on_final_answers(S, O):
    """LAPACK form (OpenBLAS dpotrf/dpotri/dgetrf/dgetri/dgemm/dgemv through scipy — the routines Julia calls) vs the scalar
    form on a config-4 sample: the same final status vectors and x to 1e-9; trip counts may differ where a Phase-1
    ratio-test tie is decided by roundoff (on the full bench shard the two forms disagree on 6.8 % of the trip counts)."""
    idx = np.linspace(0, 65535, 8).astype(int)
    c = S.workloads.config4(index=idx, total=65536)
    out = {}
    for form in ("scalar", "lapack"):
        assert O.use_lapack(form == "lapack") == form
        out[form] = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
    O.use_lapack(False)
    a, b = out["scalar"], out["lapack"]
    assert (a["status"] > 0).all() and (b["status"] > 0).all()
    assert np.array_equal(a["S"], b["S"])
    assert (np.abs(a["x"] - b["x"]).max(axis=1) / np.abs(a["x"]).max(axis=1)).max() < 1e-9


def test_full_shard_golden_is_self_consistent():
    """tests/golden/config4_shard0of8.npz: both oracle forms end on the same status vectors (optimal QPs) and agree in x."""
    g = np.load(os.path.join(HERE, "golden", "config4_shard0of8.npz"))
    ok = (g["status_lapack"] > 0) & (g["status_scalar"] > 0)
    assert ok.sum() == 8189 and np.array_equal(g["status_lapack"] <= 0, g["status_scalar"] <= 0)
    assert np.array_equal(g["S_lapack"][ok], g["S_scalar"][ok])
    P = np.linalg.norm(np.random.default_rng(20261018).standard_normal((4, 500)), axis=1)
    rel = np.abs(g["proj_lapack"] - g["proj_scalar"]) / (g["xinf_scalar"][:, None] * P[None, :])
    assert rel[ok].max() < 1e-9
    same = (g["status_lapack"] == g["status_scalar"]).mean()
    assert 0.90 < same < 0.97          # 93.2 %: the reference's own trip counts depend on LAPACK-level roundoff


def test_oracle_keeps_the_reference_status_overwrite_on_unbounded_lps_with_free_variables(S, O):
    """SimplexLP recomputes the status from the reduced costs whenever the LP has free variables (`if n > 0 ... iH = sum(ih) > 0
    ? 2 : 1`, src/Simplex.jl:1001-1019) — also after Phase 2 returned 3 (unbounded, :527-541): an unbounded LP with free
    variables comes back as status 1 / 2 with the vertex the simplex stood on.  The restatement keeps that (the device too:
    profiles/r02_v5_fuzz_replay.log); without free variables the same LPs report 3.  HiGHS confirms they are unbounded."""
    from scipy.optimize import linprog
    w = S.workloads.general_bounds_lp(nb=6, N=123, M=1, J=55, seed=256669752, bounded=False)
    i = 1
    assert (np.isinf(w["d"][i]) & np.isinf(w["u"][i])).sum() > 0
    bnds = [(None if np.isinf(a) else a, None if np.isinf(b_) else b_) for a, b_ in zip(w["d"][i], w["u"][i])]
    lp = linprog(w["c"][i], A_ub=w["G"], b_ub=w["g"][i], A_eq=w["A"], b_eq=w["b"][i], bounds=bnds, method="highs")
    assert lp.status == 3                                   # scipy: 3 = unbounded
    r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
    assert r["status"] in (1, 2)                            # the reference's overwrite
    x = r["x"]                                              # ... at a feasible point
    assert np.abs(w["A"] @ x - w["b"][i]).max() < 1e-9 and (w["G"] @ x - w["g"][i]).max() < 1e-9
    w3 = S.workloads.general_bounds_lp(nb=4, N=30, M=3, J=8, seed=5, bounded=False)
    d3 = w3["d"].copy(); d3[np.isinf(w3["d"]) & np.isinf(w3["u"])] = -1.5                # lower-only instead of free: status 3 survives
    assert [O.simplex_lp(w3["c"][k], w3["A"], w3["G"], w3["b"][k], w3["g"][k], d3[k], w3["u"][k])["status"] for k in range(4)] == [3, 3, 3, 3]
