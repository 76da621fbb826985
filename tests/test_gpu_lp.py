"""GPU parity tests of the LP path (SimplexLP, src/Simplex.jl:831-1034): CUDA (through the C ABI) vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import ssqp_b200
    if ssqp_b200.device_count() < 1:
        pytest.fail("no CUDA device visible")
    return ssqp_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ssqp_oracle
    return ssqp_oracle


def check_lp(S, O, k):
    X, St, status = S.SimplexLP_batch(k["A"], k["G"], k["c"], k["b"], k["g"], k["d"], k["u"])
    for i in range(len(status)):
        r = O.simplex_lp(k["c"][i], k["A"], k["G"], k["b"][i], k["g"][i], k["d"][i], k["u"][i])
        assert status[i] == r["status"], (i, status[i], r["status"])
        if r["status"] in (1, 2):
            assert np.array_equal(St[i], r["S"]), i
            assert np.abs(X[i] - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max()), i
    return X, St, status


def test_reference_kat_lp_unbounded(S, O):
    """test/runtests.jl:7-19: SimplexLP(LP(c, A, b; d, u, G, g)) -> status == 3."""
    k = S.workloads.kat_lp_unbounded()
    P = S.LP(k["c"][0], k["A"], k["b"][0], d=k["d"][0], u=k["u"][0], G=k["G"], g=k["g"][0])
    x, Sv, status = S.SimplexLP(P)
    assert status == 3
    check_lp(S, O, k)


def test_random_bounded_lps(S, O):
    rng = np.random.default_rng(0)
    N, M, J, nb = 12, 2, 6, 16
    A = rng.normal(size=(M, N)); x0 = rng.uniform(0.1, 0.9, N)
    G = rng.normal(size=(J, N))
    k = dict(A=A, G=G, c=rng.normal(size=(nb, N)), b=np.tile(A @ x0, (nb, 1)), g=np.tile(G @ x0 + rng.uniform(0, 0.5, J), (nb, 1)),
             d=np.zeros((nb, N)), u=np.ones((nb, N)))
    X, St, status = check_lp(S, O, k)
    assert set(status.tolist()) <= {1, 2}
    from scipy.optimize import linprog
    for i in range(4):
        lp = linprog(k["c"][i], A_ub=G, b_ub=k["g"][i], A_eq=A, b_eq=k["b"][i], bounds=[(0, 1)] * N, method="highs")
        assert abs(k["c"][i] @ X[i] - lp.fun) < 1e-9


def test_infeasible_and_inequality_only(S, O):
    # infeasible: sum(x) = 5 with 0 <= x <= 1, N = 3
    k = dict(A=np.ones((1, 3)), G=np.zeros((0, 3)), c=np.ones((1, 3)), b=np.array([[5.0]]), g=np.zeros((1, 0)), d=np.zeros((1, 3)), u=np.ones((1, 3)))
    X, St, status = check_lp(S, O, k)
    assert status[0] == 0
    # inequalities only, infinite upper bounds, bounded optimum
    G = np.array([[1.0, 1.0], [1.0, -1.0]]); 
    k = dict(A=np.zeros((0, 2)), G=G, c=np.array([[-1.0, -2.0], [1.0, 1.0]]), b=np.zeros((2, 0)), g=np.tile([4.0, 1.0], (2, 1)),
             d=np.zeros((2, 2)), u=np.full((2, 2), np.inf))
    check_lp(S, O, k)


def test_config5_shape_small(S, O):
    """BASELINE config 5's recipe at a size the oracle finishes in seconds (N=120, M=4, J=26)."""
    k = S.workloads.config5(N=120, M=4, J=26, index=np.arange(6))
    X, St, status = check_lp(S, O, k)
    assert (status > 0).all()


def test_config5_full_size_against_highs(S):
    """One full-size config-5 LP (N=1000, M=20, J=180).  The reference algorithm needs ~1.3e5 simplex loops here (it
    switches to Bland's rule for good after N loops, src/Simplex.jl:487-490) and the reference-form oracle ~15 minutes, so
    the check at full size is size-independent: feasibility, status, and the optimal value against scipy's HiGHS."""
    from scipy.optimize import linprog
    k = S.workloads.config5(index=np.arange(1))
    X, St, status = S.SimplexLP_batch(k["A"], k["G"], k["c"], k["b"], k["g"], k["d"], k["u"])
    assert status[0] in (1, 2)
    x = X[0]
    assert np.abs(k["A"] @ x - k["b"][0]).max() < 1e-9 and (k["G"] @ x <= k["g"][0] + 1e-9).all()
    assert (x >= -1e-12).all() and (x <= 1 + 1e-12).all()
    lp = linprog(k["c"][0], A_ub=k["G"], b_ub=k["g"][0], A_eq=k["A"], b_eq=k["b"][0], bounds=[(0, 1)] * 1000, method="highs")
    assert abs(k["c"][0] @ x - lp.fun) < 1e-8 * abs(lp.fun)


def test_config5_full_size_goldens(S):
    """Eight full-size config-5 LPs against the committed oracle goldens (tests/golden/config5_full_lapack_8.npz: the
    oracle's SimplexLP in its LAPACK form, ~15 minutes per LP): status, the whole status vector S, the optimal value and x,
    and the number of simplex loops (1.1e5 - 1.8e5 per LP, 99 % of them under Bland's rule) — pivot for pivot."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config5_full_lapack_8.npz"))
    k = S.workloads.config5(index=g["index"])
    X, St, status = S.SimplexLP_batch(k["A"], k["G"], k["c"], k["b"], k["g"], k["d"], k["u"])
    stats = S.context().stats(len(status))
    assert np.array_equal(status, g["status"])
    assert np.array_equal(St, g["S"].astype(np.int32))
    obj = (k["c"] * X).sum(axis=1)
    assert np.all(np.abs(obj - g["obj"]) <= 1e-9 * np.abs(g["obj"]))
    assert np.abs(X - g["x"]).max() <= 1e-9
    assert np.array_equal(stats[:, 4].astype(np.int64), g["stats"][:, 0].astype(np.int64)), (stats[:, 4], g["stats"][:, 0])


def test_free_and_upper_only_variables_lp(S, O):
    """SimplexLP's free-variable split and (-Inf,u] negation (src/Simplex.jl:861-887, 996-1032): statuses, S and x against
    the oracle; the bounded cases also against HiGHS.  (With free variables the reference recomputes the status from the
    reduced costs even after an unbounded Phase 2, so an unbounded LP reports 1 or 2 — reproduced, not repaired.)"""
    from scipy.optimize import linprog
    for bounded in (True, False):
        w = S.workloads.general_bounds_lp(nb=6, N=30, M=4, J=14, seed=3, bounded=bounded)
        X, St, status = S.SimplexLP_batch(w["A"], w["G"], w["c"], w["b"], w["g"], w["d"], w["u"])
        for i in range(6):
            r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
            assert status[i] == r["status"], (bounded, i, status[i], r["status"])
            if bounded:
                assert status[i] in (1, 2)
                assert np.array_equal(St[i], r["S"])
                assert np.abs(X[i] - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max())
                res = linprog(w["c"][i], A_ub=w["G"], b_ub=w["g"][i], A_eq=w["A"], b_eq=w["b"][i],
                              bounds=list(zip(w["d"][i], w["u"][i])), method="highs")
                assert res.status == 0 and abs(w["c"][i] @ X[i] - res.fun) <= 1e-8 * max(1.0, abs(res.fun))
    # larger shape (512-thread kernel)
    w = S.workloads.general_bounds_lp(nb=3, N=320, M=6, J=40, seed=5)
    X, St, status = S.SimplexLP_batch(w["A"], w["G"], w["c"], w["b"], w["g"], w["d"], w["u"])
    for i in range(3):
        r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
        assert status[i] == r["status"] and status[i] in (1, 2) and np.array_equal(St[i], r["S"])
        assert np.abs(X[i] - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max())


def test_basic_artificials_are_driven_out_and_redundant_rows_purged(S, O):
    """SimplexLP's drive-out of artificial variables that stay basic after Phase 1 (src/Simplex.jl:962-977, on the device) and
    its redundancy purge (:889-902, host side of the C ABI): statuses, S and x against the oracle, optimum against HiGHS."""
    from scipy.optimize import linprog
    for kind, kw in (("zero_row", {}), ("dup_row", {}), ("dup_row_inconsistent", {}), ("zero_row", dict(N=320, M=5, J=40, seed=9))):
        w = S.workloads.degenerate_lps(kind, **kw)
        X, St, status = S.SimplexLP_batch(w["A"], w["G"], w["c"], w["b"], w["g"], w["d"], w["u"])
        for i in range(w["c"].shape[0]):
            r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
            assert status[i] == r["status"], (kind, i, status[i], r["status"])
            if kind == "dup_row_inconsistent" and i % 2 == 1:
                assert status[i] == -1
                continue
            assert status[i] in (1, 2) and np.array_equal(St[i], r["S"]), (kind, i)
            assert np.abs(X[i] - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max())
            res = linprog(w["c"][i], A_ub=w["G"], b_ub=w["g"][i], A_eq=w["A"], b_eq=w["b"][i],
                          bounds=list(zip(w["d"][i], w["u"][i])), method="highs")
            assert res.status == 0 and abs(w["c"][i] @ X[i] - res.fun) <= 1e-8 * max(1.0, abs(res.fun))


def test_other_pivot_rules(S, O):
    """Settings.rule = :stpEdgeLP / :maxImprovement (src/Simplex.jl:234-416, 641-813) on the device against the oracle's
    restatement of the same functions.  stpEdgeLP: statuses, S, x and simplex loop counts through SimplexLP, iteration counts
    through solveQP.  maxImprovement ranks the candidates by |h * step|, which at a degenerate vertex is a comparison of
    roundoff-sized numbers (the reference's own inv(lu) noise decides there), so the pivot path is not reproducible between
    two correct implementations: the optimum is compared (objective, feasibility, and S / x for the strictly convex QPs)."""
    try:
        for rule in ("stpEdgeLP", "maxImprovement"):
            O.set_rule(rule)
            st = S.Settings(rule=rule)
            exact = rule == "stpEdgeLP"
            for w in (S.workloads.general_bounds_lp(nb=4, N=30, M=4, J=14, seed=3), S.workloads.degenerate_lps("zero_row"),
                      S.workloads.general_bounds_lp(nb=2, N=320, M=6, J=40, seed=5),
                      S.workloads.general_bounds_lp(nb=2, N=80, M=10, J=160, seed=8)):      # M+J = 170: the basis inverse lives in L2
                X, St, status = S.SimplexLP_batch(w["A"], w["G"], w["c"], w["b"], w["g"], w["d"], w["u"], settings=st)
                stats = S.context().stats(len(status))
                for i in range(len(status)):
                    r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
                    assert status[i] in (1, 2) and r["status"] in (1, 2), (rule, i, status[i], r["status"])
                    f, fr = w["c"][i] @ X[i], w["c"][i] @ r["x"]
                    assert abs(f - fr) <= 1e-9 * max(1.0, abs(fr)), (rule, i, f, fr)
                    assert np.abs(w["A"] @ X[i] - w["b"][i]).max() < 1e-8 and (w["G"] @ X[i] - w["g"][i]).max() < 1e-8
                    assert (X[i] - w["u"][i]).max() <= 1e-12 and (w["d"][i] - X[i]).max() <= 1e-12
                    if exact:
                        assert status[i] == r["status"] and np.array_equal(St[i], r["S"]), (rule, i)
                        assert np.abs(X[i] - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max())
                        assert stats[i, 4] == r["stats"][0], (rule, i, stats[i, 4], r["stats"][0])      # simplex loops
            for c in (S.workloads.config2(nb=3, N=40), S.workloads.config4(nb=3, N=60, J=12)):
                X, St, status = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], settingsLP=st)
                r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
                assert (status > 0).all() and np.array_equal(St, r["S"]), (rule, status, r["status"])
                assert np.abs(X - r["x"]).max() <= 1e-9 * np.abs(r["x"]).max()
                if exact:
                    assert np.array_equal(status, r["status"]), (rule, status, r["status"])
    finally:
        O.set_rule("Dantzig")
