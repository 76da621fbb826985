"""MOI-level entry (moi.Optimizer.optimize / optimize_batch, mirror of src/MOIwrapper.jl:131-171) on the device, and the
array form of SimplexLP (src/Simplex.jl:1036-1196), checked against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import ssqp_b200
    if ssqp_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the GPU tests must not silently pass (no CPU fallback)")
    return ssqp_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ssqp_oracle
    return ssqp_oracle


def test_optimize_batch_routes_many_models_through_one_device_batch(S, O):
    c = S.workloads.config2(nb=12)
    opts = [S.Optimizer().load(c["V"], c["q"][i], c["A"], c["b"][i], c["G"], c["g"][i], c["d"][i], c["u"][i]) for i in range(12)]
    lp = S.workloads.general_bounds_lp(nb=3)
    opts += [S.Optimizer().load(np.zeros((30, 30)), lp["c"][i], lp["A"], lp["b"][i], lp["G"], lp["g"][i], lp["d"][i], lp["u"][i]) for i in range(3)]
    n0 = S.context().launch_count()
    S.optimize_batch(opts)
    assert S.context().launch_count() - n0 <= 12          # 2 groups (QP, LP): set_shared + Phase-1 de-dup + solve each — not 15 solves
    r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
    for i in range(12):
        x, St, st = opts[i].Results
        assert st == r["status"][i] and np.array_equal(St, r["S"][i])
        assert np.abs(x - r["x"][i]).max() <= 1e-9 * np.abs(r["x"][i]).max()
        assert opts[i].termination_status() == (S.moi.OPTIMAL if st in (1, 2) else S.moi.INFEASIBLE_OR_UNBOUNDED if st == 3 else S.moi.ITERATION_LIMIT)
        f = 0.5 * x @ c["V"] @ x + c["q"][i] @ x
        assert abs(opts[i].objective_value() - f) <= 1e-15 + 1e-12 * abs(f)
    O.set_fix_flip(True)
    try:
        for i in range(3):
            ro = O.simplex_lp(lp["c"][i], lp["A"], lp["G"], lp["b"][i], lp["g"][i], lp["d"][i], lp["u"][i])
            x, St, st = opts[12 + i].Results
            assert st == ro["status"] and opts[12 + i].termination_status() == S.moi.OPTIMAL
            assert abs(lp["c"][i] @ x - lp["c"][i] @ ro["x"]) <= 1e-9 * max(1.0, abs(lp["c"][i] @ ro["x"]))
    finally:
        O.set_fix_flip(False)
    one = S.Optimizer().load(c["V"], c["q"][3], c["A"], c["b"][3], c["G"], c["g"][3], c["d"][3], c["u"][3])
    one.optimize()
    assert one.Results[2] == opts[3].Results[2] and np.array_equal(one.Results[0], opts[3].Results[0])


def test_simplexlp_array_form(S, O):
    """SimplexLP(c, A, b, d, u) (src/Simplex.jl:1036): equality rows only; same result as the struct form with J = 0."""
    rng = np.random.default_rng(21)
    N, Mr = 20, 5
    A = rng.standard_normal((Mr, N))
    xs = rng.uniform(0.1, 0.9, N)
    b = A @ xs
    cvec = rng.standard_normal(N)
    x, St, st = S.SimplexLP(cvec, A, b, np.zeros(N), np.ones(N))
    ro = O.simplex_lp(cvec, A, np.zeros((0, N)), b, np.zeros(0), np.zeros(N), np.ones(N))
    assert st == ro["status"] and St.shape == (N,) and np.array_equal(St, ro["S"])
    assert np.abs(x - ro["x"]).max() <= 1e-9
    xm, Sm, stm = S.SimplexLP(cvec, A, b, np.zeros(N), np.ones(N), min=False)
    assert stm in (1, 2) and cvec @ xm >= cvec @ x
