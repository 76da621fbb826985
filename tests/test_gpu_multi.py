"""Multi-GPU tests (need >= 2 visible CUDA devices; skipped otherwise): in-process sharding by QP index inside the
C ABI (one host thread + stream per device) must return exactly what one device returns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_device_context_matches_single_device():
    import ssqp_b200 as S
    if S.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    c = S.workloads.config4(index=np.arange(0, 65536, 2048), total=65536)        # 32 QPs
    one = S.Context([0])
    one.set_shared(c["V"], c["A"], c["G"])
    X1, S1, st1 = one.solve_batch(c["q"], c["b"], c["g"], c["d"], c["u"])
    two = S.Context([0, 1])
    two.set_shared(c["V"], c["A"], c["G"])
    X2, S2, st2 = two.solve_batch(c["q"], c["b"], c["g"], c["d"], c["u"])
    assert np.array_equal(st1, st2) and np.array_equal(S1, S2) and np.array_equal(X1, X2)
    stats = two.stats(32)
    assert (stats[:, 0] == st2).all()          # per-QP stats gathered back in QP order
    one.close(); two.close()


def test_two_device_sweep_and_general_bounds_match_single_device():
    """Chains of a warm-started sweep are sharded whole (chain c -> device c mod G); batches with free / (-Inf,u] variables
    and LPs take the same staging.  Two devices must return what one returns, bit for bit."""
    import ssqp_b200 as S
    if S.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    c = S.workloads.config4(nb=1, N=200, J=30)
    nb, L = 30, 5                                   # 6 chains -> 3 per device
    q = -np.logspace(-2, 0, nb)[:, None] * c["E"][None, :]
    til = lambda a: np.tile(a[0], (nb, 1))
    b, g, d, u = til(c["b"]), til(c["g"]), til(c["d"]), til(c["u"])
    res = []
    for devs in ([0], [0, 1]):
        ctx = S.Context(devs)
        ctx.set_shared(c["V"], c["A"], c["G"])
        X, St, st = ctx.solve_sweep(q, b, g, d, u, L)
        stats = ctx.stats(nb)
        assert (stats[:, 0] == st).all()
        w = S.workloads.general_bounds(nb=7, N=40, M=3, J=12, seed=11)
        ctx.set_shared(w["V"], w["A"], w["G"])
        Xg, Sg, sg = ctx.solve_batch(w["q"], w["b"], w["g"], w["d"], w["u"])
        lp = S.workloads.general_bounds_lp(nb=5, N=30, M=4, J=14, seed=3)
        ctx.set_shared(None, lp["A"], lp["G"])
        Xl, Sl, sl = ctx.solve_lp_batch(lp["c"], lp["b"], lp["g"], lp["d"], lp["u"])
        res.append((X, St, st, Xg, Sg, sg, Xl, Sl, sl))
        ctx.close()
    for a, b2 in zip(res[0], res[1]):
        assert np.array_equal(a, b2)
    assert (res[0][2] > 0).all() and (res[0][5] > 0).all() and np.isin(res[0][8], (1, 2)).all()
