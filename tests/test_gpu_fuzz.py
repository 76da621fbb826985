"""Randomised shapes through every entry point added around the hot path: QPs and LPs with mixed bound kinds (box, free,
upper-only, lower-only), with and without equality rows, odd sizes (general kernel flavour) and multiples of 4 (256-bit
flavour), against the oracle.  Fixed seeds: a failure is reproducible."""
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


SHAPES = [(5, 1, 0), (8, 0, 3), (13, 2, 5), (16, 4, 8), (24, 1, 12), (31, 3, 9), (40, 4, 16), (57, 2, 11), (64, 4, 28), (100, 3, 25)]


def test_qp_shapes_with_general_bounds(S, O):
    O.set_fix_flip(True)
    try:
        for t, (N, M, J) in enumerate(SHAPES):
            c = S.workloads.general_bounds(nb=6, N=N, M=M, J=J, seed=100 + t)
            X, St, status = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
            r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
            assert np.array_equal(status, r["status"]), (N, M, J, status, r["status"])
            ok = status > 0
            assert np.array_equal(St[ok], r["S"][ok]), (N, M, J)
            if ok.any():
                scale = np.maximum(np.abs(r["x"][ok]).max(axis=1), 1e-300)
                assert (np.abs(X[ok] - r["x"][ok]).max(axis=1) / scale).max() < 1e-9, (N, M, J)
    finally:
        O.set_fix_flip(False)


def test_lp_shapes_with_general_bounds(S, O):
    for t, (N, M, J) in enumerate(SHAPES):
        if M + J == 0:
            continue
        w = S.workloads.general_bounds_lp(nb=6, N=N, M=M, J=J, seed=200 + t, bounded=True)
        X, St, status = S.SimplexLP_batch(w["A"], w["G"], w["c"], w["b"], w["g"], w["d"], w["u"])
        for i in range(6):
            r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
            assert status[i] == r["status"], (N, M, J, i, status[i], r["status"])
            if status[i] in (1, 2):
                assert np.array_equal(St[i], r["S"]), (N, M, J, i)
                assert np.abs(X[i] - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max()), (N, M, J, i)


def test_steepest_edge_rule_on_random_shapes(S, O):
    """Settings.rule = :stpEdgeLP through initQP + solveQP and through SimplexLP on the same shapes: statuses (iteration
    counts), S and x against the oracle's restatement of stpEdgeLP (src/Simplex.jl:234-416)."""
    O.set_fix_flip(True)
    O.set_rule("stpEdgeLP")
    st = S.Settings(rule="stpEdgeLP")
    try:
        for t, (N, M, J) in enumerate(SHAPES[2:8]):
            c = S.workloads.general_bounds(nb=4, N=N, M=M, J=J, seed=300 + t)
            X, St, status = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], settingsLP=st)
            r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
            assert np.array_equal(status, r["status"]) and np.array_equal(St, r["S"]), (N, M, J, status, r["status"])
            assert np.abs(X - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max())
            w = S.workloads.general_bounds_lp(nb=4, N=N, M=M, J=J, seed=400 + t, bounded=True)
            Xl, Sl, sl = S.SimplexLP_batch(w["A"], w["G"], w["c"], w["b"], w["g"], w["d"], w["u"], settings=st)
            for i in range(4):
                rl = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
                assert sl[i] == rl["status"] and np.array_equal(Sl[i], rl["S"]), (N, M, J, i, sl[i], rl["status"])
                assert np.abs(Xl[i] - rl["x"]).max() <= 1e-9 * max(1.0, np.abs(rl["x"]).max())
    finally:
        O.set_rule("Dantzig")
        O.set_fix_flip(False)


# The differences a wide fuzz run found (scripts/gpu_fuzz_big.py, profiles/r02_v5_fuzz_big.log: 8 of 43 428 problems), pinned
# with their triage (scripts/gpu_fuzz_replay.py): every one is a pivoting tie broken differently, never another answer.
FUZZ_QPS = [(66, 3, 30, 1063817686, 2), (149, 4, 70, 26187268, 1), (154, 6, 67, 360929995, 0), (113, 4, 55, 725792724, 4)]
FUZZ_LPS = [(145, 1, 70, 890721998, 2), (161, 5, 72, 929700198, 5)]


def test_fuzz_found_qps_differ_by_the_phase1_vertex_only(S, O):
    """Four QPs whose cold-start trip count differs from the oracle's: Phase 1 (cDantzigLP, src/Simplex.jl:499-569) ends on
    another vertex (a ratio-test tie decided by 1e-17 roundoff of q = invB*b - Y*x[F]); the optimum — S and x — is the
    same, and from the oracle's own Phase-1 vertex the device needs exactly the oracle's number of trips."""
    O.set_fix_flip(True)
    try:
        for N, M, J, seed, i in FUZZ_QPS:
            c = S.workloads.general_bounds(nb=6, N=N, M=M, J=J, seed=seed)
            one = lambda a: a[i:i + 1]
            args = (c["V"], c["A"], c["G"], one(c["q"]), one(c["b"]), one(c["g"]), one(c["d"]), one(c["u"]))
            X, St, status = S.solveQP_batch(*args)
            r = O.solve_qp(c["V"], c["A"], c["G"], c["q"][i], c["b"][i], c["g"][i], c["d"][i], c["u"][i])
            assert status[0] > 0 and r["status"] > 0
            assert np.array_equal(St[0], r["S"]), (N, M, J, seed)
            assert np.abs(X[0] - r["x"]).max() <= 1e-9 * np.abs(r["x"]).max(), (N, M, J, seed)
            xo, So, sto, _ = O.init_qp(c["A"], c["G"], c["b"][i], c["g"][i], c["d"][i], c["u"][i])
            Xw, Sw, stw = S.solveQP_batch(*args, S0=So[None].astype(np.int32), x0=xo[None])
            rw = O.solve_qp(c["V"], c["A"], c["G"], c["q"][i], c["b"][i], c["g"][i], c["d"][i], c["u"][i], S0=So, x0=xo)
            assert stw[0] == rw["status"], (N, M, J, seed, stw[0], rw["status"])
            assert np.array_equal(Sw[0], rw["S"])
    finally:
        O.set_fix_flip(False)


def test_fuzz_found_lps_sit_on_a_face_of_optima(S, O):
    """Two LPs on which device and oracle stop at different vertices: both report status 2 ("infinitely many solutions",
    src/Simplex.jl:31) — the cost is a combination of the rows, the optimum is a face — with the same objective and a
    feasible point each."""
    for N, M, J, seed, i in FUZZ_LPS:
        w = S.workloads.general_bounds_lp(nb=6, N=N, M=M, J=J, seed=seed, bounded=True)
        one = lambda a: a[i:i + 1]
        X, St, status = S.SimplexLP_batch(w["A"], w["G"], one(w["c"]), one(w["b"]), one(w["g"]), one(w["d"]), one(w["u"]))
        r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
        assert status[0] == 2 and r["status"] == 2
        fo, fg = w["c"][i] @ r["x"], w["c"][i] @ X[0]
        assert abs(fg - fo) <= 1e-9 * max(1.0, abs(fo)), (N, M, J, seed, fg, fo)
        x = X[0]
        viol = max(np.abs(w["A"] @ x - w["b"][i]).max(), (w["G"] @ x - w["g"][i]).max(), (w["d"][i] - x).max(), (x - w["u"][i]).max())
        assert viol <= 1e-9, (N, M, J, seed, viol)
