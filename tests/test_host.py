"""CPU-only tests of the host logic and the C-ABI boundary (no compute calls without a GPU)."""
import ctypes
import os
import re
import warnings
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def S():
    import ssqp_b200
    ssqp_b200.build()
    return ssqp_b200


def test_library_exports_every_declared_symbol(S):
    hdr = open(os.path.join(ROOT, "include", "ssqp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ssqp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(S.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(S.EXPORTS)
    S.load()
    assert "sm_100a" in S.version()


def test_default_settings_match_reference(S):
    """src/types.jl:401-408: maxIter=7777, tol=2^-26, tolG=2^-33."""
    cs = S.CSettings()
    S.load().ssqp_default_settings(ctypes.byref(cs))
    assert (cs.max_iter, cs.tol, cs.tolG, cs.rule) == (7777, 2.0 ** -26, 2.0 ** -33, 0)
    py = S.Settings().to_c()
    assert (py.max_iter, py.tol, py.tolG, py.rule) == (7777, 2.0 ** -26, 2.0 ** -33, 0)


def test_status_codes(S):
    assert [int(s) for s in (S.IN, S.DN, S.UP, S.OE, S.EO)] == [0, 1, 2, 3, 4]     # src/types.jl:17-23


def test_qp_constructor_semantics(S):
    V = np.array([[2.0, 1.0], [0.0, 2.0]])
    Q = S.QP(V)
    assert np.array_equal(Q.V, [[2.0, 0.5], [0.5, 2.0]])                        # symmetrised, types.jl:243
    assert (Q.M, Q.J, Q.mc) == (1, 0, 1) and np.array_equal(Q.A, np.ones((1, 2))) and np.array_equal(Q.b, [1.0])
    assert np.array_equal(Q.d, [0, 0]) and np.all(np.isinf(Q.u))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert S.QP(-np.eye(2)).mc == -70                                         # not PSD
        assert S.QP(np.eye(2), d=[0, 1.0], u=[1.0, 1.0]).mc == -30                # d == u
        assert S.QP(np.eye(2), d=[-np.inf] * 2, u=[np.inf] * 2).mc == -20         # no bounds, no inequalities
        Q = S.QP(np.eye(2), d=[1.0, 0.0], u=[0.0, 1.0])                           # swap u<d
        assert np.array_equal(Q.d, [0.0, 0.0]) and np.array_equal(Q.u, [1.0, 1.0])
    with pytest.raises(ValueError):
        S.QP(np.eye(2), A=np.ones((1, 3)))
    P = S.QP(np.eye(3), u=[1.0, 1.0, 1.0])
    E = np.array([0.1, 0.2, 0.3])
    QL = S.QP.with_L(P, E, 0.5)
    assert np.array_equal(QL.q, -0.5 * E) and QL.V is P.V and QL.A is P.A            # types.jl:303-319
    Qm = S.QP.with_mu(P, 0.2, E)
    assert Qm.M == 2 and np.array_equal(Qm.A[1], E) and np.array_equal(Qm.b, [1.0, 0.2]) and not Qm.q.any()


def test_invalid_qp_returns_reference_early_exit_without_gpu(S):
    """mc <= 0 -> (zeros(N), fill(DN,N), -1) produced on the host (src/SSQP.jl:226-228)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Q = S.QP(-np.eye(3))
    z, St, status = S.solveQP(Q)
    assert status == -1 and not z.any() and list(St) == [S.DN] * 3


def test_no_cpu_fallback(S):
    if S.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(S.SsqpError):
        S.Context()
    with pytest.raises(S.SsqpError):
        S.solveQP(S.QP(np.eye(2), u=[1.0, 1.0]))


def test_workloads_are_shard_invariant(S):
    W = S.workloads
    a = W.config4(nb=4, total=64)
    b = W.config4(index=np.array([1, 3]), total=64)
    assert np.array_equal(a["q"][[1, 3]], b["q"]) and np.array_equal(a["g"][[1, 3]], b["g"])
    full = np.arange(64)
    parts = [full[r::4] for r in range(4)]
    assert sorted(np.concatenate(parts).tolist()) == full.tolist()


def test_host_row_purge_matches_reference_form(S):
    """getRowsGJr / SimplexLP's redundancy purge on the host side of the C ABI (src/utils.jl:49-86, src/Simplex.jl:889-902)
    against the oracle's restatement of the same function."""
    from oracle import ssqp_oracle as O
    rng = np.random.default_rng(3)
    for t in range(20):
        nr, nc = int(rng.integers(2, 9)), int(rng.integers(2, 12))
        X = rng.standard_normal((nr, nc))
        if t % 2 and nr > 2:
            X[-1] = X[0] - 2 * X[1]                   # a dependent row
        rows, l1 = S.solver.getRowsGJr(X, 2.0 ** -33)
        ro, lo = O.get_rows_gjr(X, 2.0 ** -33)
        assert list(rows) == [int(v) for v in ro] and l1 == lo
    w = S.workloads.degenerate_lps("dup_row")
    rows, stat = S.solver._lp_row_purge(w["A"], w["G"], w["b"][0], w["g"][0], w["d"][0], w["u"][0], 2.0 ** -26)
    assert stat is None and sorted(set(range(12)) - set(rows)) == [3]
    w = S.workloads.degenerate_lps("dup_row_inconsistent")
    assert S.solver._lp_row_purge(w["A"], w["G"], w["b"][1], w["g"][1], w["d"][1], w["u"][1], 2.0 ** -26) == (None, -1)
    w = S.workloads.degenerate_lps("zero_row")
    assert S.solver._lp_row_purge(w["A"], w["G"], w["b"][0], w["g"][0], w["d"][0], w["u"][0], 2.0 ** -26) == (None, None)


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): LAPACK-form oracle on a bounded sample,
    one JSON line with the contract's keys — no GPU, no /root/reference needed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "8"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "QPs/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["scaling"] == "strong" and line["config"]["global_batch"] == 65536 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["form"] == "port-lapack" and cb["cores"] >= 1 and "OpenBLAS" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "QPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
