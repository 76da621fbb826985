"""The MOI wrapper's optimize! / status layer (src/MOIwrapper.jl:131-240) mirrored in statusswitchingqp.jl_b200/moi.py.
CPU part: everything that never reaches the solver (presolve branches, status mapping, LP detection).  GPU part
(tests/test_gpu_moi.py) runs optimize / optimize_batch through the C ABI."""
import warnings
import numpy as np
import pytest
import ssqp_b200 as S

M = S.moi


def test_termination_status_mapping_is_the_references():
    o = S.Optimizer()
    assert o.termination_status() == M.OPTIMIZE_NOT_CALLED and o.result_count() == 0
    x = np.zeros(2)
    table = {1: M.OPTIMAL, 2: M.OPTIMAL, 3: M.INFEASIBLE_OR_UNBOUNDED, 0: M.INFEASIBLE, -1: M.NUMERICAL_ERROR,
             -7778: M.ITERATION_LIMIT, 57: M.ITERATION_LIMIT}       # (a QP's trip count > 3 falls through, as in the reference)
    for st, want in table.items():
        o.Results = (x, np.zeros(2, np.int32), st)
        assert o.termination_status() == want
        assert o.primal_status() == (M.INFEASIBLE_POINT if st == 0 else M.FEASIBLE_POINT)
        assert o.raw_status_string() == str(st)
    assert o.primal_status(2) == M.NO_SOLUTION and o.dual_status(2) == M.NO_SOLUTION and o.dual_status() == M.FEASIBLE_POINT


def test_copy_to_turns_a_model_without_V_into_an_LP_and_negates_for_max():
    N = 4
    o = S.Optimizer()
    o.load(np.zeros((N, N)), np.arange(N, dtype=float), np.ones((1, N)), [1.0], np.zeros((0, N)), [], np.zeros(N), np.ones(N), sense=M.MAX_SENSE, f0=2.5)
    assert isinstance(o.Problem, S.LP)
    assert np.array_equal(o.Problem.c, -np.arange(N, dtype=float)) and o.f0 == 2.5
    o.Results = (np.array([0, 0, 0, 1.0]), None, 1)
    assert o.objective_value() == 3.0 + 2.5           # -(c'x) + f0 with c negated
    o2 = S.Optimizer(maxIter=50)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o2.load(np.eye(N), np.zeros(N), np.ones((1, N)), [1.0], np.zeros((0, N)), [], np.zeros(N), np.ones(N))
    assert isinstance(o2.Problem, S.QP) and o2.Settings.maxIter == 50 and not o2.is_empty()
    o2.empty()
    assert o2.is_empty() and o2.Sense == M.MIN_SENSE


def test_presolve_branches_never_reach_the_solver():
    """mc == -20 (no inequalities, no bounds): answered on the host (src/MOIwrapper.jl:133-158) — no CUDA device needed."""
    N = 3
    free_d, free_u = np.full(N, -np.inf), np.full(N, np.inf)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = S.Optimizer().load(np.zeros((N, N)), np.ones(N), np.eye(N), [1.0, 2.0, 3.0], np.zeros((0, N)), [], free_d, free_u)     # LP, square A
        b = S.Optimizer().load(np.diag([1.0, 2.0, 4.0]), np.ones(N), np.zeros((0, N)), [], np.zeros((0, N)), [], free_d, free_u)    # QP, no rows
        c = S.Optimizer().load(np.zeros((N, N)), np.zeros(N), np.zeros((0, N)), [], np.zeros((0, N)), [], free_d, free_u)           # f == 0
    S.optimize_batch([a, b, c])
    assert a.Results[2] == 1 and np.allclose(a.Results[0], [1, 2, 3])
    assert b.Results[2] == 1 and np.allclose(b.Results[0], [1.0, 0.5, 0.25])          # the reference's  V \ q
    assert c.Results[2] == 1 and np.array_equal(c.Results[0], np.zeros(N))
    assert all(o.termination_status() == M.OPTIMAL for o in (a, b, c))
