"""Host-side mirror of the reference's MathOptInterface wrapper around the hot path (src/MOIwrapper.jl).

The reference's `Optimizer{T}` stores one model (`Problem::Union{QP,LP}`), and `MOI.optimize!` (:131-171) calls the scalar
`solveQP` (:165) or `SimplexLP` (:167); the status triple is then read back through `TerminationStatus` (:213-228),
`PrimalStatus` (:195-207), `ObjectiveValue` (:231-240) and `VariablePrimal`.  This module restates exactly that layer —
same field names, same presolve branches for `mc == -20` models, same (bug-compatible) status mapping — on top of the
device path, and adds the batch entry the reference lacks: `optimize_batch(opts)` solves every optimizer of a list in
ONE device batch per group of models that share (V, A, G) — what a JuMP user who builds many models over one
covariance matrix reaches the GPU through.  The Julia source of the same hook is julia/SSQPB200.jl (`optimize_batch!`,
`use_device!`); the MOI model <-> QP conversion (`MOI2QP`, :409-445) needs MathOptInterface itself and stays in Julia.
"""
import time
import numpy as np

from .types import QP, LP, Settings, DN
from . import solver

MIN_SENSE, MAX_SENSE = "MIN_SENSE", "MAX_SENSE"
# MOI.TerminationStatusCode / ResultStatusCode names used by the wrapper
OPTIMIZE_NOT_CALLED, OPTIMAL, INFEASIBLE, INFEASIBLE_OR_UNBOUNDED, NUMERICAL_ERROR, ITERATION_LIMIT = (
    "OPTIMIZE_NOT_CALLED", "OPTIMAL", "INFEASIBLE", "INFEASIBLE_OR_UNBOUNDED", "NUMERICAL_ERROR", "ITERATION_LIMIT")
NO_SOLUTION, FEASIBLE_POINT, INFEASIBLE_POINT = "NO_SOLUTION", "FEASIBLE_POINT", "INFEASIBLE_POINT"


class Optimizer:
    """mutable struct Optimizer{Float64} (src/MOIwrapper.jl:7-33): Problem, Settings, Results, Sense, Silent, f0, solTime."""

    def __init__(self, **user_settings):
        self.Problem = None
        self.Settings = Settings(**user_settings) if user_settings else Settings()
        self.Results = None
        self.Sense = MIN_SENSE
        self.Silent = True
        self.f0 = 0.0
        self.solTime = 0.01

    # MOI.empty! / MOI.is_empty (:43-51)
    def empty(self):
        self.Problem = None
        self.Results = None
        self.Sense = MIN_SENSE

    def is_empty(self):
        return self.Problem is None

    def load(self, V, q, A, b, G, g, d, u, sense=MIN_SENSE, f0=0.0):
        """What MOI.copy_to leaves in the optimizer (:119-128) once MOI2QP (:409-445) has produced V, q, A, b, G, g, d, u:
        MAX_SENSE negates V and q, and a model whose V vanishes (norm(V, Inf) == 0) becomes an LP."""
        self.Sense = sense
        self.f0 = float(f0)
        V = np.array(V, dtype=np.float64)
        q = np.array(q, dtype=np.float64).ravel()
        if sense == MAX_SENSE:
            V, q = -V, -q
        Q = QP(V, q=q, A=A, b=b, G=G, g=g, d=d, u=u)
        if np.abs(Q.V).sum(axis=1).max(initial=0.0) == 0:
            P = LP(Q.q, Q.A, Q.b, G=Q.G, g=Q.g, d=Q.d, u=Q.u)
            self.Problem = P
        else:
            self.Problem = Q
        self.Results = None
        return self

    # ---- MOI.optimize! (:131-171) ----------------------------------------------------------------------------
    def _presolve(self):
        """The `P.mc == -20` branch (:133-158): models without inequalities and bounds never reach solveQP / SimplexLP."""
        P = self.Problem
        N = P.N
        if P.M > 0:
            x = np.linalg.lstsq(P.A, P.b, rcond=None)[0] if P.A.shape[0] != P.A.shape[1] else np.linalg.solve(P.A, P.b)
            return x, np.full(N, int(DN), dtype=np.int32), 1
        if isinstance(P, QP):
            x = np.linalg.solve(P.V, P.q)                    # (the reference's `P.V \\ P.q`, sign and all)
            det = np.linalg.det(P.V)
            st = 1 if ((self.Sense == MIN_SENSE and det > 0) or (self.Sense == MAX_SENSE and det < 0)) else 3
            return x, np.full(N, int(DN), dtype=np.int32), st
        st = 1 if np.abs(P.c).max(initial=0.0) == 0 else 3
        return np.zeros(N), np.full(N, int(DN), dtype=np.int32), st

    def optimize(self, ctx=None):
        """MOI.optimize!(opt): one model through the device path (a batch of one)."""
        optimize_batch([self], ctx=ctx)

    # ---- result attributes -----------------------------------------------------------------------------------
    def result_count(self):
        return int(self.Results is not None)

    def termination_status(self):
        """MOI.TerminationStatus (:213-228), bug-compatible: the triple's third entry is read as SimplexLP's code even for
        a QP, so a QP that took exactly 3 trips reports INFEASIBLE_OR_UNBOUNDED and any iteration count above 3 falls
        through to ITERATION_LIMIT, as in the reference."""
        if self.Results is None:
            return OPTIMIZE_NOT_CALLED
        st = self.Results[2]
        if st == 3:
            return INFEASIBLE_OR_UNBOUNDED
        if st in (1, 2):
            return OPTIMAL
        if st == 0:
            return INFEASIBLE
        if st == -1:
            return NUMERICAL_ERROR
        return ITERATION_LIMIT

    def primal_status(self, result_index=1):
        if result_index != 1 or self.Results is None:
            return NO_SOLUTION
        return INFEASIBLE_POINT if self.Results[2] == 0 else FEASIBLE_POINT

    def dual_status(self, result_index=1):
        return FEASIBLE_POINT if result_index == 1 else NO_SOLUTION

    def raw_status_string(self):
        return str(self.Results[2])

    def objective_value(self):
        x = self.Results[0]
        P = self.Problem
        f = x @ (P.V @ x) / 2 + P.q @ x if isinstance(P, QP) else x @ P.c
        return (f if self.Sense == MIN_SENSE else -f) + self.f0

    def variable_primal(self, i=None):
        return self.Results[0] if i is None else self.Results[0][i]


def _same(a, b):
    return a is b or (a.shape == b.shape and np.array_equal(a, b))


def optimize_batch(opts, ctx=None):
    """`MOI.optimize!` for a list of optimizers: presolved models (mc == -20) are answered on the host like the reference
    does, the others are grouped by problem kind and shared (V, A, G) and each group is one device batch
    (ssqp_solve_batch / ssqp_solve_lp_batch).  Each optimizer's Results / solTime are filled as `optimize!` would."""
    t0 = time.time()
    groups = []
    for o in opts:
        P = o.Problem
        if P is None:
            raise ValueError("optimize! on an empty optimizer")
        if P.mc == -20:
            o.Results = o._presolve()
            continue
        for key, members in groups:
            K = key.Problem
            if type(K) is type(P) and (K.N, K.M, K.J) == (P.N, P.M, P.J) and _same(K.A, P.A) and _same(K.G, P.G) \
                    and (not isinstance(P, QP) or _same(K.V, P.V)) and _settings_equal(key.Settings, o.Settings):
                members.append(o)
                break
        else:
            groups.append((o, [o]))
    for key, members in groups:
        if isinstance(key.Problem, QP):
            res = solver.solveQP([m.Problem for m in members], settings=key.Settings, ctx=ctx)
        else:
            res = solver.SimplexLP([m.Problem for m in members], settings=key.Settings, ctx=ctx)
        for m, r in zip(members, res):
            m.Results = (r[0], r[1], int(r[2]))
    dt = time.time() - t0
    for o in opts:
        o.solTime = dt
    return opts


def _settings_equal(a, b):
    return (a.maxIter, a.tol, a.tolG, a.rule, a.pivot) == (b.maxIter, b.tol, b.tolG, b.rule, b.pivot)
