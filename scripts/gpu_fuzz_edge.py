"""Randomised parity log of edge cases around the hot path (a log, not a test; nothing asserted, the last line is the tally):
QPs made awkward on purpose — (a) infeasible (an inequality pushed below what the equalities allow), (b) redundant rows
(duplicated equality rows; inequality rows that repeat an equality), (c) fixed variables (d == u), (d) per-QP V
(`V_per_qp`), (e) Settings.rule = :maxImprovement — each against the oracle on the same inputs.
`python scripts/gpu_fuzz_edge.py [seconds] [first_seed]`"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O

W = S.workloads
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
rng = np.random.default_rng(seed0)
t_end = time.time() + budget
tally = {}


def compare(kind, tag, X, St, status, r):
    n, bad = tally.get(kind, (0, 0))
    for i in range(len(status)):
        n += 1
        both = status[i] > 0 and r["status"][i] > 0
        dx = np.abs(X[i] - r["x"][i]).max() / max(np.abs(r["x"][i]).max(), 1e-300) if both else 0.0
        sameS = np.array_equal(St[i], r["S"][i]) if both else True
        if status[i] != r["status"][i] or not sameS or dx > 1e-9:
            bad += 1
            print("%s %s qp %d: status gpu %d cpu %d, same S %s, rel dx %.1e" % (kind, tag, i, status[i], r["status"][i], sameS, dx), flush=True)
    tally[kind] = (n, bad)


O.set_fix_flip(True)
while time.time() < t_end:
    N = int(rng.integers(6, 140)); M = int(rng.integers(1, min(6, N // 2) + 1)); J = int(rng.integers(1, max(2, N // 3)))
    seed = int(rng.integers(1, 1 << 30))
    kind = ["INFEASIBLE", "REDUNDANT", "FIXED", "PERQPV", "MAXIMPR"][int(rng.integers(0, 5))]
    tag = "N=%d M=%d J=%d seed=%d" % (N, M, J, seed)
    try:
        c = W.general_bounds(nb=4, N=N, M=M, J=J, seed=seed)
        V, A, G, q, b, g, d, u = (c[k] for k in "VAGqbgdu")
        kw = {}
        if kind == "INFEASIBLE":       # G[0] := A[0] with g below b for half of the QPs: A0 x = b0 and A0 x <= b0 - 1
            G = G.copy(); g = g.copy()
            G[0] = A[0]; g[:, 0] = b[:, 0] + np.where(np.arange(4) % 2 == 0, -1.0, 0.25)
        elif kind == "REDUNDANT":      # a duplicated equality row and an inequality that repeats an equality (always tight)
            A = np.vstack([A, A[0]]); b = np.hstack([b, b[:, :1]])
            G = G.copy(); g = g.copy(); G[0] = A[0]; g[:, 0] = b[:, 0]
        elif kind == "FIXED":          # some variables fixed: d == u (inside the old bounds where they were finite)
            d = d.copy(); u = u.copy()
            fix = rng.random((4, N)) < 0.15
            val = rng.uniform(-0.5, 0.5, (4, N))
            d[fix] = val[fix]; u[fix] = val[fix]
        elif kind == "PERQPV":
            Vs = np.empty((4, N, N))
            for t in range(4):
                Bm = np.random.default_rng(seed + t).standard_normal((N, N))
                Vs[t] = Bm @ Bm.T / N + 0.1 * np.eye(N)
                Vs[t] = (Vs[t] + Vs[t].T) / 2
            V = Vs
        elif kind == "MAXIMPR":
            kw = dict(settingsLP=S.Settings(rule="maxImprovement"))
            O.set_rule("maxImprovement")
        try:
            X, St, status = S.solveQP_batch(V, A, G, q, b, g, d, u, **kw)
            r = O.solve_batch(V, A, G, q, b, g, d, u)
        finally:
            O.set_rule("Dantzig")
        compare(kind, tag, X, St, status, r)
    except Exception as e:      # noqa: BLE001 — a log: keep going
        n, bad = tally.get(kind, (0, 0)); tally[kind] = (n, bad + 1)
        print("%s %s: exception %r" % (kind, tag, e), flush=True)
O.set_fix_flip(False)
print("edge fuzz: " + ", ".join("%s %d QPs (%d differences)" % (k, v[0], v[1]) for k, v in sorted(tally.items())) + " | seed0 %d, %.0f s" % (seed0, budget), flush=True)
