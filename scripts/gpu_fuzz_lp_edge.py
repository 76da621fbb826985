"""Randomised parity log of LP edge cases through SimplexLP_batch (a log, not a test; the last line is the tally): (a) infeasible
LPs (status 0), (b) unbounded LPs WITHOUT free variables (status 3: the reference only overwrites it when free variables
exist, src/Simplex.jl:1001-1019), (c) redundant rows (rank-deficient [A 0; G I] cannot happen — a duplicated equality row makes
A0 rank-deficient: the purge of src/Simplex.jl:889-902, host side of the ABI), (d) fixed variables (d == u).  Status compared
always; S and x when both say 1; the objective when both say 2.  `python scripts/gpu_fuzz_lp_edge.py [seconds] [first_seed]`"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O

W = S.workloads
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
rng = np.random.default_rng(seed0)
t_end = time.time() + budget
tally = {}
while time.time() < t_end:
    N = int(rng.integers(5, 110)); M = int(rng.integers(1, min(5, N // 2) + 1)); J = int(rng.integers(1, max(2, N // 3)))
    seed = int(rng.integers(1, 1 << 30))
    kind = ["LP_INFEASIBLE", "LP_UNBOUNDED", "LP_REDUNDANT", "LP_FIXED"][int(rng.integers(0, 4))]
    tag = "N=%d M=%d J=%d seed=%d" % (N, M, J, seed)
    n, bad, hist = tally.get(kind, (0, 0, {}))
    try:
        w = W.general_bounds_lp(nb=4, N=N, M=M, J=J, seed=seed, bounded=(kind != "LP_UNBOUNDED"))
        A, G, c, b, g, d, u = (w[k] for k in "AGcbgdu")
        if kind == "LP_INFEASIBLE":
            G = G.copy(); g = g.copy(); G[0] = A[0]; g[:, 0] = b[:, 0] + np.where(np.arange(4) % 2 == 0, -1.0, 0.25)
        elif kind == "LP_UNBOUNDED":        # no free variables: the free ones become lower-only; random cost -> mostly unbounded
            d = d.copy(); fv = np.isinf(d) & np.isinf(u); d[fv] = -1.5
        elif kind == "LP_REDUNDANT":
            A = np.vstack([A, A[0]]); b = np.hstack([b, b[:, :1]])
        elif kind == "LP_FIXED":
            d = d.copy(); u = u.copy(); fix = rng.random((4, N)) < 0.15; val = rng.uniform(-0.5, 0.5, (4, N)); d[fix] = val[fix]; u[fix] = val[fix]
        X, St, status = S.SimplexLP_batch(A, G, c, b, g, d, u)
        for i in range(4):
            r = O.simplex_lp(c[i], A, G, b[i], g[i], d[i], u[i])
            n += 1
            hist[int(r["status"])] = hist.get(int(r["status"]), 0) + 1
            free = int((np.isinf(d[i]) & np.isinf(u[i])).sum())
            if status[i] != r["status"]:
                bad += 1; print("%s %s lp %d: status gpu %d cpu %d (free variables %d)" % (kind, tag, i, status[i], r["status"], free), flush=True)
            elif status[i] in (1, 2):
                fo, fg = c[i] @ r["x"], c[i] @ X[i]
                dx = np.abs(X[i] - r["x"]).max() / max(1.0, np.abs(r["x"]).max())
                if abs(fg - fo) > 1e-9 * max(1.0, abs(fo)):
                    bad += 1; print("%s %s lp %d: status %d, OBJECTIVE gpu %.12g cpu %.12g (free variables %d)" % (kind, tag, i, status[i], fg, fo, free), flush=True)
                elif status[i] == 1 and (not np.array_equal(St[i], r["S"]) or dx > 1e-9):
                    bad += 1; print("%s %s lp %d: status 1, same objective, another vertex: dx %.1e (free variables %d)" % (kind, tag, i, dx, free), flush=True)
    except Exception as e:      # noqa: BLE001 — a log: keep going
        bad += 1; print("%s %s: exception %r" % (kind, tag, e), flush=True)
    tally[kind] = (n, bad, hist)
print("LP edge fuzz: " + "; ".join("%s %d LPs, oracle statuses %s, %d differences" % (k, v[0], dict(sorted(v[2].items())), v[1]) for k, v in sorted(tally.items())) +
      " | seed0 %d, %.0f s" % (seed0, budget), flush=True)
