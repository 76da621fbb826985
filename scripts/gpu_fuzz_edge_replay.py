"""Triage of the REDUNDANT differences of scripts/gpu_fuzz_edge.py (QPs with a duplicated equality row and an inequality that
repeats an equality: trip count differs, final S and x agree): replay each from the ORACLE's Phase-1 vertex on both sides.
If the trip counts then agree, the difference is Phase 1's (a degenerate LP: ties) and not the redundancy purge's.
`python scripts/gpu_fuzz_edge_replay.py profiles/r02_v5_fuzz_edge.log [max cases]`"""
import os
import re
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O

W = S.workloads
cases = []
for line in open(sys.argv[1]):
    m = re.match(r"REDUNDANT N=(\d+) M=(\d+) J=(\d+) seed=(\d+) qp (\d+): status gpu (-?\d+) cpu (-?\d+)", line)
    if m:
        cases.append(tuple(int(v) for v in m.groups()))
cases = cases[:int(sys.argv[2]) if len(sys.argv) > 2 else 60]
O.set_fix_flip(True)
same = diff = samev = 0
for N, M, J, seed, i, sg, sc in cases:
    c = W.general_bounds(nb=4, N=N, M=M, J=J, seed=seed)
    V, A, G, q, b, g, d, u = (c[k] for k in "VAGqbgdu")
    A = np.vstack([A, A[0]]); b = np.hstack([b, b[:, :1]])
    G = G.copy(); g = g.copy(); G[0] = A[0]; g[:, 0] = b[:, 0]
    one = lambda a: a[i:i + 1]
    xo, So, sto, _ = O.init_qp(A, G, b[i], g[i], d[i], u[i])
    xg, Sg, stg = S.initQP_batch(A, G, one(b), one(g), one(d), one(u))[:3]
    vertex_same = np.array_equal(Sg[0], So) and np.abs(xg[0] - xo).max() <= 1e-9 * max(1.0, np.abs(xo).max())
    samev += int(vertex_same)
    Xw, Sw, stw = S.solveQP_batch(V, A, G, one(q), one(b), one(g), one(d), one(u), S0=So[None].astype(np.int32), x0=xo[None])
    rw = O.solve_qp(V, A, G, q[i], b[i], g[i], d[i], u[i], S0=So, x0=xo)
    ok = stw[0] == rw["status"] and np.array_equal(Sw[0], rw["S"])
    same += int(ok); diff += int(not ok)
    if not ok or vertex_same:
        print("N=%d M=%d J=%d seed=%d qp %d: cold trips gpu %d / oracle %d; Phase-1 vertex identical: %s; from the oracle's vertex: trips gpu %d / oracle %d, same S %s" % (
            N, M, J, seed, i, sg, sc, vertex_same, stw[0], rw["status"], np.array_equal(Sw[0], rw["S"])), flush=True)
O.set_fix_flip(False)
print("REDUNDANT replay: %d cases; from the oracle's Phase-1 vertex %d give the oracle's trip count and S, %d do not; the cold Phase-1 vertex was identical in %d" % (
    len(cases), same, diff, samev), flush=True)
