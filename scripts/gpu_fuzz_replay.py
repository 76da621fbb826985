"""Triage of the mismatches scripts/gpu_fuzz_big.py printed (seed0 1000, 240 s: 8 of 43 428 problems): replay each one on the
device and with the oracle and say WHAT differs — for the QPs, whether the final S / x agree and whether the trip count
agrees once both sides start Phase 2 from the same Phase-1 vertex; for the LPs, the status, the objective and the
feasibility of both answers (the LP cases are either a face of optima — status 2 — or unbounded LPs with free variables,
where the reference overwrites status 3 by 1 / 2, src/Simplex.jl:1001-1019, and returns the vertex it happened to stand on)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O

W = S.workloads
QPS = [(66, 3, 30, 1063817686, 2), (149, 4, 70, 26187268, 1), (154, 6, 67, 360929995, 0), (113, 4, 55, 725792724, 4)]
LPS = [(123, 1, 55, 256669752, 1, False), (145, 1, 70, 890721998, 2, True), (139, 4, 57, 811572718, 5, False), (161, 5, 72, 929700198, 5, True)]

O.set_fix_flip(True)
for N, M, J, seed, i in QPS:
    c = W.general_bounds(nb=6, N=N, M=M, J=J, seed=seed)
    sl = lambda a: a[i:i + 1]
    X, St, status = S.solveQP_batch(c["V"], c["A"], c["G"], sl(c["q"]), sl(c["b"]), sl(c["g"]), sl(c["d"]), sl(c["u"]))
    r = O.solve_qp(c["V"], c["A"], c["G"], c["q"][i], c["b"][i], c["g"][i], c["d"][i], c["u"][i])
    dx = np.abs(X[0] - r["x"]).max() / np.abs(r["x"]).max()
    xo, So, sto, _ = O.init_qp(c["A"], c["G"], c["b"][i], c["g"][i], c["d"][i], c["u"][i])
    Xw, Sw, stw = S.solveQP_batch(c["V"], c["A"], c["G"], sl(c["q"]), sl(c["b"]), sl(c["g"]), sl(c["d"]), sl(c["u"]),
                                  S0=So[None].astype(np.int32), x0=xo[None])
    rw = O.solve_qp(c["V"], c["A"], c["G"], c["q"][i], c["b"][i], c["g"][i], c["d"][i], c["u"][i], S0=So, x0=xo)
    xg, Sg, stg = S.initQP_batch(c["A"], c["G"], sl(c["b"]), sl(c["g"]), sl(c["d"]), sl(c["u"]))[:3]
    print("QP N=%d M=%d J=%d seed=%d qp %d: cold trips gpu %d / oracle %d, same final S: %s, rel dx %.1e | Phase-1 vertex: same S %s, |dx| %.1e | "
          "from the oracle's Phase-1 vertex: trips gpu %d / oracle %d, same S %s" % (
              N, M, J, seed, i, status[0], r["status"], np.array_equal(St[0], r["S"]), dx,
              np.array_equal(Sg[0], So), np.abs(xg[0] - xo).max(), stw[0], rw["status"], np.array_equal(Sw[0], rw["S"])), flush=True)
O.set_fix_flip(False)


def feas(w, i, x):
    v = 0.0
    if w["A"].shape[0]:
        v = max(v, np.abs(w["A"] @ x - w["b"][i]).max())
    if w["G"].shape[0]:
        v = max(v, (w["G"] @ x - w["g"][i]).max())
    return max(v, (w["d"][i] - x).max(), (x - w["u"][i]).max())


for N, M, J, seed, i, bounded in LPS:
    w = W.general_bounds_lp(nb=6, N=N, M=M, J=J, seed=seed, bounded=bounded)
    sl = lambda a: a[i:i + 1]
    ctx = S.context()
    Xl, Sl, st = S.SimplexLP_batch(w["A"], w["G"], sl(w["c"]), sl(w["b"]), sl(w["g"]), sl(w["d"]), sl(w["u"]))
    loops = ctx.stats(1)[0, 4]
    r = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
    print("LP N=%d M=%d J=%d seed=%d lp %d (%s cost): status gpu %d / oracle %d | objective gpu %.12g / oracle %.12g | worst violation gpu %.1e / oracle %.1e | "
          "free variables %d | simplex loops gpu %d" % (
              N, M, J, seed, i, "dual-feasible" if bounded else "random", st[0], r["status"], w["c"][i] @ Xl[0], w["c"][i] @ r["x"],
              feas(w, i, Xl[0]), feas(w, i, r["x"]), int((np.isinf(w["d"][i]) & np.isinf(w["u"][i])).sum()), loops), flush=True)
