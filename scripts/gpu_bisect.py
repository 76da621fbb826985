"""Find the first Phase-2 trip at which the device and the oracle part ways on a QP of the bench shard (both warm-started
from the oracle's Phase-1 point): bisection on Settings.maxIter (both sides return z, S at the iteration cap)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O

TOTAL, SHARDS = 65536, 8
which = [int(a) for a in sys.argv[1:]] or [7010, 7775]
for sidx in which:
    c = S.workloads.config4(index=np.array([sidx * SHARDS]), total=TOTAL)
    x0, S0, st, _ = O.init_qp(c["A"], c["G"], c["b"][0], c["g"][0], c["d"][0], c["u"][0])
    def run(t):
        sg = S.Settings(maxIter=t)
        Xg, Sg, stg = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], S0=S0[None].copy(), x0=x0[None], settings=sg)
        r = O.solve_qp(c["V"], c["A"], c["G"], c["q"][0], c["b"][0], c["g"][0], c["d"][0], c["u"][0], S0=S0.copy(), x0=x0,
                       settings=O.default_settings(max_iter=t))
        return Xg[0], Sg[0], int(stg[0]), r["x"], r["S"], r["status"]
    full = run(7777)
    print("shard idx %d: gpu status %d oracle %d" % (sidx, full[2], full[5]))
    lo, hi = 0, min(abs(full[2]), abs(full[5]))
    while hi - lo > 1:            # invariant: same S after lo trips, different after hi
        mid = (lo + hi) // 2
        Xg, Sg, stg, xo, So, sto = run(mid)
        same = np.array_equal(Sg, So) and np.abs(Xg - xo).max() < 1e-7
        if same: lo = mid
        else: hi = mid
    Xg, Sg, stg, xo, So, sto = run(hi)
    d = np.flatnonzero(Sg != So)
    print("  first difference after trip %d: S differs at %s (gpu %s oracle %s), max|dx| %.3e" % (hi, d[:8], Sg[d][:8], So[d][:8], np.abs(Xg - xo).max()))
    Xl, Sl, stl, xol, Sol, stol = run(lo)
    print("  after trip %d: same S, K=%d, EO rows=%d, max|dx| %.3e" % (lo, (Sl[:500] == 0).sum(), (Sl[500:] == 4).sum(), np.abs(Xl - xol).max()))
    r = O.solve_qp(c["V"], c["A"], c["G"], c["q"][0], c["b"][0], c["g"][0], c["d"][0], c["u"][0], S0=S0.copy(), x0=x0, trace=True)
    print("  oracle trace (K, W, kind 1 step / 2 release / 3 optimal, events) around it:", r["trace"][max(0, lo - 3):hi + 3].tolist())
    # what changed in that trip on each side
    dg = np.flatnonzero(Sg != Sl); do = np.flatnonzero(So != Sol)
    print("  trip %d: gpu switched %s -> %s ; oracle switched %s -> %s" % (hi, dg, Sg[dg], do, So[do]))
    for k in np.union1d(dg, do):
        if k < 500:
            print("    var %d: z gpu %.17g oracle %.17g (before: %.17g / %.17g) u=%.3g" % (k, Xg[k], xo[k], Xl[k], xol[k], c["u"][0][k]))
        else:
            j = k - 500
            print("    row %d: slack gpu %.6e oracle %.6e (before %.6e / %.6e)" % (j, c["g"][0][j] - c["G"][j] @ Xg, c["g"][0][j] - c["G"][j] @ xo, c["g"][0][j] - c["G"][j] @ Xl, c["g"][0][j] - c["G"][j] @ xol))
