"""Wide randomised parity sweep on one B200 (a log, not a test): random shapes (N, M, J) and seeds through solveQP_batch
(general bounds: box / free / upper-only / lower-only variables) and SimplexLP_batch, each against the oracle on the same
inputs.  Nothing is asserted: every mismatch is printed with its shape and seed so that it can be replayed, and the last
line is the tally.  `python scripts/gpu_fuzz_big.py [seconds] [first_seed] [large]`  (tests/test_gpu_fuzz.py is the pinned
subset; `large`: N = 180..520 — the 256- and 512-thread CTAs, inverses that outgrow shared memory, the column cache)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O

W = S.workloads
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
large = len(sys.argv) > 3 and sys.argv[3] == "large"
rng = np.random.default_rng(seed0)
t_end = time.time() + budget
n_qp = n_lp = bad_qp = bad_lp = cases = 0
worst = 0.0
O.set_fix_flip(True)          # the flip repaired on both sides, as in tests/test_gpu_fuzz.py
while time.time() < t_end:
    if large:
        N = int(rng.integers(180, 521))
        M = int(rng.integers(0, 7))
        J = int(rng.integers(1, max(2, N // 3)))
    else:
        N = int(rng.choice([rng.integers(3, 24), rng.integers(24, 80), rng.integers(80, 180)]))
        M = int(rng.integers(0, min(6, N // 2) + 1))
        J = int(rng.integers(0 if M > 0 else 1, max(2, N // 2)))
    seed = int(rng.integers(1, 1 << 30))
    nb = 6
    cases += 1
    tag = "N=%d M=%d J=%d seed=%d" % (N, M, J, seed)
    try:
        c = W.general_bounds(nb=nb, N=N, M=M, J=J, seed=seed)
        X, St, status = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
        r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
        for i in range(nb):
            n_qp += 1
            if status[i] != r["status"][i]:
                bad_qp += 1
                both = status[i] > 0 and r["status"][i] > 0
                print("QP  %s qp %d: status gpu %d cpu %d%s" % (tag, i, status[i], r["status"][i], "" if not both else " (final S identical: %s, rel dx %.1e)" % (
                    np.array_equal(St[i], r["S"][i]), np.abs(X[i] - r["x"][i]).max() / max(np.abs(r["x"][i]).max(), 1e-300))), flush=True)
                continue
            if status[i] > 0:
                dx = np.abs(X[i] - r["x"][i]).max() / max(np.abs(r["x"][i]).max(), 1e-300)
                worst = max(worst, dx)
                if not np.array_equal(St[i], r["S"][i]) or dx > 1e-9:
                    bad_qp += 1; print("QP  %s qp %d: S diffs at %s, rel dx %.2e" % (tag, i, np.flatnonzero(St[i] != r["S"][i])[:8], dx), flush=True)
    except Exception as e:      # noqa: BLE001 — a log: keep going
        bad_qp += 1; print("QP  %s: exception %r" % (tag, e), flush=True)
    if M + J == 0:
        continue
    try:
        w = W.general_bounds_lp(nb=nb, N=N, M=M, J=J, seed=seed, bounded=bool(rng.integers(0, 4)))
        Xl, Sl, sl = S.SimplexLP_batch(w["A"], w["G"], w["c"], w["b"], w["g"], w["d"], w["u"])
        for i in range(2 if large else nb):      # (the oracle's LP is serial: inv(lu) on every pivot)
            n_lp += 1
            rl = O.simplex_lp(w["c"][i], w["A"], w["G"], w["b"][i], w["g"][i], w["d"][i], w["u"][i])
            if sl[i] != rl["status"]:
                bad_lp += 1; print("LP  %s lp %d: status gpu %d cpu %d" % (tag, i, sl[i], rl["status"]), flush=True); continue
            if sl[i] in (1, 2):
                dx = np.abs(Xl[i] - rl["x"]).max() / max(1.0, np.abs(rl["x"]).max())
                if not np.array_equal(Sl[i], rl["S"]) or dx > 1e-9:
                    bad_lp += 1
                    print("LP  %s lp %d: S diffs at %s, dx %.2e (status %d on both sides, objective gpu %.12g / oracle %.12g, free variables %d)" % (
                        tag, i, np.flatnonzero(Sl[i] != rl["S"])[:8], dx, sl[i], w["c"][i] @ Xl[i], w["c"][i] @ rl["x"],
                        int((np.isinf(w["d"][i]) & np.isinf(w["u"][i])).sum())), flush=True)
    except Exception as e:      # noqa: BLE001
        bad_lp += 1; print("LP  %s: exception %r" % (tag, e), flush=True)
O.set_fix_flip(False)
print("fuzz%s: %d shapes, %d QPs (%d mismatches, worst rel dx of the matching ones %.2e), %d LPs (%d mismatches), seed0 %d, %.0f s" % (
    " (large)" if large else "", cases, n_qp, bad_qp, worst, n_lp, bad_lp, seed0, budget), flush=True)
