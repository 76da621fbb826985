"""Randomised parity log for the two newest device paths (a log, not a test; nothing asserted, the last line is the tally):

* warm-started sweeps (`ssqp_solve_sweep`: a chain's QP reuses the inverse its predecessor left in shared memory) on random
  portfolio shapes and random sweeps over q, against the same loop written with the oracle — solveQP(Q1), then
  solveQP(Q, S, x) (src/SSQP.jl:237) — call by call: status (trip count of the warm-started call), S, x;
* Settings.rule = :stpEdgeLP through solveQP on random shapes with general bounds, against the oracle's restatement.

`python scripts/gpu_fuzz_sweep.py [seconds] [first_seed]`"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O

W = S.workloads
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rng = np.random.default_rng(seed0)

# ---- sweeps ----------------------------------------------------------------------------------------------------------
t_end = time.time() + 0.7 * budget
n_sw = bad_sw = chains = 0
worst = 0.0
while time.time() < t_end:
    N = int(rng.integers(24, 321))
    J = int(rng.integers(0, max(1, N // 5)))
    L = int(rng.choice([2, 4, 8, 16]))
    nch = int(rng.integers(1, 5))
    seed = int(rng.integers(1, 1 << 30))
    nb = L * nch
    tag = "N=%d J=%d chain=%d x %d seed=%d" % (N, J, L, nch, seed)
    try:
        c = W.config4(nb=1, N=N, J=J, seed=seed)
        V, A, G, E = c["V"], c["A"], c["G"], c["E"]
        lo = 10.0 ** rng.uniform(-3, -1)
        Ls = np.sort(lo * 10.0 ** (rng.uniform(0, 2.5) * rng.random(nb)))
        q = -Ls[:, None] * E[None, :]
        til = lambda a: np.tile(a[0], (nb, 1))
        b, g, d, u = til(c["b"]), til(c["g"]), til(c["d"]), til(c["u"])
        X, St, status = S.solveQP_sweep(V, A, G, q, b, g, d, u, chain_len=L)
        for ch in range(nch):
            chains += 1
            prev = None
            for t in range(L):
                i = ch * L + t
                r = O.solve_qp(V, A, G, q[i], b[i], g[i], d[i], u[i], **({} if prev is None else dict(S0=prev["S"], x0=prev["x"])))
                n_sw += 1
                sameS = np.array_equal(r["S"], St[i])
                dx = np.abs(r["x"] - X[i]).max() / max(np.abs(r["x"]).max(), 1e-300) if r["status"] > 0 and status[i] > 0 else 0.0
                if r["status"] != status[i] or not sameS or dx > 1e-9:
                    bad_sw += 1
                    print("SWEEP %s qp %d (position %d of its chain): status gpu %d cpu %d, same S %s, rel dx %.1e" % (
                        tag, i, t, status[i], r["status"], sameS, dx), flush=True)
                elif r["status"] > 0:
                    worst = max(worst, dx)
                if r["status"] <= 0:
                    break           # the reference's loop would stop here too (no optimum to start the next call from)
                prev = r
    except Exception as e:      # noqa: BLE001 — a log: keep going
        bad_sw += 1; print("SWEEP %s: exception %r" % (tag, e), flush=True)
print("sweeps: %d chains, %d warm-started calls compared (%d differences, worst rel dx of the others %.1e)" % (chains, n_sw, bad_sw, worst), flush=True)

# ---- steepest-edge pivot rule ------------------------------------------------------------------------------------------
t_end = time.time() + 0.3 * budget
n_se = bad_se = 0
O.set_fix_flip(True); O.set_rule("stpEdgeLP")
st = S.Settings(rule="stpEdgeLP")
try:
    while time.time() < t_end:
        N = int(rng.integers(4, 120)); M = int(rng.integers(0, min(5, N // 2) + 1)); J = int(rng.integers(0 if M > 0 else 1, max(2, N // 3)))
        seed = int(rng.integers(1, 1 << 30))
        tag = "N=%d M=%d J=%d seed=%d" % (N, M, J, seed)
        try:
            c = W.general_bounds(nb=4, N=N, M=M, J=J, seed=seed)
            X, St, status = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], settingsLP=st)
            r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
            for i in range(4):
                n_se += 1
                dx = np.abs(X[i] - r["x"][i]).max() / max(np.abs(r["x"][i]).max(), 1e-300) if status[i] > 0 and r["status"][i] > 0 else 0.0
                sameS = np.array_equal(St[i], r["S"][i])
                if status[i] != r["status"][i] or not sameS or dx > 1e-9:
                    bad_se += 1
                    print("STPEDGE %s qp %d: status gpu %d cpu %d, same S %s, rel dx %.1e" % (tag, i, status[i], r["status"][i], sameS, dx), flush=True)
        except Exception as e:      # noqa: BLE001
            bad_se += 1; print("STPEDGE %s: exception %r" % (tag, e), flush=True)
finally:
    O.set_rule("Dantzig"); O.set_fix_flip(False)
print("stpEdgeLP: %d QPs compared (%d differences) | seed0 %d, %.0f s" % (n_se, bad_se, seed0, budget), flush=True)
