"""Throughput of the other BASELINE configs on one GPU (parity-test cases, not bench lines) + a larger parity sweep."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O
W = S.workloads


def timed(name, c, check=0):
    ctx = S.context()
    args = (c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
    S.solveQP_batch(*args)
    t = time.time(); X, St, status = S.solveQP_batch(*args); wall = time.time() - t
    kms = ctx.last_kernel_ms()
    nb = len(status)
    line = "[%s] nb=%d kernel %.1f ms -> %.0f QPs/s (wall incl. H2D/D2H + numpy staging %.0f QPs/s) | optimal %d, status<=0: %d | %s" % (
        name, nb, kms, nb / kms * 1e3, nb / wall, (status > 0).sum(), (status <= 0).sum(), ctx.last_launch_config())
    if check:
        idx = np.unique(np.concatenate([np.flatnonzero(status <= 0), np.linspace(0, nb - 1, check).astype(int)]))
        sub = lambda a: a[idx] if (a.ndim >= 2 and a.shape[0] == nb) else a
        Vs = c["V"][idx] if c["V"].ndim == 3 else c["V"]
        r = O.solve_batch(Vs, c["A"], c["G"], sub(c["q"]), sub(c["b"]), sub(c["g"]), sub(c["d"]), sub(c["u"]))
        ok = r["status"] > 0
        dx = np.abs(X[idx] - r["x"]).max(axis=1) / np.maximum(np.abs(r["x"]).max(axis=1), 1e-300)
        bad = int((status[idx] != r["status"]).sum() + ((St[idx] != r["S"]).any(axis=1) & ok).sum() + (dx[ok] > 1e-9).sum())
        ns, nS, nx = int((status[idx] != r["status"]).sum()), int(((St[idx] != r["S"]).any(axis=1) & ok).sum()), int((dx[ok] > 1e-9).sum())
        line += " | oracle check on %d QPs: mismatches %d (status %d, S %d, dx>1e-9 %d), max rel dx %.1e" % (len(idx), bad, ns, nS, nx, dx[ok].max() if ok.any() else 0.0)
        for i in np.flatnonzero((status[idx] != r["status"]) | ((St[idx] != r["S"]).any(axis=1) & ok) | ((dx > 1e-9) & ok))[:12]:
            line += "\n     qp %d: status gpu %d cpu %d, S diffs %s, dx %.2e" % (idx[i], status[idx][i], r["status"][i], np.flatnonzero(St[idx][i] != r["S"][i])[:6], dx[i])
    print(line, flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["c2", "c3", "c4"]
    if "c2" in what:
        timed("config2 shared V, 4096 x N=100", W.config2(nb=4096), check=256)
        timed("config2 per-QP V, 1024 x N=100", W.config2(nb=1024, shared_V=False), check=64)
    if "c3" in what:
        timed("config3 frontier sweep, 1024 x N=500 M=2 J=50", W.config3(nb=1024), check=64)
    if "c5" in what:        # config 5: batched LPs through SimplexLP (Phase 1 + Phase 2 of the bounded Dantzig simplex)
        from scipy.optimize import linprog
        nlp = int(os.environ.get("N5", "296"))
        k = W.config5(index=np.arange(nlp))
        ctx = S.context()
        S.SimplexLP_batch(k["A"], k["G"], k["c"][:2], k["b"][:2], k["g"][:2], k["d"][:2], k["u"][:2])
        t = time.time(); X, St, status = S.SimplexLP_batch(k["A"], k["G"], k["c"], k["b"], k["g"], k["d"], k["u"]); wall = time.time() - t
        kms = ctx.last_kernel_ms(); stats = ctx.stats(nlp)
        worst = 0.0
        for i in range(0, nlp, max(nlp // 8, 1)):
            lp = linprog(k["c"][i], A_ub=k["G"], b_ub=k["g"][i], A_eq=k["A"], b_eq=k["b"][i], bounds=[(0, 1)] * 1000, method="highs")
            worst = max(worst, abs(k["c"][i] @ X[i] - lp.fun) / abs(lp.fun))
        print("[config5 LPs, %d x N=1000 M=20 J=180] kernel %.1f ms -> %.1f LPs/s (wall %.1f LPs/s) | status counts %s | simplex loops mean %.0f max %.0f | "
              "LP 0: %d loops (reference-form oracle: 133182 loops, 909 s on one core) | objective vs HiGHS (8 samples): max rel diff %.1e | %s" % (nlp, kms, nlp / kms * 1e3, nlp / wall, dict(zip(*np.unique(status, return_counts=True))),
              stats[:, 4].mean(), stats[:, 4].max(), stats[0, 4], worst, ctx.last_launch_config()), flush=True)
    if "sweep" in what:     # frontier sweep over q = -L*E (QP(P, q, L), src/types.jl:303-319): cold batch vs warm-started chains (ssqp_solve_sweep)
        nb, L = int(os.environ.get("NSWEEP", "4736")), int(os.environ.get("CHAIN", "32"))
        c = W.config4(nb=1)
        Ls = np.logspace(-3, np.log10(3.0), nb)
        q = -Ls[:, None] * c["E"][None, :]
        til = lambda a: np.tile(a[0], (nb, 1))
        b, g, d, u = til(c["b"]), til(c["g"]), til(c["d"]), til(c["u"])
        ctx = S.context()
        S.solveQP_batch(c["V"], c["A"], c["G"], q[:8], b[:8], g[:8], d[:8], u[:8])
        Xc, Sc, sc = S.solveQP_batch(c["V"], c["A"], c["G"], q, b, g, d, u); kc = ctx.last_kernel_ms()
        Xw, Sw, sw = S.solveQP_sweep(c["V"], c["A"], c["G"], q, b, g, d, u, chain_len=L); kw = ctx.last_kernel_ms()
        print("[frontier sweep %d x N=500 M=1 J=99, q = -L*E] cold batch (Phase 1 once): kernel %.1f ms -> %.0f QPs/s, %.0f trips/QP | "
              "chains of %d (warm start from the neighbour): kernel %.1f ms -> %.0f QPs/s, %.1f trips/QP | same S: %s, max |dx| %.1e, all optimal: %s" % (
              nb, kc, nb / kc * 1e3, sc.mean(), L, kw, nb / kw * 1e3, sw.mean(), np.array_equal(Sc, Sw), np.abs(Xc - Xw).max(), bool((sw > 0).all() and (sc > 0).all())), flush=True)
        # the same loop with the oracle on a few chains: solveQP(Q1), then solveQP(Q, S, x) (src/SSQP.jl:237) — call by call
        nS = int((Sc != Sw).any(axis=1).sum()); bad = 0; worst = 0.0; ncheck = 0
        for ch in [0, nb // L // 2, nb // L - 1][:int(os.environ.get("NCHAINCHK", "3"))]:
            prev = None
            for t in range(L):
                i = ch * L + t
                r = O.solve_qp(c["V"], c["A"], c["G"], q[i], b[i], g[i], d[i], u[i], **({} if prev is None else dict(S0=prev["S"], x0=prev["x"])))
                bad += int(r["status"] != sw[i]) + int(not np.array_equal(r["S"], Sw[i]))
                worst = max(worst, np.abs(r["x"] - Xw[i]).max() / np.abs(r["x"]).max()); ncheck += 1
                prev = r
        fc, fw = 0.5 * np.einsum("bi,ij,bj->b", Xc, c["V"], Xc) + (q * Xc).sum(1), 0.5 * np.einsum("bi,ij,bj->b", Xw, c["V"], Xw) + (q * Xw).sum(1)
        print("   cold vs chained: %d of %d status vectors differ, objective rel diff max %.1e | oracle loop on %d QPs of 3 chains: mismatches (status / S) %d, max rel dx %.1e" % (
              nS, nb, np.abs(fc - fw).max() / np.abs(fc).max(), ncheck, bad, worst), flush=True)
    if "c4" in what:
        tot = int(os.environ.get("N4TOTAL", "2368"))
        timed("config4 %d x N=500 M=1 J=99 (every QP checked)" % tot, W.config4(index=np.arange(tot), total=tot), check=tot)
