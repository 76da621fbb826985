"""Developer check on a GPU box: CUDA path vs the CPU oracle on the named workloads (verbose)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
from oracle import ssqp_oracle as O
W = S.workloads


def compare(name, c, nthreads=0, show=5):
    nb = c['q'].shape[0]
    t = time.time()
    X, St, status, stats = S.solveQP_batch(c['V'], c['A'], c['G'], c['q'], c['b'], c['g'], c['d'], c['u'], return_stats=True)
    tg = time.time() - t
    kms = S.context().last_kernel_ms()
    print("   launch:", S.context().last_launch_config())
    t = time.time()
    r = O.solve_batch(c['V'], c['A'], c['G'], c['q'], c['b'], c['g'], c['d'], c['u'], nthreads=nthreads, want_stats=True)
    tc = time.time() - t
    same_status = (status == r['status'])
    same_S = (St == r['S']).all(axis=1)
    scale = np.maximum(np.abs(r['x']).max(axis=1), 1e-300)
    dx = np.abs(X - r['x']).max(axis=1) / scale
    print("[%s] nb=%d gpu %.3fs (kernel %.1f ms) cpu %.2fs (%d thr) | status equal %d/%d, S equal %d/%d, max rel dx %.2e"
          % (name, nb, tg, kms, tc, r['threads'], same_status.sum(), nb, same_S.sum(), nb, dx.max()))
    print("   trips mean %.1f max %d | maxK %d maxW %d | lp loops mean %.1f | updates %.0f rebuilds max %d degen %d | maxres %.2e | GB streamed %.3f"
          % (stats[:, 0].mean(), stats[:, 0].max(), stats[:, 2].max(), stats[:, 3].max(), stats[:, 4].mean(), stats[:, 6].mean(),
             stats[:, 7].max(), stats[:, 11].sum(), stats[:, 8].max(), stats[:, 10].sum() / 1e9))
    tot = stats[:, 9].sum()
    if tot > 0:
        names = S.STAT_NAMES
        big = stats[:, 2] > 150
        for nm, m in (("all", np.ones(nb, bool)), ("bigK", big), ("smallK", ~big)):
            if not m.any():
                continue
            cyc = stats[m, 9].sum(); trips = stats[m, 0].sum(); loops = max(stats[m, 4].sum(), 1)
            sec = {names[i]: stats[m, i].sum() for i in range(12, 23)}
            p2 = cyc - sec["cyc_p1"]
            print("   %s: n=%d cycles/QP %.3g | phase1 %.2f (%.0f cyc/loop: price %.0f; rebuild %.0f cyc each) | phase2 %.0f cyc/trip: "
                  "vpass %.0f cpass %.0f symv %.0f syr %.0f gamma %.0f ratio %.0f events %.0f kkt %.0f | symv/trip %.2f (%.0f cyc) syr/trip %.2f (%.0f cyc)"
                  % (nm, m.sum(), cyc / m.sum(), sec["cyc_p1"] / cyc, sec["cyc_p1"] / loops, sec["cyc_p1_price"] / loops,
                     sec["cyc_rebuild"] / max(stats[m, 7].sum(), 1), p2 / trips, sec["cyc_vpass"] / trips, sec["cyc_cpass"] / trips,
                     sec["cyc_symv"] / trips, sec["cyc_syr"] / trips, sec["cyc_gamma"] / trips, sec["cyc_ratio"] / trips,
                     sec["cyc_events"] / trips, sec["cyc_kkt"] / trips, stats[m, 23].sum() / trips,
                     sec["cyc_symv"] / max(stats[m, 23].sum(), 1), stats[m, 24].sum() / trips,
                     sec["cyc_syr"] / max(stats[m, 24].sum(), 1)))
            tl = ["top", "cpass", "ratio", "collect", "step", "rm_gather", "rm_check", "rm_syr", "rm_tail", "ad_gather", "ad_symv",
                  "ad_sum", "ad_syr", "ad_tail", "compact", "vpass", "cpassz", "rhs", "fsymv", "apply", "gamma", "kkt", "misc"]
            print("      timeline cyc/trip: " + " ".join("%s %.0f" % (nm2, stats[m, 13 + 16 + i].sum() / trips) for i, nm2 in enumerate(tl)))
            print("      vpass detail: load loop %.0f cyc/trip, epilogue+barrier %.0f cyc/trip" % (stats[m, 25].sum() / trips, stats[m, 26].sum() / trips))
    bad = np.flatnonzero(~(same_status & same_S) | (dx > 1e-9))
    for i in bad[:show]:
        print("   MISMATCH qp %d: status gpu %d cpu %d, S diff at %s, dx %.2e, lp loops gpu %d cpu %d" %
              (i, status[i], r['status'][i], np.flatnonzero(St[i] != r['S'][i])[:8], dx[i], stats[i, 4], r['stats'][i, 5]))
    return len(bad)


if __name__ == "__main__":
    what = sys.argv[1:] or ["kat", "c2", "c4"]
    nbad = 0
    ctx = S.context()
    if "peak" in what:
        print("fp64 peak TFLOP/s", ctx.measure_fp64_peak(), "read bw 64MB GB/s", ctx.measure_read_bw(64, 20), "read bw 2048MB", ctx.measure_read_bw(2048, 3))
    if "kat" in what:
        nbad += compare("kat3", W.kat_3asset())
        nbad += compare("config1", W.config1())
    if "c2" in what:
        nbad += compare("config2-sharedV-256", W.config2(nb=256))
        nbad += compare("config2-perQPV-64", W.config2(nb=64, shared_V=False))
    if "c4" in what:
        n4 = int(os.environ.get("N4", "32"))
        idx = np.linspace(0, 65535, n4).astype(int)
        nbad += compare("config4-%d" % n4, W.config4(index=idx, total=65536))
    if "c4all" in what:      # every QP of a small global batch (the bench's --batch B at 1 GPU): total = B
        tot = int(os.environ.get("N4TOTAL", "296"))
        nbad += compare("config4-all-%d" % tot, W.config4(index=np.arange(tot), total=tot), show=10)
    if "c4neg" in what:      # the bench batch (8192 QPs at 1 GPU): every QP the GPU does not report optimal vs the oracle
        tot = int(os.environ.get("N4TOTAL", "8192"))
        c = W.config4(index=np.arange(tot), total=tot)
        X, St, status = S.solveQP_batch(c['V'], c['A'], c['G'], c['q'], c['b'], c['g'], c['d'], c['u'])
        neg = np.flatnonzero(status <= 0)
        rng = np.random.default_rng(0)
        pick = np.unique(np.concatenate([neg, rng.choice(tot, 48, replace=False)]))
        r = O.solve_batch(c['V'], c['A'], c['G'], c['q'][pick], c['b'][pick], c['g'][pick], c['d'][pick], c['u'][pick])
        same = (status[pick] == r['status'])
        ok = r['status'] > 0
        dx = np.abs(X[pick] - r['x']).max(axis=1) / np.maximum(np.abs(r['x']).max(axis=1), 1e-300)
        sameS = (St[pick] == r['S']).all(axis=1)
        print("[config4-bench-%d] non-optimal on GPU: %d %s | checked %d QPs vs oracle: status equal %d, S equal (optimal ones) %d/%d, max rel dx %.2e"
              % (tot, len(neg), dict(zip(*np.unique(status[neg], return_counts=True))), len(pick), same.sum(), (sameS & ok).sum(), ok.sum(), dx[ok].max()))
        nbad += int((~same).sum() + (~sameS & ok).sum() + (dx[ok] > 1e-9).sum())
    if "c3" in what:
        nbad += compare("config3-32", W.config3(nb=32))
    print("TOTAL MISMATCHES", nbad)
    sys.exit(1 if nbad else 0)
