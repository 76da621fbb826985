import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
W = S.workloads
nb, L = 4736, 32
c = W.config4(nb=1)
Ls = np.logspace(-3, np.log10(3.0), nb)
q = -Ls[:, None] * c["E"][None, :]
til = lambda a: np.tile(a[0], (nb, 1))
b, g, d, u = til(c["b"]), til(c["g"]), til(c["d"]), til(c["u"])
ctx = S.context()
res = {}
for mode in ("chol", "border"):
    os.environ["SSQP_REBUILD"] = mode
    Xw, Sw, sw, stats = S.solveQP_sweep(c["V"], c["A"], c["G"], q, b, g, d, u, chain_len=L, return_stats=True)
    print(mode, "kernel %.1f ms" % ctx.last_kernel_ms(), "trips mean %.1f" % sw.mean())
    res[mode] = stats
    cyc = stats[:, 9].reshape(-1, L)
    print("   chain cycles (M): min %.0f mean %.0f max %.0f" % (cyc.sum(1).min() / 1e6, cyc.sum(1).mean() / 1e6, cyc.sum(1).max() / 1e6))
    for ch in (0, 74, 147):
        s = stats[ch * L:(ch + 1) * L]
        print("   chain %d: head %.1fM cyc %d trips | warm: mean %.2fM cyc, trips mean %.1f max %d, rebuilds mean %.2f max %d, drift %d, rebuild cyc mean %.0fk, maxres %.1e lamerr %.1e"
              % (ch, s[0, 9] / 1e6, s[0, 0], s[1:, 9].mean() / 1e6, s[1:, 0].mean(), s[1:, 0].max(), s[1:, 7].mean(), s[1:, 7].max(), s[1:, 53].sum(), s[1:, 13 + 6].mean() / 1e3, s[1:, 8].max(), s[1:, 54].max()))
del os.environ["SSQP_REBUILD"]


a, bb = res["chol"], res["border"]
print("chain: n of warm rebuilds (maxK+maxW) | rebuild kcycles chol vs border")
for ch in range(0, 148, 6):
    sl = slice(ch * L + 1, (ch + 1) * L)
    print("  %3d: K %3.0f W %2.0f | %8.0f %8.0f" % (ch, a[sl, 2].mean(), a[sl, 3].mean(), a[sl, 19].mean() / 1e3, bb[sl, 19].mean() / 1e3))
