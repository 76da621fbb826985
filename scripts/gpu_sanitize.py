"""Small cases for compute-sanitizer (memcheck / racecheck): a few config-4 QPs (column cache, factorisation, tail), a chain
with inverse reuse, general bounds, an LP.   timeout 600 compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
W = S.workloads
c = W.config4(index=np.array([0, 40000, 65535]), total=65536)
X, St, st = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
print("config4", st.tolist())
c1 = W.config4(nb=1, N=120, J=20)
nb = 6
q = -np.logspace(-2, 0, nb)[:, None] * c1["E"][None, :]
til = lambda a: np.tile(a[0], (nb, 1))
Xw, Sw, sw = S.solveQP_sweep(c1["V"], c1["A"], c1["G"], q, til(c1["b"]), til(c1["g"]), til(c1["d"]), til(c1["u"]), chain_len=3)
print("sweep", sw.tolist())
w = W.general_bounds(nb=3, N=40, M=3, J=12, seed=11)
print("general bounds", S.solveQP_batch(w["V"], w["A"], w["G"], w["q"], w["b"], w["g"], w["d"], w["u"])[2].tolist())
lp = W.general_bounds_lp(nb=2, N=30, M=4, J=14, seed=3)
print("lp", S.SimplexLP_batch(lp["A"], lp["G"], lp["c"], lp["b"], lp["g"], lp["d"], lp["u"])[2].tolist())
