"""Per-section cycles of a Phase-1 simplex loop (developer build: SSQP_TIMELINE=1 python build.py; SSQP_LIB=..._tl.so)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
c = S.workloads.config4(index=np.arange(0, 65536, 221), total=65536)
x0, S0, st = S.initQP_batch(c["A"], c["G"], c["b"], c["g"], c["d"], c["u"])
stats = S.context().stats(len(st))
loops = stats[:, 4].sum()
names = {"vpass": "pricing pass", "kkt": "pricing loop + arg-max", "ad_symv": "p = invB*A[:,k]", "ratio": "ratio test",
         "step": "x_B step + leaving row", "rm_gather": "pivot row of invB", "rm_check": "rank-1 update of invB", "rm_syr": "duals", "rm_tail": "loop end"}
tl = ["top", "cpass", "ratio", "collect", "step", "rm_gather", "rm_check", "rm_syr", "rm_tail", "ad_gather", "ad_symv",
      "ad_sum", "ad_syr", "ad_tail", "compact", "vpass", "cpassz", "rhs", "fsymv", "apply", "gamma", "kkt", "misc"]
print("Phase 1 of %d QPs: %.1f loops per QP, %.0f cycles per loop" % (len(st), loops / len(st), stats[:, 12].sum() / loops))
for i, nm in enumerate(tl):
    v = stats[:, 13 + 16 + i].sum() / loops
    if v > 1: print("   %-28s %7.0f cycles/loop" % (names.get(nm, nm), v))
