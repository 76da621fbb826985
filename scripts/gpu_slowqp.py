"""Find the most expensive QPs of a shard (diagnostics): python scripts/gpu_slowqp.py <rank> <world> [qps_per_gpu]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S
rank, world = int(sys.argv[1]), int(sys.argv[2])
per = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
total = per * world
idx = np.arange(rank, total, world)
c = S.workloads.config4(index=idx, total=total)
X, St, status, stats = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], return_stats=True)
cyc = stats[:, 9]
order = np.argsort(-cyc)[:8]
print("kernel ms", S.context().last_kernel_ms(), "median cycles/QP %.3g" % np.median(cyc))
for i in order:
    print("global qp %d (local %d): status %d cycles %.3g (%.0f ms) trips %d updates %d rebuilds %d degen %d maxK %d maxW %d lp_loops %d" % (
        idx[i], i, status[i], cyc[i], cyc[i] / 1.965e6, stats[i, 0], stats[i, 6], stats[i, 7], stats[i, 11], stats[i, 2], stats[i, 3], stats[i, 4]))
print("status<=0:", [(int(idx[i]), int(status[i])) for i in np.flatnonzero(status <= 0)])
