"""Solve the whole bench shard 0::8 of config 4 (8 192 QPs) on the GPU and store the result in the compact form of
tests/golden/config4_shard0of8.npz (see tests/golden/make_golden_shard.py) -> gpurun_out/shard_gpu.npz."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssqp_b200 as S

TOTAL, SHARDS, N = 65536, 8, 500
idx = np.arange(0, TOTAL, SHARDS, dtype=np.int64)
c = S.workloads.config4(index=idx, total=TOTAL)
ctx = S.context()
t = time.time()
X, St, status = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
wall = time.time() - t
stats = ctx.stats(len(idx))
P = np.random.default_rng(20261018).standard_normal((4, N))
obj = 0.5 * np.einsum("ij,jk,ik->i", X, c["V"], X) + np.einsum("ij,ij->i", X, c["q"])
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "shard_gpu.npz")
np.savez_compressed(out, index=idx, status=status, S=St.astype(np.int8), loops=stats[:, 4].astype(np.int64), obj=obj,
                    xinf=np.abs(X).max(axis=1), proj=X @ P.T, xs=X[::8], maxres=stats[:, 8], rebuilds=stats[:, 7],
                    drift=stats[:, 53], degen=stats[:, 11], updates=stats[:, 6], maxK=stats[:, 2], maxW=stats[:, 3], cycles=stats[:, 9])
print("shard 0::8: %d QPs in %.2f s (kernel %.1f ms) | optimal %d | status<=0 %s | maxres max %.2e | rebuilds max %d mean %.3f | drift rebuilds %d | degen %d | %s"
      % (len(idx), wall, ctx.last_kernel_ms(), (status > 0).sum(), status[status <= 0].tolist(), stats[:, 8].max(), stats[:, 7].max(),
         stats[:, 7].mean(), stats[:, 53].sum(), stats[:, 11].sum(), ctx.last_launch_config()), flush=True)
