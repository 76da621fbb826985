#!/bin/bash
# inverse rows kept in shared memory vs CTAs per SM on a mid-size shape (N=200, J=30) and on config 1's shape (N=300)
for hr in 0 160 128 100; do
for nt in 256 128; do
SSQP_HROWS=$hr SSQP_NT=$nt python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, ssqp_b200 as S
for (N,J,nb) in ((200,30,2048),):
    c=S.workloads.config4(nb=nb,N=N,J=J)
    for _ in range(2): X,St,st,stats=S.solveQP_batch(c["V"],c["A"],c["G"],c["q"],c["b"],c["g"],c["d"],c["u"],return_stats=True)
    print("HROWS=%s N=%d J=%d x %d: kernel %.1f ms -> %.0f QPs/s optimal %d maxn %d | %s"%(os.environ.get("SSQP_HROWS"),N,J,nb,S.context().last_kernel_ms(),nb/S.context().last_kernel_ms()*1e3,(st>0).sum(),(stats[:,2]+stats[:,3]).max(),S.context().last_launch_config()))
PY
done; done
