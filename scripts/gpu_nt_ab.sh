#!/bin/bash
# CTA width A/B on the small-problem configs (SSQP_NT = 128 / 256): config 2 shared V, per-QP V, and a mid-size shape
for nt in 256 128; do
  echo "== SSQP_NT=$nt"
  SSQP_NT=$nt python scripts/gpu_configs.py c2 2>&1 | tail -2 | cut -c1-330
  SSQP_NT=$nt python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, ssqp_b200 as S
for (N,J,nb) in ((200,30,2048),(60,12,4096)):
    c=S.workloads.config4(nb=nb,N=N,J=J)
    for _ in range(2): X,St,st=S.solveQP_batch(c["V"],c["A"],c["G"],c["q"],c["b"],c["g"],c["d"],c["u"])
    print("config4-shape N=%d J=%d x %d: kernel %.1f ms -> %.0f QPs/s optimal %d | %s"%(N,J,nb,S.context().last_kernel_ms(),nb/S.context().last_kernel_ms()*1e3,(st>0).sum(),S.context().last_launch_config()))
PY
done
