#!/bin/bash
bash scripts/gpu_ab_lib.sh "$@"
python scripts/gpu_shard_dump.py > gpurun_out/shard_dump.log 2>&1; tail -1 gpurun_out/shard_dump.log
python - <<'PY'
import numpy as np
g=np.load('tests/golden/config4_shard0of8.npz'); d=np.load('gpurun_out/shard_gpu.npz')
for form in ("lapack","scalar"):
    st=g['status_'+form]; ok=st>0
    pr=np.abs(g['proj_'+form]-d['proj'])/(g['xinf_'+form][:,None]*22.0)
    print(form, "status equal", (st==d['status']).sum(), "S equal", (g['S_'+form]==d['S']).all(axis=1).sum(), "proj rel max %.2e"%pr[ok].max(), "obj rel max %.2e"%(np.abs(g['obj_'+form]-d['obj'])[ok]/np.abs(g['obj_'+form][ok])).max())
PY
python -m pytest tests/test_gpu_robust.py -q -s -k "from_scratch or drift or ill_cond" 2>&1 | tail -6
