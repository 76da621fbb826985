#!/bin/bash
# perf iteration on the benchmark flavour only (512 threads, 256-bit loads): 8192-QP bench x2 + shard dump vs golden
mkdir -p gpurun_out
for r in 1 2; do
python bench.py --batch 8192 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_8192.json 2>gpurun_out/bench_8192.err; python -c "import json; d=json.loads(open('gpurun_out/bench_8192.json').readlines()[-1]); print('bench 8192:', d['value'], d['ms_per_step'], d['solved_ok'])"
done
python scripts/gpu_shard_dump.py > gpurun_out/shard_dump.log 2>&1; tail -1 gpurun_out/shard_dump.log
python - <<'PY'
import numpy as np
g=np.load('tests/golden/config4_shard0of8.npz'); d=np.load('gpurun_out/shard_gpu.npz')
for form in ("lapack","scalar"):
    st=g['status_'+form]; ok=st>0
    pr=np.abs(g['proj_'+form]-d['proj'])/(g['xinf_'+form][:,None]*22.0)
    print(form, "status equal", (st==d['status']).sum(), "S equal", (g['S_'+form]==d['S']).all(axis=1).sum(), "proj rel max %.2e"%pr[ok].max(), "obj rel max %.2e"%(np.abs(g['obj_'+form]-d['obj'])[ok]/np.abs(g['obj_'+form][ok])).max())
PY
