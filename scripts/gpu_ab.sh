#!/bin/bash
# A/B of two builds on ONE box (box-to-box spread is ~3 %, larger than most kernel changes): put a copy of an older tree —
# statusswitchingqp.jl_b200/ (with its built libssqp_b200.so), ssqp_b200.py, oracle/, and the CURRENT bench.py — under
# gpurun_ab/ (git-ignored; `git worktree add /tmp/old <commit>`, build there, copy), then `gpurun -- bash scripts/gpu_ab.sh`.
B="--steps 2 --warmup 1 --batch 2368 --no-e2e --no-cpu-baseline"
for r in 1 2; do
for d in gpurun_ab .; do
  (cd $d && python bench.py $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$d', d['value'], d['ms_per_step'], d['solved_ok'])")
done; done
