#!/bin/bash
# A/B of two builds of the library on ONE box: $1 = the variant .so (relative to the repo root)
mkdir -p gpurun_out
for r in 1 2; do
for lib in "" "$PWD/$1"; do
SSQP_LIB=$lib python bench.py --batch 8192 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('${lib:-default}', d['value'], d['ms_per_step'], d['solved_ok'])"
done; done
