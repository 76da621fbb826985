#!/bin/bash
# A/B of builds of the library on ONE box: arguments = variant .so files (relative to the repo root); the default build runs first
mkdir -p gpurun_out
for r in 1 2; do
for lib in "" "$@"; do
L=""; [ -n "$lib" ] && L="$PWD/$lib"
SSQP_LIB=$L python bench.py --batch 8192 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('${lib:-default}', d['value'], d['ms_per_step'], d['solved_ok'])"
done; done
