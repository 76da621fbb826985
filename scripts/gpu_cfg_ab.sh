#!/bin/bash
# launch-configuration A/B on one box for the benchmark shape (N=500, M+J=100): CTA width x inverse rows on chip
run() { SSQP_NT=$1 SSQP_HROWS=$2 python bench.py --batch 8192 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('NT=$1 HROWS=$2:', round(d['value'],1), round(d['ms_per_step'],1), d['solved_ok'], d['launch_config'])"; }
for r in 1 2; do
run 512 0
run 256 92
run 256 84
run 256 0
done
