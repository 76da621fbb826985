#!/bin/bash
# experiment round: parity check, then bench variants (small batch), then ncu
mkdir -p gpurun_out
N4=48 timeout 600 python scripts/gpu_check.py kat c2 c4 c3 c4all > gpurun_out/check.log 2>&1; echo "check rc $?"
grep -E "MISMATCH|TOTAL|config4-all|cycles/QP" gpurun_out/check.log
B="--steps 2 --warmup 1 --batch 1184 --no-e2e --no-cpu-baseline"
for cfg in "" "SSQP_NT=256 SSQP_HROWS=118" "SSQP_NT=256"; do
  echo "== $cfg"; env $cfg python bench.py $B 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['ms_per_step'], d['solved_ok'], d.get('launch_config'))"
done
if [ -n "$DO_NCU" ]; then bash scripts/gpu_ncu.sh > gpurun_out/ncu_sh.log 2>&1; fi
