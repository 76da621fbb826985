#!/bin/bash
# experiment round: (optional) parity tests / checks, then bench variants (small batch), optional ncu
mkdir -p gpurun_out
if [ -n "$DO_TEST" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu.log; fi
if [ -n "$DO_CHECK" ]; then
N4=48 timeout 600 python scripts/gpu_check.py c4 c4all > gpurun_out/check.log 2>&1; echo "check rc $?"
grep -E "MISMATCH|TOTAL|config4-all|cycles/QP" gpurun_out/check.log
fi
B="--steps 2 --warmup 1 --batch ${BATCH:-1184} --no-e2e --no-cpu-baseline"
run() { echo "== $*"; env "$@" python bench.py $B 2> gpurun_out/bench_var.err | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['ms_per_step'], d['solved_ok'], d.get('launch_config'))" || tail -5 gpurun_out/bench_var.err; }
run SSQP_DUMMY=1
while read -r line; do [ -n "$line" ] && run $line; done <<< "$VARIANTS"
if [ -n "$DO_NCU" ]; then bash scripts/gpu_ncu.sh > gpurun_out/ncu_sh.log 2>&1; fi
