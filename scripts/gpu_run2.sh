#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_robust.py tests/test_gpu_parity.py -q -x > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
python bench.py --batch 8192 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_8192.json 2>gpurun_out/bench_8192.err; python -c "import json; d=json.loads(open('gpurun_out/bench_8192.json').readlines()[-1]); print('bench 8192:', d['value'], d['ms_per_step'], d['solved_ok'])"
python scripts/gpu_shard_dump.py > gpurun_out/shard_dump.log 2>&1; tail -2 gpurun_out/shard_dump.log
if [ -f statusswitchingqp.jl_b200/libssqp_b200_tl.so ]; then
  SSQP_LIB=$PWD/statusswitchingqp.jl_b200/libssqp_b200_tl.so N4TOTAL=296 python scripts/gpu_check.py c4all > gpurun_out/timeline.log 2>&1; tail -12 gpurun_out/timeline.log
fi
