#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list, one ncu --set full capture of the solve kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
NCU_ARGS="--steps 1 --warmup 1 --batch 592 --no-e2e --no-cpu-baseline"
python bench.py $NCU_ARGS > gpurun_out/bench_small.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $NCU_ARGS > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ssqp_solve_kernel -c 1 -o gpurun_out/solve_full -f python bench.py $NCU_ARGS > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
