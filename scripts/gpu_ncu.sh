#!/bin/bash
# ncu full capture of the solve kernel on a small batch (2 waves); source-level counters included.
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --batch ${NCU_BATCH:-296} --no-e2e --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:ssqp_solve_kernel -c 1 -o gpurun_out/solve_full -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
cat gpurun_out/bench_small.json
