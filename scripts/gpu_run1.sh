#!/bin/bash
# round-2 GPU run: tests, full-shard dump, named-config bench, ncu launch list + full capture of the solve kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python scripts/gpu_shard_dump.py > gpurun_out/shard_dump.log 2>&1; tail -3 gpurun_out/shard_dump.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 1500 gpurun_out/bench_r2.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 800 gpurun_out/bench_r2_ref.json
NCU_CMD="python bench.py --batch 592 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launch.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    --clock-control none --import-source on -k regex:ssqp_solve_kernel -s 1 -c 1 -f -o gpurun_out/solve_full $NCU_CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
