#!/bin/bash
# final snapshot with the driver's own bench arguments
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 300 gpurun_out/bench_final_ref.json; echo
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; head -c 400 gpurun_out/bench_final.json; echo
NCU_CMD="python bench.py --batch 592 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launch.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    --clock-control none --import-source on -k regex:ssqp_solve_kernel -s 1 -c 1 -f -o gpurun_out/solve_full $NCU_CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
