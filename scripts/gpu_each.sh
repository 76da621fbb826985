#!/bin/bash
# diagnostics: the same small bench on every GPU of the box, one at a time, with the clocks nvidia-smi reports under load
n=$(nvidia-smi -L | wc -l)
for k in $(seq 0 $((n-1))); do
  (sleep 6; nvidia-smi -i $k --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader) &
  CUDA_VISIBLE_DEVICES=$k python bench.py --steps 4 --warmup 2 --batch 2368 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('gpu $k', round(d['value']), d['ms_per_step'], d['clocks'])"
  wait
done
