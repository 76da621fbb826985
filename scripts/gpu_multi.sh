#!/bin/bash
# multi-GPU box: the library's own multi-device path, the multi-device tests, and bench.py under torchrun at N = 8, 4, 2
mkdir -p gpurun_out
N=${NGPU:-8}
python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/pytest_multi.log; tail -3 gpurun_out/pytest_multi.log
python scripts/bench_multi_device.py --devices $N > gpurun_out/bench_multi_device_$N.json 2> gpurun_out/bench_multi_device.err; cat gpurun_out/bench_multi_device_$N.json
for n in $N 4 2; do
  [ $n -le $N ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  tail -c 600 gpurun_out/bench_${n}gpu.json; echo
done
