#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
