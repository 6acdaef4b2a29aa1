#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -x -q -m gpu > gpurun_out/conv_tests.log 2>&1
echo "tests exit=$?"; tail -n 5 gpurun_out/conv_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/layers.txt > gpurun_out/bench.log 2>&1
echo "bench exit=$?"; tail -c 1300 gpurun_out/bench.log
