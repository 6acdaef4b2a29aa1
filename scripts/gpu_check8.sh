#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -x -q -m gpu > gpurun_out/conv_tests.log 2>&1
echo "conv tests exit=$?"; tail -n 3 gpurun_out/conv_tests.log
timeout 1200 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/train_tests.log 2>&1
echo "train tests exit=$?"; tail -n 3 gpurun_out/train_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/layers.txt > gpurun_out/bench.log 2>&1
echo "bench exit=$?"; tail -c 1000 gpurun_out/bench.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/train_layers.txt > gpurun_out/bench_train.log 2>&1
echo "bench train exit=$?"; head -c 400 gpurun_out/bench_train.log
