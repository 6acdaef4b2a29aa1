#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "wgrad or maxpool_index or layerwise or training_step" > gpurun_out/train_tests.log 2>&1
echo "tests exit=$?"; tail -n 15 gpurun_out/train_tests.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/train_layers.txt > gpurun_out/bench_train.log 2>&1
echo "bench exit=$?"; tail -c 1800 gpurun_out/bench_train.log
