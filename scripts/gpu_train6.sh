#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/train_tests.log 2>&1
echo "tests exit=$?"; tail -n 3 gpurun_out/train_tests.log | cut -c1-300
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/train_layers.txt > gpurun_out/bench_train.log 2>&1
echo "bench exit=$?"; head -c 900 gpurun_out/bench_train.log
