#!/bin/bash
# training-path check: all training tests, then the cfg4 bench with the per-layer table
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/train_tests.log 2>&1; echo "tests exit=$?"; tail -n 2 gpurun_out/train_tests.log | cut -c1-200
echo skip-bench

