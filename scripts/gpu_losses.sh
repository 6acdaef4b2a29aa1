#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_losses.py -x -q -m gpu > gpurun_out/losses_tests.log 2>&1
echo "tests exit=$?"; tail -n 12 gpurun_out/losses_tests.log | cut -c1-300
