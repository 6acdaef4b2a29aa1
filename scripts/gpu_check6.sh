#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_tiler.py -x -q -m gpu > gpurun_out/tiler_tests.log 2>&1
echo "tests exit=$?"; tail -n 8 gpurun_out/tiler_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench exit=$?"; tail -c 1500 gpurun_out/bench.log
