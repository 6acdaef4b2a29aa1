#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "cta_pair" > gpurun_out/pair_tests.log 2>&1
echo "tests exit=$?"; tail -n 25 gpurun_out/pair_tests.log | cut -c1-220
