#!/bin/bash
# GPU check of the widened rows (losses incl. Wasserstein Dice, confusion matrices, train_transform)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_losses.py tests/test_gpu_tiler.py -q -m gpu > gpurun_out/widen.log 2>&1; echo "exit=$? widen"; tail -n 3 gpurun_out/widen.log
