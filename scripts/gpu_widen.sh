#!/bin/bash
# GPU check of the widened rows: confusion matrices, train_transform, graphed GWDICE step.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_losses.py tests/test_gpu_tiler.py tests/test_gpu_train.py -q -m gpu -k "confusion or train_transform or graphed or gwdl" > gpurun_out/widen.log 2>&1; echo "exit=$? widen"; tail -n 25 gpurun_out/widen.log
