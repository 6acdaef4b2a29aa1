#!/bin/bash
# Loss-family check on the GPU: Wasserstein Dice kernels against the reference fixtures + the SemSegment paths that use them.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1

timeout 300 python -m pytest tests/test_gpu_unet.py tests/test_gpu_train.py -q -m gpu -k "val_step_losses or training_step_fp32 or graphed or boundary" > gpurun_out/gwdl_seg.log 2>&1; echo "exit=$? seg"; tail -n 5 gpurun_out/gwdl_seg.log
