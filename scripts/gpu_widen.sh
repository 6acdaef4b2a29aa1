#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -k "graphed or channel_sum or training_step" > gpurun_out/widen.log 2>&1; echo "exit=$? widen"; grep "losses eager\|passed\|failed" gpurun_out/widen.log | tail -n 5
timeout 240 python scripts/debug_poison.py GWDICE,FOCAL > gpurun_out/poison_gw.log 2>&1; echo "exit=$? poison gwdice"; grep -v Warn gpurun_out/poison_gw.log | grep "eager\|graph\|sumsq" | tail -n 12
