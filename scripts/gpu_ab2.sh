#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for bt in 405 675 1013 2025 540; do
  timeout 200 python bench.py --batch-tiles $bt --steps 5 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ab_bt$bt.log 2>&1
  echo "batch_tiles=$bt: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_bt$bt.log | head -2 | tr '\n' ' ')"
done
