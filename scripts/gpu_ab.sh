#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 1 0 1 0; do
  DT_BN_MASK_FROM_Y=$v timeout 200 python bench.py --workload train --steps 30 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/ab_$v.log 2>&1
  echo "mask_from_y=$v: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_$v.log | head -1)"
done
