#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 1 0; do
  DT_CONV_PAIR=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/layers_pair$v.txt > gpurun_out/ab_pair$v.log 2>&1
  echo "pair=$v: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_pair$v.log | head -2 | tr '\n' ' ') $(grep -o '"frac": [0-9.]*' gpurun_out/ab_pair$v.log | head -1)"
done
