#!/bin/bash
# ncu --set full of the tiler's HBM kernels (gather + normalise, blended stitch) on one cfg2 batch
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra"
timeout 300 $SMALL > gpurun_out/hbm_plain.log 2>&1 && {
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gather_normalize|stitch_" -s 6 -c 4 -o /tmp/prof_hbm -f $SMALL > gpurun_out/hbm_ncu.log 2>&1
  echo "capture exit=$?"
  ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > gpurun_out/hbm_raw.csv 2> gpurun_out/hbm_export.log
  ncu -i /tmp/prof_hbm.ncu-rep --page details > gpurun_out/hbm_details.txt 2>> gpurun_out/hbm_export.log
  cp /tmp/prof_hbm.ncu-rep gpurun_out/
}
ls -la gpurun_out | grep hbm_
