#!/bin/bash
# ncu --set full of the tiler kernels (gather v2, two-pass stitch)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
timeout 600 $SMALL > gpurun_out/small_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gather_normalize_bf16|stitch_single|stitch_strips" -s 18 -c 6 -o gpurun_out/prof_r1c -f $SMALL > gpurun_out/ncu_r1c.log 2>&1
echo "exit=$?"; tail -n 3 gpurun_out/ncu_r1c.log | cut -c1-300
