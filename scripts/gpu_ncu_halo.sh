#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
$SMALL > gpurun_out/small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 60 -c 2 -o gpurun_out/prof_halo -f $SMALL > gpurun_out/ncu_halo.log 2>&1
tail -3 gpurun_out/ncu_halo.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
