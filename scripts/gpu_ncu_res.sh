#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
$SMALL > gpurun_out/small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_res_kernel -s 70 -c 10 -o gpurun_out/prof_res -f $SMALL > gpurun_out/ncu_res.log 2>&1
tail -2 gpurun_out/ncu_res.log | cut -c1-200
