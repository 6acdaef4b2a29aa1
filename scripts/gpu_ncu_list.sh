#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
$SMALL > gpurun_out/small_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-200
