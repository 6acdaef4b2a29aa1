#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph"
$CMD > gpurun_out/train_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1000 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1
tail -c 300 gpurun_out/train_plain.log
