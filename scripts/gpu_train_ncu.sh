#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
$CMD > gpurun_out/train_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 420 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1
tail -c 600 gpurun_out/train_plain.log
ls -la gpurun_out
