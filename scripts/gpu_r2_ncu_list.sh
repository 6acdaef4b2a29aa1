#!/bin/bash
# launch list (device time of every launch) of ONE cfg2 step: the plain run first, then the same command under ncu
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra"
$CMD > gpurun_out/r2_list_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1180 -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
echo "exit=$?"; tail -2 gpurun_out/r2_ncu_list.log | cut -c1-200; wc -l gpurun_out/r2_launches.csv
