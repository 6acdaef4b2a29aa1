#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/train_layers.txt > gpurun_out/bench_train.log 2>&1
echo "exit=$?"; tail -c 1500 gpurun_out/bench_train.log
CMD="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
$CMD > gpurun_out/train_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 420 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1
