#!/bin/bash
# launch list of one cfg4 training step (ncu gpu__time_duration, after a plain run)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TRAIN="python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline --no-profile --no-graph"
timeout 300 $TRAIN > gpurun_out/prof_train_plain.log 2>&1 && {
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1200 -c 1400 --csv --log-file gpurun_out/prof_train_launches.csv $TRAIN > gpurun_out/prof_train_ncu.log 2>&1
  echo "train launch list exit=$?"
}
ls -la gpurun_out | grep prof_train
