#!/bin/bash
# ncu launch list of the training step (eager launches, --no-graph) + the graph-vs-eager test
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "graphed" > gpurun_out/train_tests.log 2>&1
echo "tests exit=$?"; tail -n 2 gpurun_out/train_tests.log
CMD="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph"
$CMD > gpurun_out/train_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1100 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1
tail -c 300 gpurun_out/train_plain.log
