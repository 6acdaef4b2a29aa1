#!/bin/bash
# ncu launch list of the training step (eager launches, --no-graph): one step after the warm-up steps
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline --no-profile --no-graph"
timeout 200 $CMD > gpurun_out/train_plain.log 2>&1; echo "plain exit=$?"
if [ "$(tail -n1 gpurun_out/train_plain.log | head -c1)" = "{" ]; then
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1400 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1; echo "ncu exit=$?"
  python scripts/summarize_train_step.py gpurun_out/train_launches.csv > gpurun_out/train_step_table.txt 2>&1; head -n 12 gpurun_out/train_step_table.txt
fi
tail -c 300 gpurun_out/train_plain.log
