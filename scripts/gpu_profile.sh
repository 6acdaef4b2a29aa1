#!/bin/bash
# ncu evidence of one round (one gpurun call, one GPU): every command runs plainly first and only then under ncu.
#   launch list of one cfg2 batch (361 tiles = the 4096^2 mosaic), --set full of its conv launches (46: stem + maxpool and the
#   decoder tail are fused launches), launch list of one cfg4 training step.  Summaries: scripts/summarize_ncu.py, scripts/summarize_convs.py, scripts/summarize_train_step.py
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra"
timeout 300 $SMALL > gpurun_out/prof_small_plain.log 2>&1 && {
  # launches per step: gather + (stem + maxpool) + 44 convs + fused tail + 2 stitch passes = 49; skip the 3 warm-up steps
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|tail_fused|gather_normalize|maxpool|stitch_" -s 147 -c 49 --csv --log-file gpurun_out/prof_launches.csv $SMALL > gpurun_out/prof_ncu_list.log 2>&1
  echo "launch list exit=$?"
  timeout 900 ncu --set full --clock-control none -k regex:"conv_|tail_fused" -s 138 -c 46 -o /tmp/prof_convs -f $SMALL > gpurun_out/prof_ncu_full.log 2>&1
  echo "full capture exit=$?"
  ncu -i /tmp/prof_convs.ncu-rep --page raw --csv > gpurun_out/prof_convs_raw.csv 2> gpurun_out/prof_export.log
}
[ "${PROFILE_TRAIN:-1}" = "0" ] && { ls -la gpurun_out | grep prof_; exit 0; }
TRAIN="python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline --no-profile --no-graph"
timeout 300 $TRAIN > gpurun_out/prof_train_plain.log 2>&1 && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1400 --csv --log-file gpurun_out/prof_train_launches.csv $TRAIN > gpurun_out/prof_train_ncu.log 2>&1
  echo "train launch list exit=$?"
}
ls -la gpurun_out | grep prof_
