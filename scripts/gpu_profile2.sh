#!/bin/bash
# One gpurun call: plain run of the small bench, then (same command line) the ncu launch list of one whole step
# and a --set full capture of every conv launch of one 135-tile batch.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
# 4096^2 mosaic -> 19 x 19 = 361 tiles = one batch (default --batch-tiles 405); 1 timed step after 3 warm-up steps
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
TAIL=2 run small_plain $SMALL
if [ "$(tail -n1 gpurun_out/small_plain.log | head -c1)" = "{" ]; then
  # launches per step: gather + 46 conv + maxpool + head + 3 stitch passes = 52; skip the 3 warm-up steps
  TAIL=3 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|gather_normalize|maxpool|stitch_" -s 156 -c 52 --csv --log-file gpurun_out/launches.csv $SMALL
  # full capture: the 46 conv launches + head of the timed step (kernels named conv_*)
  TAIL=3 run ncu_full ncu --set full --clock-control none -k regex:conv_ -s 141 -c 47 -o /tmp/prof_convs -f $SMALL
  ncu -i /tmp/prof_convs.ncu-rep --page raw --csv > gpurun_out/prof_convs_raw.csv 2> gpurun_out/prof_export.log
fi
ls -la gpurun_out
