#!/bin/bash
# ncu --set full with source of ONE fused-tail launch (scripts/tail_probe.py), after a plain run
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/tail_probe.py > gpurun_out/tail_plain.log 2>&1 && {
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:tail_fused -s 3 -c 1 -o /tmp/prof_tail -f python scripts/tail_probe.py > gpurun_out/tail_ncu.log 2>&1
  echo "capture exit=$?"
  ncu -i /tmp/prof_tail.ncu-rep --page source --csv > gpurun_out/tail_source.csv 2> gpurun_out/tail_export.log
  ncu -i /tmp/prof_tail.ncu-rep --page details > gpurun_out/tail_details.txt 2>> gpurun_out/tail_export.log
}
ls -la gpurun_out | grep tail_
