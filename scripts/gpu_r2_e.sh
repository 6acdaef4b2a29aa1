#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/row_probe.py > gpurun_out/r2_probe.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_row -s 2 -c 1 -o gpurun_out/r2_row16 python scripts/row_probe.py 1 > gpurun_out/r2_ncu1.log 2>&1
echo "ncu1 exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_row -s 2 -c 1 -o gpurun_out/r2_row64 python scripts/row_probe.py 0 > gpurun_out/r2_ncu0.log 2>&1
echo "ncu0 exit=$?"
cat gpurun_out/r2_probe.log | tail -5
