#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DT_ROW_DEBUG=1 timeout 300 python scripts/row_probe.py 2>&1 | grep -E "^\[row|us$" | awk 'NR%4==3||NR%4==0' | cut -c1-330
timeout 300 python scripts/row_probe.py 2>&1 | grep -E "us$"
timeout 600 python -m pytest tests/test_gpu_conv_row.py -x -q -m gpu 2>&1 | tail -3
