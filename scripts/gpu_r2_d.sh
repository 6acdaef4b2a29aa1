#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash scripts/gpu_r2_c.sh
timeout 600 python -m pytest tests/test_gpu_conv_row.py -x -q -m gpu > gpurun_out/r2_row.log 2>&1
echo "row tests exit=$?"; tail -n 3 gpurun_out/r2_row.log | cut -c1-300
timeout 600 python bench.py --no-extra --no-cpu-baseline --layer-table gpurun_out/r2_layers_row.txt > gpurun_out/r2_bench_row.log 2>gpurun_out/r2_bench_row.err
echo "bench(row) exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_row.log').read().strip().splitlines()[-1])
print('row: ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['achieved'],d['roofline']['frac'])
P
paste <(awk '{print $1, $3}' gpurun_out/r2_layers_row.txt) <(awk '{print $3}' gpurun_out/r2_layers_norow.txt) | grep -E "layer1|blocks.[234]|stem|conv"
