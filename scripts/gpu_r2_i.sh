#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_parity_trained.py -x -q -m gpu > gpurun_out/r2_i_tests.log 2>&1
echo "tests exit=$?"; grep -E "pair .*@8|passed|failed|Error" gpurun_out/r2_i_tests.log | cut -c1-250 | tail -12
timeout 600 python bench.py --no-extra --no-cpu-baseline --layer-table gpurun_out/r2_layers_i.txt > gpurun_out/r2_bench_i.log 2>gpurun_out/r2_bench_i.err
echo "bench exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_i.log').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['achieved'],d['roofline']['frac'], d['clocks'])
P
grep -E "layer4|blocks.0|total" gpurun_out/r2_layers_i.txt
