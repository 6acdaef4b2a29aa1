#!/bin/bash
# round 2: elected-lane MMA issue in every forward kernel - parity suites, then the cfg2 bench with the per-layer table
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_row.py tests/test_gpu_unet.py tests/test_gpu_parity_trained.py -x -q -m gpu > gpurun_out/r2_g_tests.log 2>&1
echo "tests exit=$?"; tail -n 5 gpurun_out/r2_g_tests.log | cut -c1-300
timeout 600 python bench.py --no-extra --no-cpu-baseline --layer-table gpurun_out/r2_layers_g.txt > gpurun_out/r2_bench_g.log 2>gpurun_out/r2_bench_g.err
echo "bench exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_g.log').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['achieved'],d['roofline']['frac'], d['clocks'])
P
awk '{print $1, $3, $4}' gpurun_out/r2_layers_g.txt
