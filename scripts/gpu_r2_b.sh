#!/bin/bash
# round 2: row-streaming conv kernel - parity vs the tile kernels, then the bench with and without it
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv_row.py -x -q -m gpu -s > gpurun_out/r2_row.log 2>&1
rc=$?
echo "row tests exit=$rc"; grep -E "^\[|differs|passed|failed|Error|error|timed out" gpurun_out/r2_row.log | cut -c1-300 | tail -n 40
if [ $rc -ne 0 ]; then tail -n 30 gpurun_out/r2_row.log | cut -c1-300; exit 0; fi
timeout 600 python bench.py --no-extra --no-cpu-baseline --layer-table gpurun_out/r2_layers_row.txt > gpurun_out/r2_bench_row.log 2>gpurun_out/r2_bench_row.err
echo "bench(row) exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_row.log').read().strip().splitlines()[-1])
print('row: ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['achieved'],d['roofline']['frac'])
P
DT_CONV_ROW=0 timeout 600 python bench.py --no-extra --no-cpu-baseline --layer-table gpurun_out/r2_layers_norow.txt > gpurun_out/r2_bench_norow.log 2>gpurun_out/r2_bench_norow.err
echo "bench(no row) exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_norow.log').read().strip().splitlines()[-1])
print('norow: ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['achieved'],d['roofline']['frac'])
P
paste <(awk '{print $1, $3}' gpurun_out/r2_layers_row.txt) <(awk '{print $3}' gpurun_out/r2_layers_norow.txt) | head -60
