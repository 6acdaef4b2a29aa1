#!/bin/bash
# quick check after a kernel change: the named test files, then the cfg2 bench line without the extras
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest "$@" -x -q -m gpu > gpurun_out/quick_tests.log 2>&1
echo "tests exit=$?"; tail -n 3 gpurun_out/quick_tests.log | cut -c1-300
timeout 600 python bench.py --no-extra --no-cpu-baseline --layer-table gpurun_out/quick_layers.txt > gpurun_out/quick_bench.log 2>gpurun_out/quick_bench.err
echo "bench exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/quick_bench.log').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['achieved'],d['roofline']['frac'], d['clocks']['sm_mhz'])
print('hbm', {k:(round(v['achieved']),round(v['frac'],3)) for k,v in d['roofline_hbm'].items()})
P
tail -1 gpurun_out/quick_layers.txt
