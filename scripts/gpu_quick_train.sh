#!/bin/bash
# quick check after a training-kernel change: the named test files, then the cfg4 bench line
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest "$@" -x -q -m gpu > gpurun_out/quick_train_tests.log 2>&1
echo "tests exit=$?"; tail -n 3 gpurun_out/quick_train_tests.log | cut -c1-300
timeout 600 python bench.py --workload train --no-cpu-baseline --layer-table gpurun_out/quick_train_layers.txt > gpurun_out/quick_train_bench.log 2>gpurun_out/quick_train_bench.err
echo "bench exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/quick_train_bench.log').read().strip().splitlines()[-1])
print('cfg4 ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roofline',d['roofline']['achieved'],d['roofline']['frac'], d['clocks']['sm_mhz'], 'loss', d.get('final_loss'))
P
tail -1 gpurun_out/quick_train_layers.txt
