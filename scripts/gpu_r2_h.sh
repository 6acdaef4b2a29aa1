#!/bin/bash
# round 2: full GPU suite + smoke + default bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?"; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_full.log 2>gpurun_out/r2_bench_full.err
echo "bench exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_full.log').read().strip().splitlines()[-1])
print('cfg2 ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
print('hbm', {k:(round(v['achieved']),round(v['frac'],3)) for k,v in d['roofline_hbm'].items()})
print('cfg5', d['cfg5']['value'], d['cfg5']['ms_per_step'], d['cfg5']['roofline']['frac'])
print('cfg4', d['cfg4']['value'], d['cfg4']['ms_per_step'], d['cfg4']['roofline']['frac'], d['cfg4']['roofline']['breakdown'])
P
