#!/bin/bash
# the GPU gate on one B200 (through gpurun): the whole -m gpu suite, smoke(), the default bench.py line
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --layer-table gpurun_out/layers.txt > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench exit=$?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('cfg2 ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['achieved'],d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
print('hbm', {k:(round(v['achieved']),round(v['frac'],3)) for k,v in d['roofline_hbm'].items()})
for k in ('cfg5','cfg4'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k]['roofline']['frac'])
P
