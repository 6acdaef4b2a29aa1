#!/bin/bash
# round 2: cfg3 on N GPUs (tile-range shards).  usage: gpurun --gpus N -- bash scripts/gpu_r2_scale.sh N [extra bench args]
set -u
N=${1:-2}; shift || true
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity_trained.py -x -q -m gpu -k "tile_range" > gpurun_out/r2_shard_test.log 2>&1
  echo "shard test exit=$?"; tail -n 3 gpurun_out/r2_shard_test.log | cut -c1-300
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2_scale_n$N.log 2>gpurun_out/r2_scale_n$N.err
echo "bench N=$N exit=$?"; tail -n 3 gpurun_out/r2_scale_n$N.err | cut -c1-300
python - "$N" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2_scale_n{n}.log').read().strip().splitlines()[-1])
    print('N',n,'ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'share',d.get('roofline',{}).get('share_of_step'))
    for k in ('cfg5','cfg4'):
        if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k]['e2e']['value'])
except Exception as e:
    print('parse failed', e)
P
