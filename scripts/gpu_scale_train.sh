#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --workload train --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_train_n$N.log 2>&1
echo "cfg4 n$N exit=$?"; tail -n 4 gpurun_out/bench_train_n$N.log | cut -c1-500
