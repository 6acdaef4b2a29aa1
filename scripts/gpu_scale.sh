#!/bin/bash
# N-GPU box: cfg3 (row-sharded inference) and cfg4 (data-parallel training) at N = $1
set -u
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "cfg3 n$N exit=$?"; tail -n 1 gpurun_out/bench_n$N.log | head -c 400
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload train --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_train_n$N.log 2>&1
echo; echo "cfg4 n$N exit=$?"; tail -n 1 gpurun_out/bench_train_n$N.log | head -c 400
