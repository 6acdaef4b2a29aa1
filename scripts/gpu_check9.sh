#!/bin/bash
# 2-GPU box: cfg3 (row-sharded inference), cfg4 data-parallel training, cfg5 large tiles on one GPU
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --tile 1024 --overlap 0 --size 8192 --batch-tiles 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.log 2>&1
echo "cfg5 exit=$?"; tail -c 900 gpurun_out/bench_cfg5.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "cfg3 n2 exit=$?"; tail -n 1 gpurun_out/bench_n2.log | head -c 600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload train --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_train_n2.log 2>&1
echo "cfg4 n2 exit=$?"; tail -n 1 gpurun_out/bench_train_n2.log | head -c 600
