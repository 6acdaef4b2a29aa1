#!/bin/bash
# last check of the round: whole GPU suite, then the cfg4 bench with the unrolled / plain BatchNorm reductions (A/B)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?"; tail -n 2 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 60 python bench.py --workload train --no-cpu-baseline --no-profile > gpurun_out/bench_train_a.log 2>&1; echo "train (unrolled reductions) exit=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_train_a.log | head -1
DT_BN_REDUCE_UNROLL=0 timeout 60 python bench.py --workload train --no-cpu-baseline --no-profile > gpurun_out/bench_train_b.log 2>&1; echo "train (plain reductions) exit=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_train_b.log | head -1
