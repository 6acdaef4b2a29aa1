#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
TAIL=12 run train_tests python -m pytest tests/test_gpu_train.py tests/test_gpu_tiler.py -q -m gpu -p no:cacheprovider --tb=short -x
TAIL=3 run bench_train python bench.py --workload train --steps 5 --warmup 3
TAIL=3 run bench python bench.py --steps 3 --warmup 3 --layer-table gpurun_out/layers.txt
tail -3 gpurun_out/layers.txt
