#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
TAIL=40 run train_tests python -m pytest tests/test_gpu_train.py -q -m gpu -p no:cacheprovider --tb=short -s
grep -E "^\[|passed|failed|Error|error" gpurun_out/train_tests.log | head -120
TAIL=6 run pytest_gpu python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x --deselect tests/test_gpu_train.py
TAIL=3 run bench python bench.py --steps 3 --warmup 3 --layer-table gpurun_out/layers.txt
cat gpurun_out/layers.txt | tail -12
