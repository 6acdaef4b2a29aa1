#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
TAIL=12 run conv python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -p no:cacheprovider
TAIL=12 run unet python -m pytest tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider
TAIL=3 run bench python bench.py --steps 3 --warmup 3 --layer-table gpurun_out/layers.txt
cat gpurun_out/layers.txt
