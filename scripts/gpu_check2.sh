#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
TAIL=80 run tc_debug python scripts/tc_debug.py
TAIL=12 run conv python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=line -p no:cacheprovider
TAIL=30 run unet python -m pytest tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider -s
TAIL=5 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAIL=3 run bench python bench.py --steps 3 --warmup 3 --layer-table gpurun_out/layers.txt
cat gpurun_out/layers.txt
