#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
TAIL=15 run pytest_gpu python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x
TAIL=3 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAIL=3 run bench python bench.py --steps 3 --warmup 3 --layer-table gpurun_out/layers.txt
cat gpurun_out/layers.txt
