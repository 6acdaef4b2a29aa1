#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
TAIL=40 run tc_debug python scripts/tc_debug.py
TAIL=40 run pytest_gpu python -m pytest tests -q -m gpu -p no:cacheprovider -s --tb=short
grep -E "agreement|relative RMS|^\[unet" gpurun_out/pytest_gpu.log | head -40
TAIL=5 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAIL=3 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
