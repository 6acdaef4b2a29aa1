#!/bin/bash
# whole GPU suite + smoke + both benches with per-layer tables (what the driver runs at round end, plus the tables)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?"; tail -n 2 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"; tail -n 1 gpurun_out/smoke.log
timeout 600 python bench.py --layer-table gpurun_out/layers.txt > gpurun_out/bench.log 2>&1
echo "bench exit=$?"; head -c 300 gpurun_out/bench.log; echo
timeout 600 python bench.py --workload train --layer-table gpurun_out/train_layers.txt > gpurun_out/bench_train.log 2>&1
echo "bench train exit=$?"; head -c 300 gpurun_out/bench_train.log; echo
