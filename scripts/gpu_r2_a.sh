#!/bin/bash
# round 2, first GPU call: new parity tests + smoke + default bench (with cfg4 / cfg5 extras)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity_trained.py tests/test_server.py -x -q -m gpu -s > gpurun_out/r2_parity.log 2>&1
echo "parity exit=$?"; grep -E "^\[|agreement|passed|failed|Error|error" gpurun_out/r2_parity.log | cut -c1-400 | tail -n 30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "eval_engine or lr_changes or weight_packer_follows or rebucketing or graphed" > gpurun_out/r2_advice.log 2>&1
echo "advice exit=$?"; tail -n 8 gpurun_out/r2_advice.log | cut -c1-300
timeout 900 python bench.py --layer-table gpurun_out/r2_layers.txt > gpurun_out/r2_bench.log 2>gpurun_out/r2_bench.err
echo "bench exit=$?"; tail -c 6000 gpurun_out/r2_bench.log; tail -n 5 gpurun_out/r2_bench.err
