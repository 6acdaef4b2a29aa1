#!/bin/bash
# One gpurun call: bf16 parity re-check, per-layer table, ncu launch list, ncu full capture of conv kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
TAIL=30 run unet_bf16 python -m pytest tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider -s -k "not fp32"
TAIL=5 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAIL=3 run bench python bench.py --steps 3 --warmup 3 --layer-table gpurun_out/layers.txt
cat gpurun_out/layers.txt
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
TAIL=2 run small_plain $SMALL
if [ "$(tail -n1 gpurun_out/small_plain.log | head -c1)" = "{" ]; then
  TAIL=3 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $SMALL
  TAIL=3 run ncu_full ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 230 -c 6 -o gpurun_out/prof_conv -f $SMALL
fi
ls -la gpurun_out
