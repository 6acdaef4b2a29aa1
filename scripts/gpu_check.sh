#!/bin/bash
# One gpurun call: every GPU check in its own process (a trapping kernel must not poison the rest),
# logs under gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log; }
run tiler   python -m pytest tests/test_gpu_tiler.py -q -m gpu --tb=short -p no:cacheprovider
run losses  python -m pytest tests/test_gpu_losses.py -q -m gpu --tb=short -p no:cacheprovider
run conv_cc python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -p no:cacheprovider -k "not tcgen05 and not stem"
TAIL=60 run tc_debug python scripts/tc_debug.py
TAIL=25 run conv_tc python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=line -p no:cacheprovider -s -k "tcgen05 or stem"
TAIL=25 run unet_fp32 python -m pytest tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider -s -k "fp32"
TAIL=25 run unet_bf16 python -m pytest tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider -s -k "not fp32"
TAIL=5 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAIL=5 run bench python bench.py --steps 3 --warmup 3
