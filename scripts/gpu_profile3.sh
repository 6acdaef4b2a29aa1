#!/bin/bash
# ncu --set full of the HBM-bound tiler kernels + the two conv kernels furthest from their roofline (stem, dec3.conv1)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
timeout 600 $SMALL > gpurun_out/small_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gather_normalize_bf16|stitch_blend_v8|conv_stem_kernel|conv_halo_kernel<32" -s 30 -c 10 -o gpurun_out/prof_r1b -f $SMALL > gpurun_out/ncu_r1b.log 2>&1
echo "exit=$?"; tail -n 5 gpurun_out/ncu_r1b.log
ls -la gpurun_out | tail -5
