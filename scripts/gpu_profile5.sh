#!/bin/bash
# ncu --set full of the row-form stem and the class-fused decoder kernels (one 361-tile batch)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SMALL="python bench.py --size 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
timeout 300 $SMALL > gpurun_out/small_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_stem_rows|conv_halo_quad" -s 9 -c 3 -o gpurun_out/prof_r1d -f $SMALL > gpurun_out/ncu_r1d.log 2>&1
echo "exit=$?"; grep -c "==PROF== Profiling" gpurun_out/ncu_r1d.log
