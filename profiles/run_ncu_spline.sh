#!/bin/bash
# usage: gpurun --timeout 900 -- bash profiles/run_ncu_spline.sh
# one --set full capture of spline_level_kernel at level 0 and level 2 of the config-2 batch (launches 1 and 3 of the kernel
# with --warmup 0 --steps 1 --levels 3 are level 0, level 1, level 2)
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:spline_level_kernel" -c 3 \
    -f -o gpurun_out/ncu_spline python profiles/bench_spline.py --steps 1 --warmup 0 --levels 3 --cpu-channels 1 > gpurun_out/ncu_spline.log 2>&1
tail -3 gpurun_out/ncu_spline.log | cut -c1-200
ls -la gpurun_out/ncu_spline.ncu-rep
