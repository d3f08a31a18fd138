#!/bin/bash
# gpurun --timeout 600 -- bash profiles/run_place_sweep.sh : config-3 launch times against the grid of place_knots_kernel
mkdir -p gpurun_out
for B in 74 148 296 592 1184 4736; do
  echo "== PYITD_PLACE_BLOCKS=$B"
  PYITD_PLACE_BLOCKS=$B timeout 120 python profiles/cfg3_launch_times.py strided 2>/dev/null | cut -c1-260
done | tee gpurun_out/place_sweep.log
