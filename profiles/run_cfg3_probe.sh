#!/bin/bash
# gpurun --timeout 900 -- bash profiles/run_cfg3_probe.sh <tag> [full]
# GPU tests, then config 3 (strided path): per-level times, ncu launch list of our kernels (2^28); "full": one --set full capture (2^26)
T=${1:-probe}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_$T.txt
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$T.log
timeout 200 python profiles/cfg3_launch_times.py strided > gpurun_out/cfg3_times_$T.json 2> gpurun_out/cfg3_times_$T.err; cut -c1-600 gpurun_out/cfg3_times_$T.json
PYITD_CFG3_WARM=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:strided|place_knots|tile_prefix" -c 90 --csv \
    --log-file gpurun_out/cfg3_launches_$T.csv python profiles/cfg3_launch_times.py strided > /dev/null 2>&1
if [ "$2" = "full" ]; then
PYITD_CFG3_WARM=0 PYITD_CFG3_LOG2N=26 timeout 500 ncu --set full --clock-control none --import-source on -k "regex:strided|place_knots|tile_prefix" -c 24 \
    -f -o gpurun_out/cfg3_$T python profiles/cfg3_launch_times.py strided > gpurun_out/ncu_cfg3_$T.log 2>&1
tail -2 gpurun_out/ncu_cfg3_$T.log | cut -c1-300
fi
