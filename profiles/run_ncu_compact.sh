#!/bin/bash
# gpurun --timeout 600 -- bash profiles/run_ncu_compact.sh <tag> : strided tests, config-3 times, --set full capture of place_knots_kernel (2^27)
T=${1:-c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "strided or long or aligned" 2>&1 | tail -3
timeout 200 python profiles/cfg3_launch_times.py strided 2>/dev/null | cut -c1-260 | tee gpurun_out/cfg3_times_$T.json
PYITD_CFG3_WARM=0 PYITD_CFG3_LOG2N=27 timeout 500 ncu --set full --clock-control none --import-source on -k "regex:place_knots" -c 4 \
    -f -o gpurun_out/place_$T python profiles/cfg3_launch_times.py strided > gpurun_out/ncu_place_$T.log 2>&1
tail -2 gpurun_out/ncu_place_$T.log | cut -c1-300
