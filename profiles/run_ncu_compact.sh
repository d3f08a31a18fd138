#!/bin/bash
# gpurun --timeout 600 -- bash profiles/run_ncu_compact.sh <tag> : --set full capture of the compaction pass (config 3 at 2^27)
T=${1:-c}
mkdir -p gpurun_out
PYITD_CFG3_WARM=0 PYITD_CFG3_LOG2N=27 timeout 500 ncu --set full --clock-control none --import-source on -k "regex:compact_from" -c 8 \
    -f -o gpurun_out/compact_$T python profiles/cfg3_launch_times.py strided > gpurun_out/ncu_compact_$T.log 2>&1
tail -2 gpurun_out/ncu_compact_$T.log | cut -c1-300
