#!/bin/bash
mkdir -p gpurun_out
PROBE_S=592 PROBE_ESTAR=6 PROBE_NO_STREAM=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:resident_kernel" -s 2 -c 1 \
    -f -o gpurun_out/prof_resident_deep python profiles/hybrid_probe.py > gpurun_out/ncu_resident_deep.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_resident_deep.log | cut -c1-400
