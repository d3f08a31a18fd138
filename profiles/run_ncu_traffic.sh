#!/bin/bash
# usage: gpurun --timeout 900 -- bash profiles/run_ncu_traffic.sh [channels]
# DRAM bytes of ONE fused sweep_kernel launch (the whole decomposition of the benchmark batch) -> gpurun_out/traffic.csv
# then (here, after the run):  python profiles/write_traffic_json.py gpurun_out/traffic.csv <channels>
CH=${1:-4096}
mkdir -p gpurun_out
timeout 800 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -k regex:sweep_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv \
    python profiles/sweep_probe.py --channels $CH --reps 1 --warmup 1 > gpurun_out/traffic.log 2>&1
tail -2 gpurun_out/traffic.log | cut -c1-200
tail -4 gpurun_out/traffic.csv
