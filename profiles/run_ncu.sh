#!/bin/bash
# usage: gpurun --timeout 1200 -- bash profiles/run_ncu.sh <kernel regex> <skip> <count> <out name> [bench args...]
mkdir -p gpurun_out
K=$1; S=$2; C=$3; O=$4; shift 4
timeout 1000 ncu --set full --clock-control none --import-source on -k "regex:$K" -s $S -c $C \
    -f -o gpurun_out/$O python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$O.log 2>&1
tail -3 gpurun_out/ncu_$O.log | cut -c1-600
ls -la gpurun_out/$O.ncu-rep
