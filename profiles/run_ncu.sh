#!/bin/bash
# usage: gpurun --timeout 1200 -- bash profiles/run_ncu.sh "<ENV=..>" <kernel regex> <skip> <count> <out name> [bench args...]
mkdir -p gpurun_out
E="$1"; K=$2; S=$3; C=$4; O=$5; shift 5
env $E timeout 1000 ncu --set full --clock-control none --import-source on -k "regex:$K" -s $S -c $C \
    -f -o gpurun_out/$O python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$O.log 2>&1
tail -3 gpurun_out/ncu_$O.log | cut -c1-300
ls -la gpurun_out/$O.ncu-rep
