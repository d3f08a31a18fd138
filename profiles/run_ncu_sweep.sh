#!/bin/bash
# usage: gpurun --timeout 1500 -- bash profiles/run_ncu_sweep.sh <out name> <channels> <skip launches> <count> [ENV=..]
# ncu --set full on single stages of the sweep kernel (one launch per stage: 14 launches per decomposition)
mkdir -p gpurun_out
O=$1; CH=$2; S=$3; C=$4; shift 4
env PYITD_SWEEP_PER_STAGE=1 "$@" timeout 1300 ncu --set full --clock-control none --import-source on -k "regex:sweep_kernel" -s $S -c $C \
    -f -o gpurun_out/$O python profiles/sweep_probe.py --channels $CH --reps 1 --warmup 1 > gpurun_out/ncu_$O.log 2>&1
tail -3 gpurun_out/ncu_$O.log | cut -c1-300
ls -la gpurun_out/$O.ncu-rep
