#!/bin/bash
# fused-pair settings of the sweep kernel on config 2 (one process per setting, same box)
out=gpurun_out/${1:-fuse}; mkdir -p $out
run() { echo "== $*"; env "$@" python profiles/sweep_probe.py --reps 8 --warmup 3 | tee -a $out/fuse_sweep.jsonl | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms"], d.get("fused_pairs"), d.get("unfused_counts"), d["rows_mean"])'; }
run PYITD_SWEEP_FUSE=0
run PYITD_SWEEP_FUSE=1
run PYITD_SWEEP_PF_COUNT=5
run PYITD_SWEEP_PF_COUNT=8
run PYITD_SWEEP_PF_COUNT=0
run PYITD_SWEEP_PF_FUSED=2
run PYITD_SWEEP_PF_FUSED=4
run PYITD_SWEEP_PF_COUNT=6 PYITD_SWEEP_PF_FUSED=2
