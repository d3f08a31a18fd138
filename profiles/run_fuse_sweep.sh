#!/bin/bash
# fused-pair settings of the sweep kernel on config 2 (one process per setting, same box)
out=gpurun_out/${1:-fuse}; mkdir -p $out
run() { echo "== $*"; env "$@" python profiles/sweep_probe.py --reps 8 --warmup 3 | tee -a $out/fuse_sweep.jsonl | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms"], d.get("fused_pairs"), d.get("pairs_skipped"), d.get("pairs_failed"), d["rows_mean"])'; }
run PYITD_SWEEP_FUSE=0
run PYITD_SWEEP_FUSE=1
run PYITD_SWEEP_FUSE_MIN_A=8 PYITD_SWEEP_FUSE_MIN_B=3
run PYITD_SWEEP_PF_FUSED=3
run PYITD_SWEEP_FUSE=0
run PYITD_SWEEP_FUSE=1
