#!/bin/bash
# L2 prefetch distances of the sweep kernel on config 2 (one process per setting, same box)
out=gpurun_out/${1:-pf}; mkdir -p $out
run() { echo "== $*"; env "$@" python profiles/sweep_probe.py --reps 8 --warmup 3 | tee -a $out/pf_sweep.jsonl | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms"],3))'; }
run PYITD_SWEEP_PF_SPARSE=3
run PYITD_SWEEP_PF_SPARSE=2
run PYITD_SWEEP_PF_SPARSE=4
run PYITD_SWEEP_PF_DENSE=1
run PYITD_SWEEP_PF_DENSE=3
run PYITD_SWEEP_PF_SCAN=2
run PYITD_SWEEP_PF_SCAN=6
run PYITD_SWEEP_PF_FUSED=1
run PYITD_SWEEP_PF_FUSED=3
run PYITD_SWEEP_PF_SPARSE=3
