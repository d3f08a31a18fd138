#!/bin/bash
# gpurun --timeout 900 -- bash profiles/run_ls_probe.sh : knot_ls_kernel pre-pass on / off: parity suite, config 2 bench, config 3 times
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -4
for L in 0 1; do
  echo "== PYITD_LS=$L"
  PYITD_LS=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null > gpurun_out/bench_ls$L.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_ls$L.json')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']); print([l['ms'] for l in d['roofline']['per_level']], d['roofline']['knot_scan_ms'])"
  PYITD_LS=$L timeout 120 python profiles/cfg3_launch_times.py strided 2>/dev/null | cut -c1-260
done | tee gpurun_out/ls_probe.log
