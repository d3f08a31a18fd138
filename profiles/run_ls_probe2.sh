#!/bin/bash
# gpurun --timeout 900 -- bash profiles/run_ls_probe2.sh : LS on/off parity test, then config 2 / config 3 against the table threshold n / DIV
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "knot_ls or strided or stream" 2>&1 | tail -4
for D in 16 8 4; do
  echo "== PYITD_LS_DIV=$D"
  PYITD_LS_DIV=$D timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null > gpurun_out/bench_lsdiv$D.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_lsdiv$D.json')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']); print([l['ms'] for l in d['roofline']['per_level']], d['roofline']['knot_scan_ms'])"
  PYITD_LS_DIV=$D timeout 120 python profiles/cfg3_launch_times.py strided 2>/dev/null | cut -c1-260
done | tee gpurun_out/ls_probe2.log
