timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; tail -c 500 gpurun_out/bench_tmp.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']); print([l['ms'] for l in d['roofline']['per_level']], d['roofline']['knot_scan_ms'])"
timeout 120 python profiles/cfg3_launch_times.py strided 2>/dev/null | cut -c1-260
