timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; tail -c 1500 gpurun_out/bench_tmp.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tmp.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac']); print([ (l['e'],l['ms']) for l in d['roofline']['per_level']], d['roofline']['knot_scan_ms'])"
