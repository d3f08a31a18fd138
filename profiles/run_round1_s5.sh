#!/bin/bash
# GPU-box script of session 5: full parity suite, smoke, bench line + reference arm, ncu launch list of the bench command,
# configs 3 / 4 lines, config-3 per-level times and launch list.  usage: gpurun --timeout 2400 -- bash profiles/run_round1_s5.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:stream|scan|level|knot_ls|place_knots|tile_prefix" -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
for dt in f32_mixed f32; do timeout 300 python profiles/bench_configs.py --config 3 --dtype $dt; done > gpurun_out/bench_config3.jsonl 2> gpurun_out/bench_config3.err
timeout 300 python profiles/bench_configs.py --config 4 > gpurun_out/bench_config4.jsonl 2> gpurun_out/bench_config4.err
timeout 200 python profiles/cfg3_launch_times.py strided > gpurun_out/cfg3_times.json 2> gpurun_out/cfg3_times.err
PYITD_CFG3_WARM=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:strided|place_knots|tile_prefix" -c 90 --csv \
    --log-file gpurun_out/cfg3_launches.csv python profiles/cfg3_launch_times.py strided > /dev/null 2>&1
cut -c1-300 gpurun_out/bench.json; echo; cut -c1-200 gpurun_out/bench_reference.json; echo; cut -c1-400 gpurun_out/bench_config3.jsonl; cut -c1-300 gpurun_out/cfg3_times.json
