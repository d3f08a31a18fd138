#!/bin/bash
# GPU-box script: parity tests, the bench line, the other configs, the ncu launch list and (optionally) full captures.
# usage (from the repo root): gpurun --timeout 2400 -- bash profiles/run_round1.sh [full]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
: > gpurun_out/bench_configs.jsonl
timeout 600 python profiles/bench_configs.py --config 3 --dtype f32_mixed >> gpurun_out/bench_configs.jsonl 2>> gpurun_out/bench.err
timeout 600 python profiles/bench_configs.py --config 3 --dtype f32 >> gpurun_out/bench_configs.jsonl 2>> gpurun_out/bench.err
timeout 600 python profiles/bench_configs.py --config 4 >> gpurun_out/bench_configs.jsonl 2>> gpurun_out/bench.err
timeout 600 python profiles/bench_configs.py --config 4 --dtype f32 >> gpurun_out/bench_configs.jsonl 2>> gpurun_out/bench.err
timeout 900 python profiles/bench_configs.py --config 5 --steps 2 >> gpurun_out/bench_configs.jsonl 2>> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:stream_kernel|level_kernel|scan_kernel|resident" -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
if [ "$1" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:level_stream_kernel|scan_stream_kernel" -s 15 -c 4 \
    -f -o gpurun_out/prof_stream python bench.py --channels 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
