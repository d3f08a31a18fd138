#!/bin/bash
# GPU-box script of session 4: full parity suite, smoke, bench line + reference arm, ncu launch list of the bench command,
# fp32 report.  usage: gpurun --timeout 3000 -- bash profiles/run_round1_s4.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 python profiles/fp32_report.py > gpurun_out/fp32_report.jsonl 2> gpurun_out/fp32_report.err; echo "fp32 report rc=$?"; cut -c1-700 gpurun_out/fp32_report.jsonl
cut -c1-300 gpurun_out/bench.json; echo; cut -c1-200 gpurun_out/bench_reference.json
ls -la gpurun_out
