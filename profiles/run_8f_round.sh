#!/bin/bash
# usage: gpurun --timeout 2400 -- bash profiles/run_8f_round.sh
# tests + bench lines of the SURVEY 8f rows, then their ncu launch lists (times under ncu are never bench values)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analytics.py tests/test_gpu_sift2d.py tests/test_gpu_spline.py -m gpu -x -q > gpurun_out/pytest_8f.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_8f.log; tail -4 gpurun_out/pytest_8f.log
timeout 400 python profiles/bench_analytics.py > gpurun_out/bench_analytics.json 2> gpurun_out/bench_analytics.err; cut -c1-900 gpurun_out/bench_analytics.json
timeout 400 python profiles/bench_sift2d.py > gpurun_out/bench_sift2d.json 2> gpurun_out/bench_sift2d.err; cut -c1-600 gpurun_out/bench_sift2d.json
timeout 400 python profiles/bench_spline.py > gpurun_out/bench_spline.jsonl 2> gpurun_out/bench_spline.err; grep -o '"ms_per_call": [0-9.]*' gpurun_out/bench_spline.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sift2d.csv \
    python profiles/bench_sift2d.py --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:wpe3|column_fsum|dd_total" -c 12 --csv --log-file gpurun_out/launches_analytics.csv \
    python profiles/bench_analytics.py --channels 1024 --steps 1 --warmup 1 > /dev/null 2>&1
wc -l gpurun_out/launches_sift2d.csv gpurun_out/launches_analytics.csv
