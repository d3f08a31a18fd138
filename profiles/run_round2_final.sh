#!/bin/bash
# gpurun --timeout 1500 -- bash profiles/run_round2_final.sh : round-end evidence on the final build of round 2
# parity suite, smoke, default bench line, reference arm, ncu launch list of the bench command, DRAM traffic of the one sweep launch
O=gpurun_out/final_r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; wc -l $O/bench.json; cut -c1-300 $O/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"; cut -c1-300 $O/bench_reference.json
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
grep -c sweep_kernel $O/ncu_launches_bench.csv
# DRAM bytes of the one fused launch (hash-keyed by bench.py)
timeout 600 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -k regex:sweep_kernel -s 1 -c 1 --csv --log-file $O/traffic.csv \
    python profiles/sweep_probe.py --channels 4096 --reps 1 --warmup 1 > $O/traffic.log 2>&1; echo "ncu traffic rc=$?"
tail -4 $O/traffic.csv
