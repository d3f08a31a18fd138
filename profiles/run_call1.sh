#!/bin/bash
# GPU-box script (session 3, call 1): parity tests, launch-group variants, full ncu capture of the level kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
bash profiles/run_variants.sh "PYITD_GROUPS=1" "PYITD_GROUPS=2" "PYITD_GROUPS=4" "PYITD_GROUPS=8" "PYITD_GROUPS=16"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:level_stream_kernel" -s 43 -c 8 \
    -f -o gpurun_out/prof_stream_s3 python bench.py --channels 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
