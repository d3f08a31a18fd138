#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not_16_byte" > gpurun_out/pytest_align.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/pytest_align.log | cut -c1-200
