#!/bin/bash
# gpurun --timeout 600 -- bash profiles/run_contig_probe.sh : strided level kernel, every G-th tile vs contiguous runs per block
mkdir -p gpurun_out
PYITD_STRIDED_CONTIG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "strided or long or aligned" 2>&1 | tail -3
for C in 0 1; do
  echo "== PYITD_STRIDED_CONTIG=$C"
  PYITD_STRIDED_CONTIG=$C timeout 120 python profiles/cfg3_launch_times.py strided 2>/dev/null | cut -c1-260
done | tee gpurun_out/contig_probe.log
