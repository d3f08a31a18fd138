#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/hybrid_probe.py > gpurun_out/hybrid_probe.log 2>&1; echo rc=$?
cat gpurun_out/hybrid_probe.log | cut -c1-600
