#!/bin/bash
mkdir -p gpurun_out
timeout 200 python profiles/debug_knots2.py 2>&1 | tail -12
