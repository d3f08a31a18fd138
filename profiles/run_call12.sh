#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/cfg3_launch_times.py strided lookback > gpurun_out/cfg3_launch_times.json 2> gpurun_out/cfg3_launch_times.err; echo rc=$?
cat gpurun_out/cfg3_launch_times.json; tail -3 gpurun_out/cfg3_launch_times.err
