#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/mix_probe.py > gpurun_out/mix_probe.json 2> gpurun_out/mix_probe.err; echo rc=$?; cat gpurun_out/mix_probe.json; tail -3 gpurun_out/mix_probe.err
