#!/bin/bash
# usage: gpurun --gpus 8 --timeout 900 -- bash profiles/run_scale8.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus8.txt; free -g | head -2 >> gpurun_out/gpus8.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo rc=$?
tail -c 400 gpurun_out/bench_8gpu.err; cut -c1-500 gpurun_out/bench_8gpu.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_8gpu_reference.json 2>> gpurun_out/bench_8gpu.err; cut -c1-200 gpurun_out/bench_8gpu_reference.json
cat gpurun_out/gpus8.txt | tail -3
