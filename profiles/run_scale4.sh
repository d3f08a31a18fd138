#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus4.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo rc=$?
tail -c 400 gpurun_out/bench_4gpu.err; cut -c1-400 gpurun_out/bench_4gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 profiles/bench_configs.py --config 5 --steps 2 > gpurun_out/bench_config5_4gpu.json 2>> gpurun_out/bench_4gpu.err; cut -c1-300 gpurun_out/bench_config5_4gpu.json
