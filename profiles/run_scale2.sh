#!/bin/bash
# usage: gpurun --gpus 2 --timeout 900 -- bash profiles/run_scale2.sh : the driver's N=2 launch, stdout must be exactly one JSON line
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$?
wc -l gpurun_out/bench_2gpu.json; cut -c1-300 gpurun_out/bench_2gpu.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_reference.json 2>> gpurun_out/bench_2gpu.err; wc -l gpurun_out/bench_2gpu_reference.json; cut -c1-200 gpurun_out/bench_2gpu_reference.json
python -c "
import json
for f in ('gpurun_out/bench_2gpu.json','gpurun_out/bench_2gpu_reference.json'):
    lines=[l for l in open(f).read().splitlines() if l.strip()]
    assert len(lines)==1, (f, len(lines)); d=json.loads(lines[0]); print(f, 'one JSON line ok', d.get('value'), d.get('n_gpus'))
"
