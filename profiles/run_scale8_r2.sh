#!/bin/bash
# usage: gpurun --gpus 8 --timeout 900 -- bash profiles/run_scale8_r2.sh
# raw D2H ceiling at 1/2/4/8 concurrent GPUs (bound / unbound), then bench.py at 8 GPUs with and without NUMA binding
mkdir -p gpurun_out/r2s8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 profiles/d2h_probe.py > gpurun_out/r2s8/d2h_probe.json 2> gpurun_out/r2s8/d2h_probe.err
cut -c1-1500 gpurun_out/r2s8/d2h_probe.json
$TR --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra > gpurun_out/r2s8/bench8_bound.json 2> gpurun_out/r2s8/bench8_bound.err
$TR --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra --no-bind > gpurun_out/r2s8/bench8_unbound.json 2> gpurun_out/r2s8/bench8_unbound.err
python - <<'PY'
import json
for f in ("bench8_bound", "bench8_unbound"):
    try:
        d = json.load(open(f"gpurun_out/r2s8/{f}.json"))
        print(f, d["value"], d["ms_per_step"], {k: d["e2e"].get(k) for k in ("value", "d2h_gbs", "d2h_ceiling_gbs", "pcie_frac", "numa_bound_cpus")})
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -3 gpurun_out/r2s8/*.err
