#!/bin/bash
# usage: gpurun --gpus 8 --timeout 900 -- bash profiles/run_scale8_final.sh : the driver's N = 8 launch on the final build
mkdir -p gpurun_out/final_r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/final_r2/bench_8gpu.json 2> gpurun_out/final_r2/bench_8gpu.err; echo rc=$?
wc -l gpurun_out/final_r2/bench_8gpu.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_r2/bench_8gpu.json"))
print(d["value"], d["ms_per_step"], d["n_gpus"], d["roofline"]["frac"], d["parity"].get("all_ranks_bit_exact"))
print({k: d["e2e"].get(k) for k in ("value", "d2h_gbs", "d2h_ceiling_gbs", "pcie_frac")})
print({k: (v.get("value"), v.get("ms_per_step")) for k, v in (d.get("extra_configs") or {}).items()})
PY
tail -2 gpurun_out/final_r2/bench_8gpu.err
