#!/usr/bin/env python
"""Per-launch CUDA-event times of one 2^28-sample decomposition (config 3) for the strided and the look-back path."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyitd_b200
from pyitd_b200 import _capi, synth
from pyitd_b200.itd import clear_plan_cache, get_plan
N = 1 << int(os.environ.get('PYITD_CFG3_LOG2N', '28'))
dev = torch.device("cuda", 0)
x = synth.long_signal(n=N, seed=3, device="cuda").unsqueeze(0).contiguous()
out = {}
for path in sys.argv[1:] or ["strided", "lookback"]:
    os.environ["PYITD_FORCE_PATH"] = path
    clear_plan_cache()
    plan = get_plan(0, 1, N, _capi.F32_MIXED, 11, 2, 0)
    rows = plan.rows
    rot = torch.empty((1, rows, N), dtype=torch.float32, device=dev)
    ints = [torch.empty(1, dtype=torch.int32, device=dev) for _ in range(4)]
    counts = torch.empty((1, rows), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    def step():
        plan.decompose_device(x.data_ptr(), rot.data_ptr(), None, ints[0].data_ptr(), counts.data_ptr(), ints[1].data_ptr(),
                              ints[2].data_ptr(), ints[3].data_ptr(), st)
    for _ in range(int(os.environ.get('PYITD_CFG3_WARM', '2'))):
        step()
    plan.enable_timing(True)
    step()
    ms = plan.launch_times_ms()
    plan.enable_timing(False)
    out[path] = {"path": plan.path[0], "launch_ms": [round(v, 3) for v in ms], "total": round(sum(ms), 3),
                 "knot_counts": counts[0].cpu().tolist(), "n_rows": int(ints[0])}
    del rot
    clear_plan_cache()
print(json.dumps(out))
