#!/usr/bin/env python
"""Probe for a hybrid schedule: stream kernel for the dense first levels, then the on-chip resident kernel for the
rest.  Feeds the resident kernel the level-(e*-1) baselines of an EEG-like batch (what it would see as its
input) and times the remaining levels against the stream path's per-level times."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyitd_b200  # noqa: E402
from pyitd_b200 import _capi, synth  # noqa: E402
from pyitd_b200.itd import clear_plan_cache, get_plan  # noqa: E402

S, N = int(os.environ.get("PROBE_S", 2048)), 65536
dev = torch.device("cuda", 0)
x = synth.eeg_like(S, N, seed=1234, device=dev)
res = pyitd_b200.decompose(x, max_iteration=11, return_baselines=True)
torch.cuda.synchronize()
nr = res.n_rows.long()
out = {"S": S, "rows_mean": float(nr.double().mean())}


def time_plan(path, xin, mi, reps=3):
    os.environ["PYITD_FORCE_PATH"] = path
    clear_plan_cache()
    Sx = xin.shape[0]
    plan = get_plan(0, Sx, N, _capi.F64, mi, 2, 0)
    rows = plan.rows
    rot = torch.empty((Sx, rows, N), dtype=torch.float64, device=dev)
    ints = [torch.empty(Sx, dtype=torch.int32, device=dev) for _ in range(4)]
    counts = torch.empty((Sx, rows), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)

    def step():
        plan.decompose_device(xin.data_ptr(), rot.data_ptr(), None, ints[0].data_ptr(), counts.data_ptr(),
                              ints[1].data_ptr(), ints[2].data_ptr(), ints[3].data_ptr(), st.cuda_stream)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    r = {"path": plan.path, "ms": ms, "rows_mean": float(ints[0].double().mean()), "status_max": int(ints[3].max())}
    del rot
    clear_plan_cache()
    return r


out["full_stream"] = None if os.environ.get("PROBE_NO_STREAM") else time_plan("stream", x, 11)
for estar in [int(v) for v in os.environ.get("PROBE_ESTAR", "2,3,4,5,6").split(",")]:
    keep = nr >= estar + 2                      # signals that still have at least one more real extraction
    xin = res.baselines[keep, estar - 1, :].contiguous()
    mi = 11 - estar
    out[f"estar{estar}"] = {"signals": int(keep.sum()),
                            "resident": time_plan("resident", xin, mi),
                            "stream": None if os.environ.get("PROBE_NO_STREAM") else time_plan("stream", xin, mi)}
    print(json.dumps({estar: out[f"estar{estar}"]}), flush=True)
    del xin
os.environ.pop("PYITD_FORCE_PATH", None)
print(json.dumps(out))
