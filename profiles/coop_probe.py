#!/usr/bin/env python
"""Device time of ONE decomposition of a handful of signals (the reference's own use: one 65 536-sample signal):
the cooperative kernel against the look-back launch chain.  CUDA events around single calls, median of many.

    python profiles/coop_probe.py [--signals 1] [--samples 65536] [--reps 200]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--signals", type=int, default=1)
    ap.add_argument("--samples", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--dtype", default="f64")
    args = ap.parse_args()
    import torch

    import pyitd_b200
    from pyitd_b200 import _capi, synth
    from pyitd_b200.itd import get_plan

    dev = torch.device("cuda", 0)
    S, N = args.signals, args.samples
    code = {"f64": _capi.F64, "f32_mixed": _capi.F32_MIXED, "f32": _capi.F32}[args.dtype]
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    x = synth.eeg_like(S, N, seed=1234, device=dev).to(tdt)
    out = {"signals": S, "samples": N, "dtype": args.dtype}
    for path in ("coop", "lookback"):
        os.environ["PYITD_FORCE_PATH"] = path
        pyitd_b200.clear_plan_cache()
        plan = get_plan(0, S, N, code, 11, 2, 0)
        rows = plan.rows
        rot = torch.empty((S, rows, N), dtype=tdt, device=dev)
        ints = [torch.zeros(S * (rows if i == 1 else 1), dtype=torch.int32, device=dev) for i in range(5)]
        st = torch.cuda.current_stream(dev)

        def step():
            plan.decompose_device(x.data_ptr(), rot.data_ptr(), None, ints[0].data_ptr(), ints[1].data_ptr(),
                                  ints[2].data_ptr(), ints[3].data_ptr(), ints[4].data_ptr(), st.cuda_stream)

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            step()
            e1.record(st)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        times.sort()
        out[path] = {"path": plan.path[0], "launches": plan.launches, "us_median": times[len(times) // 2],
                     "us_min": times[0], "us_p90": times[int(len(times) * 0.9)], "rows": int(ints[0][0]),
                     "status": int(ints[4].max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
