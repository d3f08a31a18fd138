#!/usr/bin/env python
"""Mid-size batches (16 < S < 160): device time of one decomposition per kernel family (forced), same data.

    python profiles/mid_batch_probe.py [--samples 65536]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--signals", default="17,32,64,100,159")
    args = ap.parse_args()
    import torch

    import pyitd_b200
    from pyitd_b200 import _capi, synth
    from pyitd_b200.itd import get_plan

    dev = torch.device("cuda", 0)
    N = args.samples
    for S in [int(v) for v in args.signals.split(",")]:
        x = synth.eeg_like(S, N, seed=1234, device=dev)
        out = {"signals": S, "samples": N}
        os.environ.pop("PYITD_FORCE_PATH", None)
        pyitd_b200.clear_plan_cache()
        out["default"] = get_plan(0, S, N, _capi.F64, 11, 2, 0).path[0]
        for path in ("resident", "coop", "sweep", "lookback"):
            if path == "coop" and S > 64:
                continue
            os.environ["PYITD_FORCE_PATH"] = path
            pyitd_b200.clear_plan_cache()
            try:
                plan = get_plan(0, S, N, _capi.F64, 11, 2, 0)
            except Exception as ex:
                out[path] = repr(ex)[:60]
                continue
            if plan.path[0] != path:
                out[path] = "n/a (" + plan.path[0] + ")"
                continue
            rows = plan.rows
            rot = torch.empty((S, rows, N), dtype=torch.float64, device=dev)
            ints = [torch.zeros(S * (rows if i == 1 else 1), dtype=torch.int32, device=dev) for i in range(5)]
            st = torch.cuda.current_stream(dev)

            def step():
                plan.decompose_device(x.data_ptr(), rot.data_ptr(), None, ints[0].data_ptr(), ints[1].data_ptr(),
                                      ints[2].data_ptr(), ints[3].data_ptr(), ints[4].data_ptr(), st.cuda_stream)

            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(args.reps):
                step()
            e1.record(st)
            torch.cuda.synchronize()
            out[path] = round(e0.elapsed_time(e1) / args.reps * 1e3, 1)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
