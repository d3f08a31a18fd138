#!/usr/bin/env python
"""Times pyitd_decompose_device for a grid of shapes x kernel paths (CUDA events, device-resident inputs).
usage: python profiles/path_sweep.py [out.json]   -- run on the GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pyitd_b200  # noqa: E402
from pyitd_b200 import _capi, synth  # noqa: E402
from pyitd_b200.itd import get_plan  # noqa: E402

SHAPES = [  # (S, n, dtype code, max_iteration, generator)
    (1, 65536, "f64", 20, "chirp"), (8, 65536, "f64", 11, "eeg"), (64, 65536, "f64", 11, "eeg"),
    (256, 65536, "f64", 11, "eeg"), (1024, 65536, "f64", 11, "eeg"), (4096, 65536, "f64", 11, "eeg"),
    (1, 8192, "f32_mixed", 7, "audio"), (64, 8192, "f32_mixed", 7, "audio"), (3515, 8192, "f32_mixed", 7, "audio"),
    (3515, 8192, "f32", 7, "audio"), (4096, 4096, "f64", 11, "eeg"), (16384, 1024, "f64", 11, "eeg"),
]
PATHS = ["lookback", "stream", "resident", "regres"]
CODES = {"f64": _capi.F64, "f32_mixed": _capi.F32_MIXED, "f32": _capi.F32}


def make(S, n, dt, gen):
    if gen == "chirp":
        x = torch.from_numpy(synth.config1_chirp(n)).unsqueeze(0).cuda()
    elif gen == "audio":
        fr = synth.audio_frames(seconds=max(1.0, S * n / 48000.0 + 1.0), frame=n)
        x = torch.from_numpy(fr[:S].copy()).cuda()
    else:
        x = synth.eeg_like(S, n, seed=7, device="cuda")
    return x.to(torch.float64 if dt == "f64" else torch.float32).contiguous()


def main():
    out = []
    for S, n, dt, mi, gen in SHAPES:
        x = make(S, n, dt, gen)
        for path in PATHS:
            os.environ["PYITD_FORCE_PATH"] = path
            pyitd_b200.clear_plan_cache()
            try:
                plan = get_plan(0, S, n, CODES[dt], mi, 2, 0)
                got, cl = plan.path
                if got != path:
                    continue
                rows = plan.rows
                rot = torch.empty((S, rows, n), dtype=x.dtype, device="cuda")
                ints = [torch.empty(S * (rows if i == 1 else 1), dtype=torch.int32, device="cuda") for i in range(5)]
                st = torch.cuda.current_stream().cuda_stream

                def step():
                    plan.decompose_device(x.data_ptr(), rot.data_ptr(), None, ints[0].data_ptr(), ints[1].data_ptr(),
                                          ints[2].data_ptr(), ints[3].data_ptr(), ints[4].data_ptr(), st)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                reps = 20 if S * n < (1 << 26) else 5
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                nr = ints[0].long()
                rec = {"S": S, "n": n, "dtype": dt, "path": path, "cluster": cl, "ms": ms,
                       "Gsamples_per_s": S * n / ms / 1e6, "rows_mean": float(nr.double().mean()),
                       "status_max": int(ints[4].max())}
                print(json.dumps(rec), flush=True)
                out.append(rec)
                del rot
            except Exception as ex:  # noqa: BLE001
                print(json.dumps({"S": S, "n": n, "dtype": dt, "path": path, "error": str(ex)[:200]}), flush=True)
            finally:
                pyitd_b200.clear_plan_cache()
        del x
        torch.cuda.empty_cache()
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
