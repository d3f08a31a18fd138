#!/usr/bin/env python
"""Measurement of the spline-baseline level (pyitd_extract_spline_device).

    python profiles/bench_spline.py [--channels 4096] [--n 65536] [--steps 10] [--warmup 3] [--levels 4]

Level 0 = the config-2 EEG-like batch; level l > 0 = the spline baseline of level l-1 (fewer knots).  One JSON
line per level: CUDA-event time per launch (knot scan / coefficient pass / evaluation pass, from the plan's event
timing), algorithmic bytes = x read by the scan and by the evaluation + R and B written = 4 s N per signal, and the
oracle's C restatement timed on one host core on a few channels beside it (baseline only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=4096)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--cpu-channels", type=int, default=8)
    args = ap.parse_args()

    import numpy as np
    import torch

    from oracle import itd_oracle as o
    from pyitd_b200 import _capi, synth
    from pyitd_b200.itd import get_plan

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    S, N = args.channels, args.n
    x = synth.eeg_like(S, N, seed=1234, device=dev)
    plan = get_plan(0, S, N, _capi.F64, 0, 2, 0)
    plan.enable_timing(True)
    R = torch.empty_like(x)
    B = torch.empty_like(x)
    cnt = torch.empty(S, dtype=torch.int32, device=dev)
    st = torch.empty(S, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    cur = x
    for lev in range(args.levels):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        parts = []
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                torch.cuda.synchronize()
                ev0.record()
            plan.extract_spline_device(cur.data_ptr(), R.data_ptr(), B.data_ptr(), cnt.data_ptr(), st.data_ptr(), 2, stream)
            if i >= args.warmup:
                parts.append(plan.launch_times_ms())
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.steps
        K = cnt.double().mean().item()
        alg = 4.0 * 8 * S * N
        ok = bool(parts) and len(parts[0]) >= 2
        scan_ms = float(np.mean([p[0] for p in parts])) if ok else None
        spl_ms = float(np.mean([p[1] for p in parts])) if ok else None
        # CPU restatement on one core, a few channels
        xs = cur[: args.cpu_channels].cpu().numpy()
        t0 = time.perf_counter()
        for s in range(xs.shape[0]):
            o.c_spline_level(xs[s])
        cpu_s = time.perf_counter() - t0
        # parity spot check of this very output
        Ro, Bo, Ko = o.c_spline_level(xs[0])
        err = float(np.linalg.norm(B[0].cpu().numpy() - Bo) / np.linalg.norm(Bo))
        print(json.dumps({
            "metric": "input samples/s through one spline-baseline level", "level": lev,
            "value": S * N / (ms * 1e-3), "unit": "samples/s", "ms_per_call": ms,
            "config": {"workload": f"{S} x {N} fp64 channels (config-2 generator), level {lev} of the spline variant",
                       "mean_knots_per_signal": K},
            "launch_ms": {"knot_scan": scan_ms, "spline_level_kernel": spl_ms},
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes": "32 B per sample (x read twice, R and B written) + 24 B per knot"},
            "cpu_baseline": {"value": xs.size / cpu_s, "unit": "samples/s", "cores": 1, "kind": "port",
                             "sample": f"{xs.shape[0]} channels, oracle/itd_oracle.c itd_oracle_spline_level_f64"},
            "parity_rel_l2_vs_oracle_channel0": err, "status_any": bool(st.any().item()),
        }), flush=True)
        cur = B.clone()


if __name__ == "__main__":
    main()
