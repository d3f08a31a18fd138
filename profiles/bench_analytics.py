#!/usr/bin/env python
"""Measurement of the post-decomposition analytics on the rows of a config-2 decomposition.

    python profiles/bench_analytics.py [--channels 2048] [--steps 5] [--warmup 2]

One JSON line: weighted permutation entropy of every produced row and exactly rounded column sums, CUDA-event
times, achieved HBM GB/s (algorithmic bytes = one read of the S x rows x N output block; rows beyond n_rows exit /
are skipped), the oracle's C restatement on one host core on a few rows beside it.
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
    ap.add_argument("--channels", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import numpy as np
    import torch

    import pyitd_b200
    from oracle import itd_oracle as o
    from pyitd_b200 import analytics, synth

    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    torch.cuda.set_device(0)
    S, N = args.channels, 65536
    x = synth.eeg_like(S, N, seed=1234, device="cuda")
    res = pyitd_b200.decompose(x, max_iteration=11)
    torch.cuda.synchronize()
    valid_rows = int(res.n_rows.sum())
    bytes_valid = valid_rows * N * 8

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    ms_wpe = timed(lambda: analytics.wpe_rows(res))
    ms_fsum = timed(lambda: analytics.column_fsum(res.rotations, res.n_rows))
    rows = res.rows_of(0).cpu().numpy()
    t0 = time.perf_counter()
    for r in range(rows.shape[0]):
        o.c_wpe3(rows[r], True)
    cpu_wpe = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.c_column_fsum(rows)
    cpu_fsum = time.perf_counter() - t0
    w = analytics.wpe_rows(res)[0, : rows.shape[0]].cpu().numpy()
    err = max(abs(w[r] - o.c_wpe3(rows[r], True)) for r in range(rows.shape[0]))
    print(json.dumps({
        "metric": "analytics over the rows of a config-2 decomposition", "channels": S, "n_samples": N,
        "valid_rows": valid_rows,
        "wpe": {"ms": ms_wpe, "rows_per_s": valid_rows / (ms_wpe * 1e-3), "GBps": bytes_valid / (ms_wpe * 1e-3) / 1e9,
                "frac_of_peak": bytes_valid / (ms_wpe * 1e-3) / 1e9 / peak,
                "cpu_one_core_rows_per_s": rows.shape[0] / cpu_wpe, "max_abs_err_vs_oracle_signal0": float(err)},
        "column_fsum": {"ms": ms_fsum, "GBps": (bytes_valid + S * N * 8) / (ms_fsum * 1e-3) / 1e9,
                        "frac_of_peak": (bytes_valid + S * N * 8) / (ms_fsum * 1e-3) / 1e9 / peak,
                        "cpu_one_core_s_per_signal": cpu_fsum},
        "peak_GBps": peak,
    }), flush=True)


if __name__ == "__main__":
    main()
