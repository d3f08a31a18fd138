#!/usr/bin/env python
"""Measurement of the 2-D crossways ensemble ITD on the notebook's own workload:
one 512 x 512 image, 20 ensemble members = 40 960 spline extracts of 512 samples (siftED2D.ipynb cell 3 stored
output: 10.1457 s for totalextract2d, JIT included, unknown hosted CPU).

    python profiles/bench_sift2d.py [--size 512] [--steps 20] [--warmup 3]

Prints one JSON line: device time per ensemble (CUDA events, inputs resident), end-to-end time through the Python
drop-in (numpy in, numpy out), the oracle's C restatement on one host core on a bounded sample beside it.
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
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()

    import numpy as np
    import torch

    import pyitd_b200
    from oracle import itd_oracle as o
    from pyitd_b200 import _capi
    from pyitd_b200.sift2d import _plans, mad

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    H = W = args.size
    rng = np.random.default_rng(2)
    yy, xx = np.mgrid[0:H, 0:W]
    img = 128 + 40 * np.sin(xx * 0.9 + yy * 0.31) + 25 * np.sin(yy * 1.3) + 10 * rng.standard_normal((H, W))
    draws = 10
    noise = rng.normal(0, mad(img), (draws, H, W))
    rp, cp = _plans(0, 2 * draws, H, W, _capi.F64)
    L = _capi.lib()
    xt, vt = torch.from_numpy(img).to(dev), torch.from_numpy(noise).to(dev)
    scratch = torch.empty(int(L.pyitd_ensemble2d_scratch_bytes(rp.handle, draws, H, W)), dtype=torch.uint8, device=dev)
    low = torch.empty_like(xt)
    st = torch.cuda.current_stream(dev).cuda_stream

    def step():
        _capi.check(L.pyitd_ensemble2d_device(rp.handle, cp.handle, xt.data_ptr(), vt.data_ptr(), low.data_ptr(),
                                              scratch.data_ptr(), draws, H, W, 10, st), "pyitd_ensemble2d_device")

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = rp.launches + cp.launches

    t0 = time.perf_counter()
    for _ in range(3):
        both = pyitd_b200.totalextract2d(img, noise=noise)
    e2e_s = (time.perf_counter() - t0) / 3

    # CPU restatement: one ensemble member (1/20 of the workload) on one core
    t0 = time.perf_counter()
    want = o.crossways(img + noise[0])
    cpu_member_s = time.perf_counter() - t0
    got = pyitd_b200.crossways_itd_baseline_extract(img + noise[0])
    err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    extracts = 2 * draws * 2 * (H + W)
    samples = 2 * draws * 4 * H * W       # sample-levels: four 1-D passes over every member
    print(json.dumps({
        "metric": "2-D crossways ensemble ITD, one image (totalextract2d)", "value": ms * 1e-3, "unit": "s",
        "higher_is_better": False, "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup,
        "config": {"workload": f"{H} x {W} fp64 image, {2 * draws} ensemble members, {extracts} spline extracts",
                   "draws": draws},
        "sample_levels_per_s": samples / (ms * 1e-3), "gpu_launches_per_step": launches,
        "e2e": {"value": e2e_s, "unit": "s", "api": "pyitd_b200.totalextract2d (numpy in, numpy out, noise supplied)"},
        "vs_baseline": 10.145688772201538 / (ms * 1e-3),
        "baseline": "siftED2D.ipynb cell 3 stored output 10.1457 s (same workload, JIT included, unknown hosted CPU)",
        "cpu_baseline": {"value": cpu_member_s * 2 * draws, "unit": "s", "cores": 1, "kind": "port",
                         "sample": "one of the 20 ensemble members timed (oracle.crossways over the C spline level), x 20"},
        "parity_rel_l2_vs_oracle_member0": err,
    }), flush=True)


if __name__ == "__main__":
    main()
