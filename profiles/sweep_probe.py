#!/usr/bin/env python
"""Timing / profiling driver for the sweep kernel on config 2's generator.

    python profiles/sweep_probe.py [--channels 4096] [--reps 5] [--per-stage] [--path sweep|stream]

Prints one JSON line: fused-launch ms per decomposition (CUDA events), and with --per-stage the per-stage launch times
(the same kernel launched once per stage).  Used under ncu with PYITD_SWEEP_PER_STAGE=1 to profile single stages.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--per-stage", action="store_true")
    ap.add_argument("--path", default=None)
    ap.add_argument("--baselines", action="store_true")
    args = ap.parse_args()
    if args.path:
        os.environ["PYITD_FORCE_PATH"] = args.path
    import torch

    from pyitd_b200 import _capi, synth
    from pyitd_b200.itd import get_plan

    dev = torch.device("cuda", 0)
    S, N = args.channels, args.samples
    x = synth.eeg_like(S, N, seed=1234, device=dev)
    opts = _capi.OPT_BASELINES if args.baselines else 0
    plan = get_plan(0, S, N, _capi.F64, 11, 2, opts)
    rows = plan.rows
    rot = torch.empty((S, rows, N), dtype=torch.float64, device=dev)
    bas = torch.empty((S, rows, N), dtype=torch.float64, device=dev) if args.baselines else None
    ints = [torch.empty(S * (rows if i == 1 else 1), dtype=torch.int32, device=dev) for i in range(5)]
    st = torch.cuda.current_stream(dev)

    def step():
        plan.decompose_device(x.data_ptr(), rot.data_ptr(), bas.data_ptr() if bas is not None else None, ints[0].data_ptr(),
                              ints[1].data_ptr(), ints[2].data_ptr(), ints[3].data_ptr(), ints[4].data_ptr(), st.cuda_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.reps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    out = {"path": plan.path[0], "channels": S, "samples": N, "ms": e0.elapsed_time(e1) / max(args.reps, 1),
           "rows_mean": float(ints[0].double().mean()), "status_max": int(ints[4].max()),
           "env": {k: v for k, v in os.environ.items() if k.startswith("PYITD_")}}
    if plan.path[0] == "sweep" and hasattr(plan._L, "pyitd_plan_sweep_stats"):
        out["fused_pairs"], out["pairs_skipped"], out["pairs_failed"] = plan.sweep_stats()
    if args.per_stage:
        plan.enable_timing(True)
        acc = None
        for _ in range(3):
            step()
            tm = plan.launch_times_ms()
            acc = tm if acc is None else [a + b for a, b in zip(acc, tm)]
        plan.enable_timing(False)
        out["stage_ms"] = [round(a / 3, 4) for a in acc]
        out["stage_ms_sum"] = round(sum(acc) / 3, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
