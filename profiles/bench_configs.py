#!/usr/bin/env python
"""Measurements for BASELINE.json configs 3, 4 and 5 (bench.py itself is config 2, the headline).

    python profiles/bench_configs.py --config 3 [--dtype f32_mixed|f32]     one 2**28-sample signal, 1 GPU
    python profiles/bench_configs.py --config 4                             3515 x 8192 audio frames, fp32, 8 iterations
    python profiles/bench_configs.py --config 5 [--total-channels 65536]    fp64 channels, chunked, sharded over ranks
    torchrun --nproc-per-node N ... profiles/bench_configs.py --config 5    (strong scaling: total work fixed)

One JSON line per run: device-event timing, max over ranks, inputs resident in HBM, synthetic data.
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
    # stdout carries exactly one JSON line per run: NCCL prints its version banner to file descriptor 1 whatever
    # NCCL_DEBUG_FILE says, so keep a private duplicate of the real stdout and point descriptor 1 at stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4, 5])
    ap.add_argument("--dtype", default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--total-channels", type=int, default=65536)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--log2n", type=int, default=28)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from pyitd_b200 import _capi, shard, synth
    from pyitd_b200.itd import get_plan

    rank, world, local_rank = shard.env_rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log lines (e.g. the "NCCL version" banner) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    codes = {"f64": _capi.F64, "f32_mixed": _capi.F32_MIXED, "f32": _capi.F32}

    if args.config == 3:
        dt, mi, N = args.dtype or "f32_mixed", 11, 1 << args.log2n
        chunks = [synth.long_signal(n=N, seed=3, device=dev).unsqueeze(0)]
        what = f"configs[2]: single long signal of 2^{args.log2n} samples, {dt}"
    elif args.config == 4:
        dt, mi, N = args.dtype or "f32_mixed", 7, 8192
        chunks = [torch.from_numpy(synth.audio_frames()).to(dev)]
        what = f"configs[3]: 48 kHz audio, 10 min, {chunks[0].shape[0]} frames of 8192, fixed 8 iterations, {dt}"
    else:
        dt, mi, N = "f64", 11, 65536
        a, b = shard.shard_range(args.total_channels, rank, world)
        chunks = [synth.eeg_like(c1 - c0, N, seed=1234 + c0 // args.chunk, device=dev, first_channel=c0,
                                 total_channels=args.total_channels)
                  for c0, c1 in shard.chunk_ranges(a, b, args.chunk)]
        what = (f"configs[4]: {args.total_channels} x 65536-sample fp64 channels sharded over {world} GPU(s), "
                f"chunks of {args.chunk} channels recycle one output buffer")
    tdt = torch.float64 if dt == "f64" else torch.float32
    esz_io = 8 if dt == "f64" else 4
    esz_carry = 4 if dt == "f32" else 8
    Smax = max(c.shape[0] for c in chunks)
    plans = {c.shape[0]: get_plan(local_rank, c.shape[0], N, codes[dt], mi, 2, 0) for c in chunks}
    rows = next(iter(plans.values())).rows
    rot = torch.empty((Smax, rows, N), dtype=tdt, device=dev)             # recycled across chunks
    ints = [torch.empty(Smax * (rows if i == 1 else 1), dtype=torch.int32, device=dev) for i in range(5)]
    n_rows_all = [torch.empty(c.shape[0], dtype=torch.int32, device=dev) for c in chunks]
    stream = torch.cuda.current_stream(dev)

    def step():
        for c, nra in zip(chunks, n_rows_all):
            plans[c.shape[0]].decompose_device(c.data_ptr(), rot.data_ptr(), None, nra.data_ptr(), ints[1].data_ptr(),
                                               ints[2].data_ptr(), ints[3].data_ptr(), ints[4].data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = shard.max_over_ranks(ev0.elapsed_time(ev1), device=dev) / args.steps
    S_local = sum(c.shape[0] for c in chunks)
    tot = torch.tensor([S_local, sum(int(t.long().sum()) for t in n_rows_all)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    S_total, levels_total = int(tot[0]), int(tot[1])
    # algorithmic bytes per executed level: read X (carry; the first level reads the io type), write R (io), write B (carry)
    alg = levels_total * N * (esz_carry + esz_io + esz_carry) - S_total * N * (esz_carry - esz_io)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if rank == 0:
        path = plans[Smax].path
        json_out.write(json.dumps({
            "metric": "input samples/s fully decomposed (all levels)", "value": S_total * N / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "dtype": dt,
            "data": "synthetic", "scaling": "strong" if args.config == 5 else "n/a",
            "config": {"workload": what, "n_samples": N, "signals": S_total, "max_iteration": mi, "kernel_path": path[0],
                       "cluster": path[1]},
            "rows_mean": levels_total / max(S_total, 1),
            "sample_levels_per_s": levels_total * N / (ms * 1e-3),
            "roofline_whole_step": {"bound": "hbm", "algorithmic_bytes_per_step": alg, "achieved": alg / (ms * 1e-3) / 1e9 / world,
                                    "peak": peak, "unit": "GB/s per GPU", "frac": alg / (ms * 1e-3) / 1e9 / world / peak},
            "status_max": int(ints[4].max()), "t": time.strftime("%Y-%m-%dT%H:%M:%S"),
        }) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
