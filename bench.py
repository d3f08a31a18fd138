#!/usr/bin/env python
"""bench.py -- headline benchmark of the ITD sifting loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one full decomposition (all levels) of one batch of synthetic channels:
BASELINE.json configs[1] -- 4 096 x 65 536-sample fp64 EEG-like channels, knot-count stopping -- per
GPU.  Channels are independent, so N GPUs run N shards with no collective (weak scaling);
`value` = channels x samples over all ranks / max-over-ranks device time.

Keys beyond the base contract:
  roofline      achieved HBM GB/s of the level kernel by ALGORITHMIC bytes (3 x 8 B per sample per
                executed level) / its CUDA-event time, against the measured
                copy peak in MEASURED_PEAKS.json;
  e2e           the same metric through the C ABI with HOST buffers (pyitd_decompose_host: H2D of the
                inputs and D2H of every output row inside the timed region);
  cpu_baseline  the oracle's C port on this box's host cores over a bounded sample of the same
                workload (a reported baseline, not the target);
  parity        64 randomly chosen channels of the LAST timed step compared bit for bit with the oracle
                (rows, row counts, per-level knot counts); the run fails when they differ;
  extra_configs BASELINE.json configs[4] as a STRONG split (65 536 channels / N GPUs, 4096-channel
                chunks) at every N, and configs[2], configs[3] on one GPU.
`--impl reference` times that CPU port alone (the reference is Python+numba and cannot travel to the
GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SAMPLES = 65536
CHANNELS_PER_GPU = 4096
MAX_ITERATION = 11
SEED = 1234
METRIC = "input samples/s fully decomposed (all levels)"
UNIT = "samples/s"


_JSON_OUT = None


def protect_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    "NCCL version ..." banner there whatever NCCL_DEBUG_FILE says): keep a private duplicate of the real stdout for the
    JSON line and point descriptor 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=CHANNELS_PER_GPU, help="channels per GPU")
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--e2e-channels", type=int, default=2048, help="channels per e2e step (host buffers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--groups", type=int, default=0, help="stream-path launch groups (0 = the library's default)")
    ap.add_argument("--parity-channels", type=int, default=64, help="channels of the timed output checked against the oracle")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs legs (configs 3, 4, 5)")
    ap.add_argument("--config5-channels", type=int, default=65536, help="total channels of the strong-scaling leg")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the process to the GPU's NUMA node")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": "configs[1]: batched 4096 x 65536-sample synthetic EEG-like channels, fp64, "
                    "knot-count stopping (max_iteration=11), per GPU",
        "channels_per_gpu": args.channels, "n_samples": args.samples, "max_iteration": MAX_ITERATION,
        "generator": "pyitd_b200.synth.eeg_like(seed=1234+rank)", "parallelism": f"channel-shard x{n_gpus}, no collective", "kernel_path": os.environ.get("PYITD_FORCE_PATH", "auto"),
        "launch_groups": getattr(args, "groups_used", None),
        "l2": "inputs (2 GiB) and outputs (26 GiB) per step are larger than the 126 MB L2; no explicit flush",
        "options": "0 (rotation + trend rows; no baselines output)",
    }


# ---------------------------------------------------------------------------------------------
# CPU port timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------
CPU_SAMPLE_CHANNELS = 1024          # both CPU legs (cpu_baseline and --impl reference) time the same bounded sample
# SURVEY.md section 6 (survey container, one core of an 8-core Xeon, numba 0.65, JIT excluded): the reference's own
# ITD().itd on a 65 536-sample signal.  The reference is Python + numba and cannot travel to the GPU box; the C port
# timed here is ~8x faster per core, i.e. a conservative baseline.
NUMBA_REFERENCE_PER_CORE = {"value": 1.2e6, "unit": UNIT, "source": "SURVEY.md section 6 / BASELINE.md: ITD.py (numba) "
                            "on one core of the survey container, 52-57 ms per 65 536-sample signal; not measured on this box"}
try:
    # measured by tests/golden/time_reference_numba.py (imports /root/reference/ITD.py unmodified) in the build container
    _nt = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_numba_timing.json")))
    NUMBA_REFERENCE_PER_CORE = {
        "value": _nt["config2_channels_65536"]["samples_per_s_per_core"], "unit": UNIT,
        "config1_ms": _nt["config1_chirp_65536"]["ms_median"],
        "source": "the unmodified reference ITD().itd (Python + numba), one core, JIT excluded, on 8 channels of this workload's "
                  "generator; measured in the BUILD CONTAINER (%s), not on this box: tests/golden/time_reference_numba.py"
                  % _nt["host"]["cpu"]}
except Exception:
    pass


def cpu_port_throughput(n_samples: int, budget_s: float = 12.0, max_channels: int = CPU_SAMPLE_CHANNELS):
    """Times oracle/itd_oracle.c (pthreads, one channel per thread) on a bounded sample of the
    workload.  Returns (samples_per_s, threads, n_channels, seconds)."""
    import torch

    from oracle import itd_oracle
    from pyitd_b200 import synth

    threads = itd_oracle.c_max_threads()
    itd_oracle.c_decompose_batch(synth.eeg_like(threads, n_samples, seed=SEED).numpy(), MAX_ITERATION)  # warm
    chunk = max(threads * 4, 32)
    done, spent = 0, 0.0
    while spent < budget_s and done < max_channels:
        x = synth.eeg_like(chunk, n_samples, seed=SEED + 1 + done, first_channel=done,
                           total_channels=CHANNELS_PER_GPU).numpy()
        t0 = time.perf_counter()
        _, _, _, status, _ = itd_oracle.c_decompose_batch(x, MAX_ITERATION)
        spent += time.perf_counter() - t0
        done += chunk
        assert int(status.max()) == 0
    del torch
    return done * n_samples / spent, threads, done, spent


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one "step" = one bounded sample; K steps + W warm-ups stay within a few minutes
    per_step_budget = max(1.0, min(10.0, 120.0 / max(args.steps + args.warmup, 1)))
    vals = []
    threads = n_ch = 0
    for i in range(args.warmup + args.steps):
        v, threads, n_ch, secs = cpu_port_throughput(args.samples, budget_s=per_step_budget)
        if i >= args.warmup:
            vals.append((v, secs, n_ch))
    tot_samples = sum(n * args.samples for _, _, n in vals)
    tot_secs = sum(s for _, s, _ in vals)
    value = tot_samples / tot_secs
    sample = f"{vals[0][2]} channels x {args.samples} samples per step (same generator as the GPU arm), {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_secs / max(len(vals), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "numba_per_core": NUMBA_REFERENCE_PER_CORE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle/itd_oracle.c (bit-exact C port of ITD.py, pthreads over channels): the reference is "
                "Python+numba and is not shipped to the GPU box",
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(prefix="pyitd_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        # "under load" = the upper half of the samples (the sampler also sees the idle edges)
        load = sm[len(sm) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}



# ---------------------------------------------------------------------------------------------
# helpers of our arm
# ---------------------------------------------------------------------------------------------
def kernel_source_hash() -> str:
    """sha256[:16] over the CUDA sources of libpyitd_b200.so: the ncu traffic capture under profiles/ is keyed by it."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "pyitd_b200", "csrc", "*.cu*"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def parity_check(torch, x, rot, n_rows, counts, n_check: int, seed: int):
    """n_check randomly chosen channels of the timed output against the oracle, bit for bit (checker only)."""
    from oracle import itd_oracle
    S = x.shape[0]
    g = torch.Generator()
    g.manual_seed(seed)
    idx = torch.randperm(S, generator=g)[: min(n_check, S)].sort().values
    dev_idx = idx.to(x.device)
    xs = x[dev_idx].cpu().numpy()
    got_rot = rot[dev_idx].cpu().numpy()
    got_nr = n_rows[dev_idx].cpu().numpy()
    got_cnt = counts[dev_idx].cpu().numpy()
    want_rot, want_nr, want_cnt, st, _ = itd_oracle.c_decompose_batch(xs, MAX_ITERATION)
    bad = []
    for i, ch in enumerate(idx.tolist()):
        nr = int(want_nr[i])
        ok = (int(st[i]) == 0 and int(got_nr[i]) == nr and got_cnt[i, :nr].tolist() == want_cnt[i, :nr].tolist()
              and got_rot[i, :nr].tobytes() == want_rot[i, :nr].tobytes())
        if not ok:
            bad.append(ch)
    return {"channels": len(idx), "bit_exact": not bad, "mismatching_channels": bad[:8],
            "checked": "every rotation / trend row, the row count and the per-level knot counts of the last timed step "
                       "vs oracle/itd_oracle.c (ITD.py:351-433)", "seed": seed}


def timed_passes(torch, shard, step, stream, dev, steps: int, warmup: int, world: int, dist):
    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    return shard.max_over_ranks(ev0.elapsed_time(ev1), device=dev) / steps


def extra_config(torch, dist, which: int, args, rank, world, local_rank, dev, peak, rot_buf=None):
    """BASELINE.json configs[2] / [3] / [4] through the same C ABI; device-event time."""
    from pyitd_b200 import _capi, shard, synth
    from pyitd_b200.itd import get_plan
    stream = torch.cuda.current_stream(dev)
    if which == 5:
        code, tdt, mi, N, io, carry = _capi.F64, torch.float64, MAX_ITERATION, N_SAMPLES, 8, 8
        a, b = shard.shard_range(args.config5_channels, rank, world)
        chunks = [synth.eeg_like(c1 - c0, N, seed=SEED + c0 // 4096, device=dev, first_channel=c0,
                                 total_channels=args.config5_channels)
                  for c0, c1 in shard.chunk_ranges(a, b, 4096)]
        what = (f"configs[4]: {args.config5_channels} x 65536-sample fp64 channels, STRONG split over {world} GPU(s): "
                f"{b - a} channels on this rank in {len(chunks)} chunk(s) of <= 4096 that recycle one output buffer")
        steps, warm = 2, 1
    elif which == 3:
        code, tdt, mi, N, io, carry = _capi.F32_MIXED, torch.float32, MAX_ITERATION, 1 << 28, 4, 8
        chunks = [synth.long_signal(n=N, seed=3, device=dev).unsqueeze(0)]
        what = "configs[2]: single long signal of 2^28 samples, fp32 in/out around an fp64 carry (f32_mixed)"
        steps, warm = 5, 3
    else:
        code, tdt, mi, N, io, carry = _capi.F32_MIXED, torch.float32, 7, 8192, 4, 8
        chunks = [torch.from_numpy(synth.audio_frames()).to(dev)]
        what = f"configs[3]: 48 kHz audio, 10 min, {chunks[0].shape[0]} frames of 8192, fixed 8 iterations, f32_mixed"
        steps, warm = 10, 3
    if not chunks:
        chunks = []
    Smax = max((c.shape[0] for c in chunks), default=1)
    plans = {c.shape[0]: get_plan(local_rank, c.shape[0], N, code, mi, 2, 0) for c in chunks}
    rows = mi + 2
    need = Smax * rows * N
    if rot_buf is not None and rot_buf.dtype == tdt and rot_buf.numel() >= need:
        rot = rot_buf.view(-1)[:need].view(Smax, rows, N)
    else:
        rot = torch.empty((Smax, rows, N), dtype=tdt, device=dev)
    ints = [torch.empty(Smax * (rows if i == 1 else 1), dtype=torch.int32, device=dev) for i in range(5)]
    nr_all = [torch.empty(c.shape[0], dtype=torch.int32, device=dev) for c in chunks]
    st_all = [torch.empty(c.shape[0], dtype=torch.int32, device=dev) for c in chunks]

    def step():
        for c, nra, sta in zip(chunks, nr_all, st_all):
            plans[c.shape[0]].decompose_device(c.data_ptr(), rot.data_ptr(), None, nra.data_ptr(), ints[1].data_ptr(),
                                               ints[2].data_ptr(), ints[3].data_ptr(), sta.data_ptr(), stream.cuda_stream)

    ms = timed_passes(torch, shard, step, stream, dev, steps, warm, world, dist)
    tot = torch.tensor([sum(c.shape[0] for c in chunks), sum(int(t.long().sum()) for t in nr_all),
                        max((int(t.abs().max()) for t in st_all), default=0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot[:2])
        dist.all_reduce(tot[2:], op=dist.ReduceOp.MAX)
    S_total, levels_total, st_max = int(tot[0]), int(tot[1]), int(tot[2])
    # algorithmic bytes: per executed level read X (carry type; the first level reads the io type),
    # write R (io type), write B (carry type)
    alg = levels_total * N * (carry + io + carry) - S_total * N * (carry - io)
    path = next(iter(plans.values())).path if plans else ("none", 1)
    out = {"workload": what, "value": S_total * N / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "n_gpus": world, "scaling": "strong" if which == 5 else "single GPU", "signals": S_total, "n_samples": N,
           "rows_mean": levels_total / max(S_total, 1), "kernel_path": path[0], "status_max": st_max,
           "roofline_whole_step": {"bound": "hbm", "algorithmic_bytes_per_step": alg, "unit": "GB/s per GPU",
                                   "achieved": alg / (ms * 1e-3) / 1e9 / world, "peak": peak,
                                   "frac": alg / (ms * 1e-3) / 1e9 / world / peak}}
    del chunks, plans, rot, ints
    return out

def config1_leg(torch, dev):
    """BASELINE.json configs[0]: ONE 65 536-sample chirp+noise signal, the reference's own use (ITD.py:500-503).  Device time
    of one decomposition (CUDA events around single calls, median), the wall time of the drop-in ITD().itd(x) call with host
    arrays in and out, the oracle's C port on one host core beside it, and bit-exact parity of the rows."""
    import numpy as np
    import pyitd_b200
    from oracle import itd_oracle as o
    from pyitd_b200 import _capi, synth
    from pyitd_b200.itd import get_plan
    x = synth.config1_chirp()
    N, mi = int(x.shape[0]), 20
    plan = get_plan(dev.index, 1, N, _capi.F64, mi, 2, 0)
    xg = torch.from_numpy(x).to(dev).view(1, N)
    rot = torch.empty((1, plan.rows, N), dtype=torch.float64, device=dev)
    ints = [torch.zeros(plan.rows if i == 1 else 1, dtype=torch.int32, device=dev) for i in range(5)]
    st = torch.cuda.current_stream(dev)

    def step():
        plan.decompose_device(xg.data_ptr(), rot.data_ptr(), None, ints[0].data_ptr(), ints[1].data_ptr(),
                              ints[2].data_ptr(), ints[3].data_ptr(), ints[4].data_ptr(), st.cuda_stream)

    for _ in range(5):
        step()
    torch.cuda.synchronize(dev)
    us = []
    for _ in range(100):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        step()
        e1.record(st)
        torch.cuda.synchronize(dev)
        us.append(e0.elapsed_time(e1) * 1e3)
    us.sort()
    launches, path = plan.launches, plan.path[0]
    itd = pyitd_b200.ITD()
    itd.itd(x, max_iteration=mi)
    wall = []
    for _ in range(20):
        t0 = time.perf_counter()
        rows = itd.itd(x, max_iteration=mi)
        wall.append(time.perf_counter() - t0)
    wall.sort()
    t0 = time.perf_counter()
    want = o.c_decompose(x, mi)
    cpu_s = time.perf_counter() - t0
    return {"workload": "configs[0]: ONE 65536-sample chirp+noise signal, fp64, knot-count stop (max_iteration=20)",
            "kernel_path": path, "launches_per_call": launches, "rows": int(rows.shape[0]),
            "device_us_median": us[len(us) // 2], "device_us_min": us[0],
            "value": N / (us[len(us) // 2] * 1e-6), "unit": UNIT,
            "dropin_call_ms_median": wall[len(wall) // 2] * 1e3,
            "dropin_api": "pyitd_b200.ITD().itd(numpy x): host array in, host rows out (H2D + one launch + D2H inside)",
            "cpu_port_one_core_ms": cpu_s * 1e3,
            "parity": {"bit_exact": bool(rows.tobytes() == want.rotations.tobytes()),
                       "checked": "every row of the drop-in call vs oracle/itd_oracle.c (ITD.py:351-433)"}}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    protect_stdout()
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import pyitd_b200
    from pyitd_b200 import _capi, synth
    from pyitd_b200.itd import get_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # the pinned host buffers of the e2e leg should live next to this rank's GPU, not wherever torchrun started the
    # process (SCALE_r01: eight D2H streams into one NUMA node); the CPU baseline leg gets the full mask back
    from pyitd_b200 import shard
    cpu_mask = os.sched_getaffinity(0)
    bound = None if args.no_bind else shard.bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log lines (e.g. the "NCCL version" banner) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    S, N = args.channels, args.samples
    x = synth.eeg_like(S, N, seed=SEED + rank, device=dev)
    plan = get_plan(local_rank, S, N, _capi.F64, MAX_ITERATION, 2, 0)
    if args.groups > 0:
        plan.groups = args.groups
    groups = plan.groups
    args.groups_used = groups
    rows = plan.rows
    rot = torch.empty((S, rows, N), dtype=torch.float64, device=dev)
    n_rows = torch.empty(S, dtype=torch.int32, device=dev)
    counts = torch.empty((S, rows), dtype=torch.int32, device=dev)
    iknots = torch.empty(S, dtype=torch.int32, device=dev)
    kind = torch.empty(S, dtype=torch.int32, device=dev)
    status = torch.empty(S, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        plan.decompose_device(x.data_ptr(), rot.data_ptr(), None, n_rows.data_ptr(), counts.data_ptr(),
                              iknots.data_ptr(), kind.data_ptr(), status.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    assert int(status.abs().max()) == 0, "synthetic workload hit an error status"
    launches_per_step = plan.launches
    pair_stats = plan.sweep_stats() if plan.path[0] == "sweep" else None

    # ---- timed region: exactly K steps, CUDA events on the launching stream ----------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * S * N / (ms_per_step * 1e-3)

    # ---- parity of the timed output (every rank checks its own shard; rank 0 reports the worst) ----
    parity = parity_check(torch, x, rot, n_rows, counts, args.parity_channels, seed=SEED + 77 + rank)
    pbad = torch.tensor([0 if parity["bit_exact"] else 1], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(pbad, op=dist.ReduceOp.MAX)
    parity["all_ranks_bit_exact"] = int(pbad.item()) == 0
    parity["channels_all_ranks"] = parity["channels"] * world

    # ---- roofline of the dominant kernel, per-launch CUDA events on the launching stream ----------
    # (a) launch chains serialised (one group): every launch timed alone; (b) with the library's launch groups the
    # chains overlap, so the call is timed fork-to-join and the level kernel's share is taken from its bytes
    plan.enable_timing(True)
    lvl_ms = None
    reps = 3
    plan.groups = 1
    for _ in range(reps):
        step()
        tms = plan.launch_times_ms()
        lvl_ms = tms if lvl_ms is None else [a + b for a, b in zip(lvl_ms, tms)]
    lvl_ms = [v / reps for v in lvl_ms]
    plan.groups = groups
    span_ms = None
    if groups > 1:
        span_ms = 0.0
        for _ in range(reps):
            step()
            span_ms += plan.launch_times_ms()[0] / reps
    plan.enable_timing(False)
    nr = n_rows.long()
    rows_mean = float(nr.double().mean())
    active = [int((nr >= e + 1).sum()) for e in range(rows)]        # signals that execute extraction e
    level_bytes = [a * N * 24 for a in active]                      # read X + write R + write B, fp64
    alg_bytes = sum(level_bytes)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    path, cluster = plan.path
    # dram__bytes of the dominant kernel come from an ncu --set full capture (profiles/run_ncu_traffic.sh writes
    # profiles/dominant_kernel_traffic.json with the hash of the CUDA sources it profiled); a capture of another build
    # says nothing about this one, so the field is dropped (null) when the hashes differ
    src_hash = kernel_source_hash()
    traffic = None
    traffic_note = "no ncu capture of this build under profiles/ (source hash %s): not reported" % src_hash
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        if cap.get("source_hash") == src_hash and cap.get("dram_bytes_per_launch"):
            traffic = float(cap["dram_bytes_per_launch"])
            traffic_note = cap.get("how", "ncu --set full capture of this build")
        elif cap.get("source_hash"):
            traffic_note = ("the committed ncu capture is of source hash %s, this build is %s: dropped"
                            % (cap.get("source_hash"), src_hash))
    except Exception:
        pass
    sample_levels = int(nr.sum()) * N
    if path in ("resident", "regres"):
        # ONE launch decomposes the whole batch; the carry never leaves the SMs, so the bytes that MUST
        # cross HBM are: x once + every produced row once
        k_ms = lvl_ms[0]
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        compulsory = (S + int(nr.sum())) * N * 8
        roofline = {
            "bound": "hbm", "kernel": f"pyitd::{path}_kernel<double,double,double> (cluster of {cluster} CTAs per signal)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
            "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_ms, "level_launches": 1,
            "algorithmic_bytes_definition": "24 B per sample per executed level (read X, write R, write B: SURVEY.md 8d, "
                                            "an HBM-resident carry); this kernel keeps the carry on chip",
            "compulsory_bytes_per_launch": compulsory,
            "achieved_compulsory_GBps": compulsory / (k_ms * 1e-3) / 1e9,
            "frac_compulsory": compulsory / (k_ms * 1e-3) / 1e9 / peak,
            "active_signals_per_level": active,
            "sample_levels_per_s": world * sample_levels / (ms_per_step * 1e-3),
        }
    else:
        # launch 0 is the knot scan, launches 1..rows are extractions 0..rows-1, the last is the fix-up
        lv_times = lvl_ms[1:1 + rows]
        n_lv = sum(1 for a in active if a > 0)
        lv_time_ms = sum(tm for tm, a in zip(lv_times, active) if a > 0)
        achieved = alg_bytes / (lv_time_ms * 1e-3) / 1e9
        kname = {"stream": "level_stream_kernel", "sweep": "sweep_kernel"}.get(path, "level_kernel")
        roofline = {
            "bound": "hbm", "kernel": f"pyitd::{kname}<double,double,double>",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
            "algorithmic_bytes_per_launch": alg_bytes / max(n_lv, 1),
            "avg_launch_ms": lv_time_ms / max(n_lv, 1), "level_launches": n_lv,
            "per_level": [{"e": e, "active_signals": a, "ms": round(tm, 4),
                           "GBps": (round(b / (tm * 1e-3) / 1e9, 1) if a > 0 and tm > 0 else None)}
                          for e, (a, tm, b) in enumerate(zip(active, lv_times, level_bytes))],
            "knot_scan_ms": lvl_ms[0], "sample_levels_per_s": world * sample_levels / (ms_per_step * 1e-3),
        }
        if path == "sweep":
            # ONE persistent launch does the knot scan and every extraction of every signal: the launch duration is the
            # step time.  Only the extractions' bytes are credited (24 B per sample per executed level); the
            # scan stage's 8 B per input sample ride along uncredited.  per_level: the same kernel launched once per
            # stage (measurement mode), each launch timed alone.
            ach = alg_bytes / (ms_per_step * 1e-3) / 1e9
            roofline["per_stage_launches"] = {k: roofline[k] for k in ("achieved", "frac", "avg_launch_ms", "level_launches")}
            roofline.update({
                "achieved": ach, "frac": ach / peak, "frac_of_nominal_8000": ach / 8000.0,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": ms_per_step, "level_launches": 1,
                "achieved_definition": "algorithmic bytes of all extractions of the step (24 B per sample-level) / CUDA-event "
                                       "duration of the one sweep_kernel launch that runs the knot scan and every level"})
            if pair_stats is not None:
                # fused pairs of extractions never store the baseline between them: such a pair MOVES 32 B per sample for 48
                # algorithmic bytes, so `achieved` can exceed the DRAM throughput of the launch (see `traffic`)
                roofline["fused_pairs_per_step"] = {
                    "fused": pair_stats[0], "not_tried_after_prediction": pair_stats[1], "failed_check_redone": pair_stats[2],
                    "bytes_saved_per_step": pair_stats[0] * N * 16,
                    "note": "two consecutive few-knot extractions in one pass: 32 B per sample instead of 48 (DESIGN 4)"}
        if span_ms is not None:
            # grouped launch chains: all level launches + the knot scans inside one fork-to-join span
            ach = alg_bytes / (span_ms * 1e-3) / 1e9
            roofline["serial_one_group"] = {k: roofline[k] for k in ("achieved", "frac", "avg_launch_ms")}
            roofline.update({
                "achieved": ach, "frac": ach / peak, "frac_of_nominal_8000": ach / 8000.0,
                "avg_launch_ms": span_ms / max(n_lv, 1), "launch_groups": groups, "call_span_ms": span_ms,
                "achieved_definition": "level-kernel algorithmic bytes of the call / fork-to-join CUDA-event time of the "
                                       "call (the knot-scan launches run inside that span and are not credited)"})
    roofline["traffic_source"] = traffic_note
    roofline["kernel_source_hash"] = src_hash

    # ---- e2e: host buffers through the C ABI (pyitd_decompose_host) -------------------------------
    e2e = None
    if not args.no_e2e:
        Se = min(args.e2e_channels, S)
        hplan = get_plan(local_rank, Se, N, _capi.F64, MAX_ITERATION, 2, 0) if Se != S else plan
        hx = x[:Se].cpu().pin_memory()
        hrot = torch.empty((Se, rows, N), dtype=torch.float64).pin_memory()
        hn = torch.empty(Se, dtype=torch.int32).pin_memory()
        hc = torch.empty((Se, rows), dtype=torch.int32).pin_memory()
        hs = torch.empty(Se, dtype=torch.int32).pin_memory()

        def host_step():
            hplan.decompose_host(hx.data_ptr(), hrot.data_ptr(), None, hn.data_ptr(), hc.data_ptr(), None, None,
                                 hs.data_ptr())

        for _ in range(2):
            host_step()
        barrier()
        k_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            host_step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = hx.numel() * 8
        # only the rows each channel produced come back (rows beyond n_rows are unspecified by contract)
        d2h = int(hn.long().sum()) * N * 8 + (hn.numel() + hc.numel() + hs.numel()) * 4
        e2e = {"value": world * Se * N * k_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "channels_per_step": Se, "steps": k_e2e,
               "api": "pyitd_decompose_host (C ABI, pinned host buffers; chunked H2D/kernel/D2H pipeline, "
                      "every produced rotation row copied back)"}
        # ceiling of this call: the raw pinned D2H rate with ALL ranks copying at once (the call returns ~86 bytes
        # for every 8-byte input sample, so PCIe / the host memory path bounds it, not the kernels)
        n_probe = min(hrot.numel(), rot.numel(), 1 << 30)            # <= 8 GiB per rank
        src = rot.view(-1)[:n_probe]
        dst = hrot.view(-1)[:n_probe]
        dst.copy_(src, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        dtp = time.perf_counter() - t0
        tp = torch.tensor([dtp], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        ceil_gbs = world * 2 * n_probe * 8 / float(tp.item()) / 1e9
        d2h_gbs = world * d2h * k_e2e / dt / 1e9
        e2e.update({"d2h_gbs": d2h_gbs, "d2h_ceiling_gbs": ceil_gbs, "pcie_frac": d2h_gbs / ceil_gbs,
                    "d2h_ceiling_how": f"{world} rank(s) each copying {n_probe * 8 >> 20} MiB device -> pinned host twice, "
                                       "concurrently, max over ranks (aggregate GB/s)",
                    "bytes_returned_per_input_sample": d2h / (Se * N),
                    "numa_bound_cpus": (f"{bound[0]}-{bound[-1]} ({len(bound)} cores)" if bound else None)})
        del hx, hrot

    # ---- the other BASELINE.json configs, reported beside the headline ------------------------------
    extra = None
    if not args.no_extra:
        extra = {}
        # configs[4]: 65 536 channels as a STRONG split over the ranks (the weak-scaling headline above is unchanged)
        try:
            extra["config5"] = extra_config(torch, dist, 5, args, rank, world, local_rank, dev, peak, rot_buf=rot)
        except Exception as ex:                                        # never lose the headline line to an extra leg
            extra["config5"] = {"error": repr(ex)}
        if world == 1:
            del x
            try:
                pyitd_b200.clear_plan_cache()
                extra["config1"] = config1_leg(torch, dev)
            except Exception as ex:
                extra["config1"] = {"error": repr(ex)}
            for which, name in ((4, "config4"), (3, "config3")):
                try:
                    pyitd_b200.clear_plan_cache()
                    torch.cuda.empty_cache()
                    extra[name] = extra_config(torch, dist, which, args, rank, world, local_rank, dev, peak)
                except Exception as ex:
                    extra[name] = {"error": repr(ex)}
        pyitd_b200.clear_plan_cache()

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    try:
        os.sched_setaffinity(0, cpu_mask)
    except OSError:
        pass
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, n_ch, secs = cpu_port_throughput(N)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_ch} channels x {N} samples of the same generator in {secs:.1f} s "
                         f"(oracle/itd_oracle.c, one channel per pthread; the same sample as one step of --impl reference)",
               "numba_per_core": NUMBA_REFERENCE_PER_CORE}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "rows_per_channel_mean": rows_mean, "parity": parity, "extra_configs": extra,
            "notes": "value / roofline run with options = 0 (rotation + trend rows only; the drop-in ITD.itd also stores "
                     "the baselines, +8 B per sample-level); --impl reference times the oracle's C port, not numba",
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if not parity["all_ranks_bit_exact"]:
        raise SystemExit("bench.py: the timed output differs from the oracle on channels %s" % parity["mismatching_channels"])


if __name__ == "__main__":
    main()
