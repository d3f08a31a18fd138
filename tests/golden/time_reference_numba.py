#!/usr/bin/env python
"""Times the UNMODIFIED reference (/root/reference/ITD.py, Python + numba) on the workloads bench.py measures, in the build
container where it is mounted -- it cannot travel to the GPU box -- and writes tests/golden/reference_numba_timing.json, which
bench.py reports as cpu_baseline.numba_per_core (labelled with the host it was measured on).  JIT compilation excluded
(one warm-up call), one core (numba's functions are single-threaded), wall clock around ITD().itd(x).

    python tests/golden/time_reference_numba.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import platform
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import ITD as ref  # noqa: E402  (the reference itself)

from pyitd_b200 import synth  # noqa: E402


def run(x, max_iteration):
    x = np.ascontiguousarray(x, dtype=np.float64)
    ref.S = ref.T = x                      # ITD.py:375 reads module globals
    obj = ref.ITD()
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        rows = obj.itd(x.copy(), max_iteration=max_iteration)
        dt = time.perf_counter() - t0
    return dt, np.asarray(rows).shape[0]


def main():
    x1 = synth.config1_chirp()
    run(x1, 20)                            # JIT
    t1 = sorted(run(x1, 20)[0] for _ in range(5))
    xe = synth.eeg_like(8, 65536, seed=1234, device="cpu").numpy()
    te, rows = [], []
    for c in range(8):
        dt, r = run(xe[c], 11)
        te.append(dt)
        rows.append(r)
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {
        "what": "the unmodified reference ITD().itd (Python + numba, ITD.py:351-433) on one core, JIT excluded",
        "host": {"cpu": cpu, "cores": os.cpu_count(), "machine": platform.machine(), "note": "build container, not the GPU box"},
        "config1_chirp_65536": {"ms_median": 1e3 * t1[len(t1) // 2], "samples_per_s": 65536 / t1[len(t1) // 2]},
        "config2_channels_65536": {"channels": 8, "ms_mean": 1e3 * float(np.mean(te)), "rows_mean": float(np.mean(rows)),
                                   "samples_per_s_per_core": 65536 / float(np.mean(te))},
    }
    json.dump(out, open(os.path.join(HERE, "reference_numba_timing.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
