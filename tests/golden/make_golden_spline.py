"""Golden fixtures for the cubic-spline baseline variant, FROM THE REFERENCE.

Runs only in the build container (the reference is mounted at /root/reference; scipy 1.18.1 and
numba 0.65 are in the image).  Usage:  python tests/golden/make_golden_spline.py

Imports, unmodified, ``itd_baseline_extract_modified`` (numba_accelerated_itd.py:183-211) and
``itd_baseline_extract`` (MEITD.py:303-338) and records in spline_cases.npz, per case:
``x``, the baseline of the first, the (rotation, baseline) pair of the second, and the knot
count; spline_errors.json records what the reference raises on inputs with too few knots.
The config-1 chirp is followed down five levels (each level's input is the reference's own
previous baseline) with head / tail / checksum samples only, to keep the file small.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import numba_accelerated_itd as na  # noqa: E402
with contextlib.redirect_stdout(io.StringIO()):
    import MEITD as me  # noqa: E402

from pyitd_b200 import synth  # noqa: E402


def knots_of(x):
    a = na.matlab_detect_peaks(x.copy())
    b = na.matlab_detect_peaks(-x)
    return np.sort(np.hstack((a, b))).astype(np.int64)


def main():
    rng = np.random.default_rng(77)
    out = {}

    def add(name, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        R, B = me.itd_baseline_extract(x.copy())                 # MEITD.py:303
        Bm = na.itd_baseline_extract_modified(x.copy())          # numba_accelerated_itd.py:183
        out[f"{name}/x"] = x
        out[f"{name}/R"] = np.array(R)
        out[f"{name}/B"] = np.array(B)
        out[f"{name}/B_modified"] = np.array(Bm)
        out[f"{name}/K"] = np.asarray(knots_of(x).shape[0])
        print(name, x.shape, int(out[f"{name}/K"]), float(np.abs(np.array(B) - np.array(Bm)).max()))

    add("white_64", rng.standard_normal(64))
    add("white_1000", rng.standard_normal(1000))
    add("white_4096", rng.standard_normal(4096))
    add("walk_3000", np.cumsum(rng.standard_normal(3000)))
    add("plateaus_1500", np.repeat(rng.standard_normal(300), 5) + 1e-3 * np.arange(1500))
    add("two_tone_2047", np.sin(np.arange(2047) * 0.37) + 0.3 * np.sin(np.arange(2047) * 2.1))
    add("sparse_knots_5000", np.sin(2 * np.pi * 3.3 * np.linspace(0, 1, 5000)) + 1e-3 * np.linspace(0, 1, 5000))
    add("uneven_4000", np.concatenate((rng.standard_normal(2000), np.sin(np.arange(2000) * 0.01))))
    add("k2_exact_8", np.array([0.0, 1.0, 0.5, 0.25, 2.0, 3.0, 3.5, 4.0]))      # K = 2: one cubic through 4 points
    add("k3_9", np.array([0.0, 1.0, 0.5, 0.75, 0.2, 0.1, 0.3, 0.6, 1.5]))
    add("few_knots_below_10", np.array([0.0, 2.0, 1.0, 3.0, 2.5, 4.0, 5.0, 6.0, 7.5, 9.0, 11.0, 12.0]))  # modified returns x
    add("eeg_like_8192", synth.eeg_like(1, 8192, seed=5)[0] if hasattr(synth, "eeg_like") else rng.standard_normal(8192))
    np.savez_compressed(os.path.join(HERE, "spline_cases.npz"), **out)

    # config-1 chirp followed down the levels: each level's input is the reference's previous baseline
    x = synth.config1_chirp()
    chain = {"x_sha_input": np.asarray(0)}
    cur = x
    for lev in range(6):
        R, B = me.itd_baseline_extract(cur.copy())
        R, B = np.array(R), np.array(B)
        chain[f"{lev}/K"] = np.asarray(knots_of(cur).shape[0])
        chain[f"{lev}/B_head"] = B[:512].copy()
        chain[f"{lev}/B_tail"] = B[-512:].copy()
        chain[f"{lev}/B_every_64"] = B[::64].copy()
        chain[f"{lev}/B_norm"] = np.asarray(np.linalg.norm(B))
        chain[f"{lev}/R_norm"] = np.asarray(np.linalg.norm(R))
        print("config1 level", lev, int(chain[f"{lev}/K"]), float(chain[f"{lev}/B_norm"]))
        cur = B
    del chain["x_sha_input"]
    np.savez_compressed(os.path.join(HERE, "spline_config1_chain.npz"), **chain)

    errors = []
    for name, xx in (("monotone_100", np.arange(100.0)), ("one_knot", np.array([0.0, 1.0, 0.5, 0.2, 0.1])),
                     ("ones_50", np.ones(50))):
        rec = dict(name=name, x=np.asarray(xx, dtype=np.float64).tolist())
        try:
            me.itd_baseline_extract(np.asarray(xx, dtype=np.float64))
            rec["raises"] = None
        except Exception as e:  # noqa: BLE001
            rec["raises"] = type(e).__name__
            rec["message"] = str(e)[:80]
        try:
            r = na.itd_baseline_extract_modified(np.asarray(xx, dtype=np.float64))
            rec["modified_returns_input"] = bool(np.array_equal(np.array(r), np.asarray(xx, dtype=np.float64)))
        except Exception as e:  # noqa: BLE001
            rec["modified_raises"] = type(e).__name__
        errors.append(rec)
        print("error case", rec["name"], rec.get("raises"), rec.get("modified_returns_input"))
    json.dump(errors, open(os.path.join(HERE, "spline_errors.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
