"""Golden fixtures for the post-decomposition analytics, FROM THE REFERENCE.

Runs only in the build container.  Usage:  python tests/golden/make_golden_analytics.py

Imports MEITD.py unmodified and records ``weighted_permutation_entropy(x, order=3, normalize=...)``
(MEITD.py:79-128) for a set of inputs, and ``helperfunctions.shewchuk`` -- restated inline here as
``[math.fsum(a[:, i]) for i]`` because helperfunctions.py does not import (it uses numpy without importing it) --
plus ITD.py:475-481's total for the rows of the notebook golden vector's decomposition.
"""
from __future__ import annotations

import contextlib
import io
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
with contextlib.redirect_stdout(io.StringIO()):
    import MEITD as me  # noqa: E402


def main():
    rng = np.random.default_rng(404)
    out = {}
    signals = {
        "white_3": rng.standard_normal(3), "white_4": rng.standard_normal(4), "white_100": rng.standard_normal(100),
        "white_65536": rng.standard_normal(65536), "walk_5000": np.cumsum(rng.standard_normal(5000)),
        "ties_3000": np.round(rng.standard_normal(3000) * 2) / 2, "sine_2000": np.sin(np.arange(2000) * 0.05),
        "ramp_500": np.arange(500.0), "tiny_777": 1e-200 * rng.standard_normal(777),
    }
    for name, x in signals.items():
        x = np.ascontiguousarray(x, dtype=np.float64)
        with np.errstate(all="ignore"):
            out[f"{name}/x"] = x
            out[f"{name}/wpe_norm"] = np.asarray(me.weighted_permutation_entropy(x, order=3, normalize=True))
            out[f"{name}/wpe_raw"] = np.asarray(me.weighted_permutation_entropy(x, order=3, normalize=False))
        print(name, float(out[f"{name}/wpe_norm"]), float(out[f"{name}/wpe_raw"]))
    rows = np.load(os.path.join(HERE, "notebook_8000.npz"))["rotations"]
    s = np.array([math.fsum(rows[:, i]) for i in range(rows.shape[1])])       # helperfunctions.py:5-8
    out["notebook_rows/column_fsum"] = s
    out["notebook_rows/total"] = np.asarray(math.fsum(s))                      # ITD.py:481
    wide = rng.standard_normal((13, 4096)) * 10.0 ** rng.integers(-12, 12, (13, 4096))
    out["wide/rows"] = wide
    out["wide/column_fsum"] = np.array([math.fsum(wide[:, i]) for i in range(wide.shape[1])])
    out["wide/total"] = np.asarray(math.fsum(out["wide/column_fsum"]))
    np.savez_compressed(os.path.join(HERE, "analytics_cases.npz"), **out)


if __name__ == "__main__":
    main()
