"""CPU: the 2-D crossways / ensemble oracle against fixtures produced by executing code cell 1
of the reference's siftED2D.ipynb (tests/golden/make_golden_sift2d.py).  Tolerance 1e-11 relative L2: every 1-D
pass is the spline level pinned in test_oracle_spline.py (FITPACK vs moment equations, ~1e-15 per pass)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import itd_oracle as o


def cases(kind):
    z = np.load(os.path.join(GOLDEN, "sift2d_cases.npz"))
    out = {}
    for k in z.files:
        a, name, field = k.split("/")
        if a == kind:
            out.setdefault(name, {})[field] = z[k]
    return out


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_crossways_matches_notebook():
    cs = cases("crossways")
    assert len(cs) == 4
    for name, c in cs.items():
        y = o.crossways(c["x"])
        assert y.shape == c["y"].shape
        assert rel(y, c["y"]) < 1e-11, (name, rel(y, c["y"]))


def test_crossways_numpy_level_agrees():
    c = cases("crossways")["smooth_40x33"]
    y = o.crossways(c["x"], level=o.np_spline_level)
    assert rel(y, c["y"]) < 1e-11


def test_ensemble_matches_notebook_composition():
    for name, c in cases("ensemble").items():
        low = o.ensemble2d(c["x"], c["noise"])
        assert rel(low, c["lowpass"]) < 1e-11, name
        # antithetic pairs cancel the noise to first order: the low-pass stays close to a plain crossways pass
        assert rel(low, o.crossways(c["x"])) < 0.2
