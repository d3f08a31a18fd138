"""CPU: the spline-baseline oracle against fixtures generated from the reference's own
``itd_baseline_extract`` (MEITD.py:303-338) / ``itd_baseline_extract_modified`` (numba_accelerated_itd.py:183-211).

The spline solve is scipy/FITPACK's (third party): the oracle restates the same interpolation problem through
the moment equations, so the pin is a tolerance, not bit equality: 1e-12 relative L2 (measured <= 4e-15)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_cases
from oracle import itd_oracle as o
from pyitd_b200 import synth

SPLINE_TOL = 1e-12
IMPLS = [("numpy", lambda x: o.np_spline_level(x)[:2]), ("c", lambda x: o.c_spline_level(x)[:2])]


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("impl", IMPLS, ids=lambda i: i[0])
def test_spline_cases_match_reference(impl):
    cases = load_cases(os.path.join(GOLDEN, "spline_cases.npz"))
    assert len(cases) >= 10
    for name, c in cases.items():
        R, B = impl[1](c["x"])
        assert rel(B, c["B"]) < SPLINE_TOL, name
        assert np.abs(R - c["R"]).max() <= SPLINE_TOL * max(1.0, np.abs(c["x"]).max()) * 10, name
        # MEITD.py:335: rotation = x - baseline, exactly
        assert np.array_equal(R, c["x"] - B), name
        # the jitted variant returns the same baseline from 10 knots on, else its input
        # (numba_accelerated_itd.py:188-191)
        if int(c["K"]) >= 10:
            assert rel(B, c["B_modified"]) < SPLINE_TOL, name
        else:
            assert np.array_equal(c["B_modified"], c["x"]), name


def test_spline_knot_counts_and_c_equals_numpy():
    cases = load_cases(os.path.join(GOLDEN, "spline_cases.npz"))
    for name, c in cases.items():
        Rn, Bn, knots = o.np_spline_level(c["x"])
        Rc, Bc, K = o.c_spline_level(c["x"])
        assert K == len(knots) == int(c["K"]), name
        assert rel(Bc, Bn) < 1e-13, name


@pytest.mark.parametrize("impl", IMPLS, ids=lambda i: i[0])
def test_spline_config1_chain(impl):
    # the config-1 chirp followed down six levels, teacher-forced with the oracle's own baselines:
    # level inputs agree with the reference's to ~1e-15, so the knot sets must stay identical
    z = np.load(os.path.join(GOLDEN, "spline_config1_chain.npz"))
    cur = synth.config1_chirp()
    for lev in range(6):
        R, B = impl[1](cur)
        K = len(o.np_find_knots(cur))
        assert K == int(z[f"{lev}/K"]), lev
        assert rel(B[:512], z[f"{lev}/B_head"]) < 1e-10, lev
        assert rel(B[-512:], z[f"{lev}/B_tail"]) < 1e-10, lev
        assert rel(B[::64], z[f"{lev}/B_every_64"]) < 1e-11, lev
        assert abs(np.linalg.norm(B) - float(z[f"{lev}/B_norm"])) < 1e-9 * float(z[f"{lev}/B_norm"])
        cur = B


def test_spline_too_few_knots_is_the_reference_typeerror():
    errs = json.load(open(os.path.join(GOLDEN, "spline_errors.json")))
    assert errs
    for rec in errs:
        assert rec["raises"] == "TypeError"          # scipy splrep: "m > k must hold"
        x = np.asarray(rec["x"], dtype=np.float64)
        for fn in (o.np_spline_level, o.c_spline_level):
            with pytest.raises(o.OracleError) as ei:
                fn(x)
            assert ei.value.status == o.ITD_FEW_KNOTS


def test_spline_properties():
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3000)
    R, B, knots = o.np_spline_level(x)
    # the spline interpolates the knot baseline: at interior knots B equals L_k of ITD.py:106-110
    tau = np.concatenate(([0], knots, [len(x) - 1]))
    X = x[tau]
    w = (tau[1:-1] - tau[:-2]) / (tau[2:] - tau[:-2])
    L = 0.5 * (X[:-2] + w * (X[2:] - X[:-2])) + 0.5 * X[1:-1]
    assert np.abs(B[knots] - L).max() < 1e-12
    assert abs(B[0] - ((2 * x[0] - x[1]) + x[0]) / 2) < 1e-15
    assert abs(B[-1] - (x[-1] + (2 * x[-1] - x[-2])) / 2) < 1e-12
    # affine invariance: spline ITD of (a x + b) = a B + b
    R2, B2, k2 = o.np_spline_level(3.0 * x + 7.0)
    assert np.array_equal(k2, knots)
    assert rel(B2, 3.0 * B + 7.0) < 1e-12
