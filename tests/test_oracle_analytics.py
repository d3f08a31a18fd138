"""CPU: the analytics oracle against fixtures generated from the reference's own
``weighted_permutation_entropy`` (MEITD.py:79-128) and ``math.fsum`` column sums (helperfunctions.py:2-9,
ITD.py:475-481).  Both restatements are bit-exact (NaN where the reference returns NaN)."""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_cases
from oracle import itd_oracle as o


def same(a, b):
    a, b = float(a), float(b)
    return (math.isnan(a) and math.isnan(b)) or (a == b and math.copysign(1, a) == math.copysign(1, b))


@pytest.mark.parametrize("impl", [o.c_wpe3, o.np_wpe3], ids=["c", "numpy"])
def test_wpe_matches_reference_bit_for_bit(impl):
    cases = load_cases(os.path.join(GOLDEN, "analytics_cases.npz"))
    n = 0
    for name, c in cases.items():
        if "wpe_norm" not in c:
            continue
        n += 1
        assert same(impl(c["x"], True), c["wpe_norm"]), name
        assert same(impl(c["x"], False), c["wpe_raw"]), name
    assert n >= 9


def test_column_fsum_matches_math_fsum():
    cases = load_cases(os.path.join(GOLDEN, "analytics_cases.npz"))
    rows = np.load(os.path.join(GOLDEN, "notebook_8000.npz"))["rotations"]
    assert np.array_equal(o.c_column_fsum(rows), cases["notebook_rows"]["column_fsum"])
    assert np.array_equal(o.c_column_fsum(cases["wide"]["rows"]), cases["wide"]["column_fsum"])
    # the reference's reconstruction check (ITD.py:505-508) on its own golden vector: difference 0.0 (PyITD.ipynb cell 3)
    x = np.load(os.path.join(GOLDEN, "notebook_8000.npz"))["x"]
    assert abs(np.sum(x) - float(cases["notebook_rows"]["total"])) < 1e-12
