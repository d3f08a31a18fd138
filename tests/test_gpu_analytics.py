"""GPU: post-decomposition analytics through the C ABI.

* column sums: BIT-EXACT against math.fsum (the kernel runs CPython's fsum algorithm per column);
* totals: double-double, within 1 ulp of math.fsum of the column sums;
* weighted permutation entropy: the reference sums its weighted counts sequentially in float64 (rounding noise up to
  ~N eps), the kernel reduces in parallel with a double-double combine: tolerance 1e-9 absolute on the entropy (north
  star's fp64 tolerance; measured ~1e-14), NaN / -0.0 cases reproduced."""
import math
import os

import numpy as np
import pytest
import torch

import pyitd_b200
from conftest import GOLDEN, load_cases
from oracle import itd_oracle as o
from pyitd_b200 import analytics, synth

pytestmark = pytest.mark.gpu
WPE_TOL = 1e-9


def test_wpe_golden_dropin():
    cases = load_cases(os.path.join(GOLDEN, "analytics_cases.npz"))
    worst = 0.0
    for name, c in cases.items():
        if "wpe_norm" not in c:
            continue
        for norm, key in ((True, "wpe_norm"), (False, "wpe_raw")):
            got = pyitd_b200.weighted_permutation_entropy(c["x"], order=3, normalize=norm)
            want = float(c[key])
            if math.isnan(want):
                assert math.isnan(got), name
            else:
                assert abs(got - want) < WPE_TOL, (name, got, want)
                worst = max(worst, abs(got - want))
                if want == 0.0:
                    assert math.copysign(1, got) == math.copysign(1, want), name
    print("worst |wpe - reference|:", worst)


def test_wpe_rows_of_a_decomposition():
    x = synth.eeg_like(64, 8192, seed=3, device="cuda")
    res = pyitd_b200.decompose(x, max_iteration=11)
    w = analytics.wpe_rows(res)
    assert w.shape == res.rotations.shape[:2]
    wn = w.cpu().numpy()
    rot = res.rotations.cpu().numpy()
    nr = res.n_rows.cpu().numpy()
    for s in (0, 17, 63):
        for r in range(res.rotations.shape[1]):
            if r >= nr[s]:
                assert math.isnan(wn[s, r])
            else:
                want = o.c_wpe3(rot[s, r], True)
                assert (math.isnan(want) and math.isnan(wn[s, r])) or abs(wn[s, r] - want) < WPE_TOL, (s, r)


@pytest.mark.parametrize("n", [3, 4, 5, 255, 256, 257, 258, 1000, 70001])
def test_wpe_sizes_and_f32(n):
    rng = np.random.default_rng(n)
    X = np.stack([rng.standard_normal(n), np.round(rng.standard_normal(n) * 2) / 2, np.cumsum(rng.standard_normal(n))])
    got = analytics.wpe_rows(torch.from_numpy(X).cuda(), normalize=True).cpu().numpy()
    for s in range(3):
        assert abs(got[s] - o.c_wpe3(X[s], True)) < WPE_TOL, (n, s)
    X32 = X.astype(np.float32)
    got32 = analytics.wpe_rows(torch.from_numpy(X32).cuda(), normalize=False).cpu().numpy()
    for s in range(3):
        assert abs(got32[s] - o.c_wpe3(X32[s].astype(np.float64), False)) < WPE_TOL, (n, s)


def test_column_fsum_bit_exact_and_dropins():
    cases = load_cases(os.path.join(GOLDEN, "analytics_cases.npz"))
    rows = np.load(os.path.join(GOLDEN, "notebook_8000.npz"))["rotations"]
    assert np.array_equal(pyitd_b200.shewchuk(rows), cases["notebook_rows"]["column_fsum"])
    assert np.array_equal(pyitd_b200.shewchuk(cases["wide"]["rows"]), cases["wide"]["column_fsum"])
    for name in ("notebook_rows", "wide"):
        src = rows if name == "notebook_rows" else cases["wide"]["rows"]
        total, want = pyitd_b200.shewchuk_sum(src), float(cases[name]["total"])
        assert abs(total - want) <= abs(np.spacing(want)), (name, total, want)


def test_column_fsum_ragged_rows_and_cancellation():
    rng = np.random.default_rng(8)
    S, R, N = 5, 22, 3001
    a = rng.standard_normal((S, R, N)) * 10.0 ** rng.integers(-15, 15, (S, R, N))
    a[:, 1] = -a[:, 0]                                     # exact cancellation of huge terms
    nr = np.array([22, 1, 2, 13, 0], dtype=np.int32)
    sums, totals = analytics.column_fsum(torch.from_numpy(a).cuda(), torch.from_numpy(nr).cuda())
    sums = sums.cpu().numpy()
    for s in range(S):
        want = np.array([math.fsum(a[s, : nr[s], t]) for t in range(N)])
        assert np.array_equal(sums[s], want), s
        wt = math.fsum(want)
        assert abs(float(totals[s]) - wt) <= abs(np.spacing(wt)) if wt != 0 else float(totals[s]) == 0.0


def test_reconstruction_error_of_a_batch():
    # ITD.py:505-508 on a batch: |sum(x) - shewchuk_sum(rows)| stays at rounding level (the golden run records 0.0)
    x = synth.eeg_like(32, 65536, seed=9, device="cuda")
    res = pyitd_b200.decompose(x, max_iteration=11)
    err = analytics.reconstruction_error(x, res).cpu().numpy()
    scale = x.abs().sum(dim=1).cpu().numpy()
    assert (err <= 1e-15 * scale).all(), (err / scale).max()
    # and against the CPU for one signal
    rows = res.rows_of(5).cpu().numpy()
    want = math.fsum(o.c_column_fsum(rows))
    _, totals = analytics.column_fsum(res.rotations, res.n_rows)
    assert abs(float(totals[5]) - want) <= abs(np.spacing(want))
