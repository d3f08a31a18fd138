"""GPU: the spline-baseline level (pyitd_extract_spline_device) against the reference's own
outputs (tests/golden/spline_*.npz, generated from MEITD.py / numba_accelerated_itd.py) and the oracle.

Tolerance: the reference's spline solve is scipy/FITPACK (Givens QR on the B-spline collocation matrix); the
kernels solve the same not-a-knot interpolation problem by truncated parallel cyclic reduction, so parity is
1e-9 relative L2 (north star's fp64 tolerance; measured ~1e-15), knot counts exact.  The fp32-I/O variant
computes in float64 and is held to float32 rounding of the oracle (1e-6 relative L2)."""
import json
import os

import numpy as np
import pytest
import torch

import pyitd_b200
from conftest import GOLDEN, load_cases
from oracle import itd_oracle as o
from pyitd_b200 import _capi, synth

pytestmark = pytest.mark.gpu

TOL = 1e-9


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def gpu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_spline_golden_cases_dropin():
    cases = load_cases(os.path.join(GOLDEN, "spline_cases.npz"))
    worst = 0.0
    for name, c in cases.items():
        R, B = pyitd_b200.itd_baseline_extract_spline(c["x"])          # MEITD.py:303
        assert R.dtype == np.float64 and B.shape == c["x"].shape
        worst = max(worst, rel(B, c["B"]))
        assert rel(B, c["B"]) < TOL, name
        assert np.abs(R - c["R"]).max() < TOL * max(1.0, np.abs(c["x"]).max()), name
        Bm = pyitd_b200.itd_baseline_extract_modified(c["x"])          # numba_accelerated_itd.py:183
        if int(c["K"]) >= 10:
            assert rel(Bm, c["B_modified"]) < TOL, name
        else:
            assert np.array_equal(Bm, c["x"]), name                    # fewer than 10 extrema: input returned
    print("worst relative L2 vs the reference:", worst)


def test_spline_errors_dropin():
    for rec in json.load(open(os.path.join(GOLDEN, "spline_errors.json"))):
        x = np.asarray(rec["x"], dtype=np.float64)
        with pytest.raises(TypeError):
            pyitd_b200.itd_baseline_extract_spline(x)
        assert np.array_equal(pyitd_b200.itd_baseline_extract_modified(x), x)


def test_spline_config1_chain():
    z = np.load(os.path.join(GOLDEN, "spline_config1_chain.npz"))
    cur = gpu(synth.config1_chirp()).unsqueeze(0)
    for lev in range(6):
        R, B, cnt, st = pyitd_b200.extract_spline(cur)
        assert int(st[0]) == 0 and int(cnt[0]) == int(z[f"{lev}/K"]), lev
        b = B[0].cpu().numpy()
        assert rel(b[:512], z[f"{lev}/B_head"]) < TOL, lev
        assert rel(b[-512:], z[f"{lev}/B_tail"]) < TOL, lev
        assert rel(b[::64], z[f"{lev}/B_every_64"]) < TOL, lev
        assert np.array_equal((cur[0] - B[0]).cpu().numpy(), R[0].cpu().numpy())
        cur = B


@pytest.mark.parametrize("n", [4, 5, 7, 31, 32, 33, 127, 128, 129, 895, 896, 897, 898, 1023, 1024, 1025,
                               1791, 1792, 1793, 1794, 2047, 2048, 2049, 2689, 4096, 5000, 10001])
def test_spline_sizes_against_oracle(n):
    # sizes around the tile (1024 samples), mask-word (32) and PCR-window (896 knots) boundaries; alternating
    # signs make almost every sample a knot so that the knot count crosses the window boundaries too
    rng = np.random.default_rng(n)
    xs = [rng.standard_normal(n), np.cumsum(rng.standard_normal(n)),
          np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * (1.0 + rng.random(n))]
    X = np.stack(xs)
    R, B, cnt, st = pyitd_b200.extract_spline(gpu(X))
    R, B, cnt, st = R.cpu().numpy(), B.cpu().numpy(), cnt.cpu().numpy(), st.cpu().numpy()
    for s in range(X.shape[0]):
        try:
            Ro, Bo, K = o.c_spline_level(X[s])
        except o.OracleError as e:
            assert e.status == o.ITD_FEW_KNOTS and st[s] & _capi.ST_FEW_KNOTS
            assert np.array_equal(B[s], X[s]) and not R[s].any()
            continue
        assert st[s] == 0 and cnt[s] == K, (s, st[s], cnt[s], K)
        assert rel(B[s], Bo) < TOL, (n, s, rel(B[s], Bo))
        assert np.array_equal(R[s], X[s] - B[s])


@pytest.mark.parametrize("shape", [(1, 65536), (3, 40000), (40, 8192), (200, 4096), (300, 2048), (17, 1 << 18)])
def test_spline_every_scan_path(shape):
    # look-back, stream, resident-shaped and strided plans all feed the same spline kernels
    S, N = shape
    rng = np.random.default_rng(S * 7 + N)
    X = np.cumsum(rng.standard_normal((S, N)), axis=1) * 0.1 + rng.standard_normal((S, N))
    R, B, cnt, st = pyitd_b200.extract_spline(gpu(X))
    assert not st.any()
    for s in sorted({0, S // 2, S - 1}):
        Ro, Bo, K = o.c_spline_level(X[s])
        assert int(cnt[s]) == K
        assert rel(B[s].cpu().numpy(), Bo) < TOL, (shape, s)
        assert rel(R[s].cpu().numpy(), Ro) < 1e-7, (shape, s)


def test_spline_sparse_and_uneven_knots():
    # long knot-free stretches next to dense ones: the PCR rows stay diagonally dominant for any spacing
    rng = np.random.default_rng(3)
    n = 30000
    x = np.concatenate((rng.standard_normal(5000), np.sin(np.arange(20000) * 0.003) * 5, rng.standard_normal(5000)))
    x2 = np.sin(2 * np.pi * 3.3 * np.linspace(0, 1, n)) + 1e-3 * np.linspace(0, 1, n)      # 7 knots only
    X = np.stack([x, x2])
    R, B, cnt, st = pyitd_b200.extract_spline(gpu(X))
    for s in range(2):
        Ro, Bo, K = o.c_spline_level(X[s])
        assert int(cnt[s]) == K and int(st[s]) == 0
        assert rel(B[s].cpu().numpy(), Bo) < TOL, s


def test_spline_min_knots_and_baseline_only():
    rng = np.random.default_rng(11)
    X = rng.standard_normal((4, 3000))
    X[1] = np.arange(3000.0)                                  # no knots: baseline = x, FEW_KNOTS
    X[2, :] = np.sin(np.arange(3000) * 2 * np.pi * 2.2 / 3000)  # 4 knots (< 10)
    R, B, cnt, st = pyitd_b200.extract_spline(gpu(X), min_knots=10, want_rotation=False)
    assert R is None
    st = st.cpu().numpy()
    assert st[0] == 0 and st[3] == 0 and st[1] == _capi.ST_FEW_KNOTS and st[2] == 0
    assert np.array_equal(B[1].cpu().numpy(), X[1]) and np.array_equal(B[2].cpu().numpy(), X[2])
    assert rel(B[0].cpu().numpy(), o.c_spline_level(X[0])[1]) < TOL


def test_spline_f32_mixed():
    rng = np.random.default_rng(21)
    X32 = rng.standard_normal((5, 9000)).astype(np.float32)
    R, B, cnt, st = pyitd_b200.extract_spline(gpu(X32))
    assert R.dtype == torch.float32 and not st.any()
    for s in range(5):
        Ro, Bo, K = o.c_spline_level(X32[s].astype(np.float64))
        assert int(cnt[s]) == K
        # float32(oracle(float64(x32))) up to one float32 rounding of a 1e-15-accurate value
        assert rel(B[s].cpu().numpy().astype(np.float64), Bo.astype(np.float32).astype(np.float64)) < 1e-6
        assert rel(R[s].cpu().numpy().astype(np.float64), Ro.astype(np.float32).astype(np.float64)) < 1e-6


def test_spline_unaligned_rows_and_views():
    # rows that are only 8-byte aligned take the scalar load/store path
    rng = np.random.default_rng(31)
    for n in (1001, 4095):
        X = rng.standard_normal((3, n))
        R, B, cnt, st = pyitd_b200.extract_spline(gpu(X))
        for s in range(3):
            assert rel(B[s].cpu().numpy(), o.c_spline_level(X[s])[1]) < TOL


def test_spline_full_size_properties():
    # config-2 shape (512 channels here): size-independent properties + oracle spot checks
    x = synth.eeg_like(512, 65536, seed=1234, device="cuda")
    R, B, cnt, st = pyitd_b200.extract_spline(x)
    torch.cuda.synchronize()
    assert not st.any()
    assert torch.equal(R, x - B)                               # MEITD.py:335, exactly
    xs = x
    flags = ((xs[:, :-2] >= xs[:, 1:-1]) & (xs[:, 1:-1] < xs[:, 2:])) | ((xs[:, :-2] <= xs[:, 1:-1]) & (xs[:, 1:-1] > xs[:, 2:]))
    assert torch.equal(flags.sum(dim=1).to(torch.int32), cnt)
    # end knots: B[0] and B[N-1] are the odd-reflection means (MEITD.py:323-325)
    b0 = ((2 * x[:, 0] - x[:, 1]) + x[:, 0]) / 2
    assert torch.allclose(B[:, 0], b0, rtol=0, atol=1e-12)
    for s in (0, 255, 511):
        Ro, Bo, K = o.c_spline_level(x[s].cpu().numpy())
        assert rel(B[s].cpu().numpy(), Bo) < TOL
