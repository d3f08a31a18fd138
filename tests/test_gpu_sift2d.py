"""GPU: 2-D crossways / ensemble ITD against the notebook's own outputs and the oracle.
Tolerance 1e-9 relative L2 (north star's fp64 tolerance; four cascaded spline passes, measured ~1e-15)."""
import os

import numpy as np
import pytest
import torch

import pyitd_b200
from conftest import GOLDEN
from oracle import itd_oracle as o
from test_oracle_sift2d import cases, rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_crossways_golden_dropin():
    for name, c in cases("crossways").items():
        y = pyitd_b200.crossways_itd_baseline_extract(c["x"])
        assert y.dtype == np.float64 and y.shape == c["y"].shape
        assert rel(y, c["y"]) < TOL, (name, rel(y, c["y"]))


def test_ensemble_golden_dropin():
    for name, c in cases("ensemble").items():
        low = pyitd_b200.retrieve_statistical_image_component(c["x"], noise=c["noise"], iterations=2 * c["noise"].shape[0])
        assert rel(low, c["lowpass"]) < TOL, (name, rel(low, c["lowpass"]))
        both = pyitd_b200.totalextract2d(c["x"], noise=c["noise"]) if c["noise"].shape[0] == 10 else None
        if both is not None:
            assert both.shape == (2,) + c["x"].shape
            assert np.array_equal(both[1], low) and np.array_equal(both[0], c["x"] - low)


@pytest.mark.parametrize("shape", [(1, 3, 3), (2, 31, 65), (3, 100, 37), (1, 200, 200), (5, 64, 64)])
def test_crossways_batch_against_oracle(shape):
    rng = np.random.default_rng(sum(shape))
    X = rng.standard_normal(shape) * 20 + 100 + 10 * np.sin(np.arange(shape[2]) * 0.7)[None, None, :]
    Y = pyitd_b200.crossways_batch(torch.from_numpy(X).cuda()).cpu().numpy()
    for b in range(shape[0]):
        want = o.crossways(X[b])
        assert rel(Y[b], want) < TOL, (shape, b, rel(Y[b], want))


def test_crossways_transpose_symmetry():
    # crossways(x.T) == crossways(x).T up to the order of the two additions (exactly commutative)
    rng = np.random.default_rng(9)
    x = rng.standard_normal((96, 72)) * 5
    a = pyitd_b200.crossways_itd_baseline_extract(x)
    b = pyitd_b200.crossways_itd_baseline_extract(np.ascontiguousarray(x.T))
    assert rel(b.T, a) < 1e-12


def test_ensemble_512_full_size_properties():
    # the notebook's workload shape: one 512 x 512 image, 20 ensemble members = 40 960 extracts of 512 samples
    rng = np.random.default_rng(2)
    yy, xx = np.mgrid[0:512, 0:512]
    img = 128 + 40 * np.sin(xx * 0.9 + yy * 0.31) + 25 * np.sin(yy * 1.3) + 10 * rng.standard_normal((512, 512))
    noise = rng.normal(0, pyitd_b200.sift2d.mad(img), (10, 512, 512))
    both = pyitd_b200.totalextract2d(img, noise=noise)
    assert both.shape == (2, 512, 512) and np.isfinite(both).all()
    # the notebook's own check (cell 4): the two components sum back to the image
    assert np.abs(both.sum(axis=0) - img).max() < 1e-12
    # oracle spot check on one ensemble member (crossways of img + noise[0])
    want = o.crossways(img + noise[0])
    got = pyitd_b200.crossways_itd_baseline_extract(img + noise[0])
    assert rel(got, want) < TOL


def test_crossways_f32_mixed_batch():
    rng = np.random.default_rng(4)
    X = (rng.standard_normal((2, 80, 48)) * 10 + 50).astype(np.float32)
    Y = pyitd_b200.crossways_batch(torch.from_numpy(X).cuda())
    assert Y.dtype == torch.float32
    for b in range(2):
        want = o.crossways(X[b].astype(np.float64))
        # float32 storage between the four passes: held to float32 rounding, not to 1e-9
        assert rel(Y[b].cpu().numpy().astype(np.float64), want) < 5e-6
