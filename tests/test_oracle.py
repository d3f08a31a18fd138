"""CPU: both oracle restatements against the fixtures generated from the reference itself."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_cases
from oracle import itd_oracle as o
from pyitd_b200 import synth

IMPLS = [("numpy", o.np_decompose, o.np_extract_level, o.np_find_knots),
         ("c", o.c_decompose, o.c_extract_level, o.c_find_knots)]


def _check_case(case, decompose, find_knots):
    x, mi = case["x"], int(case["max_iteration"])
    r = decompose(x, mi)
    assert r.rotations.shape == case["rotations"].shape
    assert r.rotations.tobytes() == case["rotations"].tobytes()
    assert r.baselines.shape == case["baselines"].shape
    assert r.baselines.tobytes() == case["baselines"].tobytes()
    assert list(r.knot_counts) == list(case["knot_counts"])
    assert np.array_equal(find_knots(x), case["knots"])
    msg = str(case["message"])
    assert r.stop_kind == (1 if msg.startswith("No more") else 2)


@pytest.mark.parametrize("impl", IMPLS, ids=lambda i: i[0])
@pytest.mark.parametrize("name", ["notebook_8000", "demo_400"])
def test_reference_golden_vectors(impl, name):
    case = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    _check_case(case, impl[1], impl[3])


def test_notebook_vector_matches_recorded_reference_output():
    # PyITD.ipynb cell 3 stored output: 9 rotation rows / 8 baselines, reconstruction difference 0.0
    case = dict(np.load(os.path.join(GOLDEN, "notebook_8000.npz")))
    import math
    rows = case["rotations"]
    assert rows.shape == (9, 8000) and case["baselines"].shape == (8, 8000)
    total = math.fsum(math.fsum(rows[:, i]) for i in range(rows.shape[1]))
    assert abs(np.sum(case["x"]) - total) < 1e-12
    assert hashlib.sha256(rows.tobytes()).hexdigest().startswith("2ba3b6211e3a7475")
    assert np.all(case["baselines"][:, -1] == 0)


@pytest.mark.parametrize("impl", IMPLS, ids=lambda i: i[0])
def test_small_cases(impl):
    cases = load_cases(os.path.join(GOLDEN, "small_cases.npz"))
    assert len(cases) >= 15
    for name, case in cases.items():
        _check_case(case, impl[1], impl[3])


@pytest.mark.parametrize("impl", IMPLS, ids=lambda i: i[0])
def test_single_level_cases(impl):
    cases = load_cases(os.path.join(GOLDEN, "level_cases.npz"))
    for name, c in cases.items():
        R, B, k = impl[2](c["x"])
        assert R.tobytes() == c["R"].tobytes() and B.tobytes() == c["B"].tobytes()
        assert np.array_equal(k, c["knots"])
        assert np.array_equal(k, np.sort(np.concatenate([c["valleys"], c["peaks"]])))


@pytest.mark.parametrize("impl", IMPLS, ids=lambda i: i[0])
def test_config1(impl):
    g = np.load(os.path.join(GOLDEN, "config1.npz"))
    x = synth.config1_chirp()
    assert hashlib.sha256(x.tobytes()).hexdigest() == str(g["input_sha"])
    r = impl[1](x, 20)
    assert r.rotations.shape == tuple(g["shape"])
    assert hashlib.sha256(r.rotations.tobytes()).hexdigest() == str(g["rotations_sha"])
    assert hashlib.sha256(r.baselines.tobytes()).hexdigest() == str(g["baselines_sha"])
    assert list(r.knot_counts) == list(g["knot_counts"])
    assert np.array_equal(impl[3](x), g["knots"])


@pytest.mark.parametrize("impl", IMPLS, ids=lambda i: i[0])
def test_error_cases(impl):
    for e in json.load(open(os.path.join(GOLDEN, "error_cases.json"))):
        x = np.asarray(e["x"], dtype=np.float64)
        if e["raises"] == "ZeroDivisionError":
            with pytest.raises(o.OracleError) as ei:
                impl[1](x)
            assert ei.value.status == o.ITD_ZERO_DX
        else:
            assert e["raises"] is None
            impl[1](x)


def test_two_restatements_agree_on_random_inputs():
    rng = np.random.default_rng(5)
    for n in (3, 4, 5, 17, 64, 1000, 4097):
        for kind in range(3):
            x = rng.standard_normal(n)
            if kind == 1:
                x = np.cumsum(x)
            if kind == 2:
                x = np.round(x * 2) / 2 + 1e-9 * np.arange(n)
            for mi in (0, 2, 11):
                try:
                    a = o.np_decompose(x, mi)
                except o.OracleError as ea:
                    with pytest.raises(o.OracleError) as eb:
                        o.c_decompose(x, mi)
                    assert ea.status == eb.value.status
                    continue
                b = o.c_decompose(x, mi)
                assert a.rotations.tobytes() == b.rotations.tobytes()
                assert a.baselines.tobytes() == b.baselines.tobytes()
                assert list(a.knot_counts) == list(b.knot_counts) and a.stop_kind == b.stop_kind


def test_pure_fp32_restatements_agree():
    rng = np.random.default_rng(6)
    for n in (5, 333, 4096):
        x = rng.standard_normal(n).astype(np.float32)
        a, b = o.np_decompose(x, 11), o.c_decompose(x, 11)
        assert a.rotations.dtype == np.float32
        assert a.rotations.tobytes() == b.rotations.tobytes()
        assert list(a.knot_counts) == list(b.knot_counts)


def test_batch_driver_matches_single_calls():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((9, 2000))
    x[3] = np.arange(2000.0)          # monotone: single zero row
    x[5] = 1.0                        # constant: zero delta-X
    rot, n_rows, counts, status, bas = o.c_decompose_batch(x, 5, want_baselines=True, nthreads=3)
    for s in range(x.shape[0]):
        if s == 5:
            assert status[s] == o.ITD_ZERO_DX
            continue
        r = o.c_decompose(x[s], 5)
        assert status[s] == 0 and n_rows[s] == r.rotations.shape[0]
        assert rot[s, :n_rows[s]].tobytes() == r.rotations.tobytes()
        assert bas[s, :r.baselines.shape[0]].tobytes() == r.baselines.tobytes()


def test_reconstruction_invariant():
    # ITD.py:505-508: the rows sum back to the input
    x = synth.config1_chirp(n=16384, seed=3)
    r = o.c_decompose(x, 20)
    assert np.max(np.abs(r.rotations.sum(axis=0) - x)) < 1e-14


def test_extract_with_knots_restatements():
    """ITD.py:95-119 with the knot list given: with a signal's own knots it IS the pinned
    itd_baseline_extract; with another channel's knots the C and numpy restatements agree bit for bit."""
    rng = np.random.default_rng(77)
    for n in (3, 8, 100, 5000):
        x = rng.standard_normal(n)
        y = np.cumsum(rng.standard_normal(n)) + 0.1 * rng.standard_normal(n)
        k = o.c_find_knots(x)
        R, B, _ = o.c_extract_level(x)
        R2, B2 = o.c_extract_with_knots(x, k)
        assert R.tobytes() == R2.tobytes() and B.tobytes() == B2.tobytes()
        try:
            Ra, Ba = o.c_extract_with_knots(y, k)
        except o.OracleError as e:
            with pytest.raises(o.OracleError):
                o.np_extract_with_knots(y, k)
            assert e.status == o.ITD_ZERO_DX
            continue
        Rn, Bn = o.np_extract_with_knots(y, k)
        assert Ra.tobytes() == Rn.tobytes() and Ba.tobytes() == Bn.tobytes()
        assert Ba[-1] == 0.0 and np.array_equal(Ra + Ba, (y - Ba) + Ba)
    for bad in ([5, 5, 9], [0, 3], [3, 99]):
        with pytest.raises(o.OracleError) as ei:
            o.c_extract_with_knots(np.arange(100.0) ** 2, np.array(bad))
        assert ei.value.status == o.ITD_BAD_KNOTS
        with pytest.raises(o.OracleError):
            o.np_extract_with_knots(np.arange(100.0) ** 2, np.array(bad))
