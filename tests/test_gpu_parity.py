"""GPU: the CUDA path (through the C ABI) against the oracle and the reference's golden fixtures.

fp64 is held to BIT-EXACT equality on every level (stronger than the 1e-9 relative L2 the
north star asks for); the fp32 variants are held to the tolerances stated next to each test.
"""
import hashlib
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


def gpu(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def check_against_oracle(x2d, max_iteration=11, min_extrema=2, **kw):
    """decompose a [S, N] float64 batch on the GPU and compare every signal bit-for-bit."""
    res = pyitd_b200.decompose(gpu(x2d), max_iteration=max_iteration, min_extrema=min_extrema,
                               return_baselines=True, **kw)
    torch.cuda.synchronize()
    status = res.status.cpu().numpy()
    for s in range(x2d.shape[0]):
        try:
            want = o.c_decompose(x2d[s], max_iteration, min_extrema)
        except o.OracleError as e:
            assert status[s] & e.status, (s, status[s], e.status)
            continue
        assert status[s] == 0, (s, status[s])
        got = res.rows_of(s).cpu().numpy()
        assert got.shape == want.rotations.shape, (s, got.shape, want.rotations.shape)
        assert got.tobytes() == want.rotations.tobytes(), f"signal {s}: rotations differ"
        assert res.baselines_of(s).cpu().numpy().tobytes() == want.baselines.tobytes(), f"signal {s}: baselines differ"
        nr = want.rotations.shape[0]
        assert res.knot_counts[s, :nr].cpu().tolist() == list(want.knot_counts), s
        assert int(res.input_knots[s]) == want.input_knots
        assert int(res.stop_kind[s]) == want.stop_kind
    return res


# ---------------------------------------------------------------------------------------------
# the reference's own golden vectors, through the drop-in class
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["notebook_8000", "demo_400"])
def test_reference_golden_vectors_dropin(name):
    case = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    itd = pyitd_b200.ITD()
    rows = itd.itd(case["x"], max_iteration=int(case["max_iteration"]))
    assert rows.dtype == np.float64 and rows.shape == case["rotations"].shape
    assert rows.tobytes() == case["rotations"].tobytes()
    assert itd.get_baselines().tobytes() == case["baselines"].tobytes()
    assert itd.get_baselines().shape == case["baselines"].shape
    assert itd.get_rotations() is rows
    assert list(itd.knot_counts) == list(case["knot_counts"])
    # level-1 knot indices bit-exact (north star), via both kinds like ITD.py:87-88,97
    v, p = pyitd_b200.detect_peaks(case["x"]), pyitd_b200.detect_peaks(-case["x"])
    assert v.dtype == np.int64
    assert np.array_equal(np.sort(np.concatenate([v, p])), case["knots"])


def test_notebook_vector_hash_and_reconstruction():
    case = dict(np.load(os.path.join(GOLDEN, "notebook_8000.npz")))
    rows = pyitd_b200.ITD()(case["x"], max_iterations=11)          # __call__ works here (ITD.py:189 bug fixed)
    assert hashlib.sha256(rows.tobytes()).hexdigest().startswith("2ba3b6211e3a7475")
    import math
    total = math.fsum(math.fsum(rows[:, i]) for i in range(rows.shape[1]))
    assert abs(np.sum(case["x"]) - total) < 1e-12                  # ITD.py:505-508 / notebook output 0.0


def test_config1_against_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "config1.npz"))
    x = synth.config1_chirp()
    itd = pyitd_b200.ITD()
    rows = itd.itd(x, max_iteration=20)
    assert rows.shape == tuple(g["shape"])
    assert hashlib.sha256(rows.tobytes()).hexdigest() == str(g["rotations_sha"])
    assert hashlib.sha256(itd.get_baselines().tobytes()).hexdigest() == str(g["baselines_sha"])
    assert list(itd.knot_counts) == list(g["knot_counts"])
    knots, cnt, _ = pyitd_b200.find_knots(gpu(x))
    assert np.array_equal(knots[0, :int(cnt[0])].cpu().numpy(), g["knots"])


def test_small_cases_dropin():
    for name, case in load_cases(os.path.join(GOLDEN, "small_cases.npz")).items():
        itd = pyitd_b200.ITD()
        rows = itd.itd(case["x"], max_iteration=int(case["max_iteration"]))
        assert rows.shape == case["rotations"].shape, name
        assert rows.tobytes() == case["rotations"].tobytes(), name
        assert itd.get_baselines().shape == case["baselines"].shape, name
        assert itd.get_baselines().tobytes() == case["baselines"].tobytes(), name
        assert list(itd.knot_counts) == list(case["knot_counts"]), name


def test_single_level_functions():
    for name, c in load_cases(os.path.join(GOLDEN, "level_cases.npz")).items():
        R, B = pyitd_b200.itd_baseline_extract(c["x"])
        assert R.tobytes() == c["R"].tobytes() and B.tobytes() == c["B"].tobytes(), name
        assert np.array_equal(pyitd_b200.detect_peaks(c["x"]), c["valleys"]), name
        assert np.array_equal(pyitd_b200.detect_peaks(-c["x"]), c["peaks"]), name


def test_error_behaviour_matches_reference():
    for e in json.load(open(os.path.join(GOLDEN, "error_cases.json"))):
        x = np.asarray(e["x"], dtype=np.float64)
        if e["raises"] == "ZeroDivisionError":
            with pytest.raises(ZeroDivisionError):
                pyitd_b200.ITD().itd(x)
            with pytest.raises(ZeroDivisionError):
                pyitd_b200.itd_baseline_extract(x)
        else:
            pyitd_b200.ITD().itd(x)
    with pytest.raises(ValueError):
        pyitd_b200.ITD().itd(np.array([0.0, np.nan, 1.0, 2.0, 0.5]))
    with pytest.raises(ValueError):
        pyitd_b200.ITD().itd(np.array([0.0, 1.0]))
    # batched: errors are per-signal status words, raised only on request
    x = np.random.default_rng(0).standard_normal((4, 500))
    x[2] = 3.0
    res = pyitd_b200.decompose(gpu(x))
    assert res.status.cpu().tolist() == [0, 0, _capi.ST_ZERO_DX, 0]
    with pytest.raises(ZeroDivisionError):
        pyitd_b200.decompose(gpu(x), strict=True)


# ---------------------------------------------------------------------------------------------
# seeded batches against the oracle: ragged sizes around every tile boundary, every tile config
# ---------------------------------------------------------------------------------------------
def _mixed_batch(rng, S, n):
    x = rng.standard_normal((S, n))
    for s in range(S):
        m = s % 6
        if m == 1:
            x[s] = np.cumsum(x[s])
        elif m == 2:
            x[s] = np.round(x[s] * 2) / 2 + 1e-7 * np.arange(n)            # plateaus and ties
        elif m == 3:
            x[s] = np.sin(np.arange(n) * (0.01 + 0.3 * rng.random())) + 0.01 * x[s]
        elif m == 4:
            x[s] = np.arange(n, dtype=np.float64) * (1 if s % 2 else -1)   # monotone
        elif m == 5:
            x[s, : n // 2] = np.linspace(0, 1, n // 2)                     # long knot-free stretch
    return x


@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
def test_tile_boundaries_all_configs(cfg, monkeypatch):
    monkeypatch.setenv("PYITD_TILE_CFG", str(cfg))
    monkeypatch.setenv("PYITD_FORCE_PATH", "lookback")
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(100 + cfg)
    T = (512, 1024, 2048, 2048)[cfg]
    try:
        for n in (3, 4, 5, 31, 33, T - 1, T, T + 1, T + 2, 2 * T - 1, 2 * T + 1, 3 * T + 7, 7777):
            check_against_oracle(_mixed_batch(rng, 7, n), max_iteration=11)
    finally:
        pyitd_b200.clear_plan_cache()


@pytest.mark.parametrize("path", ["stream", "lookback"])
def test_both_level_kernels_agree_with_oracle(path, monkeypatch):
    """The one-CTA-per-signal TMA-pipelined kernel and the multi-CTA look-back kernel are forced in
    turn over sizes around the 1024-sample tile and the 128-sample warp span."""
    monkeypatch.setenv("PYITD_FORCE_PATH", path)
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(300)
    try:
        for n in (4, 8, 32, 36, 124, 128, 132, 1020, 1024, 1028, 2044, 2048, 2052, 3076, 5120, 8192, 10004):
            check_against_oracle(_mixed_batch(rng, 9, n), max_iteration=11)
        for mi in (0, 1, 4):
            check_against_oracle(_mixed_batch(rng, 9, 6000), max_iteration=mi)
        x32 = _mixed_batch(rng, 6, 4096).astype(np.float32)
        for dt in ("f32_mixed", "f32"):
            res = pyitd_b200.decompose(gpu(x32), max_iteration=7, dtype=dt)
            for s in range(6):
                src = x32[s].astype(np.float64) if dt == "f32_mixed" else x32[s]
                try:
                    want = o.c_decompose(src, 7)
                except o.OracleError:
                    assert int(res.status[s]) != 0
                    continue
                assert res.rows_of(s).cpu().numpy().tobytes() == want.rotations.astype(np.float32).tobytes(), (dt, s)
    finally:
        pyitd_b200.clear_plan_cache()


@pytest.mark.parametrize("groups", [2, 3, 7])
def test_stream_launch_groups_do_not_change_results(groups, monkeypatch):
    """The stream path may cut the batch into signal ranges with their own launch chains on internal streams
    (pyitd_plan_set_groups); every signal must still match the oracle bit for bit, ragged stops included."""
    monkeypatch.setenv("PYITD_FORCE_PATH", "stream")
    monkeypatch.setenv("PYITD_GROUPS", str(groups))
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(310 + groups)
    try:
        from pyitd_b200.itd import get_plan
        assert get_plan(0, 11, 5000, _capi.F64, 11, 2, _capi.OPT_BASELINES).groups == groups
        for n in (1024, 5000, 9000):
            check_against_oracle(_mixed_batch(rng, 11, n), max_iteration=11)
        check_against_oracle(_mixed_batch(rng, 2, 3000), max_iteration=3)        # fewer signals than groups
        x = synth.eeg_like(40, 16384, seed=77, device="cuda").cpu().numpy()
        check_against_oracle(x, max_iteration=11)
    finally:
        pyitd_b200.clear_plan_cache()


@pytest.mark.parametrize("path", ["stream", "lookback"])
def test_extract_with_supplied_knots(path, monkeypatch):
    """SURVEY 8f rank 1 (itd.cpp:41-44, :156-169 compute_extrema=false): knots detected once and reused.
    (a) a signal's own knots reproduce itd_baseline_extract bit for bit; (b) channel 0's knots applied to every
    other channel (one shared list) and (c) per-signal lists equal the oracle's ITD.py:95-119 with tau given;
    (d) a list that is not strictly increasing inside [1, n-2] is reported and replaced by the empty list."""
    monkeypatch.setenv("PYITD_FORCE_PATH", path)
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(330)
    try:
        for n in (8, 1000, 4096, 6148):
            x = rng.standard_normal((7, n)).cumsum(axis=1) + 0.3 * rng.standard_normal((7, n))
            xg = gpu(x)
            knots, cnt, st = pyitd_b200.find_knots(xg)
            # (a) own knots
            R, B, st2 = pyitd_b200.extract_with_knots(xg, knots, cnt)
            R0, B0, _, st0 = pyitd_b200.extract_level(xg)
            assert torch.equal(st2, st0) and torch.equal(R, R0) and torch.equal(B, B0)
            # (b) channel 0's list shared by all channels
            k0 = knots[0, : int(cnt[0])]
            R, B, st3 = pyitd_b200.extract_with_knots(xg, k0)
            for s in range(7):
                try:
                    wr, wb = o.c_extract_with_knots(x[s], k0.cpu().numpy())
                except o.OracleError as e:
                    assert int(st3[s]) & e.status
                    continue
                assert int(st3[s]) == 0
                assert R[s].cpu().numpy().tobytes() == wr.tobytes() and B[s].cpu().numpy().tobytes() == wb.tobytes()
            # (c) per-signal lists: signal s uses the knots of signal (s + 1) % 7
            perm = [(s + 1) % 7 for s in range(7)]
            R, B, st4 = pyitd_b200.extract_with_knots(xg, knots[perm], cnt[perm])
            for s in range(7):
                kk = knots[perm[s], : int(cnt[perm[s]])].cpu().numpy()
                try:
                    wr, wb = o.c_extract_with_knots(x[s], kk)
                except o.OracleError as e:
                    assert int(st4[s]) & e.status
                    continue
                assert R[s].cpu().numpy().tobytes() == wr.tobytes() and B[s].cpu().numpy().tobytes() == wb.tobytes()
            # numpy restatement agrees with the C one
            wr2, wb2 = o.np_extract_with_knots(x[1], k0.cpu().numpy())
            wr, wb = o.c_extract_with_knots(x[1], k0.cpu().numpy())
            assert wr.tobytes() == wr2.tobytes() and wb.tobytes() == wb2.tobytes()
        # (d) invalid lists
        x = rng.standard_normal((3, 512))
        bad = torch.tensor([[5, 5, 9, 0], [0, 3, 4, 5], [7, 9, 511, 0]], dtype=torch.int32)
        R, B, st5 = pyitd_b200.extract_with_knots(gpu(x), bad, torch.tensor([3, 4, 3], dtype=torch.int32))
        assert [int(v) & _capi.ST_BAD_KNOTS for v in st5.cpu()] == [8, 8, 8]
        wr, wb = o.c_extract_with_knots(x[0], np.empty(0, dtype=np.int64))       # processed with the empty list
        assert R[0].cpu().numpy().tobytes() == wr.tobytes()
        # empty list: the monotone case (two end knots only)
        R, B, st6 = pyitd_b200.extract_with_knots(gpu(x), torch.zeros((1, 1), dtype=torch.int32),
                                                  torch.zeros(1, dtype=torch.int32))
        assert int(st6.abs().max()) == 0
        for s in range(3):
            wr, wb = o.c_extract_with_knots(x[s], np.empty(0, dtype=np.int64))
            assert B[s].cpu().numpy().tobytes() == wb.tobytes()
        # fp32 I/O around the fp64 arithmetic
        x32 = rng.standard_normal((4, 3000)).astype(np.float32)
        kn, c, _ = pyitd_b200.find_knots(gpu(x32), dtype="f32_mixed")
        R, B, _ = pyitd_b200.extract_with_knots(gpu(x32), kn[0, : int(c[0])], dtype="f32_mixed")
        wr, wb = o.c_extract_with_knots(x32[2].astype(np.float64), kn[0, : int(c[0])].cpu().numpy())
        assert R[2].cpu().numpy().tobytes() == wr.astype(np.float32).tobytes()
    finally:
        pyitd_b200.clear_plan_cache()


RES_SIZES = (3, 4, 5, 31, 33, 127, 128, 129, 255, 256, 257, 258, 511, 513, 1023, 1024, 1025, 2047, 2049, 4096,
             4097, 7777, 8192, 10004, 16385, 20000)


@pytest.mark.parametrize("cfg", [0, 1])
@pytest.mark.parametrize("cl", [0, 1, 2, 4, 8])
def test_resident_kernel_cluster_shapes(cfg, cl, monkeypatch):
    """The on-chip kernel (whole decomposition in one launch, signal resident in a cluster's shared memory)
    for both warp configurations and every cluster size, over sizes around the unit (128/256 samples) and
    the per-warp / per-CTA range boundaries.  A cluster wider than the signal leaves warps without samples."""
    from pyitd_b200.itd import get_plan
    monkeypatch.setenv("PYITD_FORCE_PATH", "resident")
    monkeypatch.setenv("PYITD_RES_CFG", str(cfg))
    if cl:
        monkeypatch.setenv("PYITD_RES_CL", str(cl))
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(500 + 10 * cfg + cl)
    try:
        used = 0
        for n in RES_SIZES:
            plan = get_plan(0, 7, n, _capi.F64, 11, 2, _capi.OPT_BASELINES)
            path, csize = plan.path
            if path != "resident":
                continue                                  # this cluster size cannot hold n samples on chip
            assert cl == 0 or csize == cl
            used += 1
            check_against_oracle(_mixed_batch(rng, 7, n), max_iteration=11)
        assert used >= 10, used
    finally:
        pyitd_b200.clear_plan_cache()


@pytest.mark.parametrize("cfg", [0, 1])
def test_resident_kernel_many_signals_per_cluster(cfg, monkeypatch):
    """More signals than resident clusters: every cluster walks several signals (buffer parity, halo
    values and backup area are reused across signals)."""
    monkeypatch.setenv("PYITD_FORCE_PATH", "resident")
    monkeypatch.setenv("PYITD_RES_CFG", str(cfg))
    monkeypatch.setenv("PYITD_RES_CLUSTERS", "3")
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(520 + cfg)
    try:
        for n, mi in ((700, 11), (4100, 3), (9000, 0)):
            check_against_oracle(_mixed_batch(rng, 23, n), max_iteration=mi)
            check_against_oracle(_mixed_batch(rng, 23, n), max_iteration=mi, zero_tail=True)
    finally:
        pyitd_b200.clear_plan_cache()


@pytest.mark.parametrize("cfg", [0, 1])
def test_resident_kernel_fp32_variants(cfg, monkeypatch):
    monkeypatch.setenv("PYITD_FORCE_PATH", "resident")
    monkeypatch.setenv("PYITD_RES_CFG", str(cfg))
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(540 + cfg)
    try:
        for n in (5, 130, 1000, 8192, 8195, 30000):
            x32 = _mixed_batch(rng, 6, n).astype(np.float32)
            for dt in ("f32_mixed", "f32"):
                res = pyitd_b200.decompose(gpu(x32), max_iteration=7, dtype=dt, return_baselines=True)
                for s in range(6):
                    src = x32[s].astype(np.float64) if dt == "f32_mixed" else x32[s]
                    try:
                        want = o.c_decompose(src, 7)
                    except o.OracleError:
                        assert int(res.status[s]) != 0
                        continue
                    assert int(res.status[s]) == 0
                    assert res.rows_of(s).cpu().numpy().tobytes() == want.rotations.astype(np.float32).tobytes(), (dt, n, s)
                    assert res.baselines_of(s).cpu().numpy().tobytes() == want.baselines.astype(np.float32).tobytes(), (dt, n, s)
    finally:
        pyitd_b200.clear_plan_cache()


def test_default_path_by_shape():
    from pyitd_b200.itd import get_plan
    assert get_plan(0, 4096, 65536, _capi.F64, 11, 2, 0).path[0] == "sweep"
    assert get_plan(0, 64, 65536, _capi.F64, 11, 2, 0).path == ("resident", 4)
    assert get_plan(0, 1, 65536, _capi.F64, 11, 2, 0).path[0] == "coop"
    assert get_plan(0, 3515, 8192, _capi.F32_MIXED, 7, 2, 0).path[0] == "sweep"
    assert get_plan(0, 1, 1 << 22, _capi.F32, 11, 2, 0).path[0] == "strided"
    assert get_plan(0, 2, 1 << 22, _capi.F32, 11, 2, 0).path[0] == "strided"
    assert get_plan(0, 2, 1 << 16, _capi.F32, 11, 2, 0).path[0] == "coop"
    assert get_plan(0, 4, 1 << 17, _capi.F64, 11, 2, 0).path[0] == "coop"
    assert get_plan(0, 20, 8192, _capi.F64, 11, 2, 0).path[0] == "coop"           # the chip holds them all at once
    assert get_plan(0, 12, 65536, _capi.F64, 11, 2, 0).path[0] == "resident"      # ... here it would need a second round
    pyitd_b200.clear_plan_cache()


@pytest.mark.parametrize("max_iteration", [0, 1, 3, 7, 20])
def test_iteration_cap(max_iteration):
    rng = np.random.default_rng(11)
    res = check_against_oracle(_mixed_batch(rng, 12, 8192), max_iteration=max_iteration)
    assert int(res.n_rows.max()) <= max_iteration + 2


@pytest.mark.parametrize("min_extrema", [0, 1, 3, 10])
def test_min_extrema(min_extrema):
    rng = np.random.default_rng(12)
    check_against_oracle(_mixed_batch(rng, 6, 3000), max_iteration=11, min_extrema=min_extrema)


@pytest.mark.parametrize("path", ["lookback", "strided"])
def test_long_single_signal_multi_cta(path, monkeypatch):
    # one signal over thousands of tiles: exercises the decoupled look-back chain and deep levels
    # where almost every tile is knot-free (one CTA per tile / persistent CTAs striding over the tiles)
    monkeypatch.setenv("PYITD_FORCE_PATH", path)
    pyitd_b200.clear_plan_cache()
    try:
        x = synth.long_signal(n=1 << 21, seed=3).double().numpy()
        check_against_oracle(x[None, :], max_iteration=11)
    finally:
        pyitd_b200.clear_plan_cache()


@pytest.mark.parametrize("ctas", [0, 1, 3])
def test_strided_long_signal_kernel(ctas, monkeypatch):
    """level_strided_kernel (one long signal, persistent CTAs striding over its tiles, look-back rank bases, halo
    samples riding with the TMA slice): sizes around the 1024-sample tile, grids of 1 / 3 / all-that-fit CTAs (one,
    several or hundreds of tiles per CTA), every stop kind, baselines, zero tail, fp32 variants -- bit for bit."""
    monkeypatch.setenv("PYITD_FORCE_PATH", "strided")
    if ctas:
        monkeypatch.setenv("PYITD_STRIDED_CTAS", str(ctas))
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(340 + ctas)
    try:
        from pyitd_b200.itd import get_plan
        assert get_plan(0, 1, 4096, _capi.F64, 11, 2, _capi.OPT_BASELINES).path[0] == "strided"
        for n in (4, 8, 36, 128, 1020, 1024, 1028, 2044, 2048, 2052, 3076, 5120, 10004, 20000):
            for kind in range(6):
                x = _mixed_batch(rng, 6, n)[kind:kind + 1]
                check_against_oracle(x, max_iteration=11)
        for mi in (0, 1, 4):
            check_against_oracle(_mixed_batch(rng, 4, 6000)[3:4], max_iteration=mi)
        check_against_oracle(_mixed_batch(rng, 2, 9000)[1:2], max_iteration=11, min_extrema=5)
        xz = _mixed_batch(rng, 4, 5000)[3:4]
        rz = pyitd_b200.decompose(gpu(xz), max_iteration=11, return_baselines=True, zero_tail=True)
        ref = check_against_oracle(xz, max_iteration=11)
        nr = int(rz.n_rows[0])
        assert torch.equal(rz.rotations[0, :nr], ref.rotations[0, :nr])
        assert float(rz.rotations[0, nr:].abs().max() if nr < rz.rotations.shape[1] else 0) == 0
        # config 3's generator, 2^20 samples (1024 tiles: 2-3 per CTA at the default grid)
        x32 = synth.long_signal(n=1 << 20, seed=3).numpy()
        for dt in ("f32_mixed", "f32"):
            res = pyitd_b200.decompose(gpu(x32[None, :]), max_iteration=11, dtype=dt, return_baselines=True)
            want = o.c_decompose(x32.astype(np.float64) if dt == "f32_mixed" else x32, 11)
            assert res.rows_of(0).cpu().numpy().tobytes() == want.rotations.astype(np.float32).tobytes(), dt
            assert res.baselines_of(0).cpu().numpy().tobytes() == want.baselines.astype(np.float32).tobytes(), dt
            assert res.knot_counts[0, : want.rotations.shape[0]].cpu().tolist() == list(want.knot_counts)
        # a few long signals: one after the other through the same launches
        check_against_oracle(_mixed_batch(rng, 3, 6148), max_iteration=11)
        # supplied knots and the single-level entry go through the same kernel
        xs = rng.standard_normal((1, 7000)).cumsum(axis=1)
        kn, c, _ = pyitd_b200.find_knots(gpu(xs))
        R, B, st = pyitd_b200.extract_with_knots(gpu(xs), kn, c)
        wr, wb, _ = o.c_extract_level(xs[0])
        assert R[0].cpu().numpy().tobytes() == wr.tobytes() and B[0].cpu().numpy().tobytes() == wb.tobytes()
    finally:
        pyitd_b200.clear_plan_cache()


def test_eeg_like_batch_config2_subset():
    # config 2's generator, 64 channels of 65 536 samples, knot-count stopping (ragged row counts)
    x = synth.eeg_like(64, 65536, seed=1234, device="cuda").cpu().numpy()
    res = check_against_oracle(x, max_iteration=11)
    assert len(set(res.n_rows.cpu().tolist())) > 1


def check_batch_against_c_oracle(res, x, max_iteration=11, baselines=True):
    """every channel of a big batch bit-for-bit against the multi-threaded C oracle (rows, baselines, knot counts,
    stop kind) -- one oracle call for the whole batch, so hundreds of 65 536-sample channels take a second or two."""
    rot, n_rows, counts, status, bas = o.c_decompose_batch(x, max_iteration, want_baselines=baselines)
    assert int(np.abs(status).max()) == 0
    assert res.status.cpu().tolist() == [0] * x.shape[0]
    assert res.n_rows.cpu().tolist() == n_rows.tolist()
    got_rot = res.rotations.cpu().numpy()
    got_bas = res.baselines.cpu().numpy() if baselines else None
    got_cnt = res.knot_counts.cpu().numpy()
    kind = res.stop_kind.cpu().numpy()
    emax = max_iteration + 1
    for s in range(x.shape[0]):
        nr = int(n_rows[s])
        assert got_rot[s, :nr].tobytes() == rot[s, :nr].tobytes(), f"signal {s}: rotations differ"
        assert got_cnt[s, :nr].tolist() == counts[s, :nr].tolist(), f"signal {s}: knot counts differ"
        want_kind = _capi.STOP_KNOTS if counts[s, nr - 1] < 2 else _capi.STOP_ITER
        assert int(kind[s]) == want_kind and (want_kind == _capi.STOP_KNOTS or nr == emax + 1), s
        if baselines:
            nb = nr if want_kind == _capi.STOP_ITER else nr - 1
            assert got_bas[s, :nb].tobytes() == bas[s, :nb].tobytes(), f"signal {s}: baselines differ"


@pytest.mark.parametrize("groups", [1, 2])
def test_benchmarked_stream_path_256_channels_bit_exact(groups, monkeypatch):
    """The kernel the headline number comes from (one CTA per signal, N = 65 536 = 64 tiles per signal, launch groups 1
    and 2, bench.py's generator): 256 channels, EVERY channel bit for bit against the oracle -- rotations, baselines,
    per-level knot counts, stop kind (ITD.py:351-433)."""
    monkeypatch.setenv("PYITD_FORCE_PATH", "stream")
    monkeypatch.setenv("PYITD_GROUPS", str(groups))
    pyitd_b200.clear_plan_cache()
    try:
        from pyitd_b200.itd import get_plan
        plan = get_plan(0, 256, 65536, _capi.F64, 11, 2, _capi.OPT_BASELINES)
        assert plan.path[0] == "stream" and plan.groups == groups
        xg = synth.eeg_like(256, 65536, seed=1234, device="cuda")
        res = pyitd_b200.decompose(xg, max_iteration=11, return_baselines=True)
        torch.cuda.synchronize()
        check_batch_against_c_oracle(res, xg.cpu().numpy(), 11)
        assert len(set(res.n_rows.cpu().tolist())) > 2            # ragged stops
        # the bench itself runs without baselines (options = 0): same rows
        res0 = pyitd_b200.decompose(xg, max_iteration=11)
        torch.cuda.synchronize()
        check_batch_against_c_oracle(res0, xg.cpu().numpy(), 11, baselines=False)
    finally:
        pyitd_b200.clear_plan_cache()


def test_zero_tail_option():
    rng = np.random.default_rng(13)
    x = _mixed_batch(rng, 6, 5000)
    res = pyitd_b200.decompose(gpu(x), max_iteration=11, return_baselines=True, zero_tail=True)
    ref = check_against_oracle(x, max_iteration=11)
    for s in range(6):
        nr = int(res.n_rows[s])
        assert torch.equal(res.rotations[s, :nr], ref.rotations[s, :nr])
        assert float(res.rotations[s, nr:].abs().max() if nr < res.rotations.shape[1] else 0) == 0
        nb = res.baselines_of(s).shape[0]
        assert float(res.baselines[s, nb:].abs().max() if nb < res.baselines.shape[1] else 0) == 0


def test_host_entry_point_matches_device_entry():
    rng = np.random.default_rng(14)
    x = _mixed_batch(rng, 5, 4100)
    a = pyitd_b200.decompose(x, max_iteration=6, return_baselines=True)            # numpy in -> host ABI
    b = pyitd_b200.decompose(gpu(x), max_iteration=6, return_baselines=True)
    assert not a.rotations.is_cuda
    for s in range(5):
        assert torch.equal(a.rows_of(s), b.rows_of(s).cpu())
        assert torch.equal(a.baselines_of(s), b.baselines_of(s).cpu())
    assert torch.equal(a.n_rows, b.n_rows.cpu()) and torch.equal(a.status, b.status.cpu())


@pytest.mark.parametrize("chunk,zero_tail", [(3, False), (3, True), (1, False), (8, False)])
def test_host_entry_point_chunked_pipeline(chunk, zero_tail, monkeypatch):
    """pyitd_decompose_host walks the batch in chunks through two device slots (H2D / kernels / D2H of
    neighbouring chunks overlap) and copies back only the rows each signal produced."""
    monkeypatch.setenv("PYITD_HOST_CHUNK", str(chunk))
    pyitd_b200.clear_plan_cache()
    try:
        rng = np.random.default_rng(140 + chunk)
        x = _mixed_batch(rng, 8, 5000)
        for mi in (11, 2):
            a = pyitd_b200.decompose(x, max_iteration=mi, return_baselines=True, zero_tail=zero_tail)
            st = a.status.numpy()
            for s in range(8):
                try:
                    want = o.c_decompose(x[s], mi)
                except o.OracleError as e:
                    assert st[s] & e.status
                    continue
                assert a.rows_of(s).numpy().tobytes() == want.rotations.tobytes(), (mi, s)
                assert a.baselines_of(s).numpy().tobytes() == want.baselines.tobytes(), (mi, s)
                assert int(a.stop_kind[s]) == want.stop_kind
                if zero_tail:
                    nr = int(a.n_rows[s])
                    assert float(a.rotations[s, nr:].abs().sum()) == 0.0
    finally:
        pyitd_b200.clear_plan_cache()


def test_sharded_driver_on_gpu():
    """pyitd_b200.shard.decompose_sharded on one process (world 1): same driver code bench.py --gpus N runs."""
    from pyitd_b200 import shard
    rng = np.random.default_rng(150)
    x = _mixed_batch(rng, 6, 3000)
    res, summ = shard.decompose_sharded(lambda a, b: gpu(x[a:b]), 6, max_iteration=11, return_baselines=True)
    torch.cuda.synchronize()
    for s in range(6):
        try:
            want = o.c_decompose(x[s], 11)
        except o.OracleError:
            assert int(summ.status[s]) != 0
            continue
        assert res.rows_of(s).cpu().numpy().tobytes() == want.rotations.tobytes()
    assert torch.equal(summ.n_rows, res.n_rows)


def test_repeated_calls_reuse_the_plan_and_stay_exact():
    rng = np.random.default_rng(15)
    x = _mixed_batch(rng, 8, 6000)
    first = pyitd_b200.decompose(gpu(x))
    for _ in range(5):
        again = pyitd_b200.decompose(gpu(x))
        assert torch.equal(first.n_rows, again.n_rows)
        for s in range(8):
            assert torch.equal(first.rows_of(s), again.rows_of(s))


@pytest.mark.parametrize("path,S", [("stream", 5), ("strided", 1), ("strided", 3)])
def test_input_pointer_that_is_not_16_byte_aligned(path, S, monkeypatch):
    """The TMA bulk copies of the stream / strided kernels need 16-byte aligned rows.  A caller's x that is only
    8-byte aligned must still decompose exactly: the scan and the first level fall back to the look-back kernels
    (which leave complete knot tables), later levels read the library's own aligned carry."""
    from pyitd_b200.itd import get_plan
    monkeypatch.setenv("PYITD_FORCE_PATH", path)
    monkeypatch.setenv("PYITD_GROUPS", "2")
    pyitd_b200.clear_plan_cache()
    rng = np.random.default_rng(350)
    try:
        n = 5000
        x = _mixed_batch(rng, 6, n)[:S] if S > 1 else _mixed_batch(rng, 6, n)[3:4]
        buf = torch.zeros(S * n + 1, dtype=torch.float64, device="cuda")
        buf[1:] = torch.from_numpy(x).reshape(-1).cuda()
        assert (buf.data_ptr() + 8) % 16 == 8
        plan = get_plan(0, S, n, _capi.F64, 11, 2, _capi.OPT_BASELINES)
        assert plan.path[0] == path
        rows = plan.rows
        rot = torch.empty((S, rows, n), dtype=torch.float64, device="cuda")
        bas = torch.empty_like(rot)
        ints = [torch.empty(S, dtype=torch.int32, device="cuda") for _ in range(4)]
        counts = torch.empty((S, rows), dtype=torch.int32, device="cuda")
        plan.decompose_device(buf.data_ptr() + 8, rot.data_ptr(), bas.data_ptr(), ints[0].data_ptr(), counts.data_ptr(),
                              ints[1].data_ptr(), ints[2].data_ptr(), ints[3].data_ptr(),
                              torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for s in range(S):
            try:
                want = o.c_decompose(x[s], 11)
            except o.OracleError as e:
                assert int(ints[3][s]) & e.status
                continue
            nr = int(ints[0][s])
            assert rot[s, :nr].cpu().numpy().tobytes() == want.rotations.tobytes(), s
            nb = want.baselines.shape[0]
            assert bas[s, :nb].cpu().numpy().tobytes() == want.baselines.tobytes(), s
            assert int(ints[1][s]) == want.input_knots
    finally:
        pyitd_b200.clear_plan_cache()


# ---------------------------------------------------------------------------------------------
# fp32 variants
# ---------------------------------------------------------------------------------------------
def test_fp32_mixed_equals_rounded_reference():
    """fp32 in/out around an fp64 carry must equal float32(reference(float64(x32))) exactly: zero knot
    mismatches on every level and rel-L2 at rounding level (stated tolerance 1e-4; measured ~3e-8)."""
    rng = np.random.default_rng(16)
    x32 = _mixed_batch(rng, 8, 8192).astype(np.float32)
    res = pyitd_b200.decompose(gpu(x32), max_iteration=7, dtype="f32_mixed", return_baselines=True)
    assert res.rotations.dtype == torch.float32
    for s in range(8):
        try:
            want = o.c_decompose(x32[s].astype(np.float64), 7)
        except o.OracleError:
            assert int(res.status[s]) != 0
            continue
        got = res.rows_of(s).cpu().numpy()
        assert got.shape == want.rotations.shape
        assert got.tobytes() == want.rotations.astype(np.float32).tobytes()
        assert res.knot_counts[s, : got.shape[0]].cpu().tolist() == list(want.knot_counts)
        num = np.linalg.norm(got.astype(np.float64) - want.rotations, axis=1)
        den = np.maximum(np.linalg.norm(want.rotations, axis=1), 1e-300)
        assert np.all(num / den < 1e-4)


def test_fp32_pure_is_bit_exact_against_fp32_oracle_and_reports_mismatch_rate():
    """pure fp32 storage + arithmetic: pinned bit-for-bit by an IEEE binary32 execution of the same
    operation sequence; against the fp64 oracle only level 1 is within 1e-4 in general (SURVEY.md
    section 0 item 9), so deeper levels are reported, not asserted."""
    x32 = synth.audio_frames(seconds=2.0)[:10]
    res = pyitd_b200.decompose(gpu(x32), max_iteration=7, dtype="f32")
    for s in range(x32.shape[0]):
        want = o.c_decompose(x32[s], 7)
        got = res.rows_of(s).cpu().numpy()
        assert got.shape == want.rotations.shape and got.tobytes() == want.rotations.tobytes()
    # level 1 against the fp64 reference arithmetic
    R64, _, k64 = o.c_extract_level(x32[0].astype(np.float64))
    k32, cnt, _ = pyitd_b200.find_knots(gpu(x32[:1]), dtype="f32")
    assert np.array_equal(k32[0, :int(cnt[0])].cpu().numpy(), k64)        # comparisons only: exact
    r1 = res.rotations[0, 0].cpu().numpy().astype(np.float64)
    assert np.linalg.norm(r1 - R64) / np.linalg.norm(R64) < 1e-4


# ---------------------------------------------------------------------------------------------
# full-size properties (no oracle: size-independent invariants)
# ---------------------------------------------------------------------------------------------
def test_full_size_batch_properties():
    """512 x 65 536 fp64 EEG-like channels (one eighth of config 2; same kernels and grid shape)."""
    x = synth.eeg_like(512, 65536, seed=99, device="cuda")
    res = pyitd_b200.decompose(x, max_iteration=11, zero_tail=True)
    torch.cuda.synchronize()
    assert int(res.status.abs().max()) == 0
    recon = res.rotations.sum(dim=1)                     # rows sum back to the input (ITD.py:505-508)
    err = (recon - x).abs().max().item()
    assert err < 1e-12, err
    nr = res.n_rows.long()
    assert int(nr.min()) >= 2 and int(nr.max()) <= 13
    # knot counts shrink monotonically until the stop, and the stop rule holds on the device
    kc = res.knot_counts.long()
    last = kc.gather(1, (nr - 1).unsqueeze(1)).squeeze(1)
    knot_stop = res.stop_kind == _capi.STOP_KNOTS
    assert bool((last[knot_stop] < 2).all())
    assert bool((res.input_knots.long() > kc[:, 0]).all())
    # every rotation's last sample is that level's input sample: baseline[N-1] == 0 (ITD.py:112)
    assert float((res.rotations[:, 0, -1] - x[:, -1]).abs().max()) == 0.0
    # spot-check 4 channels against the oracle
    for s in (0, 17, 300, 511):
        want = o.c_decompose(x[s].cpu().numpy(), 11)
        assert res.rows_of(s).cpu().numpy().tobytes() == want.rotations.tobytes()


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs 3, 4, 5 (config 1 and 2 are covered above)
# ---------------------------------------------------------------------------------------------
def _level1_knots_torch(x):
    """ITD.py:44-59 on x and -x, unioned (ITD.py:97), as a torch stencil on the device: the knot mask."""
    a, b, c = x[:-2], x[1:-1], x[2:]
    return ((a >= b) & (b < c)) | ((a <= b) & (b > c))


@pytest.mark.parametrize("dt", ["f32_mixed", "f32"])
def test_config3_long_signal_fp32_at_oracle_size(dt):
    """Config 3's generator at 2**22 samples (the oracle needs seconds there): the strided long-signal path.
    f32_mixed must equal float32(reference(float64(x32))) bit for bit; pure f32 equals a binary32 execution."""
    x32 = synth.long_signal(n=1 << 22, seed=3).numpy()
    res = pyitd_b200.decompose(gpu(x32[None, :]), max_iteration=11, dtype=dt)
    from pyitd_b200.itd import get_plan
    assert get_plan(0, 1, 1 << 22, _capi.F32_MIXED if dt == "f32_mixed" else _capi.F32, 11, 2, 0).path[0] == "strided"
    want = o.c_decompose(x32.astype(np.float64) if dt == "f32_mixed" else x32, 11)
    got = res.rows_of(0).cpu().numpy()
    assert got.shape == want.rotations.shape
    assert got.tobytes() == want.rotations.astype(np.float32).tobytes()
    assert res.knot_counts[0, : got.shape[0]].cpu().tolist() == list(want.knot_counts)
    if dt == "f32":
        # knot-index mismatch rate and rel-L2 of the pure-fp32 path against the fp64 arithmetic, per level
        # (reported, SURVEY.md section 0 item 9: only level 1 is bounded by 1e-4)
        ref = o.c_decompose(x32.astype(np.float64), 11)
        k32 = np.flatnonzero(o.np_knot_flags(got[0].astype(np.float32)))
        r1 = np.linalg.norm(got[0].astype(np.float64) - ref.rotations[0]) / np.linalg.norm(ref.rotations[0])
        assert r1 < 1e-4, r1
        assert int(res.input_knots[0]) == ref.input_knots            # level-1 knot indices: zero mismatches
        del k32


def test_config3_full_size_properties():
    """One 2**28-sample fp32 signal (f32_mixed: fp64 carry): size-independent checks at the full size."""
    n = 1 << 28
    x = synth.long_signal(n=n, seed=3, device="cuda")
    res = pyitd_b200.decompose(x.unsqueeze(0), max_iteration=11, dtype="f32_mixed")
    torch.cuda.synchronize()
    assert int(res.status[0]) == 0
    nr = int(res.n_rows[0])
    assert 2 <= nr <= 13
    # level-1 knot count equals the reference stencil evaluated on the device
    assert int(res.input_knots[0]) == int(_level1_knots_torch(x).sum())
    # rows sum back to the input (ITD.py:505-508) within fp32 rounding of each stored row
    recon = res.rotations[0, :nr].double().sum(dim=0)
    err = (recon - x.double()).abs().max().item()
    assert err < 1e-5, err
    # baseline[N-1] == 0 on every level (ITD.py:112) => rotation row 0 ends with x[N-1]
    assert float(res.rotations[0, 0, -1]) == float(x[-1])
    # (knot counts usually shrink level by level but need not: rounding can create new plateaus in a baseline)
    kc = res.knot_counts[0, :nr].cpu().tolist()
    assert kc[0] < int(res.input_knots[0]) and kc[-1] < kc[0]
    del res, recon
    torch.cuda.empty_cache()


def test_config4_audio_frames():
    """48 kHz audio, 8192-sample frames, fixed 8 iterations (max_iteration=7), fp32 in/out, fp64 carry."""
    fr = synth.audio_frames(seconds=60.0)                      # 351 frames here; the bench runs all 3515
    assert fr.shape[1] == 8192
    res = pyitd_b200.decompose(gpu(fr), max_iteration=7, dtype="f32_mixed", return_baselines=True)
    torch.cuda.synchronize()
    assert int(res.status.abs().max()) == 0
    assert int(res.n_rows.max()) <= 9
    for s in (0, 1, 100, 350):
        want = o.c_decompose(fr[s].astype(np.float64), 7)
        assert res.rows_of(s).cpu().numpy().tobytes() == want.rotations.astype(np.float32).tobytes()
        assert res.baselines_of(s).cpu().numpy().tobytes() == want.baselines.astype(np.float32).tobytes()
        assert int(res.stop_kind[s]) == want.stop_kind
    x = gpu(fr)
    nr = res.n_rows.long()
    rows = torch.arange(res.rotations.shape[1], device="cuda")[None, :, None]
    recon = torch.where(rows < nr[:, None, None], res.rotations.double(), torch.zeros((), device="cuda", dtype=torch.float64)).sum(dim=1)
    assert float((recon - x.double()).abs().max()) < 1e-5


def test_config5_chunked_shard_matches_unchunked():
    """Config 5 walks a shard of the channel axis in chunks that recycle one output buffer; chunking must
    not change any channel's result (same generator, same seeds, per-chunk plans)."""
    from pyitd_b200 import shard
    x = synth.eeg_like(640, 65536, seed=1234, device="cuda")
    whole = pyitd_b200.decompose(x, max_iteration=11)
    torch.cuda.synchronize()
    for c0, c1 in shard.chunk_ranges(0, 640, 256):
        part = pyitd_b200.decompose(x[c0:c1], max_iteration=11)
        assert torch.equal(part.n_rows, whole.n_rows[c0:c1])
        for s in (0, (c1 - c0) // 2, c1 - c0 - 1):
            assert torch.equal(part.rows_of(s), whole.rows_of(c0 + s))
    want = o.c_decompose(x[639].cpu().numpy(), 11)
    assert whole.rows_of(639).cpu().numpy().tobytes() == want.rotations.tobytes()


@pytest.mark.parametrize("path,S,n", [("stream", 170, 6000), ("stream", 161, 4116), ("strided", 1, 70000),
                                      ("strided", 2, 33000)])
def test_knot_ls_prepass_equals_in_kernel_knot_baseline(path, S, n, monkeypatch):
    """knot_ls_kernel (one thread per knot: L_k and the slope of its segment, ITD.py:106-110 / :116) feeds the tiles with
    few knots from level 1 on.  With the pre-pass switched off (PYITD_LS=0) every warp evaluates its own knots inside the
    level kernel: both must give the same bytes and the same status words, in every precision variant, and match the
    oracle."""
    monkeypatch.setenv("PYITD_FORCE_PATH", path)
    rng = np.random.default_rng(S * 31 + n)
    x = _mixed_batch(rng, S, n)
    x32 = x.astype(np.float32)
    try:
        out = {}
        for ls in ("0", "1"):
            monkeypatch.setenv("PYITD_LS", ls)
            pyitd_b200.clear_plan_cache()
            from pyitd_b200.itd import get_plan
            assert get_plan(0, S, n, _capi.F64, 11, 2, _capi.OPT_BASELINES).path[0] == path
            for dt, xin in (("f64", x), ("f32_mixed", x32), ("f32", x32)):
                r = pyitd_b200.decompose(gpu(xin), max_iteration=11, dtype=dt, return_baselines=True, zero_tail=True)
                torch.cuda.synchronize()
                out[ls, dt] = (r.rotations.cpu().numpy().tobytes(), r.baselines.cpu().numpy().tobytes(),
                               r.n_rows.cpu().tolist(), r.status.cpu().tolist(), r.knot_counts.cpu().tolist())
        for dt in ("f64", "f32_mixed", "f32"):
            assert out["0", dt] == out["1", dt], dt
        check_against_oracle(x[: min(S, 12)], max_iteration=11) if path == "strided" else check_against_oracle(x, max_iteration=11)
    finally:
        pyitd_b200.clear_plan_cache()
