"""GPU: the cooperative kernel (itd_coop.cuh) -- a handful of signals, each kept in the shared memory of a group of CTAs
for the whole decomposition, ONE launch, one group barrier per extraction -- against the oracle, bit for bit, through the
C ABI.  This is the path the reference's own use takes (ITD().itd(x) on one signal, ITD.py:500-503); the golden
vectors of the reference run through it in test_gpu_parity.py (the drop-in class).

Covered: lengths around the 256-sample round and the chunk boundaries, one CTA per signal and many, more signals than
groups (a group works through several signals: barrier counters and summary slots carry over), plateaus / ties /
monotone / knot-free stretches (chunks without knots: the neighbour search walks over them), every stop kind,
min_extrema, baselines, zero tails, the fp32 variants, error statuses, repeated calls on one plan.
"""
import numpy as np
import pytest
import torch

import pyitd_b200
from oracle import itd_oracle as o
from pyitd_b200 import _capi, synth
from test_gpu_parity import _mixed_batch, check_against_oracle, gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_coop(monkeypatch):
    monkeypatch.setenv("PYITD_FORCE_PATH", "coop")
    pyitd_b200.clear_plan_cache()
    yield
    pyitd_b200.clear_plan_cache()


def _plan_path():
    from pyitd_b200 import itd as _itd
    return next(reversed(_itd._PLAN_CACHE.values())).path[0]


@pytest.mark.parametrize("n", [3, 4, 5, 31, 33, 255, 256, 257, 258, 511, 512, 513, 767, 769, 1000, 1025, 4097, 20001])
@pytest.mark.parametrize("chunk", [None, "256"])
def test_lengths_around_chunks(n, chunk, monkeypatch):
    if chunk:
        monkeypatch.setenv("PYITD_COOP_CHUNK", chunk)          # many CTAs per signal even for short signals
    rng = np.random.default_rng(4200 + n)
    check_against_oracle(_mixed_batch(rng, 7, n), max_iteration=11)
    assert _plan_path() == "coop"


@pytest.mark.parametrize("n", [150001, 262143])
def test_longest_signals_of_the_path(n):
    """Just below 2^18 samples (from there on the strided path takes a single signal): 1792-sample chunks on 147 CTAs."""
    x = synth.eeg_like(2, n, seed=n, device="cpu").numpy()
    check_against_oracle(x, max_iteration=11)
    assert _plan_path() == "coop"


@pytest.mark.parametrize("S", [1, 2, 5, 16, 40])
def test_signals_per_group(S, monkeypatch):
    """S = 40 forced onto this path: fewer groups than signals on 65 536-sample signals (128 CTAs each), so a group works
    through several signals one after the other."""
    n = 65536 if S >= 16 else 30000
    x = synth.eeg_like(S, n, seed=S, device="cpu").numpy()
    check_against_oracle(x, max_iteration=11)
    assert _plan_path() == "coop"


def test_single_signal_default_path_and_golden_shape(monkeypatch):
    monkeypatch.delenv("PYITD_FORCE_PATH")
    pyitd_b200.clear_plan_cache()
    x = synth.eeg_like(1, 65536, seed=5, device="cpu").numpy()
    res = check_against_oracle(x, max_iteration=11)
    assert _plan_path() == "coop"
    # the drop-in class on the same signal
    rows = pyitd_b200.ITD().itd(x[0])
    assert rows.tobytes() == res.rows_of(0).cpu().numpy().tobytes()


@pytest.mark.parametrize("max_iteration", [0, 1, 3, 20])
def test_iteration_cap(max_iteration):
    rng = np.random.default_rng(4300)
    check_against_oracle(_mixed_batch(rng, 6, 5000), max_iteration=max_iteration)


@pytest.mark.parametrize("min_extrema", [0, 1, 3, 10])
def test_min_extrema(min_extrema):
    rng = np.random.default_rng(4400)
    check_against_oracle(_mixed_batch(rng, 6, 3000), max_iteration=11, min_extrema=min_extrema)


def test_sparse_levels_walk_over_empty_chunks(monkeypatch):
    """A slow sine on 256-sample chunks: most chunks hold no knot at all, the knots a chunk needs lie many chunks away."""
    monkeypatch.setenv("PYITD_COOP_CHUNK", "256")
    t = np.arange(60000, dtype=np.float64)
    x = np.stack([np.sin(t * 2 * np.pi / 23000.0) + 1e-3 * np.sin(t * 0.7),
                  np.sin(t * 2 * np.pi / 9000.0),
                  np.cos(t * 2 * np.pi / 59000.0) + 1e-9 * t])
    check_against_oracle(x, max_iteration=11)


def test_zero_tail_and_baselines():
    x = synth.eeg_like(3, 10000, seed=3, device="cpu").numpy()
    for mi in (2, 11, 20):
        res = pyitd_b200.decompose(gpu(x), max_iteration=mi, return_baselines=True, zero_tail=True)
        torch.cuda.synchronize()
        rot, bas = res.rotations.cpu().numpy(), res.baselines.cpu().numpy()
        for s_ in range(x.shape[0]):
            want = o.c_decompose(x[s_], mi)
            nr = want.rotations.shape[0]
            assert int(res.n_rows[s_]) == nr and int(res.stop_kind[s_]) == want.stop_kind
            assert rot[s_, :nr].tobytes() == want.rotations.tobytes()
            assert not rot[s_, nr:].any()
            nb = want.baselines.shape[0]
            assert bas[s_, :nb].tobytes() == want.baselines.tobytes()
            assert not bas[s_, nb:].any()
        res0 = pyitd_b200.decompose(gpu(x), max_iteration=mi)
        torch.cuda.synchronize()
        for s_ in range(x.shape[0]):
            assert res0.rows_of(s_).cpu().numpy().tobytes() == o.c_decompose(x[s_], mi).rotations.tobytes()


def test_fp32_variants():
    x32 = synth.eeg_like(4, 12000, seed=8, device="cpu").numpy().astype(np.float32)
    r = pyitd_b200.decompose(gpu(x32), max_iteration=11, dtype="f32_mixed", return_baselines=True)
    torch.cuda.synchronize()
    for s_ in range(4):
        want = o.c_decompose(x32[s_].astype(np.float64), 11)
        assert r.rows_of(s_).cpu().numpy().tobytes() == want.rotations.astype(np.float32).tobytes()
        assert r.baselines_of(s_).cpu().numpy().tobytes() == want.baselines.astype(np.float32).tobytes()
    r = pyitd_b200.decompose(gpu(x32), max_iteration=11, dtype="f32", return_baselines=True)
    torch.cuda.synchronize()
    for s_ in range(4):
        want = o.c_decompose(x32[s_], 11)
        assert r.rows_of(s_).cpu().numpy().tobytes() == want.rotations.tobytes()
        assert r.baselines_of(s_).cpu().numpy().tobytes() == want.baselines.tobytes()


def test_error_statuses():
    x = np.random.default_rng(1).standard_normal((5, 3000))
    x[1] = 2.5                                   # constant: zero dx
    x[3, 1700] = np.nan
    res = pyitd_b200.decompose(gpu(x), max_iteration=11)
    torch.cuda.synchronize()
    st = res.status.cpu().tolist()
    assert st[0] == st[2] == st[4] == 0
    assert st[1] & _capi.ST_ZERO_DX and st[3] & _capi.ST_NONFINITE
    for s_ in (0, 2, 4):
        assert res.rows_of(s_).cpu().numpy().tobytes() == o.c_decompose(x[s_], 11).rotations.tobytes()


def test_repeated_calls_and_side_stream():
    """The barrier counters go back to zero at the end of every launch; a plan is reused across calls and streams."""
    x = synth.eeg_like(3, 40000, seed=9, device="cuda")
    ref = pyitd_b200.decompose(x, max_iteration=11)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    for i in range(6):
        if i % 2:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = pyitd_b200.decompose(x, max_iteration=11)
            torch.cuda.current_stream().wait_stream(side)
        else:
            out = pyitd_b200.decompose(x, max_iteration=11)
        torch.cuda.synchronize()
        assert torch.equal(out.n_rows, ref.n_rows)
        for s_ in range(3):                                   # (rows beyond n_rows are not written without zero_tail)
            assert torch.equal(out.rows_of(s_), ref.rows_of(s_))
    check_against_oracle(x.cpu().numpy(), max_iteration=11)
