"""GPU: the sweep kernel (itd_sweep.cuh) -- the whole decomposition of a batch in ONE persistent launch -- against
the oracle, bit for bit, through the C ABI.  This is the kernel the headline number of bench.py comes from.

Covered: lengths around the 128-sample span and the eight-region split (including lengths that are not a multiple of
4: the kernel has no alignment requirement), plateaus / ties / monotone / knot-free stretches, dense (per-warp scratch)
and sparse (shared-memory table) knot modes and the switch between them inside one decomposition, every stop kind,
min_extrema, baselines, zero tails, the fp32 variants, one launch per stage vs one launch for everything, the ticket
scheduler with fewer signals than CTAs and with more, and the full benchmark shape.
"""
import numpy as np
import pytest
import torch

import pyitd_b200
from oracle import itd_oracle as o
from pyitd_b200 import _capi, synth
from test_gpu_parity import _mixed_batch, check_against_oracle, check_batch_against_c_oracle, gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_sweep(monkeypatch):
    monkeypatch.setenv("PYITD_FORCE_PATH", "sweep")
    pyitd_b200.clear_plan_cache()
    yield
    pyitd_b200.clear_plan_cache()


def test_sweep_is_the_default_for_big_batches(monkeypatch):
    monkeypatch.delenv("PYITD_FORCE_PATH")
    pyitd_b200.clear_plan_cache()
    from pyitd_b200.itd import get_plan
    assert get_plan(0, 4096, 65536, _capi.F64, 11, 2, 0).path[0] == "sweep"
    assert get_plan(0, 3515, 8192, _capi.F32_MIXED, 7, 2, 0).path[0] == "sweep"
    assert get_plan(0, 200, 4098, _capi.F64, 11, 2, 0).path[0] == "sweep"          # rows need no alignment
    assert get_plan(0, 64, 65536, _capi.F64, 11, 2, 0).path[0] == "resident"
    assert get_plan(0, 1, 65536, _capi.F64, 11, 2, 0).path[0] == "coop"


@pytest.mark.parametrize("n", [3, 4, 5, 31, 33, 127, 128, 129, 130, 255, 257, 1023, 1024, 1025, 1026, 1027, 2049, 4100,
                               8191, 8193, 10007, 16384, 20001])
def test_lengths_around_spans_and_regions(n):
    rng = np.random.default_rng(7000 + n)
    check_against_oracle(_mixed_batch(rng, 13, n), max_iteration=11)


@pytest.mark.parametrize("max_iteration", [0, 1, 3, 7, 20])
def test_iteration_cap(max_iteration):
    rng = np.random.default_rng(7100)
    res = check_against_oracle(_mixed_batch(rng, 12, 8192), max_iteration=max_iteration)
    assert int(res.n_rows.max()) <= max_iteration + 2


@pytest.mark.parametrize("min_extrema", [0, 1, 3, 10])
def test_min_extrema(min_extrema):
    rng = np.random.default_rng(7200)
    check_against_oracle(_mixed_batch(rng, 6, 3000), max_iteration=11, min_extrema=min_extrema)


def test_dense_to_sparse_switch_inside_one_decomposition():
    """White noise has ~0.66 N knots on the first level (per-warp scratch mode) and drops below the 2046-knot table
    within two or three levels (block table mode); a slow sine stays sparse from the start; a zig-zag has a knot on
    EVERY interior sample (128 knots per span, the scratch capacity)."""
    rng = np.random.default_rng(7300)
    n = 40000
    x = rng.standard_normal((6, n))
    x[1] = np.sin(np.arange(n) * 0.001)
    x[2] = np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * (1 + 0.1 * rng.random(n))
    x[3] = np.cumsum(x[3])
    x[4, 5000:30000] = np.linspace(0, 3, 25000)            # one region sees no knot at all
    x[5] = np.round(x[5] * 3) / 3 + 1e-9 * np.arange(n)
    res = check_against_oracle(x, max_iteration=11)
    assert int(res.input_knots[2]) == n - 2 and int(res.input_knots[0]) > 2046 > int(res.input_knots[1])


def test_zero_tail_and_baselines():
    rng = np.random.default_rng(7400)
    x = _mixed_batch(rng, 8, 5000)
    res = pyitd_b200.decompose(gpu(x), max_iteration=11, return_baselines=True, zero_tail=True)
    ref = check_against_oracle(x, max_iteration=11)
    for s in range(x.shape[0]):
        nr = int(res.n_rows[s])
        assert torch.equal(res.rotations[s, :nr], ref.rotations[s, :nr])
        assert float(res.rotations[s, nr:].abs().max() if nr < res.rotations.shape[1] else 0) == 0
        nb = res.baselines_of(s).shape[0]
        assert torch.equal(res.baselines_of(s), ref.baselines_of(s))
        assert float(res.baselines[s, nb:].abs().max() if nb < res.baselines.shape[1] else 0) == 0
    # without the baselines (bench.py's options): same rows
    r0 = pyitd_b200.decompose(gpu(x), max_iteration=11)
    for s in range(x.shape[0]):
        assert torch.equal(r0.rows_of(s), ref.rows_of(s))


def test_fp32_variants():
    rng = np.random.default_rng(7500)
    for n in (130, 4096, 9001):
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


def test_error_statuses():
    x = np.random.default_rng(7600).standard_normal((5, 3000))
    x[1] = 3.0                                             # constant: zero delta-X (ITD.py:116 raises ZeroDivisionError)
    x[3, 100] = np.nan
    res = pyitd_b200.decompose(gpu(x))
    st = res.status.cpu().tolist()
    assert st[0] == 0 and st[2] == 0 and st[4] == 0
    assert st[1] & _capi.ST_ZERO_DX and st[3] & _capi.ST_NONFINITE


@pytest.mark.parametrize("S", [1, 3, 700])
def test_ticket_scheduler_with_few_and_many_signals(S):
    """S = 1: every item waits for the previous stage of the same signal; S = 700: more items per stage than resident
    CTAs (592), stopped signals skipped, ragged stops."""
    x = synth.eeg_like(S, 6000, seed=50 + S).numpy()
    x[0] = np.arange(6000.0)                               # stops at once
    res = pyitd_b200.decompose(gpu(x), max_iteration=11, return_baselines=True)
    torch.cuda.synchronize()
    check_batch_against_c_oracle(res, x, 11)


@pytest.mark.parametrize("n", [5, 129, 1000, 4100, 8192, 20001, 40000])
def test_separate_scan_stage_equals_scan_fused_into_extraction_0(n, monkeypatch):
    """PYITD_SWEEP_FUSED_SCAN=1: extraction 0 finds the knots of the raw input itself, chunk by chunk, straight into shared
    memory (no scan stage, no knot lists of the input).  Same bytes as with the separate scan stage (the default).
    The variant lost (profiles/r2/README.md) and is compiled only with -DPYITD_SWEEP_WITH_FUSED_SCAN."""
    if not _capi.has_feature("sweep_fused_scan"):
        pytest.skip("this build does not carry the scan-free extraction 0 (make EXTRA=-DPYITD_SWEEP_WITH_FUSED_SCAN)")
    rng = np.random.default_rng(7800 + n)
    x = _mixed_batch(rng, 13, n)
    a = check_against_oracle(x, max_iteration=11)
    monkeypatch.setenv("PYITD_SWEEP_FUSED_SCAN", "1")
    b = check_against_oracle(x, max_iteration=11)
    for s_ in range(x.shape[0]):
        if int(a.status[s_]) == 0:
            assert torch.equal(a.rows_of(s_), b.rows_of(s_)) and torch.equal(a.baselines_of(s_), b.baselines_of(s_))
    assert torch.equal(a.input_knots, b.input_knots) and torch.equal(a.status, b.status)


def test_one_launch_equals_one_launch_per_stage(monkeypatch):
    x = synth.eeg_like(40, 16384, seed=77, device="cuda")
    a = pyitd_b200.decompose(x, max_iteration=11, return_baselines=True, zero_tail=True)
    torch.cuda.synchronize()
    monkeypatch.setenv("PYITD_SWEEP_PER_STAGE", "1")
    b = pyitd_b200.decompose(x, max_iteration=11, return_baselines=True, zero_tail=True)
    torch.cuda.synchronize()
    assert torch.equal(a.rotations, b.rotations) and torch.equal(a.baselines, b.baselines)
    assert torch.equal(a.n_rows, b.n_rows) and torch.equal(a.knot_counts, b.knot_counts)
    check_batch_against_c_oracle(a, x.cpu().numpy(), 11)


@pytest.mark.parametrize("depth", ["0", "1"])
def test_item_order_does_not_change_results(depth, monkeypatch):
    """Stage-major tickets (an item waits for its signal's previous stage) and signal-major tickets (a CTA runs all
    stages of a signal back to back: the default for short signals) must give the same bytes; both against the oracle."""
    monkeypatch.setenv("PYITD_SWEEP_DEPTH", depth)
    rng = np.random.default_rng(7700)
    for S, n in ((700, 2048), (9, 8192), (37, 20000)):
        x = _mixed_batch(rng, S, n)
        res = pyitd_b200.decompose(gpu(x), max_iteration=7, return_baselines=True, zero_tail=True)
        torch.cuda.synchronize()
        rot, n_rows, counts, status, bas = o.c_decompose_batch(x, 7, want_baselines=True)
        ok = status == 0
        assert (res.status.cpu().numpy() != 0).tolist() == (~ok).tolist()
        assert res.n_rows.cpu().numpy()[ok].tolist() == n_rows[ok].tolist()
        got = res.rotations.cpu().numpy()
        for s_ in np.flatnonzero(ok):
            nr = int(n_rows[s_])
            assert got[s_, :nr].tobytes() == rot[s_, :nr].tobytes(), (depth, S, n, s_)
            assert float(np.abs(got[s_, nr:]).max() if nr < got.shape[1] else 0) == 0
    # config 4's shape (framed audio, fp32 in / out around an fp64 carry)
    fr = synth.audio_frames(seconds=30.0)
    r = pyitd_b200.decompose(gpu(fr), max_iteration=7, dtype="f32_mixed", return_baselines=True)
    torch.cuda.synchronize()
    for s_ in (0, 7, fr.shape[0] - 1):
        want = o.c_decompose(fr[s_].astype(np.float64), 7)
        assert r.rows_of(s_).cpu().numpy().tobytes() == want.rotations.astype(np.float32).tobytes()
        assert r.baselines_of(s_).cpu().numpy().tobytes() == want.baselines.astype(np.float32).tobytes()


def test_benchmark_shape_256_channels_bit_exact():
    """bench.py's workload shape and generator: 256 channels x 65 536 samples, EVERY channel bit for bit -- rotations,
    baselines, per-level knot counts, stop kind (ITD.py:351-433) -- with and without the baselines."""
    from pyitd_b200.itd import get_plan
    assert get_plan(0, 256, 65536, _capi.F64, 11, 2, _capi.OPT_BASELINES).path[0] == "sweep"
    xg = synth.eeg_like(256, 65536, seed=1234, device="cuda")
    res = pyitd_b200.decompose(xg, max_iteration=11, return_baselines=True)
    torch.cuda.synchronize()
    check_batch_against_c_oracle(res, xg.cpu().numpy(), 11)
    assert len(set(res.n_rows.cpu().tolist())) > 2
    res0 = pyitd_b200.decompose(xg, max_iteration=11)
    torch.cuda.synchronize()
    check_batch_against_c_oracle(res0, xg.cpu().numpy(), 11, baselines=False)


def test_repeated_calls_and_side_stream():
    """A plan is reused across calls (ticket / done counters are reset per call) and across streams (the library orders
    a call after the plan's previous call on another stream)."""
    x = synth.eeg_like(170, 4096, seed=9, device="cuda")
    ref = pyitd_b200.decompose(x, max_iteration=11)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    outs = []
    for i in range(4):
        if i % 2:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                outs.append(pyitd_b200.decompose(x, max_iteration=11))
        else:
            outs.append(pyitd_b200.decompose(x, max_iteration=11))
    torch.cuda.synchronize()
    for r in outs:
        assert torch.equal(r.n_rows, ref.n_rows)
        for s in (0, 85, 169):
            assert torch.equal(r.rows_of(s), ref.rows_of(s))
    check_batch_against_c_oracle(ref, x.cpu().numpy(), 11, baselines=False)


def _last_plan():
    from pyitd_b200 import itd as _itd
    return next(reversed(_itd._PLAN_CACHE.values()))


@pytest.mark.parametrize("thr", [None, ("4", "1640", "2"), ("4", "60", "2"), ("200", "2200", "50")])
@pytest.mark.parametrize("depth", ["0", "1"])
def test_fused_pairs_equal_single_extractions(thr, depth, monkeypatch):
    """Two consecutive few-knot extractions run as ONE item: the knots of B_e are predicted from the knot table (the extrema
    stencil at the knots of X_e), one pass reads X_e and writes R_e, R_{e+1}, B_{e+1} and checks the prediction on every
    sample; B_e is never stored.  Same bytes as one item per extraction and as the oracle (ITD.py:79-121 applied twice, the
    stop test ITD.py:404 in between).  The batch holds plateaus, ties and flat steps (predictions that FAIL the check: the
    extraction is redone on its own) next to smooth signals (predictions that hold); the thresholds are moved so that every
    way out of a pair is taken: fused, the second extraction is the discarded last one (row e + 1 = B_e recomputed), too
    few / too many knots for a pair."""
    monkeypatch.setenv("PYITD_SWEEP_DEPTH", depth)
    rng = np.random.default_rng(9100)
    fused_total = skipped_total = failed_total = 0
    for S, n, mi in ((200, 16384, 11), (21, 65536, 11), (300, 2048, 11), (40, 5001, 5), (64, 8192, 20)):
        x = _mixed_batch(rng, S, n) if n < 10000 else synth.eeg_like(S, n, seed=n + S, device="cpu").numpy()
        if n == 16384:
            x[::5] = np.round(x[::5] * 64) / 64                      # quantised channels: flat steps inside the baselines
        monkeypatch.setenv("PYITD_SWEEP_FUSE", "0")
        pyitd_b200.clear_plan_cache()
        a = pyitd_b200.decompose(gpu(x), max_iteration=mi, return_baselines=True, zero_tail=True)
        torch.cuda.synchronize()
        assert _last_plan().sweep_stats() == (0, 0, 0)
        monkeypatch.setenv("PYITD_SWEEP_FUSE", "2")               # (2: also in signal-major order)
        if thr is not None:
            for k, v in zip(("MIN_A", "MAX_A", "MIN_B"), thr):
                monkeypatch.setenv("PYITD_SWEEP_FUSE_" + k, v)
        pyitd_b200.clear_plan_cache()
        b = pyitd_b200.decompose(gpu(x), max_iteration=mi, return_baselines=True, zero_tail=True)
        torch.cuda.synchronize()
        f, u, bad = _last_plan().sweep_stats()
        fused_total += f
        skipped_total += u
        failed_total += bad
        assert torch.equal(a.status, b.status)
        ok = (a.status == 0).cpu().numpy()
        okt = torch.from_numpy(ok).to(a.rotations.device)
        assert torch.equal(a.rotations[okt], b.rotations[okt]), (S, n)
        assert torch.equal(a.n_rows[okt], b.n_rows[okt]) and torch.equal(a.stop_kind[okt], b.stop_kind[okt])
        assert torch.equal(a.knot_counts[okt], b.knot_counts[okt])
        ab, bb = a.baselines.cpu().numpy(), b.baselines.cpu().numpy()
        nr, kind = a.n_rows.cpu().numpy(), a.stop_kind.cpu().numpy()
        for s_ in np.flatnonzero(ok):
            nb = nr[s_] if kind[s_] == _capi.STOP_ITER else nr[s_] - 1          # the discarded extraction's baseline row is not defined
            assert ab[s_, :nb].tobytes() == bb[s_, :nb].tobytes(), (S, n, s_)
            assert not ab[s_, nr[s_]:].any() and not bb[s_, nr[s_]:].any()
        # and the fused run against the oracle
        for s_ in list(np.flatnonzero(ok)[:12]):
            want = o.c_decompose(x[s_], mi)
            assert b.rows_of(int(s_)).cpu().numpy().tobytes() == want.rotations.tobytes(), (S, n, s_)
            assert b.baselines_of(int(s_)).cpu().numpy().tobytes() == want.baselines.tobytes(), (S, n, s_)
    assert fused_total > 0
    if thr is not None and thr[0] == "4":
        assert skipped_total > 0 and failed_total > 0


@pytest.mark.parametrize("dt", ["f32_mixed", "f32"])
def test_fused_pairs_fp32_variants(dt, monkeypatch):
    """The fused pairs in the two float32 variants (fp64 carry around float32 rows; everything in binary32): same bytes as one
    item per extraction, and as the rounded / binary32 oracle."""
    monkeypatch.setenv("PYITD_SWEEP_DEPTH", "0")
    x32 = synth.eeg_like(170, 20000, seed=77, device="cpu").numpy().astype(np.float32)
    x32[::7] = np.round(x32[::7] * 32) / 32                        # flat steps: failed predictions
    monkeypatch.setenv("PYITD_SWEEP_FUSE", "0")
    pyitd_b200.clear_plan_cache()
    a = pyitd_b200.decompose(gpu(x32), max_iteration=11, dtype=dt, return_baselines=True, zero_tail=True)
    torch.cuda.synchronize()
    monkeypatch.setenv("PYITD_SWEEP_FUSE", "2")
    pyitd_b200.clear_plan_cache()
    b = pyitd_b200.decompose(gpu(x32), max_iteration=11, dtype=dt, return_baselines=True, zero_tail=True)
    torch.cuda.synchronize()
    f, _, bad = _last_plan().sweep_stats()
    assert f > 0 and bad > 0
    ok = ((a.status == 0) & (b.status == 0))
    assert torch.equal(a.status, b.status) and int(ok.sum()) > 100
    assert torch.equal(a.rotations[ok], b.rotations[ok]) and torch.equal(a.n_rows[ok], b.n_rows[ok])
    assert torch.equal(a.knot_counts[ok], b.knot_counts[ok])
    for s_ in [int(i) for i in torch.nonzero(ok).flatten()[:10]]:
        if dt == "f32":
            want = o.c_decompose(x32[s_], 11)
            assert b.rows_of(s_).cpu().numpy().tobytes() == want.rotations.tobytes(), s_
        else:
            want = o.c_decompose(x32[s_].astype(np.float64), 11)
            assert b.rows_of(s_).cpu().numpy().tobytes() == want.rotations.astype(np.float32).tobytes(), s_


@pytest.mark.parametrize("shape", [(200, 20000, "sweep"), (1, 65536, "coop"), (3, 30000, "coop")])
def test_cuda_graph_capture_and_replay(shape, monkeypatch):
    """A decomposition is a fixed sequence of memsets and launches with no host round trip (the level loop and the stop test
    run on the device: ONE persistent / cooperative launch), so it can be captured into a CUDA graph once the plan's
    workspace exists, and replayed; warm-up on another stream than the capturing one on purpose."""
    monkeypatch.delenv("PYITD_FORCE_PATH")
    pyitd_b200.clear_plan_cache()
    from pyitd_b200.itd import get_plan
    S, N, path = shape
    x = synth.eeg_like(S, N, seed=3, device="cuda")
    plan = get_plan(0, S, N, _capi.F64, 11, 2, 0)
    assert plan.path[0] == path
    rows = plan.rows
    rot = torch.zeros((S, rows, N), dtype=torch.float64, device="cuda")
    ints = [torch.zeros(S * (rows if i == 1 else 1), dtype=torch.int32, device="cuda") for i in range(5)]

    def step(st):
        plan.decompose_device(x.data_ptr(), rot.data_ptr(), None, ints[0].data_ptr(), ints[1].data_ptr(), ints[2].data_ptr(),
                              ints[3].data_ptr(), ints[4].data_ptr(), st.cuda_stream)

    step(torch.cuda.current_stream())
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step(side)
    for _ in range(3):
        rot.zero_()
        for t in ints:
            t.fill_(-7)
        g.replay()
        torch.cuda.synchronize()
        assert int(ints[4].abs().max()) == 0
        xs = x.cpu().numpy()
        for s_ in range(min(S, 4)):
            want = o.c_decompose(xs[s_], 11)
            nr = int(ints[0][s_])
            assert nr == want.rotations.shape[0]
            assert rot[s_, :nr].cpu().numpy().tobytes() == want.rotations.tobytes()
