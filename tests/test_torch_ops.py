"""torch.ops.pyitd.* (csrc/torch_ops.cpp): the thin PyTorch extension over the C ABI.

CPU: the extension builds against this interpreter's torch, registers the ops with the documented schemas, infers shapes
on the meta device and refuses CPU tensors.  GPU: every op equals the ctypes path and the oracle bit for bit."""
import os

import numpy as np
import pytest
import torch

import pyitd_b200
from pyitd_b200 import _capi, torch_ops


@pytest.fixture(scope="module")
def ops():
    if not os.path.exists(torch_ops.LIB_PATH):
        torch_ops.build()
    return torch_ops.load()


def test_ops_are_registered_with_the_documented_schemas(ops):
    assert ops.abi_version() == 1
    s = str(torch.ops.pyitd.decompose.default._schema)
    assert "int max_iteration=11" in s and "int min_extrema=2" in s and "bool return_baselines=False" in s
    assert "Tensor rotations, Tensor n_rows, Tensor knot_counts, Tensor baselines, Tensor status" in s
    assert "Tensor knots, Tensor count, Tensor status" in str(torch.ops.pyitd.find_knots.default._schema)
    assert "Tensor rotation, Tensor baseline, Tensor count, Tensor status" in str(torch.ops.pyitd.extract_level.default._schema)


def test_meta_shapes(ops):
    x = torch.empty(5, 300, dtype=torch.float64, device="meta")
    rot, n_rows, counts, bas, status, kind, iknots = ops.decompose(x, 7, 2, True)
    assert rot.shape == (5, 9, 300) and bas.shape == (5, 9, 300) and counts.shape == (5, 9)
    assert n_rows.shape == status.shape == kind.shape == iknots.shape == (5,) and n_rows.dtype == torch.int32
    assert ops.decompose(x)[3].numel() == 0 and ops.decompose(x)[0].shape == (5, 13, 300)
    R, B, c, st = ops.extract_level(x[0])
    assert R.shape == B.shape == (1, 300) and c.shape == (1,)
    assert ops.find_knots(x, 3, 64)[0].shape == (5, 64)


def test_cpu_tensors_are_refused_not_computed(ops):
    # no CPU dispatch key is registered: there is no fallback to fall back to
    x = torch.zeros(2, 100, dtype=torch.float64)
    for call in (lambda: ops.decompose(x), lambda: ops.extract_level(x), lambda: ops.find_knots(x)):
        with pytest.raises(NotImplementedError):
            call()


def test_extension_links_the_c_abi_and_not_the_oracle(ops):
    import subprocess
    out = subprocess.run(["ldd", torch_ops.LIB_PATH], capture_output=True, text=True).stdout
    assert "libpyitd_b200.so" in out and "itd_oracle" not in out
    src = open(os.path.join(os.path.dirname(torch_ops.LIB_PATH), "csrc", "torch_ops.cpp")).read()
    assert "__global__" not in src and "<<<" not in src            # no kernel lives in the torch layer


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape,max_iteration", [((3, 8192), 11), ((200, 4096), 5), ((1, 70000), 20), ((33, 1000), 0)])
def test_decompose_op_equals_ctypes_path_and_oracle(ops, shape, max_iteration):
    from oracle import itd_oracle
    g = torch.Generator(device="cuda").manual_seed(shape[0] * 7 + max_iteration)
    x = torch.randn(shape, dtype=torch.float64, device="cuda", generator=g).cumsum(1) * 0.1 \
        + torch.randn(shape, dtype=torch.float64, device="cuda", generator=g)
    rot, n_rows, counts, bas, status, kind, iknots = ops.decompose(x, max_iteration, 2, True, True)
    ref = pyitd_b200.decompose(x, max_iteration=max_iteration, return_baselines=True, zero_tail=True)
    torch.cuda.synchronize()
    assert torch.equal(n_rows, ref.n_rows) and torch.equal(kind, ref.stop_kind) and torch.equal(status, ref.status)
    assert torch.equal(iknots, ref.input_knots)
    assert rot.cpu().numpy().tobytes() == ref.rotations.cpu().numpy().tobytes()
    xs = x.cpu().numpy()
    for s in range(0, shape[0], max(1, shape[0] // 5)):
        want = itd_oracle.c_decompose(xs[s], max_iteration)
        nr = int(n_rows[s])
        assert nr == want.rotations.shape[0]
        assert rot[s, :nr].cpu().numpy().tobytes() == want.rotations.tobytes()
        nb = want.baselines.shape[0]
        assert bas[s, :nb].cpu().numpy().tobytes() == want.baselines.tobytes()
        assert counts[s, :len(want.knot_counts)].cpu().tolist() == list(want.knot_counts)


@pytest.mark.gpu
def test_level_and_knot_ops_equal_oracle(ops):
    from oracle import itd_oracle
    rng = np.random.default_rng(5)
    xs = rng.standard_normal((6, 5000))
    x = torch.from_numpy(xs).cuda()
    R, B, cnt, st = ops.extract_level(x)
    knots, kc, kst = ops.find_knots(x)
    valleys, vc, _ = ops.find_knots(x, _capi.KNOTS_VALLEYS)
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and torch.equal(cnt, kc)
    for s in range(xs.shape[0]):
        Ro, Bo = itd_oracle.c_extract_level(xs[s])[:2]
        assert R[s].cpu().numpy().tobytes() == Ro.tobytes() and B[s].cpu().numpy().tobytes() == Bo.tobytes()
        v = xs[s]
        want = np.flatnonzero((v[:-2] >= v[1:-1]) & (v[1:-1] < v[2:])) + 1          # detect_peaks, ITD.py:59
        assert valleys[s, : int(vc[s])].cpu().tolist() == want.tolist()
        assert knots[s, : int(kc[s])].cpu().tolist() == itd_oracle.c_find_knots(v).tolist()


@pytest.mark.gpu
def test_ops_run_on_the_current_stream_and_precisions(ops):
    x32 = torch.randn(8, 4096, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        a = ops.decompose(x32, 6)[0]                        # float32 -> f32_mixed
        b = ops.decompose(x32, 6, 2, False, False, "f32")[0]
    side.synchronize()
    ref = pyitd_b200.decompose(x32, max_iteration=6)
    ref32 = pyitd_b200.decompose(x32, max_iteration=6, dtype="f32")
    torch.cuda.synchronize()
    nr = ref.n_rows
    for s in range(8):
        assert torch.equal(a[s, : int(nr[s])], ref.rotations[s, : int(nr[s])])
        assert torch.equal(b[s, : int(ref32.n_rows[s])], ref32.rotations[s, : int(ref32.n_rows[s])])
    with pytest.raises(RuntimeError):
        ops.decompose(x32, 6, 2, False, False, "f64")
    with pytest.raises(RuntimeError):
        ops.decompose(torch.zeros(2, 2, device="cuda"))
    ops.clear_plans()
