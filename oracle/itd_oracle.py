"""CPU oracle for the PyITD sifting loop -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module.  Nothing under ``pyitd_b200/`` does; the product path
raises when its CUDA library is missing instead of falling back to anything in here.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference's own numba code
from ``/root/reference/ITD.py`` in the build container and stores its outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks both restatements in this directory (the numpy
one below and the scalar C one in ``itd_oracle.c``) against those files bit for bit.

Two restatements, written independently of each other:

* ``np_*``  -- whole-array numpy (prefix sums instead of the reference's per-segment slices);
* ``c_*``   -- ctypes wrappers over ``libitd_oracle.so`` (scalar loops, pthreads over channels),
               fast enough for 2**22-sample signals and for the CPU-baseline timing.

All ``file:line`` citations are into ``/root/reference``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

ITD_OK, ITD_ZERO_DX, ITD_NONFINITE, ITD_TOO_SHORT, ITD_BAD_KNOTS, ITD_FEW_KNOTS = 0, 1, 2, 3, 8, 16


class OracleError(Exception):
    """Carries the status code the reference would have turned into an exception."""

    def __init__(self, status: int):
        super().__init__({1: "zero delta-X (ZeroDivisionError, ITD.py:116)",
                          2: "non-finite input", 3: "signal shorter than 3 samples",
                          8: "supplied knots are not strictly increasing inside [1, n-2]",
                          16: "fewer than 2 interior knots (splrep: m > k must hold)"}.get(status, str(status)))
        self.status = status


# --------------------------------------------------------------------------------------------
# numpy restatement
# --------------------------------------------------------------------------------------------
def np_knot_flags(x: np.ndarray) -> np.ndarray:
    """Boolean mask of the union ``detect_peaks(x) | detect_peaks(-x)``.

    ITD.py:44 forms ``dx = x[1:] - x[:-1]``; ITD.py:59 keeps ``i`` with ``dx[i] > 0 and
    dx[i-1] <= 0``; ITD.py:87-88 runs that on ``x`` and ``-x``; ITD.py:70-73 drop both ends.
    """
    n = x.shape[0]
    flags = np.zeros(n, dtype=bool)
    if n < 3:
        return flags
    dx = x[1:] - x[:-1]
    left, right = dx[:-1], dx[1:]          # dx[i-1], dx[i] for i = 1 .. n-2
    flags[1:-1] = ((right > 0) & (left <= 0)) | ((-right > 0) & (-left <= 0))
    return flags


def np_find_knots(x: np.ndarray) -> np.ndarray:
    """Sorted interior knot indices (int64) -- ITD.py:97 ``sort(unique(hstack(...)))``."""
    return np.flatnonzero(np_knot_flags(x)).astype(np.int64)


def np_extract_with_knots(x: np.ndarray, knots: np.ndarray):
    """ITD.py:95-119 with the interior knots GIVEN (ascending, inside [1, n-2]): the body of
    ``itd_baseline_extract`` when ``knots`` are x's own, and the reference's "reuse the extrema along
    multiple channels" mode (itd.cpp:41-44, ``compute_extrema=false`` at itd.cpp:156-169) on ITD.py's
    interpolant when they come from another channel.  Returns ``(R, B)``; dtype follows ``x``."""
    ft = x.dtype.type
    n = x.shape[0]
    if n < 3:
        raise OracleError(ITD_TOO_SHORT)
    if not np.all(np.isfinite(x)):
        raise OracleError(ITD_NONFINITE)
    knots = np.asarray(knots, dtype=np.int64)
    if knots.size and (knots[0] < 1 or knots[-1] > n - 2 or np.any(np.diff(knots) <= 0)):
        raise OracleError(ITD_BAD_KNOTS)
    flags = np.zeros(n, dtype=bool)
    flags[knots] = True
    tau = np.concatenate(([0], knots, [n - 1])).astype(np.int64)              # ITD.py:95-98
    X = x[tau]
    L = np.empty(tau.shape[0], dtype=x.dtype)
    L[0] = ((ft(0.0) + x[0]) + x[1]) / ft(2.0)                                # ITD.py:101
    L[-1] = ((ft(0.0) + x[-2]) + x[-1]) / ft(2.0)                             # ITD.py:102
    if knots.size:                                                            # ITD.py:106-110
        w = ((tau[1:-1] - tau[:-2]).astype(np.float64)
             / (tau[2:] - tau[:-2]).astype(np.float64)).astype(x.dtype)
        d = X[2:] - X[:-2]
        q = X[:-2] + w * d
        L[1:-1] = ft(0.5) * q + ft(0.5) * X[1:-1]
    den = X[1:] - X[:-1]                                                      # ITD.py:116
    if np.any(den == 0):
        raise OracleError(ITD_ZERO_DX)
    slope = (L[1:] - L[:-1]) / den
    seg = np.cumsum(flags)                  # sample t lies in [tau_seg, tau_seg+1)
    seg[-1] = min(seg[-1], slope.shape[0] - 1)
    u = x - X[seg]
    v = slope[seg] * u
    B = L[seg] + v                                                            # ITD.py:115-117
    B[-1] = 0                                                                 # ITD.py:112
    R = x - B                                                                 # ITD.py:119
    return R, B


def np_extract_level(x: np.ndarray):
    """One sifting level, ITD.py:79-121.  Returns ``(R, B, knots)``; dtype follows ``x``
    (float64 = the reference; float32 = the product's pure-fp32 variant, same operation order
    with the knot weight rounded once from an exact double quotient)."""
    if x.shape[0] < 3:
        raise OracleError(ITD_TOO_SHORT)
    if not np.all(np.isfinite(x)):
        raise OracleError(ITD_NONFINITE)
    knots = np.flatnonzero(np_knot_flags(x)).astype(np.int64)                 # ITD.py:87-97
    R, B = np_extract_with_knots(x, knots)
    return R, B, knots


def np_spline_level(x: np.ndarray):
    """SURVEY 8f rank 2 -- one level of the cubic-spline baseline variant: ``itd_baseline_extract``
    of MEITD.py:303-338 (returns rotation and baseline) == ``itd_baseline_extract_modified`` of
    numba_accelerated_itd.py:183-211 (returns the baseline).  Knots and knot baseline as in ITD.py; end
    knots = mean of the odd-reflected pad (MEITD.py:323-325); baseline = ``splev(arange(n),
    splrep(tau, L, k=3))`` (MEITD.py:330-333), scipy's FITPACK interpolating spline with not-a-knot
    ends, restated through the moment equations (see oracle/itd_oracle.c).  float64 only.
    Returns ``(R, B, knots)``."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    if n < 3:
        raise OracleError(ITD_TOO_SHORT)
    if not np.all(np.isfinite(x)):
        raise OracleError(ITD_NONFINITE)
    knots = np_find_knots(x)
    K = knots.shape[0]
    if K < 2:
        raise OracleError(ITD_FEW_KNOTS)
    tau = np.concatenate(([0], knots, [n - 1])).astype(np.int64)
    X = x[tau]
    y = np.empty(K + 2)
    y[0] = ((2.0 * x[0] - x[1]) + x[0]) / 2.0
    y[-1] = (x[-1] + (2.0 * x[-1] - x[-2])) / 2.0
    w = (tau[1:-1] - tau[:-2]) / (tau[2:] - tau[:-2])
    y[1:-1] = 0.5 * (X[:-2] + w * (X[2:] - X[:-2])) + 0.5 * X[1:-1]
    h = np.diff(tau).astype(np.float64)
    a = h[:-1] / (h[:-1] + h[1:])
    c = h[1:] / (h[:-1] + h[1:])
    b = np.full(K, 2.0)
    d = 6.0 * ((y[2:] - y[1:-1]) / h[1:] - (y[1:-1] - y[:-2]) / h[:-1]) / (h[:-1] + h[1:])
    r0, rK = h[0] / h[1], h[K] / h[K - 1]
    b[0] += a[0] * (1.0 + r0); c[0] -= a[0] * r0; a[0] = 0.0
    b[K - 1] += c[K - 1] * (1.0 + rK); a[K - 1] -= c[K - 1] * rK; c[K - 1] = 0.0
    cp = np.empty(K); dp = np.empty(K)
    cp[0] = c[0] / b[0]; dp[0] = d[0] / b[0]
    for i in range(1, K):
        m = b[i] - a[i] * cp[i - 1]
        cp[i] = c[i] / m
        dp[i] = (d[i] - a[i] * dp[i - 1]) / m
    M = np.empty(K + 2)
    M[K] = dp[K - 1]
    for i in range(K - 2, -1, -1):
        M[i + 1] = dp[i] - cp[i] * M[i + 2]
    M[0] = (1.0 + r0) * M[1] - r0 * M[2]
    M[K + 1] = (1.0 + rK) * M[K] - rK * M[K - 1]
    t = np.arange(n)
    seg = np.minimum(np.searchsorted(tau, t, side="right") - 1, K)
    u = (t - tau[seg]).astype(np.float64)
    hj = h[seg]
    c1 = (y[seg + 1] - y[seg]) / hj - hj * (2.0 * M[seg] + M[seg + 1]) / 6.0
    c2 = M[seg] / 2.0
    c3 = (M[seg + 1] - M[seg]) / (6.0 * hj)
    B = y[seg] + u * (c1 + u * (c2 + u * c3))
    return x - B, B, knots


# --------------------------------------------------------------------------------------------
# 2-D crossways ensemble ITD (siftED2D.ipynb code cell 1, raw JSON :233-278)
# --------------------------------------------------------------------------------------------
def spline_baseline_min10(x: np.ndarray, level=None) -> np.ndarray:
    """The notebook's 1-D ``itd_baseline_extract`` (== numba_accelerated_itd.py:183-211): the spline baseline,
    or the input itself when it has fewer than 10 extrema."""
    level = level or c_spline_level
    x = np.ascontiguousarray(x, dtype=np.float64)
    if np_find_knots(x).shape[0] < 10:
        return x.copy()
    return level(x)[1]


def crossways(data: np.ndarray, level=None) -> np.ndarray:
    """``crossways_itd_baseline_extract`` (siftED2D.ipynb cell 1): rows, columns, rows of the column result,
    columns of the row result, mean of the two."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    H, W = data.shape
    lengthwise = np.zeros((H, W))
    crosswise = np.zeros((H, W))
    for r in range(H):
        lengthwise[r, :] = spline_baseline_min10(data[r, :], level)
    for c in range(W):
        crosswise[:, c] = spline_baseline_min10(data[:, c], level)
    for r in range(H):
        crosswise[r, :] = spline_baseline_min10(crosswise[r, :], level)
    for c in range(W):
        lengthwise[:, c] = spline_baseline_min10(lengthwise[:, c], level)
    return (lengthwise + crosswise) / 2.0


def ensemble2d(data: np.ndarray, noise: np.ndarray, level=None) -> np.ndarray:
    """``retrieve_statistical_image_component`` (siftED2D.ipynb cell 1) with the noise draws ``noise[e]`` given
    instead of drawn: members ``v + data`` and ``v * -1 + data``, crossways on each, pair means, mean over draws
    accumulated in draw order.  Returns the low-pass image; ``totalextract2d`` is ``[data - lowpass, lowpass]``."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    draws = noise.shape[0]
    x = np.zeros_like(data)
    for e in range(draws):
        a = crossways(noise[e] + data, level)
        b = crossways((noise[e] * -1) + data, level)
        x += (a + b) / 2.0
    return x / (draws * 1.0)


# --------------------------------------------------------------------------------------------
# post-decomposition analytics
# --------------------------------------------------------------------------------------------
def c_wpe3(x: np.ndarray, normalize: bool = True) -> float:
    """``weighted_permutation_entropy(x, order=3, normalize)`` of MEITD.py:79-128 (see itd_oracle.c)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    return float(c_lib().itd_oracle_wpe3_f64(_ptr(x), x.shape[0], int(bool(normalize))))


def np_wpe3(x: np.ndarray, normalize: bool = True) -> float:
    """numpy restatement of the same function (vectorised pattern / weight computation, sequential sums)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    if n < 3:
        return -0.0
    a, b, c = x[:-2], x[1:-1], x[2:]
    slot = np.where(a <= b, np.where(b <= c, 5, np.where(a <= c, 3, 2)), np.where(a <= c, 4, np.where(b <= c, 1, 0)))
    mean = ((a + b) + c) / 3.0
    w = (((a - mean) ** 2 + (b - mean) ** 2) + (c - mean) ** 2) / 3.0
    wc = []
    for k in range(6):
        sel = w[slot == k]
        if sel.shape[0]:
            acc = 0.0
            for v in sel.tolist():          # MEITD.py:113-117 accumulates in window order
                acc += v
            wc.append(acc)
    total = wc[0]
    for v in wc[1:]:
        total += v
    acc = None
    with np.errstate(all="ignore"):
        for v in wc:
            p = np.float64(v) / np.float64(total)
            term = float(p * np.log2(p))
            acc = term if acc is None else acc + term
    pe = -acc
    if normalize:
        pe /= float(np.log2(6.0))
    return pe


def c_column_fsum(rows: np.ndarray) -> np.ndarray:
    """``shewchuk(a)`` of helperfunctions.py:2-9 (== inner loop of ITD.py:475-481): exactly rounded column sums."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    out = np.empty(rows.shape[1])
    c_lib().itd_oracle_column_fsum_f64(_ptr(rows), rows.shape[0], rows.shape[1], _ptr(out))
    return out


@dataclass
class OracleResult:
    rotations: np.ndarray        # (n_rows, N): proper rotations then the final trend row
    baselines: np.ndarray        # (n_baselines, N) as ITD.get_baselines() returns them
    knot_counts: np.ndarray      # what ITD.py:403 prints, one per loop pass
    input_knots: int
    stop_kind: int               # 1 knot-count stop, 2 iteration stop


def np_decompose(x: np.ndarray, max_iteration: int = 11, min_extrema: int = 2) -> OracleResult:
    """The level loop ITD.py:389-432 (row cap generalised to max_iteration + 2)."""
    x = np.ascontiguousarray(x)
    n = x.shape[0]
    rots, bases, counts = [], [], []
    cur = x
    R, B, knots = np_extract_level(cur)                                       # ITD.py:389
    input_knots = int(knots.size)
    c = 0
    while True:
        ne = int(np.count_nonzero(np_knot_flags(B)))                          # ITD.py:400-402
        counts.append(ne)
        if ne < min_extrema:                                                  # ITD.py:404-416
            final = cur.copy() if c > 0 else np.zeros(n, dtype=x.dtype)
            return OracleResult(np.stack(rots + [final]),
                                np.stack(bases) if bases else np.zeros((0, n), x.dtype),
                                np.asarray(counts), input_knots, 1)
        if c > max_iteration:                                                 # ITD.py:418-426
            return OracleResult(np.stack(rots + [R + B]),
                                np.stack(bases + [np.zeros(n, dtype=x.dtype)]),
                                np.asarray(counts), input_knots, 2)
        rots.append(R)                                                        # ITD.py:429-432
        bases.append(B)
        cur = B
        R, B, _ = np_extract_level(cur)
        c += 1


# --------------------------------------------------------------------------------------------
# C restatement (ctypes)
# --------------------------------------------------------------------------------------------
_LIB: Optional[ctypes.CDLL] = None


def build_c_oracle(force: bool = False) -> str:
    so = os.path.join(HERE, "libitd_oracle.so")
    src = os.path.join(HERE, "itd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s", "libitd_oracle.so"], check=True)
    return so


def c_lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        i64, ci = ctypes.c_int64, ctypes.c_int
        vp = ctypes.c_void_p
        lib.itd_oracle_find_knots_f64.restype = i64
        lib.itd_oracle_find_knots_f64.argtypes = [vp, i64, vp]
        lib.itd_oracle_find_knots_f32.restype = i64
        lib.itd_oracle_find_knots_f32.argtypes = [vp, i64, vp]
        for name in ("itd_oracle_extract_level_f64", "itd_oracle_extract_level_f32"):
            f = getattr(lib, name)
            f.restype = ci
            f.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        for name in ("itd_oracle_extract_with_knots_f64", "itd_oracle_extract_with_knots_f32"):
            f = getattr(lib, name)
            f.restype = ci
            f.argtypes = [vp, i64, vp, i64, vp, vp, vp]
        for name in ("itd_oracle_decompose_f64", "itd_oracle_decompose_f32"):
            f = getattr(lib, name)
            f.restype = ci
            f.argtypes = [vp, i64, ci, ci, vp, vp, vp, vp, vp, vp, vp]
        for name in ("itd_oracle_decompose_batch_f64", "itd_oracle_decompose_batch_f32"):
            f = getattr(lib, name)
            f.restype = ci
            f.argtypes = [vp, i64, i64, ci, ci, vp, vp, vp, vp, vp, ci]
        lib.itd_oracle_max_threads.restype = ci
        lib.itd_oracle_wpe3_f64.restype = ctypes.c_double
        lib.itd_oracle_wpe3_f64.argtypes = [vp, i64, ci]
        lib.itd_oracle_column_fsum_f64.restype = None
        lib.itd_oracle_column_fsum_f64.argtypes = [vp, i64, i64, vp]
        lib.itd_oracle_spline_level_f64.restype = ci
        lib.itd_oracle_spline_level_f64.argtypes = [vp, i64, vp, vp, vp]
        _LIB = lib
    return _LIB


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _suffix(x: np.ndarray) -> str:
    if x.dtype == np.float64:
        return "f64"
    if x.dtype == np.float32:
        return "f32"
    raise TypeError(f"oracle supports float64/float32, got {x.dtype}")


def c_find_knots(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    idx = np.empty(max(x.shape[0], 1), dtype=np.int64)
    k = getattr(c_lib(), "itd_oracle_find_knots_" + _suffix(x))(_ptr(x), x.shape[0], _ptr(idx))
    return idx[:k].copy()


def c_extract_level(x: np.ndarray):
    x = np.ascontiguousarray(x)
    n = x.shape[0]
    R = np.empty_like(x)
    B = np.empty_like(x)
    tau = np.empty(n + 2, dtype=np.int64)
    L = np.empty(n + 2, dtype=x.dtype)
    K = ctypes.c_int64(0)
    st = getattr(c_lib(), "itd_oracle_extract_level_" + _suffix(x))(
        _ptr(x), n, _ptr(R), _ptr(B), _ptr(tau), _ptr(L), ctypes.byref(K))
    if st != ITD_OK:
        raise OracleError(st)
    return R, B, tau[1:K.value + 1].copy()


def c_extract_with_knots(x: np.ndarray, knots: np.ndarray):
    x = np.ascontiguousarray(x)
    n = x.shape[0]
    knots = np.asarray(knots, dtype=np.int64)
    K = knots.shape[0]
    R = np.empty_like(x)
    B = np.empty_like(x)
    tau = np.empty(K + 2, dtype=np.int64)
    tau[1:K + 1] = knots
    L = np.empty(K + 2, dtype=x.dtype)
    st = getattr(c_lib(), "itd_oracle_extract_with_knots_" + _suffix(x))(
        _ptr(x), n, _ptr(tau), K, _ptr(R), _ptr(B), _ptr(L))
    if st != ITD_OK:
        raise OracleError(st)
    return R, B


def c_spline_level(x: np.ndarray):
    """C restatement of the spline-baseline level (see np_spline_level).  Returns ``(R, B, K)``."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    R = np.empty_like(x)
    B = np.empty_like(x)
    K = ctypes.c_int64(0)
    st = c_lib().itd_oracle_spline_level_f64(_ptr(x), x.shape[0], _ptr(R), _ptr(B), ctypes.byref(K))
    if st != ITD_OK:
        raise OracleError(st)
    return R, B, int(K.value)


def c_decompose(x: np.ndarray, max_iteration: int = 11, min_extrema: int = 2) -> OracleResult:
    x = np.ascontiguousarray(x)
    n = x.shape[0]
    rmax = max_iteration + 2
    rot = np.empty((rmax, n), dtype=x.dtype)
    bas = np.empty((rmax, n), dtype=x.dtype)
    counts = np.zeros(rmax, dtype=np.int32)
    n_rows, n_bas, kind = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    ik = ctypes.c_int64(0)
    st = getattr(c_lib(), "itd_oracle_decompose_" + _suffix(x))(
        _ptr(x), n, max_iteration, min_extrema, _ptr(rot), _ptr(bas), ctypes.byref(n_rows),
        ctypes.byref(n_bas), _ptr(counts), ctypes.byref(ik), ctypes.byref(kind))
    if st != ITD_OK:
        raise OracleError(st)
    return OracleResult(rot[:n_rows.value].copy(), bas[:n_bas.value].copy(),
                        counts[:n_rows.value].astype(np.int64), ik.value, kind.value)


def c_decompose_batch(x: np.ndarray, max_iteration: int = 11, min_extrema: int = 2,
                      want_baselines: bool = False, nthreads: int = 0):
    """(nsig, N) -> rotations (nsig, rmax, N), n_rows, knot_counts (nsig, rmax), status,
    [baselines].  One whole channel per pthread."""
    x = np.ascontiguousarray(x)
    nsig, n = x.shape
    rmax = max_iteration + 2
    rot = np.empty((nsig, rmax, n), dtype=x.dtype)
    bas = np.empty((nsig, rmax, n), dtype=x.dtype) if want_baselines else None
    n_rows = np.zeros(nsig, dtype=np.int32)
    counts = np.zeros((nsig, rmax), dtype=np.int32)
    status = np.zeros(nsig, dtype=np.int32)
    getattr(c_lib(), "itd_oracle_decompose_batch_" + _suffix(x))(
        _ptr(x), nsig, n, max_iteration, min_extrema, _ptr(rot), _ptr(bas), _ptr(n_rows),
        _ptr(counts), _ptr(status), nthreads)
    return rot, n_rows, counts, status, bas


def c_max_threads() -> int:
    return int(c_lib().itd_oracle_max_threads())
