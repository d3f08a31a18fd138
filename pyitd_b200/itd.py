"""Host-side mirror of the reference's interface for the ITD sifting loop.

Same names, arguments and error behaviour as ``/root/reference/ITD.py``:

* ``ITD(extrema_detection="matlab")``            ITD.py:157
* ``ITD.itd(data, max_iteration=11)``            ITD.py:351   -> ``ndarray (rows, N)`` float64
* ``ITD.__call__(S, max_iterations=12)``         ITD.py:189   (the reference forwards a wrong keyword
                                                               and raises TypeError; fixed here)
* ``ITD.get_baselines()`` / ``get_rotations()``  ITD.py:436 / ITD.py:449
* ``itd_baseline_extract(data)``                 ITD.py:79    -> ``(rotation, baseline)``
* ``detect_peaks(x)``                            ITD.py:33    -> ``int64[:]``

plus the batched entry ``decompose(x[S, N])`` the B200 build adds.  Everything runs through the C
ABI in ``libpyitd_b200.so`` (``_capi.py``); torch is used for device memory and streams only.
"""
from __future__ import annotations

import collections
from dataclasses import dataclass
from typing import Optional, Union

import numpy as np
import torch

from . import _capi

ArrayLike = Union[np.ndarray, torch.Tensor]

_DTYPES = {"f64": _capi.F64, "f32_mixed": _capi.F32_MIXED, "f32": _capi.F32}


# ---------------------------------------------------------------------------------------------
# plan cache (a plan owns GBs of workspace for big batches: reuse it across calls)
# ---------------------------------------------------------------------------------------------
_PLAN_CACHE: "collections.OrderedDict[tuple, _capi.Plan]" = collections.OrderedDict()
_PLAN_CACHE_MAX = 4


def get_plan(device: int, n_signals: int, n_samples: int, dtype: int, max_iteration: int,
             min_extrema: int, options: int) -> _capi.Plan:
    key = (device, n_signals, n_samples, dtype, max_iteration, min_extrema, options)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        while len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _, old = _PLAN_CACHE.popitem(last=False)
            old.close()
        plan = _capi.Plan(device, n_signals, n_samples, dtype, max_iteration, min_extrema, options)
        _PLAN_CACHE[key] = plan
    else:
        _PLAN_CACHE.move_to_end(key)
    return plan


def clear_plan_cache() -> None:
    while _PLAN_CACHE:
        _, p = _PLAN_CACHE.popitem()
        p.close()


def _cuda_device_index(device) -> int:
    if not torch.cuda.is_available():
        raise _capi.PyITDLibraryError(
            "pyitd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    d = torch.device(device)
    if d.type != "cuda":
        raise ValueError(f"device must be a CUDA device, got {d}")
    return torch.cuda.current_device() if d.index is None else d.index


def _resolve_dtype(x_dtype, dtype: Optional[str]) -> tuple[int, torch.dtype]:
    """-> (library dtype code, torch dtype of the I/O buffers)."""
    if dtype is None:
        dtype = "f64" if x_dtype in (torch.float64, np.float64) else (
            "f32_mixed" if x_dtype in (torch.float32, np.float32) else "f64")
    if dtype not in _DTYPES:
        raise ValueError(f"dtype must be one of {sorted(_DTYPES)}, got {dtype!r}")
    return _DTYPES[dtype], (torch.float64 if dtype == "f64" else torch.float32)


# ---------------------------------------------------------------------------------------------
# batched result
# ---------------------------------------------------------------------------------------------
@dataclass
class ITDResult:
    """Output of :func:`decompose` for ``S`` signals of ``N`` samples (``rows = max_iteration + 2``).

    ``rotations[s, :n_rows[s]]`` is exactly what ``ITD().itd(x[s])`` returns in the reference:
    proper rotations followed by the trend row.  Rows at and beyond ``n_rows[s]`` are unspecified
    unless ``zero_tail=True`` was requested."""

    rotations: torch.Tensor               # [S, rows, N]
    n_rows: torch.Tensor                  # [S] int32
    knot_counts: torch.Tensor             # [S, rows] int32: extrema of each new baseline (ITD.py:403)
    input_knots: torch.Tensor             # [S] int32
    stop_kind: torch.Tensor               # [S] int32: 1 knot stop, 2 iteration stop
    status: torch.Tensor                  # [S] int32: 0 ok, bit0 zero delta-X, bit1 non-finite input
    baselines: Optional[torch.Tensor] = None   # [S, rows, N] when requested

    def rows_of(self, s: int) -> torch.Tensor:
        return self.rotations[s, : int(self.n_rows[s])]

    def baselines_of(self, s: int) -> torch.Tensor:
        if self.baselines is None:
            raise ValueError("baselines were not requested (return_baselines=True)")
        nr, kind = int(self.n_rows[s]), int(self.stop_kind[s])
        return self.baselines[s, : (nr if kind == _capi.STOP_ITER else max(nr - 1, 0))]

    def raise_for_status(self) -> None:
        st = self.status
        if bool((st != 0).any()):
            bad = torch.nonzero(st != 0).flatten()
            first = int(bad[0])
            code = int(st[first])
            where = f"signal {first}" + (f" (+{bad.numel() - 1} more)" if bad.numel() > 1 else "")
            if code & _capi.ST_NONFINITE:
                raise ValueError(f"{where}: input contains NaN or Inf (unsupported, ITD.py:46-51)")
            if code & _capi.ST_ZERO_DX:
                # numba raises exactly this at ITD.py:116
                raise ZeroDivisionError(f"{where}: division by zero (equal signal values at adjacent knots)")
            raise RuntimeError(f"{where}: status {code}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def decompose(x: ArrayLike, max_iteration: int = 11, min_extrema: int = 2, dtype: Optional[str] = None,
              return_baselines: bool = False, zero_tail: bool = False, strict: bool = False,
              device=None) -> ITDResult:
    """Decompose a batch of independent signals ``x[S, N]`` (or one signal ``x[N]``).

    * CUDA tensor in  -> everything stays on that device, launched on the current stream, no host
      synchronisation (unless ``strict=True``, which has to read the status words);
    * numpy array / CPU tensor in -> ``pyitd_decompose_host`` (copies in, runs, copies out) and CPU
      tensors come back.

    ``dtype``: ``'f64'`` (reference arithmetic, bit-exact), ``'f32_mixed'`` (float32 in/out around a
    float64 carry: equals ``float32(reference(float64(x)))``), ``'f32'`` (pure float32).  Default:
    by input dtype (float64 -> f64, float32 -> f32_mixed).
    """
    if isinstance(x, np.ndarray):
        xt = torch.from_numpy(np.ascontiguousarray(x))
    elif isinstance(x, torch.Tensor):
        xt = x
    else:
        xt = torch.as_tensor(np.asarray(x, dtype=np.float64))
    if xt.dim() == 1:
        xt = xt.unsqueeze(0)
    if xt.dim() != 2:
        raise ValueError(f"expected a 1-D signal or a 2-D [signals, samples] batch, got shape {tuple(xt.shape)}")
    code, io_dtype = _resolve_dtype(xt.dtype, dtype)
    if xt.dtype != io_dtype:
        xt = xt.to(io_dtype)
    xt = xt.contiguous()
    S, N = xt.shape
    if N < 3:
        raise ValueError("signal shorter than 3 samples (undefined in the reference, ITD.py:42-43)")
    options = (_capi.OPT_BASELINES if return_baselines else 0) | (_capi.OPT_ZERO_TAIL if zero_tail else 0)
    on_device = xt.is_cuda
    dev = xt.device.index if on_device else _cuda_device_index(device)
    plan = get_plan(dev, S, N, code, max_iteration, min_extrema, options)
    rows = plan.rows
    out_dev = xt.device if on_device else torch.device("cpu")
    pin = not on_device
    mk = dict(device=out_dev, pin_memory=pin) if pin else dict(device=out_dev)
    rot = torch.empty((S, rows, N), dtype=io_dtype, **mk)
    bas = torch.empty((S, rows, N), dtype=io_dtype, **mk) if return_baselines else None
    n_rows = torch.empty(S, dtype=torch.int32, **mk)
    counts = torch.empty((S, rows), dtype=torch.int32, **mk)
    iknots = torch.empty(S, dtype=torch.int32, **mk)
    kind = torch.empty(S, dtype=torch.int32, **mk)
    status = torch.empty(S, dtype=torch.int32, **mk)
    if on_device:
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            plan.decompose_device(_ptr(xt), _ptr(rot), _ptr(bas), _ptr(n_rows), _ptr(counts), _ptr(iknots),
                                  _ptr(kind), _ptr(status), stream)
    else:
        plan.decompose_host(_ptr(xt), _ptr(rot), _ptr(bas), _ptr(n_rows), _ptr(counts), _ptr(iknots),
                            _ptr(kind), _ptr(status))
    res = ITDResult(rot, n_rows, counts, iknots, kind, status, bas)
    if strict:
        res.raise_for_status()
    return res


def extract_level(x: torch.Tensor, dtype: Optional[str] = None):
    """One sifting level for a CUDA batch ``x[S, N]`` -> ``(rotation, baseline, knot_count, status)``."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise TypeError("extract_level expects a CUDA tensor; use itd_baseline_extract for numpy input")
    xt = x if x.dim() == 2 else x.unsqueeze(0)
    code, io_dtype = _resolve_dtype(xt.dtype, dtype)
    xt = xt.to(io_dtype).contiguous()
    S, N = xt.shape
    dev = xt.device.index
    plan = get_plan(dev, S, N, code, 0, 2, 0)
    R = torch.empty_like(xt)
    B = torch.empty_like(xt)
    cnt = torch.empty(S, dtype=torch.int32, device=xt.device)
    st = torch.empty(S, dtype=torch.int32, device=xt.device)
    with torch.cuda.device(dev):
        plan.extract_level_device(_ptr(xt), _ptr(R), _ptr(B), _ptr(cnt), _ptr(st),
                                  torch.cuda.current_stream(dev).cuda_stream)
    return R, B, cnt, st


def find_knots(x: torch.Tensor, kinds: int = _capi.KNOTS_BOTH, dtype: Optional[str] = None,
               capacity: Optional[int] = None):
    """Knot indices of a CUDA batch ``x[S, N]`` -> ``(knots[S, capacity] int32, count[S], status[S])``."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise TypeError("find_knots expects a CUDA tensor; use detect_peaks for numpy input")
    xt = x if x.dim() == 2 else x.unsqueeze(0)
    code, io_dtype = _resolve_dtype(xt.dtype, dtype)
    xt = xt.to(io_dtype).contiguous()
    S, N = xt.shape
    cap = N if capacity is None else int(capacity)
    dev = xt.device.index
    plan = get_plan(dev, S, N, code, 0, 2, 0)
    knots = torch.empty((S, cap), dtype=torch.int32, device=xt.device)
    cnt = torch.empty(S, dtype=torch.int32, device=xt.device)
    st = torch.empty(S, dtype=torch.int32, device=xt.device)
    with torch.cuda.device(dev):
        plan.find_knots_device(_ptr(xt), kinds, _ptr(knots), cap, _ptr(cnt), _ptr(st),
                               torch.cuda.current_stream(dev).cuda_stream)
    return knots, cnt, st


def extract_with_knots(x: torch.Tensor, knots: torch.Tensor, knot_count: Optional[torch.Tensor] = None,
                       dtype: Optional[str] = None):
    """One sifting level of a CUDA batch ``x[S, N]`` with SUPPLIED knots: detect the
    knots once (:func:`find_knots` on a reference channel), then apply the knot baseline and the interpolation of
    ITD.py:95-119 to other channels or to updated data -- the "retain and reuse the extrema ... along multiple
    channels" mode of the reference's C++ port (itd.cpp:41-44, ``compute_extrema=false`` at itd.cpp:156-169).

    ``knots``: int32 ``[S, cap]`` (one ascending list per signal) or ``[cap]`` / ``[1, cap]`` (one list shared by
    every signal); ``knot_count``: valid entries per list (default: all ``cap``).
    Returns ``(rotation, baseline, status)``."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise TypeError("extract_with_knots expects a CUDA tensor")
    xt = x if x.dim() == 2 else x.unsqueeze(0)
    code, io_dtype = _resolve_dtype(xt.dtype, dtype)
    xt = xt.to(io_dtype).contiguous()
    S, N = xt.shape
    kt = knots if knots.dim() == 2 else knots.unsqueeze(0)
    kt = kt.to(device=xt.device, dtype=torch.int32).contiguous()
    rows, cap = kt.shape
    if rows not in (1, S):
        raise ValueError(f"knots must have 1 or {S} rows, got {rows}")
    if knot_count is None:
        kc = torch.full((rows,), cap, dtype=torch.int32, device=xt.device)
    else:
        kc = knot_count.to(device=xt.device, dtype=torch.int32).reshape(rows).contiguous()
    dev = xt.device.index
    plan = get_plan(dev, S, N, code, 0, 2, 0)
    R = torch.empty_like(xt)
    B = torch.empty_like(xt)
    st = torch.empty(S, dtype=torch.int32, device=xt.device)
    with torch.cuda.device(dev):
        plan.extract_with_knots_device(_ptr(xt), _ptr(kt), cap, _ptr(kc), rows, _ptr(R), _ptr(B), _ptr(st),
                                       torch.cuda.current_stream(dev).cuda_stream)
    return R, B, st


# ---------------------------------------------------------------------------------------------
# the reference's two module-level functions
# ---------------------------------------------------------------------------------------------
def _require_float64(x, name: str) -> np.ndarray:
    a = np.asarray(x)
    if isinstance(x, np.ndarray) and x.dtype != np.float64:
        # the reference's eager numba signatures accept float64[:] only (ITD.py:33, ITD.py:79)
        raise TypeError(f"{name}: no matching definition for argument type(s) array({x.dtype}, 1d)")
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 1:
        raise TypeError(f"{name}: expected a 1-D array")
    return a


def detect_peaks(x) -> np.ndarray:
    """Drop-in for ``detect_peaks`` (ITD.py:33-76): indices ``i`` with ``x[i-1] >= x[i] < x[i+1]``
    (despite its name the reference function returns valleys; call it on ``-x`` for peaks)."""
    a = _require_float64(x, "detect_peaks")
    if a.shape[0] < 3:
        return np.empty(0, dtype=np.int64)
    dev = _cuda_device_index(None)
    knots, cnt, st = find_knots(torch.from_numpy(a).to(f"cuda:{dev}"), kinds=_capi.KNOTS_VALLEYS)
    if int(st[0]) & _capi.ST_NONFINITE:
        raise ValueError("detect_peaks: NaN/Inf input is not supported (reference NaN path ITD.py:46-51)")
    return knots[0, : int(cnt[0])].to(torch.int64).cpu().numpy()


def itd_baseline_extract(data):
    """Drop-in for ``itd_baseline_extract`` (ITD.py:79-121) -> ``(rotation, baseline)`` float64."""
    a = _require_float64(data, "itd_baseline_extract")
    if a.shape[0] < 3:
        raise ValueError("signal shorter than 3 samples (undefined in the reference, ITD.py:42-43)")
    dev = _cuda_device_index(None)
    R, B, _, st = extract_level(torch.from_numpy(a).to(f"cuda:{dev}"), dtype="f64")
    code = int(st[0])
    if code & _capi.ST_NONFINITE:
        raise ValueError("itd_baseline_extract: NaN/Inf input is not supported")
    if code & _capi.ST_ZERO_DX:
        raise ZeroDivisionError("division by zero")
    return R[0].cpu().numpy(), B[0].cpu().numpy()


# ---------------------------------------------------------------------------------------------
# the reference's class
# ---------------------------------------------------------------------------------------------
class ITD:
    """Intrinsic Time-scale Decomposition (Frei & Osorio 2007), B200 implementation behind the
    reference's ``ITD`` interface (ITD.py:123-465).

    >>> itd = ITD()
    >>> rows = itd.itd(signal)            # (n_rotations + 1, N): rotations, then the trend
    >>> baselines = itd.get_baselines()
    """

    def __init__(self, extrema_detection: str = "matlab", device=None, dtype: Optional[str] = None,
                 min_extrema: int = 2, verbose: bool = False):
        self.extrema_detection = extrema_detection
        assert self.extrema_detection in (
            "simple", "parabol", "matlab",
        ), "Only 'simple', 'matlab', and 'parabol' values supported"     # ITD.py:177-181
        self.DTYPE = np.float64
        self.device = device
        self.dtype = dtype
        self.min_extrema = min_extrema
        self.verbose = verbose
        self.rotations = None
        self.baselines = None
        self.knot_counts = None
        self.result: Optional[ITDResult] = None

    def __call__(self, S, max_iterations: int = 12):
        return self.itd(S, max_iteration=max_iterations)

    def itd(self, data: ArrayLike, max_iteration: int = 11):
        """1-D input: the reference's return value, ``ndarray (rows, N)`` (float64 unless the
        instance was built with an fp32 ``dtype``).  2-D input ``[S, N]``: an :class:`ITDResult`."""
        batched = getattr(data, "ndim", 1) == 2
        if not batched:
            arr = data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
            self.DTYPE = arr.dtype                                       # ITD.py:368
            if self.dtype in (None, "f64"):
                arr = np.asarray(arr, dtype=np.float64)                  # ITD.py:389 casts to float64
            else:
                arr = np.asarray(arr, dtype=np.float32)
            data = arr
        res = decompose(data, max_iteration=max_iteration, min_extrema=self.min_extrema,
                        dtype=self.dtype, return_baselines=True, device=self.device)
        self.result = res
        res.raise_for_status()
        if batched:
            self.rotations, self.baselines = res.rotations, res.baselines
            self.knot_counts = res.knot_counts
            return res
        nr = int(res.n_rows[0])
        self.knot_counts = res.knot_counts[0, :nr].cpu().numpy()
        if self.verbose:                                                 # ITD.py:403, :406, :419
            for c in self.knot_counts:
                print(int(c))
            print("No more decompositions possible" if int(res.stop_kind[0]) == _capi.STOP_KNOTS
                  else "Out of time!")
        self.rotations = res.rows_of(0).cpu().numpy()
        self.baselines = res.baselines_of(0).cpu().numpy()
        return self.rotations

    def get_baselines(self):
        if self.baselines is None:
            raise ValueError("No baselines found. Please, run ITD method or its variant first.")
        return self.baselines

    def get_rotations(self):
        if self.rotations is None:
            raise ValueError("No IPR found. Please, run ITD method or its variant first.")
        return self.rotations
