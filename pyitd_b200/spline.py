"""Host-side mirror of the reference's cubic-spline baseline variant.

Reference interfaces (paths relative to /root/reference):

* ``itd_baseline_extract(data) -> (rotation, baseline)``          MEITD.py:303-338
* ``itd_baseline_extract_modified(x) -> baseline``                numba_accelerated_itd.py:183-211

Both detect the knots with the ITD.py stencil, form the Frei-Osorio knot baseline, take the end knots from
the odd-reflected pad and pass a cubic spline (scipy ``splrep(k=3)`` / ``splev``: interpolating, not-a-knot
ends) through ``(tau_k, L_k)``.  Here the whole level is three kernel launches behind
``pyitd_extract_spline_device`` (C ABI); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _capi
from .itd import _cuda_device_index, _ptr, _require_float64, _resolve_dtype, get_plan

__all__ = ["extract_spline", "itd_baseline_extract_spline", "itd_baseline_extract_modified"]


def extract_spline(x: torch.Tensor, min_knots: int = 2, dtype: Optional[str] = None,
                   want_rotation: bool = True):
    """One spline-baseline level for a CUDA batch ``x[S, N]`` -> ``(rotation, baseline, knot_count, status)``.

    ``min_knots``: signals with fewer interior knots get ``baseline = x`` and ``rotation = 0`` (the rule of
    numba_accelerated_itd.py:188-191 with 10); ``status`` carries ``ST_FEW_KNOTS`` where the reference's
    ``splrep`` would raise (fewer than 2 interior knots).  float64, or float32 in/out around float64
    arithmetic (``dtype='f32_mixed'``, the default for float32 input)."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise TypeError("extract_spline expects a CUDA tensor; use itd_baseline_extract_spline for numpy input")
    xt = x if x.dim() == 2 else x.unsqueeze(0)
    code, io_dtype = _resolve_dtype(xt.dtype, dtype)
    if code == _capi.F32:
        raise ValueError("the spline variant computes in float64: dtype must be 'f64' or 'f32_mixed'")
    xt = xt.to(io_dtype).contiguous()
    S, N = xt.shape
    dev = xt.device.index
    plan = get_plan(dev, S, N, code, 0, 2, 0)
    R = torch.empty_like(xt) if want_rotation else None
    B = torch.empty_like(xt)
    cnt = torch.empty(S, dtype=torch.int32, device=xt.device)
    st = torch.empty(S, dtype=torch.int32, device=xt.device)
    with torch.cuda.device(dev):
        plan.extract_spline_device(_ptr(xt), _ptr(R), _ptr(B), _ptr(cnt), _ptr(st), min_knots,
                                   torch.cuda.current_stream(dev).cuda_stream)
    return R, B, cnt, st


def _run_one(a: np.ndarray, min_knots: int, want_rotation: bool):
    if a.shape[0] < 3:
        raise ValueError("signal shorter than 3 samples")
    dev = _cuda_device_index(None)
    R, B, cnt, st = extract_spline(torch.from_numpy(a).to(f"cuda:{dev}"), min_knots=min_knots, dtype="f64",
                                   want_rotation=want_rotation)
    code = int(st[0])
    if code & _capi.ST_NONFINITE:
        raise ValueError("NaN/Inf input is not supported")
    return R, B, int(cnt[0]), code


def itd_baseline_extract_spline(data):
    """Drop-in for MEITD.py's ``itd_baseline_extract(data)`` (MEITD.py:303-338) -> ``(rotation, baseline)``
    float64.  Raises ``TypeError`` where scipy's ``splrep`` does (fewer than 4 spline points)."""
    a = np.ascontiguousarray(np.asarray(data, dtype=np.float64))      # MEITD.py:305 casts its input
    if a.ndim != 1:
        raise TypeError("itd_baseline_extract: expected a 1-D array")
    R, B, _, code = _run_one(a, 2, True)
    if code & _capi.ST_FEW_KNOTS:
        raise TypeError("m > k must hold")
    return R[0].cpu().numpy(), B[0].cpu().numpy()


def itd_baseline_extract_modified(x):
    """Drop-in for ``itd_baseline_extract_modified(x)`` (numba_accelerated_itd.py:183-211) -> baseline float64;
    with fewer than 10 extrema the input itself is returned (numba_accelerated_itd.py:188-191)."""
    a = _require_float64(x, "itd_baseline_extract_modified")
    _, B, _, _ = _run_one(a, 10, False)
    return B[0].cpu().numpy()
