"""Host-side mirror of the reference's post-decomposition analytics.

Reference interfaces (paths relative to /root/reference):

* ``weighted_permutation_entropy(time_series, order=3, normalize=False)``   MEITD.py:79-128 -- evaluated per rotation by
  the MEITD / XITD drivers (MEITD.py:346, :374, :547);
* ``shewchuk(a)``                                                          helperfunctions.py:2-9 -- exactly rounded
  column sums (``math.fsum``) of the output rows, and ``shewchuk_sum`` of ITD.py:475-481, the reconstruction check
  of ITD.py:505-508.

The batched forms work on the decomposition's rows while they are still in device memory.  All compute goes
through ``pyitd_wpe_device`` / ``pyitd_column_fsum_device`` of the C ABI; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _capi
from .itd import ITDResult, _cuda_device_index, _ptr

__all__ = ["wpe_rows", "weighted_permutation_entropy", "column_fsum", "shewchuk", "shewchuk_sum",
           "reconstruction_error"]


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.float64:
        return _capi.F64
    if t.dtype == torch.float32:
        return _capi.F32
    raise TypeError(f"expected float64 or float32 rows, got {t.dtype}")


def wpe_rows(rows, n_rows: Optional[torch.Tensor] = None, order: int = 3, normalize: bool = True) -> torch.Tensor:
    """Weighted permutation entropy of every row of a CUDA tensor ``[R, N]`` or ``[S, rows, N]`` (or of an
    :class:`ITDResult`): float64 ``[R]`` / ``[S, rows]``; rows at or beyond ``n_rows[s]`` are NaN."""
    if isinstance(rows, ITDResult):
        rows, n_rows = rows.rotations, rows.n_rows
    if not (isinstance(rows, torch.Tensor) and rows.is_cuda):
        raise TypeError("wpe_rows expects a CUDA tensor or an ITDResult")
    t = rows.contiguous()
    shape = t.shape[:-1]
    N = t.shape[-1]
    R = int(np.prod(shape)) if len(shape) else 1
    per = t.shape[-2] if t.dim() == 3 else 1
    out = torch.empty(R, dtype=torch.float64, device=t.device)
    valid = None
    if n_rows is not None:
        if t.dim() != 3:
            raise ValueError("n_rows needs rows of shape [S, rows, N]")
        valid = n_rows.to(device=t.device, dtype=torch.int32).contiguous()
    dev = t.device.index
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().pyitd_wpe_device(_ptr(t), R, N, _code(t), int(order), int(bool(normalize)), _ptr(valid), per,
                                                 _ptr(out), torch.cuda.current_stream(dev).cuda_stream), "pyitd_wpe_device")
    return out.reshape(shape) if len(shape) else out[0]


def weighted_permutation_entropy(time_series, order: int = 3, normalize: bool = False) -> float:
    """Drop-in for ``weighted_permutation_entropy`` (MEITD.py:79-128); float64 arithmetic."""
    a = np.ascontiguousarray(np.array(time_series), dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("expected a 1-D time series")
    dev = _cuda_device_index(None)
    return float(wpe_rows(torch.from_numpy(a).to(f"cuda:{dev}").unsqueeze(0), order=order, normalize=normalize)[0])


def column_fsum(rows: torch.Tensor, n_rows: Optional[torch.Tensor] = None):
    """Exactly rounded column sums of CUDA rows ``[S, rows, N]`` (or ``[rows, N]``) over the first ``n_rows[s]`` rows:
    ``(sums[S, N], totals[S])`` float64 -- ``math.fsum`` per column, double-double total."""
    if not (isinstance(rows, torch.Tensor) and rows.is_cuda):
        raise TypeError("column_fsum expects a CUDA tensor")
    t = (rows if rows.dim() == 3 else rows.unsqueeze(0)).contiguous()
    S, R, N = t.shape
    sums = torch.empty((S, N), dtype=torch.float64, device=t.device)
    totals = torch.empty(S, dtype=torch.float64, device=t.device)
    valid = None if n_rows is None else n_rows.to(device=t.device, dtype=torch.int32).contiguous()
    dev = t.device.index
    with torch.cuda.device(dev):
        for s0 in range(0, S, 65535):
            s1 = min(S, s0 + 65535)
            _capi.check(_capi.lib().pyitd_column_fsum_device(
                _ptr(t[s0:s1]), s1 - s0, R, N, _code(t), _ptr(valid[s0:s1]) if valid is not None else None,
                _ptr(sums[s0:s1]), _ptr(totals[s0:s1]), torch.cuda.current_stream(dev).cuda_stream),
                "pyitd_column_fsum_device")
    return (sums, totals) if rows.dim() == 3 else (sums[0], totals[0])


def shewchuk(a, axis: int = 0) -> np.ndarray:
    """Drop-in for ``shewchuk(a)`` (helperfunctions.py:2-9): ``s[i] = math.fsum(a[:, i])``."""
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    dev = _cuda_device_index(None)
    sums, _ = column_fsum(torch.from_numpy(arr).to(f"cuda:{dev}"))
    return sums.cpu().numpy()


def shewchuk_sum(a, axis: int = 0) -> float:
    """Drop-in for ``shewchuk_sum(a)`` (ITD.py:475-481): the sum of the exactly rounded column sums."""
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    dev = _cuda_device_index(None)
    _, total = column_fsum(torch.from_numpy(arr).to(f"cuda:{dev}"))
    return float(total)


def reconstruction_error(x: torch.Tensor, result: ITDResult) -> torch.Tensor:
    """The reference's reconstruction check (ITD.py:505-508) for a whole batch on the device:
    ``abs(sum(x[s]) - shewchuk_sum(rows of s))`` float64 ``[S]`` (both sums in extended precision)."""
    _, totals = column_fsum(result.rotations, result.n_rows)
    xt = x if x.dim() == 2 else x.unsqueeze(0)
    _, xsum = column_fsum(xt.unsqueeze(1).contiguous())
    return (xsum - totals).abs()
