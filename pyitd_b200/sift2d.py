"""Host-side mirror of the reference's 2-D "crossways" ensemble ITD.

Reference: siftED2D.ipynb code cell 1 (raw JSON :233-290): ``crossways_itd_baseline_extract(data)``,
``retrieve_statistical_image_component(data)`` and ``totalextract2d(data)``.  Same names, arguments and return
shapes; numpy in, numpy float64 out (CUDA tensors in, CUDA tensors out for the batched entry).  Everything
runs through ``pyitd_crossways_device`` / ``pyitd_ensemble2d_device`` of the C ABI; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _capi
from .itd import _cuda_device_index, _ptr, get_plan

__all__ = ["crossways_batch", "crossways_itd_baseline_extract", "retrieve_statistical_image_component",
           "totalextract2d", "mad"]

_MIN_KNOTS = 10      # siftED2D.ipynb cell 1: "if num_extrema < 10: return x"


def _plans(dev: int, n_images: int, H: int, W: int, code: int):
    return (get_plan(dev, n_images * H, W, code, 0, 2, 0), get_plan(dev, n_images * W, H, code, 0, 2, 0))


def _code_of(t: torch.Tensor) -> int:
    if t.dtype == torch.float64:
        return _capi.F64
    if t.dtype == torch.float32:
        return _capi.F32_MIXED
    raise TypeError(f"expected float64 or float32, got {t.dtype}")


def crossways_batch(images: torch.Tensor, min_knots: int = _MIN_KNOTS) -> torch.Tensor:
    """``crossways_itd_baseline_extract`` for a CUDA batch ``images[B, H, W]`` (or ``[H, W]``)."""
    if not (isinstance(images, torch.Tensor) and images.is_cuda):
        raise TypeError("crossways_batch expects a CUDA tensor")
    x = images if images.dim() == 3 else images.unsqueeze(0)
    x = x.contiguous()
    B, H, W = x.shape
    dev = x.device.index
    rp, cp = _plans(dev, B, H, W, _code_of(x))
    L = _capi.lib()
    scratch = torch.empty(int(L.pyitd_crossways_scratch_bytes(rp.handle, B, H, W)), dtype=torch.uint8, device=x.device)
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        _capi.check(L.pyitd_crossways_device(rp.handle, cp.handle, _ptr(x), _ptr(out), _ptr(scratch), B, H, W,
                                             int(min_knots), torch.cuda.current_stream(dev).cuda_stream),
                    "pyitd_crossways_device")
    return out if images.dim() == 3 else out[0]


def _image64(data) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(data), dtype=np.float64)
    if a.ndim != 2:
        raise TypeError("expected a 2-D array")
    if not np.isfinite(a).all():
        raise ValueError("NaN/Inf input is not supported")
    return a


def crossways_itd_baseline_extract(data) -> np.ndarray:
    """Drop-in for ``crossways_itd_baseline_extract(data)`` (siftED2D.ipynb cell 1) -> float64 ``[H, W]``."""
    a = _image64(data)
    dev = _cuda_device_index(None)
    return crossways_batch(torch.from_numpy(a).to(f"cuda:{dev}")).cpu().numpy()


def mad(arr) -> float:
    """Median absolute deviation as the notebook's ``mad`` computes it (numpy.median twice)."""
    a = np.asarray(arr, dtype=np.float64)
    med = np.median(a)
    return float(np.median(np.abs(a - med)))


def retrieve_statistical_image_component(data, noise=None, iterations: int = 20, seed: Optional[int] = None) -> np.ndarray:
    """Drop-in for ``retrieve_statistical_image_component(data)`` -> the low-pass image, float64 ``[H, W]``.

    The notebook draws ``iterations // 2`` noise fields ``normal(0, mad(data))`` from numba's unseeded generator;
    here they come from ``noise[iterations // 2, H, W]`` when given (reproducible), else from
    ``numpy.random.default_rng(seed)``."""
    a = _image64(data)
    draws = iterations // 2
    if noise is None:
        noise = np.random.default_rng(seed).normal(0.0, mad(a), (draws,) + a.shape)
    v = np.ascontiguousarray(np.asarray(noise), dtype=np.float64)
    if v.shape != (draws,) + a.shape:
        raise ValueError(f"noise must have shape {(draws,) + a.shape}, got {v.shape}")
    H, W = a.shape
    dev = _cuda_device_index(None)
    device = torch.device("cuda", dev)
    rp, cp = _plans(dev, 2 * draws, H, W, _capi.F64)
    L = _capi.lib()
    xt, vt = torch.from_numpy(a).to(device), torch.from_numpy(v).to(device)
    scratch = torch.empty(int(L.pyitd_ensemble2d_scratch_bytes(rp.handle, draws, H, W)), dtype=torch.uint8, device=device)
    low = torch.empty_like(xt)
    with torch.cuda.device(dev):
        _capi.check(L.pyitd_ensemble2d_device(rp.handle, cp.handle, _ptr(xt), _ptr(vt), _ptr(low), _ptr(scratch), draws, H, W,
                                              _MIN_KNOTS, torch.cuda.current_stream(dev).cuda_stream),
                    "pyitd_ensemble2d_device")
    return low.cpu().numpy()


def totalextract2d(data, noise=None, seed: Optional[int] = None) -> np.ndarray:
    """Drop-in for ``totalextract2d(data)`` -> ``asarray([data - lowpass, lowpass])`` (without the notebook's timing print)."""
    a = np.asarray(data).astype(dtype=np.float64)
    low = retrieve_statistical_image_component(a, noise=noise, seed=seed)
    return np.asarray([a - low, low])
