"""pyitd_b200 -- B200-native Intrinsic Time-scale Decomposition behind PyITD's ``ITD`` interface.

``from pyitd_b200 import ITD`` is the drop-in for ``from ITD import ITD`` of the reference
(/root/reference/ITD.py); ``decompose`` is the batched entry point.  All compute goes through the
C ABI of ``libpyitd_b200.so`` (hand-written sm_100a kernels); there is no CPU fallback.
"""
from ._capi import PyITDLibraryError, Plan  # noqa: F401
from .itd import (ITD, ITDResult, clear_plan_cache, decompose, detect_peaks, extract_level,  # noqa: F401
                  extract_with_knots, find_knots, itd_baseline_extract)

from . import torch_ops  # noqa: F401  (torch.ops.pyitd.*: torch_ops.load())

from .spline import extract_spline, itd_baseline_extract_modified, itd_baseline_extract_spline  # noqa: F401

from .sift2d import (crossways_batch, crossways_itd_baseline_extract, retrieve_statistical_image_component,  # noqa: F401
                     totalextract2d)

from .analytics import (column_fsum, reconstruction_error, shewchuk, shewchuk_sum, weighted_permutation_entropy,  # noqa: F401
                        wpe_rows)

__all__ = ["wpe_rows", "weighted_permutation_entropy", "column_fsum", "shewchuk", "shewchuk_sum", "reconstruction_error",
           "crossways_batch", "crossways_itd_baseline_extract", "retrieve_statistical_image_component", "totalextract2d",
           "extract_spline", "itd_baseline_extract_spline", "itd_baseline_extract_modified",
           "ITD", "ITDResult", "decompose", "detect_peaks", "itd_baseline_extract", "extract_level",
           "find_knots", "extract_with_knots", "Plan", "PyITDLibraryError", "clear_plan_cache"]
__version__ = "0.1.0"
