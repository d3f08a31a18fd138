"""``torch.ops.pyitd.*``: the thin PyTorch extension over the C ABI.

``pyitd_b200/_torch_ops.so`` (``csrc/torch_ops.cpp``, host code only) registers

* ``pyitd::find_knots(x, kinds=3, capacity=0) -> (knots, count, status)``            detect_peaks, ITD.py:33-76
* ``pyitd::extract_level(x, precision="auto") -> (rotation, baseline, count, status)``  itd_baseline_extract, ITD.py:79
* ``pyitd::decompose(x, max_iteration=11, min_extrema=2, return_baselines=False, zero_tail=False,
  precision="auto") -> (rotations, n_rows, knot_counts, baselines, status, stop_kind, input_knots)``   ITD.itd, ITD.py:351

for the CUDA dispatch key (and Meta for shape inference).  Every op launches on torch's current stream of the input's
device, allocates with torch's caching allocator and never synchronises.  CPU tensors are rejected by the dispatcher:
there is no CPU fallback.
"""
from __future__ import annotations

import os
import subprocess

import torch

from . import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_torch_ops.so")
_loaded = False


def build(verbose: bool = False) -> str:
    """g++ the extension in-tree against this interpreter's torch (needs libpyitd_b200.so, built first)."""
    _capi.build(verbose=verbose)
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "torch_ops", f"TORCH_DIR={os.path.dirname(torch.__file__)}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise _capi.PyITDLibraryError("building _torch_ops.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def load():
    """Register the ops (idempotent) and return ``torch.ops.pyitd``."""
    global _loaded
    if not _loaded:
        if not os.path.exists(LIB_PATH):
            raise _capi.PyITDLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        torch.ops.load_library(LIB_PATH)
        _loaded = True
    return torch.ops.pyitd


OP_NAMES = ("find_knots", "extract_level", "decompose", "clear_plans", "abi_version")
