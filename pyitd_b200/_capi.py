"""ctypes binding of ``libpyitd_b200.so`` (the C ABI declared in ``include/pyitd_b200.h``).

This is the only way the Python layer reaches the kernels.  There is no CPU fallback: if the
library is missing or there is no B200-class device the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpyitd_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "pyitd_b200.h")

F64, F32_MIXED, F32 = 0, 1, 2
ST_ZERO_DX, ST_NONFINITE, ST_TOO_SHORT, ST_BAD_KNOTS, ST_FEW_KNOTS = 1, 2, 4, 8, 16
STOP_KNOTS, STOP_ITER = 1, 2
OPT_BASELINES, OPT_ZERO_TAIL = 1, 2
KNOTS_VALLEYS, KNOTS_PEAKS, KNOTS_BOTH = 1, 2, 3
E_INVALID, E_CUDA, E_NOMEM, E_NODEVICE = -1, -2, -3, -4


class PyITDLibraryError(RuntimeError):
    """libpyitd_b200.so is missing, failed to load, or a call into it failed."""


_lib: Optional[ctypes.CDLL] = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree with nvcc for sm_100a (no GPU needed to build)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise PyITDLibraryError("building libpyitd_b200.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PyITDLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pyitd_b200/csrc`.  pyitd_b200 has no CPU fallback.")
    try:
        L = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the box
        raise PyITDLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.pyitd_abi_version.restype = ci
    L.pyitd_abi_version.argtypes = []
    L.pyitd_last_error.restype = ctypes.c_char_p
    L.pyitd_last_error.argtypes = []
    L.pyitd_device_count.restype = ci
    L.pyitd_device_count.argtypes = []
    L.pyitd_plan_create.restype = ci
    L.pyitd_plan_create.argtypes = [ctypes.POINTER(vp), ci, i64, i64, ci, ci, ci, ci]
    L.pyitd_plan_destroy.restype = None
    L.pyitd_plan_destroy.argtypes = [vp]
    L.pyitd_plan_rows.restype = ci
    L.pyitd_plan_rows.argtypes = [vp]
    L.pyitd_plan_workspace_bytes.restype = i64
    L.pyitd_plan_workspace_bytes.argtypes = [vp]
    L.pyitd_plan_launches.restype = ci
    L.pyitd_plan_launches.argtypes = [vp]
    if hasattr(L, "pyitd_has_feature"):
        L.pyitd_has_feature.restype = ci
        L.pyitd_has_feature.argtypes = [ctypes.c_char_p]
    if hasattr(L, "pyitd_plan_sweep_stats"):           # (absent from older builds used in A/B measurements)
        L.pyitd_plan_sweep_stats.restype = ci
        L.pyitd_plan_sweep_stats.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.pyitd_plan_path.restype = ci
    L.pyitd_plan_path.argtypes = [vp, ctypes.POINTER(ci)]
    L.pyitd_plan_set_groups.restype = ci
    L.pyitd_plan_set_groups.argtypes = [vp, ci]
    L.pyitd_plan_groups.restype = ci
    L.pyitd_plan_groups.argtypes = [vp]
    L.pyitd_probe_mixed_traffic.restype = ci
    L.pyitd_probe_mixed_traffic.argtypes = [vp, vp, vp, i64, ci, i64, vp]
    L.pyitd_plan_enable_timing.restype = ci
    L.pyitd_plan_enable_timing.argtypes = [vp, ci]
    L.pyitd_plan_launch_times.restype = ci
    L.pyitd_plan_launch_times.argtypes = [vp, vp, ci]
    L.pyitd_decompose_device.restype = ci
    L.pyitd_decompose_device.argtypes = [vp] * 10
    L.pyitd_decompose_host.restype = ci
    L.pyitd_decompose_host.argtypes = [vp] * 9
    L.pyitd_extract_level_device.restype = ci
    L.pyitd_extract_level_device.argtypes = [vp] * 7
    L.pyitd_extract_with_knots_device.restype = ci
    L.pyitd_extract_with_knots_device.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp, vp, vp]
    L.pyitd_extract_spline_device.restype = ci
    L.pyitd_extract_spline_device.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp]
    L.pyitd_crossways_scratch_bytes.restype = i64
    L.pyitd_crossways_scratch_bytes.argtypes = [vp, i64, i64, i64]
    L.pyitd_crossways_device.restype = ci
    L.pyitd_crossways_device.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, ci, vp]
    L.pyitd_ensemble2d_scratch_bytes.restype = i64
    L.pyitd_ensemble2d_scratch_bytes.argtypes = [vp, i64, i64, i64]
    L.pyitd_ensemble2d_device.restype = ci
    L.pyitd_ensemble2d_device.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, ci, vp]
    L.pyitd_wpe_device.restype = ci
    L.pyitd_wpe_device.argtypes = [vp, i64, i64, ci, ci, ci, vp, i64, vp, vp]
    L.pyitd_column_fsum_device.restype = ci
    L.pyitd_column_fsum_device.argtypes = [vp, i64, i64, i64, ci, vp, vp, vp, vp]
    L.pyitd_find_knots_device.restype = ci
    L.pyitd_find_knots_device.argtypes = [vp, vp, ci, vp, i64, vp, vp, vp]
    _lib = L
    return L


def has_feature(name: str) -> bool:
    """Optional code paths of the loaded build (see pyitd_has_feature in the header)."""
    L = lib()
    return bool(hasattr(L, "pyitd_has_feature") and L.pyitd_has_feature(name.encode()))


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = lib().pyitd_last_error().decode("utf-8", "replace")
    if rc == E_NOMEM:
        raise MemoryError(f"{what}: {msg}")
    if rc == E_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise PyITDLibraryError(f"{what} failed ({rc}): {msg}")


class Plan:
    """Owns one ``pyitd_plan`` (device workspace for a fixed batch shape)."""

    def __init__(self, device: int, n_signals: int, n_samples: int, dtype: int = F64,
                 max_iteration: int = 11, min_extrema: int = 2, options: int = 0):
        self._h = ctypes.c_void_p()
        self._L = lib()
        check(self._L.pyitd_plan_create(ctypes.byref(self._h), device, n_signals, n_samples, dtype,
                                        max_iteration, min_extrema, options), "pyitd_plan_create")
        self.device, self.n_signals, self.n_samples = device, n_signals, n_samples
        self.dtype, self.max_iteration, self.min_extrema, self.options = dtype, max_iteration, min_extrema, options
        self.rows = self._L.pyitd_plan_rows(self._h)

    @property
    def handle(self):
        if not self._h:
            raise PyITDLibraryError("plan already destroyed")
        return self._h

    @property
    def workspace_bytes(self) -> int:
        return int(self._L.pyitd_plan_workspace_bytes(self.handle))

    @property
    def launches(self) -> int:
        return int(self._L.pyitd_plan_launches(self.handle))

    def sweep_stats(self) -> tuple[int, int, int]:
        """(pairs of extractions fused into one item, pairs not tried after the knot prediction, fused passes whose check of
        the prediction failed) of the last call."""
        a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        check(self._L.pyitd_plan_sweep_stats(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)),
              "pyitd_plan_sweep_stats")
        return int(a.value), int(b.value), int(c.value)

    @property
    def path(self) -> tuple[str, int]:
        """('resident' | 'stream' | 'lookback' | 'strided' | 'sweep' | 'coop', CTAs per cluster)."""
        cl = ctypes.c_int(1)
        code = int(self._L.pyitd_plan_path(self.handle, ctypes.byref(cl)))
        return {0: "lookback", 1: "stream", 2: "resident", 3: "strided", 4: "sweep", 5: "coop"}[code], int(cl.value)

    @property
    def groups(self) -> int:
        return int(self._L.pyitd_plan_groups(self.handle))

    @groups.setter
    def groups(self, g: int) -> None:
        check(self._L.pyitd_plan_set_groups(self.handle, int(g)), "pyitd_plan_set_groups")

    def enable_timing(self, on: bool = True) -> None:
        check(self._L.pyitd_plan_enable_timing(self.handle, int(on)), "pyitd_plan_enable_timing")

    def launch_times_ms(self) -> list[float]:
        buf = (ctypes.c_float * (self.rows + 8))()
        n = self._L.pyitd_plan_launch_times(self.handle, buf, len(buf))
        if n < 0:
            check(n, "pyitd_plan_launch_times")
        return [float(buf[i]) for i in range(n)]

    def close(self) -> None:
        if self._h:
            self._L.pyitd_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def decompose_device(self, x, rotations, baselines, n_rows, knot_counts, input_knots, stop_kind,
                         status, stream) -> None:
        check(self._L.pyitd_decompose_device(self.handle, x, rotations, baselines, n_rows, knot_counts,
                                             input_knots, stop_kind, status, stream),
              "pyitd_decompose_device")

    def decompose_host(self, x, rotations, baselines, n_rows, knot_counts, input_knots, stop_kind,
                       status) -> None:
        check(self._L.pyitd_decompose_host(self.handle, x, rotations, baselines, n_rows, knot_counts,
                                           input_knots, stop_kind, status), "pyitd_decompose_host")

    def extract_level_device(self, x, rotation, baseline, knot_count, status, stream) -> None:
        check(self._L.pyitd_extract_level_device(self.handle, x, rotation, baseline, knot_count, status,
                                                 stream), "pyitd_extract_level_device")

    def extract_with_knots_device(self, x, knots, capacity, knot_count, n_knot_rows, rotation, baseline, status,
                                  stream) -> None:
        check(self._L.pyitd_extract_with_knots_device(self.handle, x, knots, capacity, knot_count, n_knot_rows,
                                                      rotation, baseline, status, stream),
              "pyitd_extract_with_knots_device")

    def extract_spline_device(self, x, rotation, baseline, knot_count, status, min_knots, stream) -> None:
        check(self._L.pyitd_extract_spline_device(self.handle, x, rotation, baseline, knot_count, status,
                                                  int(min_knots), stream), "pyitd_extract_spline_device")

    def find_knots_device(self, x, kinds, knots, capacity, knot_count, status, stream) -> None:
        check(self._L.pyitd_find_knots_device(self.handle, x, kinds, knots, capacity, knot_count, status,
                                              stream), "pyitd_find_knots_device")


def declared_symbols() -> list[str]:
    """Every function name ``include/pyitd_b200.h`` declares (used by the CPU-side ABI test)."""
    import re

    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"PYITD_API[^;(]*?\b(pyitd_\w+)\s*\(", text)))
