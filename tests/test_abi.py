"""CPU: the C-ABI library loads and exports what include/pyitd_b200.h declares; host-side logic."""
import ctypes
import os

import numpy as np
import pytest
import torch

import pyitd_b200
from pyitd_b200 import _capi


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        _capi.build()
    return _capi.lib()


def test_header_symbols_are_exported(lib):
    names = _capi.declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pyitd_b200.h but not exported"


def test_abi_version_and_constants(lib):
    assert lib.pyitd_abi_version() == 1
    hdr = open(_capi.HEADER_PATH).read()
    for name, val in (("PYITD_F64", _capi.F64), ("PYITD_F32_MIXED", _capi.F32_MIXED), ("PYITD_F32", _capi.F32),
                      ("PYITD_ST_ZERO_DX", _capi.ST_ZERO_DX), ("PYITD_ST_NONFINITE", _capi.ST_NONFINITE),
                      ("PYITD_STOP_KNOTS", _capi.STOP_KNOTS), ("PYITD_STOP_ITER", _capi.STOP_ITER),
                      ("PYITD_OPT_BASELINES", _capi.OPT_BASELINES), ("PYITD_OPT_ZERO_TAIL", _capi.OPT_ZERO_TAIL),
                      ("PYITD_KNOTS_BOTH", _capi.KNOTS_BOTH)):
        import re
        m = re.search(rf"#define\s+{name}\s+\(?(-?\d+)\)?", hdr)
        assert m and int(m.group(1)) == val, name


def test_no_torch_types_in_abi():
    hdr = open(_capi.HEADER_PATH).read()
    assert "torch" not in hdr.lower() and "at::" not in hdr and "Tensor" not in hdr
    assert 'extern "C"' in hdr


def test_library_does_not_link_the_oracle_or_torch():
    import subprocess
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "itd_oracle" not in out and "torch" not in out
    syms = subprocess.run(["nm", "-D", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "itd_oracle" not in syms


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_a_gpu(lib):
    h = ctypes.c_void_p()
    rc = lib.pyitd_plan_create(ctypes.byref(h), 0, 1, 1024, 0, 11, 2, 0)
    assert rc == _capi.E_NODEVICE and not h
    assert b"CUDA" in lib.pyitd_last_error()
    with pytest.raises(pyitd_b200.PyITDLibraryError):
        pyitd_b200.ITD().itd(np.random.default_rng(0).standard_normal(100))
    with pytest.raises(pyitd_b200.PyITDLibraryError):
        pyitd_b200.decompose(np.zeros((2, 100)))


def test_plan_argument_validation(lib):
    h = ctypes.c_void_p()
    assert lib.pyitd_plan_create(ctypes.byref(h), 0, 0, 1024, 0, 11, 2, 0) == _capi.E_INVALID
    assert lib.pyitd_plan_create(ctypes.byref(h), 0, 1, 2, 0, 11, 2, 0) == _capi.E_INVALID
    assert lib.pyitd_plan_create(ctypes.byref(h), 0, 1, 1024, 7, 11, 2, 0) == _capi.E_INVALID
    assert lib.pyitd_plan_create(ctypes.byref(h), 0, 1, 1024, 0, -1, 2, 0) == _capi.E_INVALID
    assert lib.pyitd_plan_create(None, 0, 1, 1024, 0, 11, 2, 0) == _capi.E_INVALID


def test_reference_interface_surface():
    # ITD.py:157, :177-181, :436-465
    itd = pyitd_b200.ITD()
    assert itd.extrema_detection == "matlab"
    with pytest.raises(AssertionError):
        pyitd_b200.ITD(extrema_detection="cubic")
    with pytest.raises(ValueError):
        itd.get_baselines()
    with pytest.raises(ValueError):
        itd.get_rotations()
    for name in ("itd", "get_baselines", "get_rotations", "__call__"):
        assert callable(getattr(itd, name))
    # eager float64-only signatures of the two jitted functions, ITD.py:33 / ITD.py:79
    with pytest.raises(TypeError):
        pyitd_b200.detect_peaks(np.zeros(10, dtype=np.float32))
    with pytest.raises(TypeError):
        pyitd_b200.itd_baseline_extract(np.zeros(10, dtype=np.int64))


def test_product_never_imports_the_oracle():
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, "pyitd_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
            text = open(path, errors="replace").read()
            assert "itd_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path
