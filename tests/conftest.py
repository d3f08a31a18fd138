import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when there is no device, so that a plain `pytest tests/`
    # on the CPU box stays green; `-m gpu` on the GPU box runs them for real.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_cases(npz_path):
    """npz with keys 'name/field' -> {name: {field: array}}"""
    import numpy as np

    z = np.load(npz_path)
    out = {}
    for k in z.files:
        name, field = k.split("/", 1)
        out.setdefault(name, {})[field] = z[k]
    return out
