"""Golden fixtures for the 2-D crossways ensemble ITD, FROM THE REFERENCE.

Runs only in the build container.  Usage:  python tests/golden/make_golden_sift2d.py

Executes code cell 1 of /root/reference/siftED2D.ipynb unmodified (it defines, among others,
``crossways_itd_baseline_extract`` and ``retrieve_statistical_image_component``) and records in sift2d_cases.npz:

* ``crossways/<name>/{x,y}``: images and the notebook's ``crossways_itd_baseline_extract`` output;
* ``ensemble/<name>/{x,noise,lowpass}``: the notebook's ensemble driver draws its noise from numba's unseeded
  generator and cannot be replayed, so the fixture is built from the notebook's own crossways function applied to
  ``noise[e] + x`` and ``noise[e] * -1 + x`` and averaged in the driver's order (cell 1, raw :262-272), with the
  draws stored next to it; the scale of the draws is the notebook's ``mad(x)``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def notebook_namespace():
    nb = json.load(open(os.path.join(REF, "siftED2D.ipynb")))
    cells = [c for c in nb["cells"] if c["cell_type"] == "code"]
    # numba's cache=True needs a real file: the cell source goes, byte for byte, into a scratch module under /tmp
    import importlib.util
    import tempfile

    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "numba_cache_sifted2d"))
    path = os.path.join(tempfile.gettempdir(), "sifted2d_cell1.py")
    with open(path, "w") as f:
        f.write("".join(cells[0]["source"]))
    spec = importlib.util.spec_from_file_location("sifted2d_cell1", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sifted2d_cell1"] = mod          # numba's objmode blocks resolve globals through sys.modules
    spec.loader.exec_module(mod)
    return vars(mod)


def main():
    ns = notebook_namespace()
    cross = ns["crossways_itd_baseline_extract"]
    mad = ns["mad"]
    rng = np.random.default_rng(512)
    out = {}

    def image(h, w, kind):
        yy, xx = np.mgrid[0:h, 0:w]
        if kind == "texture":
            return 128 + 40 * np.sin(xx * 0.9 + yy * 0.31) + 25 * np.sin(yy * 1.3) + 10 * rng.standard_normal((h, w))
        if kind == "smooth":      # few extrema per line: exercises the "< 10 extrema -> return x" rule
            return 100 + 50 * np.sin(xx * 2 * np.pi * 1.5 / w) * np.cos(yy * 2 * np.pi / h) + 0.5 * rng.standard_normal((h, w)) * (xx > w // 2)
        return rng.standard_normal((h, w)) * 30 + 100

    for name, (h, w, kind) in {"texture_64x64": (64, 64, "texture"), "noise_48x80": (48, 80, "noise"),
                               "smooth_40x33": (40, 33, "smooth"), "texture_128x96": (128, 96, "texture")}.items():
        x = np.ascontiguousarray(image(h, w, kind), dtype=np.float64)
        y = np.array(cross(x.copy()))
        out[f"crossways/{name}/x"] = x
        out[f"crossways/{name}/y"] = y
        print("crossways", name, x.shape, float(np.abs(y).mean()))

    for name, (h, w, draws) in {"texture_64x48_d3": (64, 48, 3), "noise_32x32_d10": (32, 32, 10)}.items():
        x = np.ascontiguousarray(image(h, w, "texture" if "texture" in name else "noise"), dtype=np.float64)
        m = float(mad(x.copy()))
        noise = rng.normal(0, m, (draws, h, w))
        acc = np.zeros_like(x)
        for e in range(draws):
            a = np.array(cross(noise[e] + x))
            b = np.array(cross((noise[e] * -1) + x))
            acc += (a + b) / 2.0
        low = acc / (draws * 1.0)
        out[f"ensemble/{name}/x"] = x
        out[f"ensemble/{name}/noise"] = noise
        out[f"ensemble/{name}/lowpass"] = low
        out[f"ensemble/{name}/mad"] = np.asarray(m)
        print("ensemble", name, x.shape, m, float(np.abs(low).mean()))
    np.savez_compressed(os.path.join(HERE, "sift2d_cases.npz"), **out)


if __name__ == "__main__":
    main()
