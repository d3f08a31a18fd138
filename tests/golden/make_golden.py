"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only in the build container, where the reference is mounted at /root/reference; the GPU box
never sees that path, which is why the outputs are committed.  Usage:

    python tests/golden/make_golden.py

What it records (all produced by /root/reference/ITD.py's own numba functions, imported
unmodified; the only work-around is ``ITD.S = ITD.T = data`` for the NameError at ITD.py:375):

* notebook_8000.npz  -- the reference's single golden vector (PyITD.ipynb code cell 2) with the
                        full ``ITD().itd`` output, ``get_baselines()``, level-1 knots and the
                        per-pass knot counts the reference prints (ITD.py:403);
* demo_400.npz       -- the ``__main__`` demo signal ITD.py:491-495, same contents;
* config1.npz        -- BASELINE.json configs[0] (65 536-sample chirp + noise): knots, counts,
                        SHA-256 of the outputs and the first/last 256 samples of every row;
* small_cases.npz    -- plateaus, quantised ramps, random walks, monotone input, iteration-cap
                        cases: full outputs for N <= 4096;
* level_cases.npz    -- ``detect_peaks`` / ``itd_baseline_extract`` called directly;
* error_cases.json   -- inputs on which the reference raises, with the exception type.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

sys.path.insert(0, REF)
sys.path.insert(0, REPO)
import ITD as ref  # noqa: E402  (the reference module, eager-JITs on import)

from pyitd_b200 import synth  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference(x: np.ndarray, max_iteration: int = 11):
    """ITD().itd through the reference's public entry point; returns rotations, baselines,
    printed knot counts and the final message."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    ref.S = ref.T = x                      # ITD.py:375 reads module globals
    obj = ref.ITD()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        rows = obj.itd(x.copy(), max_iteration=max_iteration)
    lines = buf.getvalue().split("\n")
    counts = np.asarray([int(s) for s in lines if s.strip().lstrip("-").isdigit()], dtype=np.int64)
    msg = [s for s in lines if s and not s.strip().isdigit()]
    return (np.array(rows, copy=True), np.array(obj.get_baselines(), copy=True), counts,
            msg[-1] if msg else "")


def ref_knots(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    a = ref.detect_peaks(x.copy())
    b = ref.detect_peaks(-x)
    return np.sort(np.unique(np.hstack((a, b)))).astype(np.int64)   # ITD.py:97


def notebook_vector() -> np.ndarray:
    nb = json.load(open(os.path.join(REF, "PyITD.ipynb")))
    cells = [c for c in nb["cells"] if c["cell_type"] == "code"]
    src = "".join(cells[1]["source"])
    ns: dict = {}
    exec(src, {"numpy": np, "np": np}, ns)   # defines `inputarray`
    return np.asarray(ns["inputarray"], dtype=np.float64)


def full_case(x, max_iteration=11):
    rows, bases, counts, msg = run_reference(x, max_iteration)
    return dict(x=np.asarray(x, dtype=np.float64), rotations=rows, baselines=bases,
                knot_counts=counts, knots=ref_knots(x), message=np.asarray(msg),
                max_iteration=np.asarray(max_iteration))


def main():
    # 1. the reference's golden vector ------------------------------------------------------
    x = notebook_vector()
    assert x.shape == (8000,)
    case = full_case(x)
    assert case["rotations"].shape == (9, 8000) and case["baselines"].shape == (8, 8000)
    np.savez_compressed(os.path.join(HERE, "notebook_8000.npz"), **case)
    print("notebook_8000", case["rotations"].shape, case["knot_counts"], sha(case["rotations"])[:16])

    # 2. ITD.py __main__ demo ---------------------------------------------------------------
    T = np.linspace(0, 2 * np.pi, 400, dtype=np.float64)
    S = np.sin(20 * T * (1 + 0.2 * T)) + T ** 2 + np.sin(13 * T)
    case = full_case(S)
    np.savez_compressed(os.path.join(HERE, "demo_400.npz"), **case)
    print("demo_400", case["rotations"].shape, case["knot_counts"])

    # 3. config 1 ---------------------------------------------------------------------------
    x = synth.config1_chirp()
    rows, bases, counts, msg = run_reference(x, max_iteration=20)
    np.savez_compressed(
        os.path.join(HERE, "config1.npz"),
        knots=ref_knots(x).astype(np.int32), knot_counts=counts,
        rotations_sha=np.asarray(sha(rows)), baselines_sha=np.asarray(sha(bases)),
        input_sha=np.asarray(sha(x)), shape=np.asarray(rows.shape),
        head=rows[:, :256].copy(), tail=rows[:, -256:].copy(),
        row_sums=rows.sum(axis=1), message=np.asarray(msg), max_iteration=np.asarray(20))
    print("config1", rows.shape, counts, msg)

    # 4. small cases -------------------------------------------------------------------------
    rng = np.random.default_rng(2024)
    small = {}

    def add(name, x, mi=11):
        c = full_case(x, mi)
        for k, v in c.items():
            small[f"{name}/{k}"] = v
        print("small", name, c["rotations"].shape, c["knot_counts"], str(c["message"]))

    add("white_1024", rng.standard_normal(1024))
    add("white_4096_cap0", rng.standard_normal(4096), 0)
    add("white_4096_cap1", rng.standard_normal(4096), 1)
    add("white_4096_cap3", rng.standard_normal(4096), 3)
    add("walk_3000", np.cumsum(rng.standard_normal(3000)))
    add("quantised_2048", np.round(np.cumsum(rng.standard_normal(2048)) * 2) / 2 + 0.001 * np.arange(2048))
    add("plateaus_1500", np.repeat(rng.standard_normal(300), 5) + 1e-3 * np.arange(1500))
    add("monotone_1000", np.arange(1000.0))
    add("monotone_dec_777", -np.arange(777.0) ** 1.5)
    add("sine_100", np.sin(2 * 2 * np.pi * np.linspace(0, 1, 100)))
    add("short_3", np.array([0.0, 1.0, 0.5]))
    add("short_4", np.array([0.0, 1.0, 0.5, 2.0]))
    add("short_5", np.array([0.3, -1.0, 0.5, 0.25, 2.0]))
    add("two_tone_2047", np.sin(np.arange(2047) * 0.37) + 0.3 * np.sin(np.arange(2047) * 2.1))
    add("tiny_amp_2049", 1e-300 * rng.standard_normal(2049))
    add("huge_amp_2050", 1e300 * rng.standard_normal(2050))
    add("negzero_64", np.where(rng.standard_normal(64) > 0, 1.0, -1.0) * np.abs(rng.standard_normal(64)) * np.array([0.0 if i % 7 == 3 else 1.0 for i in range(64)]) - 0.0)
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **small)

    # 5. direct calls of the two jitted functions --------------------------------------------
    lv = {}
    for i, n in enumerate((3, 4, 7, 33, 257, 2048, 5000)):
        x = rng.standard_normal(n)
        if i % 2:
            x = np.round(x * 3) / 3 + 1e-6 * np.arange(n)      # ties between neighbours
        try:
            R, B = ref.itd_baseline_extract(x.copy())
        except ZeroDivisionError:
            continue
        lv[f"{n}/x"] = x
        lv[f"{n}/R"] = np.array(R)
        lv[f"{n}/B"] = np.array(B)
        lv[f"{n}/valleys"] = np.array(ref.detect_peaks(x.copy()))
        lv[f"{n}/peaks"] = np.array(ref.detect_peaks(-x))
        lv[f"{n}/knots"] = ref_knots(x)
    np.savez_compressed(os.path.join(HERE, "level_cases.npz"), **lv)
    print("level cases", sorted({k.split('/')[0] for k in lv}))

    # 6. inputs on which the reference raises ------------------------------------------------
    errors = []
    for name, x in (("ones_1000", np.ones(1000)), ("flat_start", np.array([1.0, 1, 3, 0, 2, 5, 1, 4])),
                    ("flat_end", np.array([1.0, 4, 2, 5, 0, 3, 3])), ("zeros_16", np.zeros(16)),
                    ("equal_ends_monotone_then_back", np.array([0.0, 1.0, 2.0, 1.0, 0.0]))):
        try:
            run_reference(x)
            errors.append(dict(name=name, x=x.tolist(), raises=None))
        except Exception as e:   # noqa: BLE001 - we record whatever the reference raises
            errors.append(dict(name=name, x=x.tolist(), raises=type(e).__name__))
        print("error case", errors[-1]["name"], errors[-1]["raises"])
    json.dump(errors, open(os.path.join(HERE, "error_cases.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
