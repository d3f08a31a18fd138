#!/usr/bin/env python
"""Small decompositions on every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyitd_b200
from oracle import itd_oracle as o
rng = np.random.default_rng(5)
cases = [("stream", 5, 5000, {}), ("stream", 3, 1028, {"PYITD_GROUPS": "2"}), ("lookback", 2, 3000, {}),
         ("strided", 1, 5000, {"PYITD_STRIDED_CTAS": "2"}), ("strided", 1, 2052, {}), ("resident", 3, 4097, {})]
for path, S, n, env in cases:
    os.environ["PYITD_FORCE_PATH"] = path
    for k in ("PYITD_GROUPS", "PYITD_STRIDED_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pyitd_b200.clear_plan_cache()
    x = rng.standard_normal((S, n)).cumsum(axis=1) + 0.5 * rng.standard_normal((S, n))
    res = pyitd_b200.decompose(torch.from_numpy(x).cuda(), max_iteration=5, return_baselines=True)
    torch.cuda.synchronize()
    for s in range(S):
        want = o.c_decompose(x[s], 5)
        assert res.rows_of(s).cpu().numpy().tobytes() == want.rotations.tobytes(), (path, s)
    if path != "resident":
        kn, c, _ = pyitd_b200.find_knots(torch.from_numpy(x).cuda())
        R, B, st = pyitd_b200.extract_with_knots(torch.from_numpy(x).cuda(), kn[0, : int(c[0])])
        torch.cuda.synchronize()
    print("ok", path, S, n, env, flush=True)
pyitd_b200.clear_plan_cache()
