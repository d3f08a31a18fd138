#!/usr/bin/env python
"""Round-2 kernels under compute-sanitizer (memcheck / racecheck / synccheck): the sweep kernel (stage-major with fused pairs,
failed predictions and fallbacks; signal-major) and the cooperative kernel (one and several CTAs per signal, several signals
per group), each checked against the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyitd_b200
from oracle import itd_oracle as o
rng = np.random.default_rng(5)
cases = [("sweep", 6, 9000, 11, {"PYITD_SWEEP_FUSE": "2", "PYITD_SWEEP_DEPTH": "0", "PYITD_SWEEP_FUSE_MIN_A": "4", "PYITD_SWEEP_FUSE_MIN_B": "2"}),
         ("sweep", 5, 4100, 7, {"PYITD_SWEEP_FUSE": "2", "PYITD_SWEEP_DEPTH": "1"}),
         ("sweep", 4, 2050, 11, {"PYITD_SWEEP_FUSE": "0"}),
         ("coop", 3, 5000, 11, {}), ("coop", 2, 3000, 11, {"PYITD_COOP_CHUNK": "256"}), ("coop", 1, 258, 5, {"PYITD_COOP_CHUNK": "256"})]
keys = ("PYITD_SWEEP_FUSE", "PYITD_SWEEP_DEPTH", "PYITD_SWEEP_FUSE_MIN_A", "PYITD_SWEEP_FUSE_MIN_B", "PYITD_COOP_CHUNK")
for path, S, n, mi, env in cases:
    os.environ["PYITD_FORCE_PATH"] = path
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    pyitd_b200.clear_plan_cache()
    x = rng.standard_normal((S, n)).cumsum(axis=1) + 0.5 * rng.standard_normal((S, n))
    x[0] = np.round(x[0] * 8) / 8                                   # flat steps: predictions that fail their check
    res = pyitd_b200.decompose(torch.from_numpy(x).cuda(), max_iteration=mi, return_baselines=True, zero_tail=True)
    torch.cuda.synchronize()
    for s in range(S):
        try:
            want = o.c_decompose(x[s], mi)
        except o.OracleError:
            continue
        assert res.rows_of(s).cpu().numpy().tobytes() == want.rotations.tobytes(), (path, s)
    print("ok", path, S, n, env, flush=True)
pyitd_b200.clear_plan_cache()
