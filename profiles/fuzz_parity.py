#!/usr/bin/env python
"""Randomised parity run: random batch shapes, signal kinds, options and kernel-path switches against the C oracle, bit for
bit, until the time budget is used up.  python profiles/fuzz_parity.py [--seconds 150] [--seed 1]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(rng, S, n):
    import numpy as np
    x = rng.standard_normal((S, n))
    for s in range(S):
        m = int(rng.integers(0, 8))
        if m == 1:
            x[s] = np.cumsum(x[s])
        elif m == 2:
            x[s] = np.round(x[s] * 4) / 4 + 1e-7 * np.arange(n)
        elif m == 3:
            x[s] = np.sin(np.arange(n) * (0.001 + 0.3 * rng.random())) + 0.01 * x[s]
        elif m == 4:
            x[s] = np.cumsum(np.cumsum(x[s])) * 1e-3
        elif m == 5:
            x[s, : n // 2] = np.linspace(0, 1, n // 2)
        elif m == 6:
            x[s] = np.round(np.cumsum(x[s]) * 16) / 16
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=150)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import numpy as np
    import torch
    import pyitd_b200
    from oracle import itd_oracle as o
    from pyitd_b200 import _capi
    rng = np.random.default_rng(args.seed)
    t_end = time.time() + args.seconds
    cases = fails = 0
    by_path = {}
    keys = ("PYITD_FORCE_PATH", "PYITD_SWEEP_FUSE", "PYITD_SWEEP_DEPTH", "PYITD_SWEEP_FUSE_MIN_A", "PYITD_SWEEP_FUSE_MIN_B",
            "PYITD_COOP_CHUNK")
    while time.time() < t_end:
        for k in keys:
            os.environ.pop(k, None)
        kind = rng.choice(["sweep", "sweep", "coop", "auto"])
        if kind == "sweep":
            S, n = int(rng.integers(1, 400)), int(rng.choice([2048, 3000, 4097, 8192, 12345, 20000, 40000, 65536]))
            if S * n > 12_000_000:
                S = max(1, 12_000_000 // n)
            os.environ["PYITD_FORCE_PATH"] = "sweep"
            os.environ["PYITD_SWEEP_FUSE"] = str(rng.choice([0, 1, 2, 2]))
            os.environ["PYITD_SWEEP_DEPTH"] = str(rng.integers(0, 2))
            if rng.random() < 0.5:
                os.environ["PYITD_SWEEP_FUSE_MIN_A"] = str(rng.choice([4, 8, 30]))
                os.environ["PYITD_SWEEP_FUSE_MIN_B"] = str(rng.choice([1, 2, 6]))
        elif kind == "coop":
            S, n = int(rng.integers(1, 17)), int(rng.integers(3, 70000))
            os.environ["PYITD_FORCE_PATH"] = "coop"
            if rng.random() < 0.5:
                os.environ["PYITD_COOP_CHUNK"] = str(rng.choice([256, 512, 1024]))
        else:
            S, n = int(rng.integers(1, 200)), int(rng.integers(3, 30000))
        mi = int(rng.choice([0, 1, 3, 7, 11, 20]))
        me = int(rng.choice([2, 2, 2, 1, 3, 5]))
        x = make(rng, S, n)
        pyitd_b200.clear_plan_cache()
        res = pyitd_b200.decompose(torch.from_numpy(x).cuda(), max_iteration=mi, min_extrema=me, return_baselines=True,
                                   zero_tail=bool(rng.integers(0, 2)))
        torch.cuda.synchronize()
        from pyitd_b200 import itd as _itd
        path = next(reversed(_itd._PLAN_CACHE.values())).path[0]
        by_path[path] = by_path.get(path, 0) + 1
        status = res.status.cpu().numpy()
        for s in range(S):
            try:
                want = o.c_decompose(x[s], mi, me)
            except o.OracleError as e:
                if not (status[s] & e.status):
                    fails += 1
                    print("STATUS MISMATCH", dict(os.environ.items() & {}.items()), kind, S, n, mi, me, s, flush=True)
                continue
            good = (status[s] == 0 and res.rows_of(s).cpu().numpy().tobytes() == want.rotations.tobytes()
                    and res.baselines_of(s).cpu().numpy().tobytes() == want.baselines.tobytes()
                    and res.knot_counts[s, : want.rotations.shape[0]].cpu().tolist() == list(want.knot_counts)
                    and int(res.stop_kind[s]) == want.stop_kind)
            if not good:
                fails += 1
                print("MISMATCH", {k: os.environ.get(k) for k in keys}, kind, S, n, mi, me, s, flush=True)
                break
        cases += 1
    print(json.dumps({"cases": cases, "failures": fails, "by_path": by_path, "seed": args.seed, "seconds": args.seconds}))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
