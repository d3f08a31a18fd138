#!/usr/bin/env python
"""Small runs of the SURVEY 8f kernels (spline level, 2-D crossways / ensemble, analytics), meant to run under
compute-sanitizer (memcheck / racecheck / synccheck); every result is checked against the oracle as well."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyitd_b200
from oracle import itd_oracle as o
from pyitd_b200 import analytics
rng = np.random.default_rng(6)
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
for S, n in ((3, 2500), (2, 1024), (200, 2048), (1, 4099)):
    x = np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * (1 + rng.random((S, n))) if S == 3 else rng.standard_normal((S, n))
    R, B, cnt, st = pyitd_b200.extract_spline(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert rel(B[0].cpu().numpy(), o.c_spline_level(x[0])[1]) < 1e-9
    print("ok spline", S, n, int(cnt[0]), flush=True)
img = rng.standard_normal((2, 40, 70)) * 10 + 50
y = pyitd_b200.crossways_batch(torch.from_numpy(img).cuda())
torch.cuda.synchronize()
assert rel(y[1].cpu().numpy(), o.crossways(img[1])) < 1e-9
print("ok crossways", flush=True)
noise = rng.standard_normal((3, 40, 70))
low = pyitd_b200.retrieve_statistical_image_component(img[0], noise=noise, iterations=6)
assert rel(low, o.ensemble2d(img[0], noise)) < 1e-9
print("ok ensemble", flush=True)
x = rng.standard_normal((4, 3000)).cumsum(axis=1)
res = pyitd_b200.decompose(torch.from_numpy(x).cuda(), max_iteration=5)
w = analytics.wpe_rows(res).cpu().numpy()
assert abs(w[2, 1] - o.c_wpe3(res.rows_of(2)[1].cpu().numpy(), True)) < 1e-9
sums, tot = analytics.column_fsum(res.rotations, res.n_rows)
assert np.array_equal(sums[3].cpu().numpy(), o.c_column_fsum(res.rows_of(3).cpu().numpy()))
print("ok analytics", flush=True)
pyitd_b200.clear_plan_cache()
