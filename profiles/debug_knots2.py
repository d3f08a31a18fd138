import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["PYITD_FORCE_PATH"] = "strided"
import pyitd_b200
from pyitd_b200 import _capi, synth
from oracle import itd_oracle as o
import test_gpu_parity as T
rng = np.random.default_rng(340)
for n in (4, 8, 36, 128, 1020, 1024, 1028, 2044, 2048, 2052, 3076, 5120, 10004, 20000):
    for kind in range(6):
        x = T._mixed_batch(rng, 6, n)[kind:kind + 1]
        T.check_against_oracle(x, max_iteration=11)
for mi in (0, 1, 4):
    T.check_against_oracle(T._mixed_batch(rng, 4, 6000)[3:4], max_iteration=mi)
T.check_against_oracle(T._mixed_batch(rng, 2, 9000)[1:2], max_iteration=11, min_extrema=5)
xz = T._mixed_batch(rng, 4, 5000)[3:4]
rz = pyitd_b200.decompose(T.gpu(xz), max_iteration=11, return_baselines=True, zero_tail=True)
ref = T.check_against_oracle(xz, max_iteration=11)
if "--skip-long" not in sys.argv:
    x32 = synth.long_signal(n=1 << 20, seed=3).numpy()
    for dt in ("f32_mixed", "f32"):
        res = pyitd_b200.decompose(T.gpu(x32[None, :]), max_iteration=11, dtype=dt, return_baselines=True)
print("pre-checks ok")
xs = rng.standard_normal((1, 7000)).cumsum(axis=1)
xg = T.gpu(xs)
kn, c, st = pyitd_b200.find_knots(xg)
K = int(c[0]); ko = o.c_find_knots(xs[0])
print("K", K, len(ko), "knots equal", np.array_equal(kn[0, :K].cpu().numpy(), ko), "status", int(st[0]))
if not np.array_equal(kn[0, :K].cpu().numpy(), ko):
    a = kn[0, :K].cpu().numpy(); m = min(len(a), len(ko)); dd = np.flatnonzero(a[:m] != ko[:m]); print("first knot diffs at", dd[:10], a[dd[:5]], ko[dd[:5]])
R, B, st2 = pyitd_b200.extract_with_knots(xg, kn, c)
wr, wb, _ = o.c_extract_level(xs[0])
d = np.flatnonzero(R[0].cpu().numpy() != wr)
print("with_knots: ndiff", d.size, d[:10], "status", int(st2[0]))
R2, B2, cnt, st3 = pyitd_b200.extract_level(xg)
d2 = np.flatnonzero(R2[0].cpu().numpy() != wr)
print("extract_level: ndiff", d2.size, d2[:10])
R3, B3, st4 = pyitd_b200.extract_with_knots(xg, torch.from_numpy(ko.astype(np.int32)).cuda())
d3 = np.flatnonzero(R3[0].cpu().numpy() != wr)
print("with oracle knots: ndiff", d3.size, d3[:10], int(st4[0]))
