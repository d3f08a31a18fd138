import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PYITD_FORCE_PATH"] = sys.argv[1] if len(sys.argv) > 1 else "strided"
import pyitd_b200
from oracle import itd_oracle as o
rng = np.random.default_rng(340)
for n in (7000, 5000, 1024):
    xs = rng.standard_normal((1, n)).cumsum(axis=1)
    xg = torch.from_numpy(xs).cuda()
    kn, c, st = pyitd_b200.find_knots(xg)
    K = int(c[0]); ko = o.c_find_knots(xs[0])
    print(n, "K", K, len(ko), "knots equal", np.array_equal(kn[0, :K].cpu().numpy(), ko), "status", int(st[0]))
    R, B, st2 = pyitd_b200.extract_with_knots(xg, kn, c)
    wr, wb, _ = o.c_extract_level(xs[0])
    d = np.flatnonzero(R[0].cpu().numpy() != wr)
    print("  with_knots: ndiff", d.size, d[:10], "status", int(st2[0]))
    R2, B2, cnt, st3 = pyitd_b200.extract_level(xg)
    d2 = np.flatnonzero(R2[0].cpu().numpy() != wr)
    print("  extract_level: ndiff", d2.size, d2[:10])
    if d.size:
        i = d[0]; print("   first diff", i, R[0, i].item(), wr[i], "knots near", ko[:6])
