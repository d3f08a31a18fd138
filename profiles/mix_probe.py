#!/usr/bin/env python
"""Ceiling of the memory system for the level kernel's traffic mix (1 read : 2 write streams of float64)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyitd_b200 import _capi
L = _capi.lib()
n = 1 << 28                                   # 2 GiB per stream
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x); z = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
out = {}
for ctas, chunk in ((148 * 4, 0), (148 * 8, 0), (148 * 16, 0), (148 * 32, 0),
                    (444, 65536), (888, 65536), (1184, 65536), (4096, 65536), (444, 1024), (1184, 1024)):
    for _ in range(3):
        _capi.check(L.pyitd_probe_mixed_traffic(x.data_ptr(), y.data_ptr(), z.data_ptr(), n, ctas, chunk, st), "probe")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.pyitd_probe_mixed_traffic(x.data_ptr(), y.data_ptr(), z.data_ptr(), n, ctas, chunk, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[f"{ctas}x256 chunk {chunk}"] = {"ms": ms, "GBps": 24.0 * n / ms / 1e6}
# plain copy for reference (1 : 1)
for _ in range(3):
    y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out["torch_copy"] = {"ms": ms, "GBps": 16.0 * n / ms / 1e6}
for _ in range(3):
    y.zero_()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    y.zero_()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out["torch_fill"] = {"ms": ms, "GBps": 8.0 * n / ms / 1e6}
print(json.dumps(out))
