#!/usr/bin/env python
"""Turn `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` output for ONE launch of
the dominant kernel into profiles/dominant_kernel_traffic.json, keyed by the hash of the CUDA sources that were
profiled (bench.py reports roofline.traffic only when that hash matches the sources it runs).

usage: write_traffic_json.py <ncu.csv> <channels> [note]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
vals = {}
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    name, unit, v = r[ci["Metric Name"]], r[ci["Metric Unit"]], float(r[ci["Metric Value"]].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)
    vals.setdefault(name, []).append(v * scale)
    kernel = r[ci["Kernel Name"]]
rd, wr, tm = vals["dram__bytes_read.sum"][-1], vals["dram__bytes_write.sum"][-1], vals["gpu__time_duration.sum"][-1]
out = {
    "kernel": kernel, "source_hash": bench.kernel_source_hash(), "channels": int(sys.argv[2]),
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "gpu_time_ms_under_ncu": tm * 1e3,
    "how": "ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum on the one sweep_kernel launch that "
           "decomposes %s x 65536 fp64 channels (profiles/run_ncu_traffic.sh); keyed by the hash of pyitd_b200/csrc/*.cu*" % sys.argv[2],
    "note": sys.argv[3] if len(sys.argv) > 3 else "",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"), "w"), indent=1)
print(json.dumps(out))
