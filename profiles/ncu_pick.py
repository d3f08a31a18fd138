#!/usr/bin/env python
"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output (one column per launch)."""
import csv
import re
import sys

PAT = sys.argv[2] if len(sys.argv) > 2 else (
    r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|launch__registers|"
    r"launch__occupancy_limit|launch__waves|achieved_occupancy|sm__warps_active.avg.pct|"
    r"warp_issue_stalled.*per_warp_active|issue_active.avg.pct|pipe_fp64.*pct|l1tex__data_bank_conflicts|"
    r"lts__t_sector_hit_rate|smsp__inst_executed.sum$|launch__shared_mem_per_block|sm__throughput.avg.pct|"
    r"l1tex__throughput.avg.pct|lts__throughput.avg.pct|launch__grid_size|launch__block_size")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
for i, h in enumerate(hdr):
    if re.search(PAT, h):
        print(f"{h} [{units[i]}]: " + " | ".join(r[i] for r in data))
