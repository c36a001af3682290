"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: total time, share, launches.
Usage: python tools/launch_summary.py gpurun_out/launches.csv "<header comment>" > profiles/launches_summary.txt"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0][:70]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[k]:10.3f} ms {100 * tot[k] / total:6.2f}% {cnt[k]:5d} launches  {k}")
print(f"total {total:.3f} ms")
