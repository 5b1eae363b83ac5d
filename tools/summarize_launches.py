"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.
usage: python tools/summarize_launches.py launches.csv [skip_first_n_launches]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mi = hdr.index("Metric Name")
for r in rd:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((r[ki].split("(")[0], us))
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for k, us in rows:
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
print("launches %d  total %.3f ms" % (len(rows), tot / 1000.0))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-46s n=%5d  total %9.3f ms  share %5.1f%%  avg %8.2f us" % (k[:46], n, us / 1000.0, 100.0 * us / tot, us / n))
