"""Per-kernel shares of ONE training step from an `ncu --metrics gpu__time_duration.sum` launch list (csv or csv.gz):
the step is everything from the last pack_weights launch on. usage: python tools/launch_summary.py launches.csv.gz"""
import collections
import csv
import gzip
import re
import sys

path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ix = {h: i for i, h in enumerate(hdr)}
data = list(r)
names = [d[ix["Kernel Name"]] for d in data]
start = max(i for i, nm in enumerate(names) if "pack_weights" in nm and i < len(names) - 100)
step = data[start:]
agg = collections.defaultdict(lambda: [0, 0.0])
for d in step:
    nm = re.sub(r"\(.*", "", d[ix["Kernel Name"]]).replace("b200seg::", "")
    nm = re.sub(r"void |at::native::", "", nm)[:60]
    a = agg[nm]
    a[0] += 1
    a[1] += float(d[ix["Metric Value"]].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
print("one step: %d launches, %.3f ms of serialised kernel time (ncu: cold caches, one kernel at a time)" % (len(step), tot))
print("%-62s %6s %10s %7s %9s" % ("kernel", "n", "ms", "share", "avg us"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %6d %10.3f %6.1f%% %9.1f" % (k, v[0], v[1], 100 * v[1] / tot, 1e3 * v[1] / v[0]))
