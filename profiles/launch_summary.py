"""Summarise the LAST pass of an ncu launch list (csv of gpu__time_duration.sum) of a script that runs its workload `passes` times:
per kernel name + grid, launches and total microseconds.    python profiles/launch_summary.py launches.csv [passes=3]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
data = [(r[ik], r[ig], r[ib], float(r[iv].replace(",", ""))) for r in rows[1:] if len(r) > iv and r[hdr.index("Metric Name")] == "gpu__time_duration.sum"]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
unit = rows[1][hdr.index("Metric Unit")]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
n = len(data) // passes
last = data[-n:]
agg = OrderedDict()
for name, grid, block, v in last:
    key = (name.split("(")[0][-60:], grid)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v * scale
tot = sum(a[1] for a in agg.values())
print("last pass: %d launches, %.1f us" % (len(last), tot))
for (name, grid), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%7.1f us %5.1f%%  x%-3d %-62s %s" % (t, 100 * t / tot, c, name, grid))
