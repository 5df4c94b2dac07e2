"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python tools/launch_summary.py launches.csv [skip_regex]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    if skip and skip.search(name):
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"{'total us':>12} {'count':>6} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:12.1f} {n:6d} {100 * t / tot:5.1f}%  {k[:100]}")
print(f"{tot:12.1f} us total")
