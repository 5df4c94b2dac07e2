"""Per-kernel device time of ONE bench step from an `ncu --metrics gpu__time_duration.sum -k regex:^k_ --csv` launch list
(one step = from one k_mesh_tables launch, the first kernel of a step, to the next; the 4th such step of the list). Usage: python tools/step_kernels.py launches.csv [other.csv]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
    data = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        data.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("msm::", ""), v))
    b = [i for i, (n, _) in enumerate(data) if n.startswith("k_mesh_tables")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in data[b[3]:b[4]]:
        agg[k][0] += 1
        agg[k][1] += v
    return agg


aggs = [load(p) for p in sys.argv[1:]]
names = sorted(set().union(*[a.keys() for a in aggs]), key=lambda n: -aggs[0].get(n, [0, 0])[1])
tot = [sum(v[1] for v in a.values()) for a in aggs]
print("%-44s" % "kernel" + "".join("%16s" % p.split("/")[-1][:15] for p in sys.argv[1:]))
for n in names:
    print("%-44s" % n[:44] + "".join("%10.1f x%-4d" % (a.get(n, [0, 0])[1], a.get(n, [0, 0])[0]) for a in aggs))
print("%-44s" % "total us" + "".join("%10.1f      " % t for t in tot))
