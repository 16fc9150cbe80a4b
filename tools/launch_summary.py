"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`):
launches, total and average duration per kernel, in order of total time."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rd:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    name = re.sub(r"void at::.*?(distribution_elementwise|vectorized_elementwise|elementwise|reduce)_kernel.*", r"torch \1 kernel", name)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:90]:90s} n={n:5d} total={ms:10.3f} ms  avg={ms / n:10.4f} ms  share={ms / tot:6.1%}")
