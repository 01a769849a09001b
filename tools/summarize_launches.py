"""Aggregates an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"void ", "", name)
    rows.append((name.split("(")[0][:90], ns))
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, total {tot / 1e6:.3f} ms (cold-cache, serialised: compare SHARES)")
print(f"{'share':>7} {'ms':>9} {'n':>5} {'us/launch':>10}  kernel")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * ns / tot:6.2f}% {ns / 1e6:9.3f} {c:5d} {ns / c / 1e3:10.1f}  {n}")
