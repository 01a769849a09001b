"""Per-kernel durations of the LAST complete step in an ncu launch list (gpu__time_duration.sum CSV).
usage: python tools/step_launches.py gpurun_out/mm_launches.csv [marker-substring]"""
import csv
import re
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "FillFunctor"
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = []
r = csv.reader(lines)
next(r)
for x in r:
    try:
        rows.append((x[4], x[8], float(x[-1])))
    except ValueError:
        pass
idx = [i for i, (k, g, t) in enumerate(rows) if marker in k]
a, b = idx[-2], idx[-1]
tot = 0.0
for k, g, t in rows[a:b]:
    k = re.sub(r"\(.*", "", k).replace("void <unnamed>::", "").replace("<unnamed>::", "")
    print(f"{t / 1000:8.1f} us  {g:14s} {k[:80]}")
    tot += t
print(f"sum {tot / 1000:.1f} us over {b - a} launches")
