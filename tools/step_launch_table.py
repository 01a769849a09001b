import csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
out = []
for r in rows:
    k0 = r["Kernel Name"]
    k = re.sub(r"\(anonymous namespace\)::", "", k0)
    k = re.sub(r"\(.*", "", k).replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")
    out.append((int(r["ID"]), r["Stream"], r["Grid Size"], k[:60], float(r["Metric Value"].replace(",", "")) / 1000))
idx = [i for i, o in enumerate(out) if "dropout_mask_kernel" in o[3]]
print("markers", idx[:10], len(out))
a = idx[-1] if idx else 0
step = out[idx[-2]:idx[-1]] if len(idx) > 1 else out[a:]
print(len(step), "launches, sum", sum(o[4] for o in step))
from collections import defaultdict
bys = defaultdict(float); cnt = defaultdict(int)
for o in step:
    bys[o[1]] += o[4]; cnt[o[1]] += 1
for s in bys: print("stream", s, cnt[s], "launches", round(bys[s], 1), "us")
if len(sys.argv) > 2:
    for o in step:
        if sys.argv[2] == "all" or o[1] == sys.argv[2]: print(f"{o[1]:4s} {o[4]:8.1f} {o[2]:16s} {o[3]}")
