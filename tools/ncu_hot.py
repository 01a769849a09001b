"""Hot SASS instructions of one kernel from an `ncu --page source --csv` dump.  usage: python tools/ncu_hot.py file.csv <kernel substring> [topN] [context]"""
import csv, sys
path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ctxn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
kern, rows, cur = None, {}, None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "Kernel Name":
        cur = r[1]; rows.setdefault(cur, []); hdr = None; continue
    if r[0] == "Address":
        hdr = r; continue
    rows[cur].append(r)
names = [k for k in rows if key in k]
for k in names[:1]:
    rs = rows[k]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in rs)
    print(k[:120]); print("instructions:", len(rs), "samples:", tot)
    order = sorted(range(len(rs)), key=lambda i: -int(rs[i][si]))[:top]
    shown = set()
    for i in sorted(order):
        for j in range(max(0, i - ctxn), min(len(rs), i + ctxn + 1)):
            if j in shown: continue
            shown.add(j)
            r = rs[j]
            st = sorted(((int(r[c]), h[6:]) for c, h in stall_cols if r[c] not in ("", "0")), reverse=True)[:3]
            mark = "*" if j in order else " "
            print(f"{mark}{j:5d} {100*int(r[si])/max(tot,1):5.1f}% exec {r[ii]:>8s}  {r[1].strip()[:70]:70s} {' '.join(f'{n}:{v}' for v, n in st)}")
        if ctxn: print("      ...")
