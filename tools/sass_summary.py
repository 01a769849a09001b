"""Per-kernel SASS evidence of the Blackwell paths in libmml_b200.so (cuobjdump -sass): counts of tcgen05 MMA (UTCHMMA), TMEM loads
(LDTM), TMA loads / stores (UTMALDG / UTMASTG), mbarrier waits (SYNCS), fp64 atomics (RED/ATOM .F64) per kernel.
usage: python tools/sass_summary.py > profiles/r2_sass_tcgen05.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "task-specific-pretraining-multimodal_b200", "libmml_b200.so")
PATS = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
                                ("UTMAPF", r"\bUTMAPF|UTMACCTL"), ("SYNCS", r"\bSYNCS"), ("ATOM/RED.F64", r"\b(RED|ATOM|ATOMG)\.[A-Z0-9.]*F64"),
                                ("HMMA(legacy)", r"\bHMMA\b"), ("instr", r"^\s+/\*[0-9a-f]{4}\*/")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    dem = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k, p in PATS.items():
            if re.search(p, line):
                kernels[cur][k] += 1
    names = list(kernels)
    d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip(names, d))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a); one row per kernel that uses tensor cores, TMEM, TMA or fp64 atomics")
    print("# " + "  ".join(f"{k:>12s}" for k in PATS) + "  kernel")
    tot = collections.Counter()
    for n, c in kernels.items():
        tot.update(c)
        if not any(c[k] for k in list(PATS)[:8]):
            continue
        name = re.sub(r"\(anonymous namespace\)::", "", dem.get(n, n))
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)
        print("  " + "  ".join(f"{c[k]:12d}" for k in PATS) + "  " + name[:110])
    print("# total over %d kernels" % len(kernels))
    print("  " + "  ".join(f"{tot[k]:12d}" for k in PATS))


if __name__ == "__main__":
    sys.exit(main())
