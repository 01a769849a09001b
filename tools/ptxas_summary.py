"""Registers / stack / spills / static shared memory of every kernel from the `-Xptxas -v` logs the Makefile keeps next to the objects
(csrc/*.ptxas.log).  CPU only:  python tools/ptxas_summary.py > profiles/r2_ptxas_summary.txt"""
import glob
import os
import re
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "task-specific-pretraining-multimodal_b200", "csrc")
ENTRY = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'")
FRAME = re.compile(r"^\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads")
USED = re.compile(r"Used (\d+) registers")
SMEM = re.compile(r"(\d+) bytes smem")


def demangle(names):
    if not names:
        return []
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names) + "\n", stdout=subprocess.PIPE, text=True, check=True, timeout=30).stdout.splitlines()
        out = [re.sub(r"^void ", "", o).replace("(anonymous namespace)::", "") for o in out]
        return [o[: o.rfind(">") + 1] if ">(" in o else o.split("(", 1)[0] for o in out]
    except Exception:
        return names


def main():
    total = spilled = 0
    for path in sorted(glob.glob(os.path.join(CSRC, "*.ptxas.log"))):
        print("==", os.path.basename(path))
        rows, cur = [], None
        for line in open(path):  # line by line: entry -> frame line -> "Used ..." line
            m = ENTRY.search(line)
            if m:
                cur = {"name": m.group(1), "stack": 0, "st": 0, "ld": 0}
                continue
            if cur is None:
                continue
            m = FRAME.match(line)
            if m:
                cur["stack"], cur["st"], cur["ld"] = (int(x) for x in m.groups())
                continue
            m = USED.search(line)
            if m:
                s = SMEM.search(line)
                rows.append((cur["name"], int(m.group(1)), cur["stack"], cur["st"], cur["ld"], int(s.group(1)) if s else 0))
                cur = None
        for (name, regs, stack, st, ld, smem), nice in zip(rows, demangle([r[0] for r in rows])):
            print(f"  {nice[:110]:<110} regs {regs:>3}  stack {stack:>4} B  spill {st}/{ld} B  static smem {smem} B")
            total += 1
            spilled += 1 if (st or ld) else 0
    print(f"== {total} kernels, {spilled} with spill stores / loads")


if __name__ == "__main__":
    main()
