#!/usr/bin/env python3
"""profiles/ncu_regions.py — executed warp instructions / stall samples of an .ncu-rep grouped by source FUNCTION
(the enclosing `__device__`/`__global__` definition of each line), from ncu's source/SASS correlation.
usage: python profiles/ncu_regions.py rep.ncu-rep src.cu [units]   (units = how many work items the launch had,
to print instructions per unit)"""
import collections
import csv
import re
import subprocess
import sys


def regions(path):
    """[(first_line, name)] for every function definition in the file (crude: lines that start a definition)."""
    out = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|extern)\b.*?\b(\w+)\s*\(")
    for i, l in enumerate(open(path), 1):
        m = pat.match(l)
        if m and not l.rstrip().endswith(";"):
            out.append((i, m.group(1)))
    return out


def main():
    rep, src = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    regs = regions(src)
    base = src.split("/")[-1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    cur, hdr = None, None
    inst, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue
        extra = len(r) - len(hdr)
        if extra > 0:
            r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
        elif extra < 0:
            continue
        name = cur
        if cur == base:
            name = "?"
            for first, fn in regs:
                if first <= int(r[0]):
                    name = fn
        inst[name] += int(r[hdr.index("Instructions Executed")] or 0)
        thr[name] += int(r[hdr.index("Thread Instructions Executed")] or 0)
        smp[name] += int(r[hdr.index("# Samples")] or 0)
    T, S = sum(inst.values()) or 1, sum(smp.values()) or 1
    print(f"warp instructions {T}, samples {S}")
    for k, v in inst.most_common():
        per = f"  {v / units:9.1f} warp-inst/unit" if units else ""
        print(f"{k:28s} inst {100 * v / T:5.1f}%  samples {100 * smp[k] / S:5.1f}%  active threads/inst {thr[k] / max(v, 1):4.1f}{per}")


if __name__ == "__main__":
    main()
