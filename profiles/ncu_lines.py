#!/usr/bin/env python3
"""profiles/ncu_lines.py — stall samples of an .ncu-rep aggregated per CUDA source line.

ncu's CSV source page carries metrics only for SASS; this script disassembles the matching cubin with
line info (nvdisasm -g) and aligns the two listings by instruction order.
usage: python profiles/ncu_lines.py rep.ncu-rep lib.so kernel_substring [topN]"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel not in out:
            continue
        lines, cur, infunc = [], None, False
        for l in out.splitlines():
            if l.startswith(".text.") and kernel in l:
                infunc = True
                continue
            if l.startswith(".text.") and kernel not in l:
                infunc = False
            if not infunc:
                continue
            m = re.search(r"//## File \"([^\"]+)\", line (\d+)", l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
                lines.append(cur)
        if lines:
            return lines
    return []


def main():
    rep, so, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    ls = out.splitlines()
    start = next(i for i, l in enumerate(ls) if l.startswith('"Address"'))
    rows = list(csv.reader(ls[start:]))
    hdr = rows[0]
    isamp, iinst, ithr = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    lmap = sass_lines(so, kernel)
    body = rows[1:]
    if len(lmap) != len(body):
        print(f"warning: {len(lmap)} disassembled instructions vs {len(body)} profiled rows; alignment may drift")
    agg = defaultdict(lambda: [0, 0, 0.0])
    tot = 0
    for i, r in enumerate(body):
        try:
            s = int(r[isamp] or 0)
            n = int(r[iinst] or 0)
            t = float(r[ithr] or 0)
        except ValueError:
            continue
        key = lmap[i] if i < len(lmap) else None
        agg[key][0] += s
        agg[key][1] += n
        agg[key][2] += n * t
        tot += s
    src = {}
    print(f"total samples {tot}")
    for key, (s, n, nt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
            if path not in src and os.path.exists(path):
                src[path] = open(path).read().splitlines()
            if path in src and key[1] - 1 < len(src[path]):
                text = src[path][key[1] - 1].strip()[:100]
        print(f"{100.0 * s / max(tot, 1):6.2f}%  inst={n:>11d} thr={nt / max(n, 1):5.1f}  {key}  {text}")


if __name__ == "__main__":
    main()
