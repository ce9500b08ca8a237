#!/usr/bin/env python3
"""profiles/ncu_source.py — hottest SASS instructions (by stall samples) of an .ncu-rep, grouped by CUDA source line.
usage: python profiles/ncu_source.py rep.ncu-rep [topN]  (needs -lineinfo and --import-source on at capture time)"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(lines[start:]))
    hdr = rows[0]
    ia, isrc, isamp, iinst, ithr = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
    iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    data = []
    tot = 0
    for r in rows[1:]:
        try:
            s = int(r[isamp] or 0)
        except ValueError:
            continue
        tot += s
        data.append((s, r[ia], r[isrc], r[iinst], r[ithr], r[iw] if iw is not None else ""))
    print(f"total samples {tot}")
    for s, a, src, inst, thr, wf in sorted(data, reverse=True)[:top]:
        print(f"{100.0 * s / max(tot, 1):6.2f}%  {s:7d}  inst={inst:>10s} thr={thr:>5s} wf={wf:>10s}  {src[:110]}")


if __name__ == "__main__":
    main()
