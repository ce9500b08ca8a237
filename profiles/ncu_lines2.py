#!/usr/bin/env python3
"""profiles/ncu_lines2.py — warp-stall samples and executed instructions per CUDA source line, straight from
ncu's own source/SASS correlation (needs -lineinfo at build time and --import-source on at capture time).
usage: python profiles/ncu_lines2.py rep.ncu-rep [topN]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    cur, hdr, out = None, None, []
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
        extra = len(r) - len(hdr)  # unescaped quotes/commas in the source text split it into several fields
        if extra > 0:
            r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
        elif extra < 0:
            continue
        g = lambda name: r[hdr.index(name)]
        samp, inst = int(g("# Samples") or 0), int(g("Instructions Executed") or 0)
        stalls = {h[6:]: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "0")}
        topst = ",".join(f"{k}={v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        out.append((samp, inst, g("Avg. Threads Executed"), cur, int(r[0]), r[1].strip()[:90], topst))
    tot = sum(o[0] for o in out) or 1
    toti = sum(o[1] for o in out) or 1
    print(f"total samples {tot}, warp instructions {toti}")
    for o in sorted(out, reverse=True)[:top]:
        print(f"{100*o[0]/tot:5.2f}% smp {100*o[1]/toti:5.2f}% inst thr={o[2]:>3} {o[3]}:{o[4]:<4} {o[5]}   [{o[6]}]")


if __name__ == "__main__":
    main()
