#!/usr/bin/env python3
"""profiles/sass_static.py — static SASS instruction count of one kernel per source function (needs -lineinfo).
usage: python profiles/sass_static.py lib.so kernel_substring src.cu"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, kernel, src = sys.argv[1], sys.argv[2], sys.argv[3]
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|extern)\b.*?\b(\w+)\s*\(")
    regs = [(i + 1, pat.match(l).group(1)) for i, l in enumerate(open(src)) if pat.match(l) and not l.rstrip().endswith(";")]
    base = os.path.basename(src)
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cnt = collections.Counter()
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        infunc, cur = False, "?"
        for l in out.splitlines():
            if l.startswith(".text."):
                infunc = kernel in l
                continue
            if not infunc:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                fn, ln = os.path.basename(m.group(1)), int(m.group(2))
                cur = fn
                if fn == base:
                    cur = "?"
                    for first, name in regs:
                        if first <= ln:
                            cur = name
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
                cnt[cur] += 1
    tot = sum(cnt.values())
    print(f"{kernel}: {tot} SASS instructions ({tot * 16 / 1024:.0f} KB)")
    for k, v in cnt.most_common():
        print(f"  {k:28s} {v:6d}")


if __name__ == "__main__":
    main()
