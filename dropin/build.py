#!/usr/bin/env python3
"""dropin/build.py — build the four drop-in executables the reference's timing harnesses spawn
(LZ4_seq.exe, LZ4_par.exe, JPEG_seq.exe, JPEG_par.exe) and, where /root/reference is mounted, the UNMODIFIED
harnesses themselves (Experiment/*_experiment.c) so that tests can show them working against the drop-ins.

Everything lands in dropin/_bin/ (git-ignored; it travels to the GPU box with the repo snapshot).  The JPEG programs
need the stb single-header PNG library, which the reference vendors (third-party code): it is compiled from where
it lies under /root/reference and never copied into this repository, so JPEG_*.exe and the harness binaries can
only be (re)built where the reference is mounted.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "_bin")
REF = os.environ.get("LJB_REFERENCE_DIR", "/root/reference")
LIBDIR = os.path.join(ROOT, "lz4-jpeg_b200")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("drop-in build failed")


def _stale(target, sources):
    return not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources if os.path.exists(s))


def build_all(force: bool = False) -> dict[str, str]:
    os.makedirs(BIN, exist_ok=True)
    link = ["-I" + os.path.join(ROOT, "include"), "-L" + LIBDIR, "-llz4jpeg_b200", "-Wl,-rpath,$ORIGIN/../../lz4-jpeg_b200", "-lm"]
    out = {}
    lz4_src = os.path.join(HERE, "lz4_main.c")
    hdr = os.path.join(ROOT, "include", "lz4jpeg_b200.h")
    for name in ("LZ4_seq.exe", "LZ4_par.exe"):  # one program: the GPU path is always block-parallel
        exe = os.path.join(BIN, name)
        if force or _stale(exe, [lz4_src, hdr]):
            _run(["gcc", "-O2", "-o", exe, lz4_src, *link])
        out[name] = exe
    stb_dir = os.path.join(REF, "Algorithms", "sequential", "JPEG")
    jpg_src = os.path.join(HERE, "jpeg_main.c")
    for name in ("JPEG_seq.exe", "JPEG_par.exe"):
        exe = os.path.join(BIN, name)
        if os.path.exists(os.path.join(stb_dir, "stb_image.h")) and (force or _stale(exe, [jpg_src, hdr])):
            _run(["gcc", "-O2", "-ffp-contract=off", "-w", "-I" + stb_dir, "-o", exe, jpg_src, *link])
        if os.path.exists(exe):
            out[name] = exe
    exp = os.path.join(REF, "Experiment")
    for src, name in (("LZ4_sequential_experiment.c", "harness_LZ4_seq"), ("LZ4_parallel_experiment.c", "harness_LZ4_par"),
                      ("JPEG_sequential_experiment.c", "harness_JPEG_seq"), ("JPEG_parallel_experiment.c", "harness_JPEG_par")):
        exe = os.path.join(BIN, name)
        s = os.path.join(exp, src)
        if os.path.exists(s) and (force or _stale(exe, [s])):
            _run(["gcc", "-O2", "-w", "-I" + exp, "-o", exe, s, "-lm"])  # the reference's own harness, unmodified
        if os.path.exists(exe):
            out[name] = exe
    return out


if __name__ == "__main__":
    for k, v in build_all(force="--force" in sys.argv).items():
        print(f"{k:18s} {v}")
