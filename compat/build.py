#!/usr/bin/env python3
"""compat/build.py — compat/libljb_compat.so: the reference's own C entry points (include/ljb_compat.h) over
lz4-jpeg_b200/liblz4jpeg_b200.so.  Plain C; built in-tree (git-ignored, travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libljb_compat.so")
LIBDIR = os.path.join(ROOT, "lz4-jpeg_b200")
SRCS = [os.path.join(HERE, "ljb_compat.c")]
DEPS = SRCS + [os.path.join(ROOT, "include", "ljb_compat.h"), os.path.join(ROOT, "include", "lz4jpeg_b200.h"), os.path.abspath(__file__)]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    cmd = ["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", LIB, *SRCS, "-L" + LIBDIR,
           "-llz4jpeg_b200", "-Wl,-rpath,$ORIGIN/../lz4-jpeg_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("compat build failed")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return LIB


def build_test(force: bool = False) -> str:
    """compat/test_compat: a C program written the way a user of the reference's LZ4.c would write it."""
    exe = os.path.join(HERE, "test_compat")
    src = os.path.join(HERE, "test_compat.c")
    build(force)
    if force or not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe) or os.path.getmtime(LIB) > os.path.getmtime(exe):
        cmd = ["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", exe, src, "-L" + HERE, "-lljb_compat", "-L" + LIBDIR,
               "-llz4jpeg_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../lz4-jpeg_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("compat test build failed")
    return exe


def build_comm_test(force: bool = False) -> str:
    """compat/test_comm: a C host on N GPUs (include/ljb_comm.h, lz4-jpeg_b200/libljb_comm.so)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ljb_build", os.path.join(LIBDIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    comm = mod.build_comm(force)
    exe = os.path.join(HERE, "test_comm")
    src = os.path.join(HERE, "test_comm.c")
    if force or not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe) or os.path.getmtime(comm) > os.path.getmtime(exe):
        cmd = ["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", exe, src, "-L" + LIBDIR, "-lljb_comm", "-llz4jpeg_b200",
               "-Wl,-rpath,$ORIGIN/../lz4-jpeg_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("comm test build failed")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_test(force="--force" in sys.argv))
    print(build_comm_test(force="--force" in sys.argv))
