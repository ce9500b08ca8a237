"""tests/emu/build.py — TEST INFRASTRUCTURE: compiles the LZ4 kernel sources for the host under the lock-step CUDA
emulator (cuda_emu.h) into tests/emu/libemu_lz4.so.  g++ only; no GPU, no nvcc."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libemu_lz4.so")
SRCS = [os.path.join(HERE, "cuda_emu.cpp"), os.path.join(HERE, "emu_lz4.cpp")]
DEPS = SRCS + [os.path.join(HERE, "cuda_emu.h")] + [os.path.join(ROOT, "lz4-jpeg_b200", "csrc", f)
                                                     for f in ("lz4_encode.cu", "lz4_lazy.cuh", "lz4_decode.cu", "common.cuh")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS if os.path.exists(d)):
        return LIB
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-I", HERE, "-o", LIB, *SRCS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
