"""lz4-jpeg_b200/build.py — compile the CUDA kernels + C ABI into lz4-jpeg_b200/liblz4jpeg_b200.so.

sm_100a only (``-gencode arch=compute_100a,code=sm_100a``); nvcc cross-compiles without a GPU.
The .so is built in-tree (git-ignored) so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblz4jpeg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["api.cu", "lz4_encode.cu", "lz4_decode.cu", "jpeg_encode.cu", "jpeg_decode.cu", "jpeg_entropy.cu", "jfif_encode.cu"]
C_SOURCES = ["synth.c"]


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in CU_SOURCES + C_SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "lz4jpeg_b200.h"),
                                                                os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if src.endswith(".cu"):
            cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-c", src, "-o", obj]
        else:
            cmd = ["gcc", "-O2", "-fPIC", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError(f"compile failed: {src}")
        if verbose:
            sys.stderr.write(r.stderr)
        objs.append(obj)
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


COMM_LIB = os.path.join(HERE, "libljb_comm.so")


def build_comm(force: bool = False) -> str:
    """lz4-jpeg_b200/libljb_comm.so: the multi-GPU entry points for a C host (include/ljb_comm.h).  The only object that links
    NCCL (system libnccl, /usr/include/nccl.h); it sits next to liblz4jpeg_b200.so and resolves the single-GPU entry points there."""
    build(force)
    src = os.path.join(CSRC, "comm.cu")
    deps = [src, os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(HERE), "include", "ljb_comm.h"), LIB, os.path.abspath(__file__)]
    if not force and os.path.exists(COMM_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(COMM_LIB) for d in deps):
        return COMM_LIB
    cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", COMM_LIB, src, "-L" + HERE, "-llz4jpeg_b200",
           "-lnccl", "-lcudart", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("comm build failed")
    return COMM_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_comm(force="--force" in sys.argv))
