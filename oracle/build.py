#!/usr/bin/env python3
"""oracle/build.py — build the CPU checkers (TEST INFRASTRUCTURE, not product code).

* ``oracle/liboracle.so``      : this repo's C restatement (lz4_oracle.c + jpeg_oracle.c + jfif_oracle.c).
* ``oracle/_ref/libref_*.so``  : the REFERENCE's own C sources, compiled from where they lie under
  /root/reference (never copied into the repo), with ref_glue_*.c as a buffer-level driver.
  Built only where /root/reference exists (this container); the GPU box uses the prebuilt files
  that travel with the snapshot (oracle/_ref/ is git-ignored but not gpurun-ignored).

Variants of the reference LZ4 build:
  libref_lz4_verbatim.so  unmodified source — golden-vector test only (its match extension reads
                          past the block: undefined behaviour).
  libref_lz4.so           "bounded" parity target: three textual edits applied to a temporary copy
                          (deleted after compilation) that stop the extension at the block end.

The reference's vendored baseline-JPEG writer (stb_image_write.h, never called by the reference programs):
  libref_jfif.so          stbi_write_jpg_to_func, compiled from a temporary copy of the header in which the
                          chroma-subsampling rule (`quality <= 90`) reads an override first (default: stb's rule).

Flags follow SURVEY.md Appendix E: gcc -O2 -ffp-contract=off, x86-64 SSE2, no -march=native.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LJB_REFERENCE_DIR", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
CFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w"]


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("oracle build failed")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.exists(s) and os.path.getmtime(s) <= t for s in sources)


def build_restatement(force: bool = False) -> str:
    out = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, "lz4_oracle.c"), os.path.join(HERE, "jpeg_oracle.c"), os.path.join(HERE, "jfif_oracle.c")]
    if force or not _newer(out, srcs + [os.path.abspath(__file__)]):
        _run(["gcc", *CFLAGS, "-Wall", "-o", out, *srcs, "-lm"])
    return out


def _patch_bounded(src_text: str) -> str:
    """The three edits of SURVEY.md Appendix E(b); each anchor must occur exactly once."""
    a1 = "uint8_t find_longest_match(uint8_t *input, size_t current_index, uint16_t *match_distance)"
    a2 = "while (current_match_length < MAX_MATCH_LENGTH &&"
    a3 = "void block_encode(const char *block_entry, size_t block_length, LZ4Block *block, FILE *log_file, FILE *output_file, LZ4Frame *frame)\n{"
    for a in (a1, a2, a3):
        if src_text.count(a) != 1:
            raise RuntimeError(f"reference LZ4.c changed: anchor not unique: {a!r}")
    src_text = src_text.replace(a1, "static __thread size_t g_block_length;\n" + a1)
    src_text = src_text.replace(a2, a2 + " current_index + current_match_length < g_block_length &&")
    src_text = src_text.replace(a3, a3 + "\n    g_block_length = block_length;")
    return src_text


def _patch_subsample(src_text: str) -> str:
    """One edit: let the caller force 4:4:4 or 4:2:0 (BASELINE.json's "quality 75, 4:4:4"); -1 keeps stb's rule."""
    a = "subsample = quality <= 90 ? 1 : 0;"
    if src_text.count(a) != 1:
        raise RuntimeError(f"reference stb_image_write.h changed: anchor not unique: {a!r}")
    return src_text.replace(a, "subsample = ljb_force_subsample >= 0 ? (ljb_force_subsample != 0) : (quality <= 90 ? 1 : 0);")


def build_reference(force: bool = False) -> dict[str, str]:
    """Compile the reference where it lies; returns {name: path} of what exists afterwards."""
    names = {
        "lz4": os.path.join(REF_OUT, "libref_lz4.so"),
        "lz4_verbatim": os.path.join(REF_OUT, "libref_lz4_verbatim.so"),
        "jpeg": os.path.join(REF_OUT, "libref_jpeg.so"),
        "jfif": os.path.join(REF_OUT, "libref_jfif.so"),
        "lz4_par": os.path.join(REF_OUT, "libref_lz4_par.so"),    # the parallel builds (windows.h shim), CPU baselines only
        "jpeg_par": os.path.join(REF_OUT, "libref_jpeg_par.so"),
    }
    lz4_src = os.path.join(REF, "Algorithms/sequential/LZ4/LZ4.c")
    jpg_dir = os.path.join(REF, "Algorithms/sequential/JPEG")
    jpg_src = os.path.join(jpg_dir, "JPEG.c")
    if not (os.path.exists(lz4_src) and os.path.exists(jpg_src)):
        return {k: v for k, v in names.items() if os.path.exists(v)}
    os.makedirs(REF_OUT, exist_ok=True)
    glue_lz4 = os.path.join(HERE, "ref_glue_lz4.c")
    glue_jpg = os.path.join(HERE, "ref_glue_jpeg.c")
    shim = os.path.join(HERE, "shim")
    me = os.path.abspath(__file__)
    if force or not _newer(names["lz4_verbatim"], [glue_lz4, lz4_src, me]):
        _run(["gcc", *CFLAGS, f"-I{shim}", f'-DREF_SRC="{lz4_src}"', "-o", names["lz4_verbatim"], glue_lz4, "-lpthread"])
    if force or not _newer(names["lz4"], [glue_lz4, lz4_src, me]):
        tmp = tempfile.mkdtemp(prefix="ljb_ref_")
        try:
            patched = os.path.join(tmp, "LZ4_bounded.c")
            with open(lz4_src, "r", encoding="latin-1") as f:
                text = f.read()
            with open(patched, "w", encoding="latin-1") as f:
                f.write(_patch_bounded(text))
            _run(["gcc", *CFLAGS, f"-I{shim}", f'-DREF_SRC="{patched}"', "-o", names["lz4"], glue_lz4, "-lpthread"])
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if force or not _newer(names["jpeg"], [glue_jpg, jpg_src, me]):
        _run(["gcc", *CFLAGS, f"-I{jpg_dir}", f'-DREF_SRC="{jpg_src}"', "-o", names["jpeg"], glue_jpg, "-lm", "-lpthread"])
    stbw = os.path.join(jpg_dir, "stb_image_write.h")
    glue_jfif = os.path.join(HERE, "ref_glue_jfif.c")
    if os.path.exists(stbw) and (force or not _newer(names["jfif"], [glue_jfif, stbw, me])):
        tmp = tempfile.mkdtemp(prefix="ljb_ref_")
        try:
            patched = os.path.join(tmp, "stbw_subsample.h")
            with open(stbw, "r", encoding="latin-1") as f:
                text = f.read()
            with open(patched, "w", encoding="latin-1") as f:
                f.write(_patch_subsample(text))
            _run(["gcc", *CFLAGS, f'-DREF_STBW="{patched}"', "-o", names["jfif"], glue_jfif, "-lm", "-lpthread"])
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    # ---- the parallel builds (Algorithms/parallel/*), against oracle/shim/windows.h: bench.py's "parallel build" CPU baselines
    plz4_src = os.path.join(REF, "Algorithms/parallel/LZ4/LZ4.c")
    glue_plz4 = os.path.join(HERE, "ref_glue_lz4_par.c")
    shim_w = os.path.join(shim, "windows.h")
    if os.path.exists(plz4_src) and (force or not _newer(names["lz4_par"], [glue_plz4, plz4_src, shim_w, me])):
        tmp = tempfile.mkdtemp(prefix="ljb_ref_")
        try:
            patched = os.path.join(tmp, "LZ4_par_bounded.c")
            with open(plz4_src, "r", encoding="latin-1") as f:
                text = f.read()
            with open(patched, "w", encoding="latin-1") as f:
                f.write(_patch_bounded_parallel(text))
            _run(["gcc", *CFLAGS, "-Wl,-Bsymbolic", f"-I{shim}", f'-DREF_SRC="{patched}"', "-o", names["lz4_par"], glue_plz4, "-lpthread"])
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    pjpg_dir = os.path.join(REF, "Algorithms/parallel/JPEG")
    pjpg_src = os.path.join(pjpg_dir, "JPEG.c")
    glue_pjpg = os.path.join(HERE, "ref_glue_jpeg_par.c")
    if os.path.exists(pjpg_src) and (force or not _newer(names["jpeg_par"], [glue_pjpg, pjpg_src, shim_w, me])):
        _run(["gcc", *CFLAGS, "-Wl,-Bsymbolic", f"-I{shim}", f"-I{pjpg_dir}", f'-DREF_SRC="{pjpg_src}"', "-o", names["jpeg_par"], glue_pjpg,
              "-lm", "-lpthread"])
    return {k: v for k, v in names.items() if os.path.exists(v)}


def _patch_bounded_parallel(src_text: str) -> str:
    """The same three edits as _patch_bounded, on the parallel source's spelling of the anchors."""
    a1 = "uint8_t find_longest_match(uint8_t* input, size_t current_index, uint16_t* match_distance) {"
    a2 = "while (current_match_length < MAX_MATCH_LENGTH &&"
    a3 = "    size_t block_length = args->block_length;\n    LZ4Block* block = args->block;"
    for a in (a1, a2, a3):
        if src_text.count(a) != 1:
            raise RuntimeError(f"reference parallel LZ4.c changed: anchor not unique: {a!r}")
    src_text = src_text.replace(a1, "static __thread size_t g_block_length;\n" + a1)
    src_text = src_text.replace(a2, a2 + " current_index + current_match_length < g_block_length &&")
    src_text = src_text.replace(a3, a3 + "\n    g_block_length = block_length;")
    return src_text


def build_all(force: bool = False) -> dict[str, str]:
    libs = {"oracle": build_restatement(force)}
    libs.update({"ref_" + k: v for k, v in build_reference(force).items()})
    return libs


if __name__ == "__main__":
    for k, v in build_all(force="--force" in sys.argv).items():
        print(f"{k:18s} {v}")
