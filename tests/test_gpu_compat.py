"""The reference's own C entry points (include/ljb_compat.h, compat/libljb_compat.so) driven by a C program written the way the
reference's lz4_encode() uses them (compat/test_compat.c): divide_input -> block_encode -> write_output."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compat"))
ALL = {name: (data, bl) for name, data, bl in cases.lz4_cases()}


@pytest.fixture(scope="module")
def exe():
    import build as compat_build  # compat/build.py

    return compat_build.build_test()


@pytest.mark.parametrize("mode", ["blocks", "par"])
@pytest.mark.parametrize("name", ["golden_input", "extract_30000", "wrap_257", "wrap_513", "lit_271", "lit_526", "repeats_ge1024_b4096",
                                  "zero_containing", "synth_64k_x3"])
def test_block_encode_write_output(exe, oracle, tmp_path, name, mode):
    """The frame written by write_output() from the LZ4Frame that block_encode() filled == the oracle's stream."""
    data, bl = ALL[name]
    inp = tmp_path / "in.bin"
    out = tmp_path / "out.bin"
    data.tofile(inp)
    r = subprocess.run([exe, mode, str(inp), str(bl), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    s, _, _ = oracle.lz4_compress(data, bl, 1)
    got = np.fromfile(out, dtype=np.uint8)
    assert np.array_equal(got, s), (name, mode, r.stdout)


@pytest.mark.parametrize("name", ["golden_input", "long_runs_b3000", "repeats_ge1024_b4096"])
def test_find_longest_match(exe, tmp_path, name):
    data, bl = ALL[name]
    inp = tmp_path / "in.bin"
    data.tofile(inp)
    r = subprocess.run([exe, "match", str(inp), str(bl)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout


def test_file_contract(exe, tmp_path):
    """lz4_encode() + LZ4_decode() on the reference's fixed relative paths reproduce its committed files."""
    run = tmp_path / "Experiment"
    for d in ("Experiment", "Output-Input/input", "Output-Input/out", "Output-Input/log"):
        (tmp_path / d).mkdir(parents=True, exist_ok=True)
    shutil.copy(os.path.join(cases.GOLDEN, "lz4_input.txt"), tmp_path / "Output-Input/input/input.txt")
    r = subprocess.run([exe, "files"], cwd=run, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "", r.stdout + r.stderr
    for got, want in (("Output-Input/out/compressed.bin", "lz4_compressed.bin"), ("Output-Input/out/compressed.txt", "lz4_compressed_hex.txt"),
                      ("Output-Input/out/uncompressed.txt", "lz4_uncompressed.txt")):
        assert (tmp_path / got).read_bytes() == open(os.path.join(cases.GOLDEN, want), "rb").read(), got
