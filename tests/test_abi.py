"""CPU-only checks of the drop-in boundary: the shared object builds, loads, and exports every symbol that
include/lz4jpeg_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lz4jpeg_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g

    g.build()
    p = os.path.join(ROOT, "lz4-jpeg_b200", "liblz4jpeg_b200.so")
    assert os.path.exists(p)
    return p


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ljb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("ljb_ctx_create", "ljb_lz4_compress", "ljb_lz4_compress_dev", "ljb_lz4_decompress", "ljb_lz4_block_matches",
                 "ljb_jpeg_encode_rgba", "ljb_jpeg_encode_rgba_dev", "ljb_synth_text", "ljb_synth_image"):
        assert must in names


def test_shared_object_exports_every_declared_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(ljb_[a-z0-9_]+)", out))
    missing = [n for n in _declared() if n not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    L = ctypes.CDLL(lib_path)
    for n in _declared():
        assert getattr(L, n) is not None


def test_python_binding_covers_every_symbol(lib_path):
    import lz4jpeg_b200 as ljb

    assert sorted(ljb._native.SIGNATURES) == _declared()
    ljb._native.lib()  # resolves all of them; raises on a missing one


def test_no_gpu_means_loud_failure(lib_path):
    """Without a usable B200 the library must fail, not fall back to a CPU path."""
    import torch

    import lz4jpeg_b200 as ljb

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ljb.LjbError):
        ljb.Context(0)


def test_host_side_helpers(lib_path):
    import lz4jpeg_b200 as ljb

    L = ljb._native.lib()
    assert L.ljb_lz4_block_count(350, 300) == 2
    assert L.ljb_lz4_block_count(65536, 65536) == 1
    assert L.ljb_lz4_bound(100, 50) >= 100 + 7
    assert L.ljb_jpeg_group_count(1200, 630) == 11813  # ceil(w*h/64), JPEG.c:1131
    assert L.ljb_jpeg_group_count(16, 8) == 2
    assert ljb.lz4.divide_input(350, 300) == [(0, 300), (300, 50)]
    assert L.ljb_strerror(-3).decode() == "output buffer too small"


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under lz4-jpeg_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lz4-jpeg_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle_" not in text, os.path.join(dp, f)
