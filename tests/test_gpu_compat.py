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


def _compat_build():
    import importlib.util

    spec = importlib.util.spec_from_file_location("ljb_compat_build", os.path.join(ROOT, "compat", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

ALL = {name: (data, bl) for name, data, bl in cases.lz4_cases()}


@pytest.fixture(scope="module")
def exe():
    return _compat_build().build_test()


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


def test_process_entry_point(exe, oracle, tmp_path):
    """process() (Algorithms/parallel/JPEG/JPEG.c:1103) from C: quantised coefficients == the oracle's for the same groups, samples and
    dequantised coefficients == what the Python mirror (the same CUDA path) returns."""
    import lz4jpeg_b200 as ljb

    img = cases.synth_image(11, 64, 32)
    ng = 32
    smp = np.zeros((ng, 128), np.uint8)
    for g in range(ng):
        st = oracle.jpeg_group_stages(img, g)[0]  # oracle order: lum | r | b
        smp[g, :64], smp[g, 64:96], smp[g, 96:] = st[:64], st[96:], st[64:96]  # PixelGroup order: lum | b | r
    inp, out = tmp_path / "smp.bin", tmp_path / "out.bin"
    smp.tofile(inp)
    r = subprocess.run([exe, "process", str(inp), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, dtype=np.uint8).reshape(ng, 128 + 128 * 8)
    rec_c = raw[:, :128]
    deq_c = raw[:, 128:].copy().view(np.float64).reshape(ng, 128)
    rec, coefs = ljb.jpeg.process_groups(smp)
    ref = oracle.jpeg_encode(img, 0, ng)["coefs"]
    assert np.array_equal(coefs, ref)
    assert np.array_equal(rec_c, rec)
    ql = np.array([8, 6, 6, 8, 10, 14, 18, 22, 6, 6, 7, 9, 12, 20, 22, 20, 6, 7, 8, 10, 14, 22, 25, 22, 8, 9, 10, 14, 18, 28, 27, 22, 10, 12, 14,
                   18, 22, 35, 33, 26, 14, 18, 22, 22, 27, 33, 36, 30, 18, 22, 26, 28, 33, 40, 40, 34, 22, 26, 28, 30, 36, 34, 35, 33], np.float64)
    qc = np.array([17, 18, 24, 47, 18, 21, 26, 66, 24, 26, 56, 99, 47, 66, 99, 99] + [66] + [99] * 15, np.float64)  # JPEG.c:12-27
    deq = coefs.astype(np.float64) * np.concatenate([ql, qc, qc])
    assert np.array_equal(deq_c, deq)
    # the reconstructed samples, put through assemble_image's colour formula (JPEG.c:552-619), are the oracle's reconstructed pixels
    want = oracle.jpeg_decode(ref, 64, 32)
    got = np.zeros((32, 64, 3), np.int64)
    for g in range(ng):
        br, bc = divmod(g, 8)
        lum = rec[g, :64].reshape(8, 8).astype(np.int64)
        cb = np.repeat(rec[g, 64:96].reshape(8, 4).astype(np.int64) - 128, 2, axis=1)
        cr = np.repeat(rec[g, 96:].reshape(8, 4).astype(np.int64) - 128, 2, axis=1)
        tr = lambda a: np.trunc(a).astype(np.int64)
        blk = np.stack([lum + tr(1.402 * cr), lum - tr(0.344136 * cb) - tr(0.714136 * cr), lum + tr(1.772 * cb)], axis=-1)
        got[br * 8:br * 8 + 8, bc * 8:bc * 8 + 8] = np.clip(blk, 0, 255)
    assert np.array_equal(got.astype(np.uint8), want[:, :, :3])


@pytest.mark.parametrize("ngpus", [1, 2])
def test_c_host_on_n_gpus(tmp_path, ngpus):
    """include/ljb_comm.h from a C program: one process, N GPUs, one NCCL all-gather of shard totals; the stream and the offset
    tables equal the single-GPU ones (compat/test_comm.c compares them itself)."""
    import torch

    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    exe = _compat_build().build_comm_test()
    inp = tmp_path / "in.txt"
    cases.synth_text(23 * 65536 + 999, seed=5).tofile(inp)
    r = subprocess.run([exe, str(ngpus), str(inp), "65536", "640", "328"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("identical to 1 GPU") == 3, r.stdout
