"""GPU: the drop-in executables (dropin/) behave like the reference programs the timing harnesses spawn, and the
reference's UNMODIFIED harness (compiled from /root/reference into dropin/_bin by dropin/build.py) runs against them."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "dropin", "_bin")


def _tree(tmp_path):
    """The directory layout the reference programs assume: cwd = <root>/Experiment, data under ../Output-Input, ../Assets."""
    exp = tmp_path / "Experiment"
    for d in (exp / "results", tmp_path / "Output-Input" / "input", tmp_path / "Output-Input" / "out", tmp_path / "Output-Input" / "log",
              tmp_path / "Output-Input" / "Images", tmp_path / "Assets" / "Images"):
        d.mkdir(parents=True)
    return exp


def _need(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (python dropin/build.py; the JPEG programs and harnesses need /root/reference for stb and the harness sources)")
    return p


@pytest.mark.parametrize("exe", ["LZ4_seq.exe", "LZ4_par.exe"])
def test_lz4_exe_reproduces_reference_files(tmp_path, exe):
    """input.txt -> compressed.bin / compressed.txt / uncompressed.txt identical to the files committed in the reference repo."""
    path = _need(exe)
    exp = _tree(tmp_path)
    shutil.copy(os.path.join(cases.GOLDEN, "lz4_input.txt"), tmp_path / "Output-Input" / "input" / "input.txt")
    r = subprocess.run([path], cwd=exp, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout == "" and r.stderr == ""  # the harness echoes any output as an error
    out = tmp_path / "Output-Input" / "out"
    for got, want in (("compressed.bin", "lz4_compressed.bin"), ("compressed.txt", "lz4_compressed_hex.txt"),
                      ("uncompressed.txt", "lz4_uncompressed.txt")):
        assert (out / got).read_bytes() == open(os.path.join(cases.GOLDEN, want), "rb").read(), got


def test_lz4_exe_rejects_short_input_like_the_reference(tmp_path):
    path = _need("LZ4_seq.exe")
    exp = _tree(tmp_path)
    (tmp_path / "Output-Input" / "input" / "input.txt").write_bytes(b"too short")
    r = subprocess.run([path], cwd=exp, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "default block length is too high" in r.stdout  # LZ4.c:632-637


def test_jpeg_exe_writes_the_reference_pictures(tmp_path, oracle):
    from PIL import Image

    path = _need("JPEG_seq.exe")
    exp = _tree(tmp_path)
    img = cases.synth_image(42, 72, 52)  # last group row partial: exercises the unprocessed-groups behaviour
    Image.fromarray(img, "RGBA").save(tmp_path / "Assets" / "Images" / "rand_8X8.png")
    r = subprocess.run([path], cwd=exp, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout == "" and r.stderr == ""
    pics = tmp_path / "Output-Input" / "Images"
    got = {n: np.array(Image.open(pics / f"{n}.png").convert("RGBA")) for n in
           ("original", "luminance", "rChrominance", "bChrominance", "reconstructed")}
    assert np.array_equal(got["original"], img)
    Y, Cr, Cb = oracle.jpeg_planes(img)
    assert np.array_equal(got["luminance"][..., 0], Y) and np.array_equal(got["luminance"][..., 1], Y)
    coefs = oracle.jpeg_encode(img)["coefs"]
    assert np.array_equal(got["reconstructed"], oracle.jpeg_decode(coefs, 72, 52, img))


def test_unmodified_reference_harness_runs_against_the_dropin(tmp_path):
    """Experiment/LZ4_parallel_experiment.c, compiled unmodified, spawns LZ4_par.exe from its cwd (10 runs at one size)
    and writes results/LZ4_par.exe_execution_times.json in the reference's schema."""
    harness = _need("harness_LZ4_par")
    exp = _tree(tmp_path)
    shutil.copy(_need("LZ4_par.exe"), exp / "LZ4_par.exe")
    # the copy resolves its library through LD_LIBRARY_PATH (the rpath is relative to dropin/_bin)
    env = dict(os.environ, PATH=f"{exp}:{os.environ['PATH']}", LD_LIBRARY_PATH=os.path.join(ROOT, "lz4-jpeg_b200"))
    shutil.copy(os.path.join(cases.GOLDEN, "Metamorphosis.txt"), tmp_path / "Output-Input" / "input" / "Metamorphosis.txt")
    r = subprocess.run([harness], cwd=exp, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "[LZ4 ERROR]" not in r.stdout or "Number of cores" in r.stdout
    res = json.load(open(exp / "results" / "LZ4_par.exe_execution_times.json"))
    assert res[0]["exe_name"] == "LZ4_par.exe" and len(res[0]["execution_times_sec"]) == 10
    data = (tmp_path / "Output-Input" / "input" / "input.txt").read_bytes()
    comp = (tmp_path / "Output-Input" / "out" / "compressed.bin").read_bytes()
    assert comp[0] == (len(data) + 299) // 300 and len(comp) > 0
