"""GPU parity: the CUDA baseline-JPEG encoder called through the C ABI, checked byte for byte against the committed
hashes of the reference's stb_image_write.h build and against the oracle."""
import ctypes as C
import hashlib
import io
import json
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

VEC = json.load(open(os.path.join(cases.GOLDEN, "jfif_ref_vectors.json")))
ALL = {name: (px, q, sub) for name, px, q, sub in cases.jfif_cases()}


@pytest.fixture(scope="module")
def ljb():
    import lz4jpeg_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(ljb):
    c = ljb.Context(0)
    yield c
    c.close()


def _first_diff(a, b):
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return (int(d[0]) if d.size else n, a.size, b.size)


@pytest.mark.parametrize("name", sorted(ALL))
def test_file_matches_the_stb_build(ljb, ctx, oracle, name):
    px, q, sub = ALL[name]
    jpg = ljb.jfif.write_jpg(px, q, sub, ctx=ctx)
    if hashlib.sha256(jpg.tobytes()).hexdigest() != VEC[name]["sha256"]:
        want = oracle.jfif_encode(px, q, sub)
        pytest.fail(f"first difference (byte, got size, want size): {_first_diff(jpg, want)}")
    assert jpg.size == VEC[name]["size"]


def test_committed_file(ljb, ctx):
    want = np.fromfile(os.path.join(cases.GOLDEN, "og_crop_q75.jpg"), dtype=np.uint8)
    assert np.array_equal(ljb.jfif.write_jpg(cases.og_crop(), 75, ctx=ctx), want)


def _device_encode(ljb, ctx, px, q, sub, want_coefs=True, cap=None):
    import torch

    a = np.ascontiguousarray(px)
    h, w = a.shape[:2]
    comp = 1 if a.ndim == 2 else a.shape[2]
    units = ((w + 15) // 16) * ((h + 15) // 16) * 6 if (sub == 1 or (sub < 0 and (q or 90) <= 90)) else ((w + 7) // 8) * ((h + 7) // 8) * 3
    d_px = torch.from_numpy(a).cuda()
    cap = cap if cap is not None else 607 + 2 + 4 * w * h + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(3, dtype=torch.int64, device="cuda")
    d_coefs = torch.zeros(units * 64, dtype=torch.int16, device="cuda") if want_coefs else None
    torch.cuda.synchronize()
    ljb.jfif.encode_device(d_px, w, h, comp, q, sub, d_out, d_res, ctx, d_coefs=d_coefs)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream)
    stream.synchronize()
    res = d_res.cpu().numpy()
    n = int(res[0])
    return d_out[: min(n, cap)].cpu().numpy(), res, (d_coefs.cpu().numpy().reshape(units, 64) if want_coefs else None)


@pytest.mark.parametrize("sub", [-1, 0])
def test_noise_1024_coefficients_and_file_vs_oracle(ljb, ctx, oracle, sub):
    """1024x1024 uniform noise at quality 75 (the benchmark distribution), both chroma layouts: every quantised
    coefficient and every byte of the file equal the oracle's (tolerance 0)."""
    img = ljb.synth.random_image(1024, 1024, seed=42)
    want, want_coefs = oracle.jfif_encode(img, 75, sub, want_coefs=True)
    got, res, coefs = _device_encode(ljb, ctx, img, 75, sub)
    assert int(res[2]) == 0
    mism = int((coefs != want_coefs).sum())
    assert mism == 0, f"{mism} of {want_coefs.size} coefficients differ"
    assert np.array_equal(got, want), _first_diff(got, want)


@pytest.mark.parametrize("quality,sub", [(100, -1), (100, 1), (1, -1), (50, 0), (90, -1)])
def test_noise_512_other_qualities(ljb, ctx, oracle, quality, sub):
    """quality 100 on noise overflows the per-warp bit buffer inside a tile (early flush path)."""
    img = ljb.synth.random_image(512, 384, seed=7)
    assert np.array_equal(ljb.jfif.write_jpg(img, quality, sub, ctx=ctx), oracle.jfif_encode(img, quality, sub))


def test_spilled_tiles(ljb, ctx, oracle, monkeypatch):
    """Eight rounds per tile at quality 75 (the library would choose three): every tile overflows the 2.5 KB bit buffer
    several times and goes through the spill area in global memory; the file does not change."""
    monkeypatch.setenv("LJB_JFIF_ROUNDS", "8")
    img = ljb.synth.random_image(640, 480, seed=21)
    for q, sub in ((75, -1), (75, 0), (100, 1)):
        assert np.array_equal(ljb.jfif.write_jpg(img, q, sub, ctx=ctx), oracle.jfif_encode(img, q, sub)), (q, sub)


@pytest.mark.parametrize("chunk", [1, 20000, 300000])
def test_banded_upload(ljb, ctx, oracle, monkeypatch, chunk):
    """The host-buffer entry point uploads the rows in bands and encodes the tiles of a band as soon as it has arrived; a
    small LJB_PIPE_CHUNK_BYTES drives up to twenty-four bands (16-row bands, bands that hold no complete tile, ragged last
    bands) through small images; the file does not change."""
    monkeypatch.setenv("LJB_PIPE_CHUNK_BYTES", str(chunk))
    rng = np.random.default_rng(5)
    for (w, h, comp) in ((640, 480, 4), (333, 257, 3), (1000, 50, 4), (48, 700, 1), (16, 16, 4)):
        px = rng.integers(0, 256, size=(h, w, comp), dtype=np.uint8)
        for q, sub in ((75, -1), (75, 0), (92, 1)):
            got, want = ljb.jfif.write_jpg(px, q, sub, ctx=ctx), oracle.jfif_encode(px, q, sub)
            assert np.array_equal(got, want), (w, h, comp, q, sub, _first_diff(got, want))
    crop = cases.og_crop()
    img = np.tile(crop, (6, 3, 1))
    assert np.array_equal(ljb.jfif.write_jpg(img, 75, -1, ctx=ctx), oracle.jfif_encode(img, 75, -1))


def test_natural_image_tiles(ljb, ctx, oracle):
    """The og.png crop tiled to 2048 x 1500: smooth content, most coefficients zero, long runs, many tiles."""
    crop = cases.og_crop()
    img = np.tile(crop, (15, 8, 1))
    for q, sub in ((75, -1), (75, 0), (95, -1)):
        assert np.array_equal(ljb.jfif.write_jpg(img, q, sub, ctx=ctx), oracle.jfif_encode(img, q, sub)), (q, sub)


def test_random_shapes_and_components(ljb, ctx, oracle):
    rng = np.random.default_rng(99)
    for i in range(40):
        h, w, comp = int(rng.integers(1, 200)), int(rng.integers(1, 200)), int(rng.integers(1, 5))
        px = rng.integers(0, 256, size=(h, w, comp), dtype=np.uint8)
        if i % 3 == 0:
            px = (px // 64 * 64).astype(np.uint8)
        q, sub = int(rng.integers(0, 101)), int(rng.integers(-1, 2))
        got = ljb.jfif.write_jpg(px, q, sub, ctx=ctx)
        want = oracle.jfif_encode(px, q, sub)
        assert np.array_equal(got, want), (h, w, comp, q, sub, _first_diff(got, want))


def test_row_stride_and_unaligned_base(ljb, ctx, oracle):
    """Rows at a pitch larger than w*comp, and a base pointer that is not 16-byte aligned (generic load path)."""
    import torch

    img = ljb.synth.random_image(100, 60, seed=3)
    want = oracle.jfif_encode(img, 75, -1)
    pitch = 100 * 4 + 48
    host = np.zeros((60, pitch), np.uint8)
    host[:, :400] = img.reshape(60, 400)
    for shift in (0, 4):
        flat = torch.zeros(60 * pitch + 64, dtype=torch.uint8, device="cuda")
        flat[shift:shift + 60 * pitch] = torch.from_numpy(host.reshape(-1)).cuda()
        d_px = flat[shift:]
        d_out = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
        d_res = torch.zeros(3, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ljb.jfif.encode_device(d_px, 100, 60, 4, 75, -1, d_out, d_res, ctx, stride=pitch)
        torch.cuda.ExternalStream(ctx.stream).synchronize()
        n = int(d_res.cpu()[0])
        assert np.array_equal(d_out[:n].cpu().numpy(), want), shift


def test_capacity_error(ljb, ctx):
    img = ljb.synth.random_image(256, 256, seed=1)
    N = ljb._native
    out = np.empty(4096, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = N.lib().ljb_jfif_encode(ctx.handle, img.ctypes.data, 256, 256, 4, 1024, 75, -1, out.ctypes.data, out.size, C.byref(n))
    assert rc == N.E_CAPACITY and n.value > 4096
    rc = N.lib().ljb_jfif_encode(ctx.handle, img.ctypes.data, 256, 256, 4, 1024, 75, -1, out.ctypes.data, 100, C.byref(n))
    assert rc == N.E_CAPACITY
    full = ljb.jfif.write_jpg(img, 75, ctx=ctx)
    out = np.empty(full.size, dtype=np.uint8)  # exactly enough
    rc = N.lib().ljb_jfif_encode(ctx.handle, img.ctypes.data, 256, 256, 4, 1024, 75, -1, out.ctypes.data, out.size, C.byref(n))
    assert rc == N.OK and n.value == full.size and np.array_equal(out, full)


def test_argument_errors(ljb, ctx):
    N = ljb._native
    img = np.zeros((8, 8, 4), np.uint8)
    out = np.empty(4096, dtype=np.uint8)
    n = C.c_size_t(0)
    f = N.lib().ljb_jfif_encode
    assert f(ctx.handle, img.ctypes.data, 0, 8, 4, 32, 75, -1, out.ctypes.data, out.size, C.byref(n)) == N.E_ARG
    assert f(ctx.handle, img.ctypes.data, 8, 8, 5, 40, 75, -1, out.ctypes.data, out.size, C.byref(n)) == N.E_ARG
    assert f(ctx.handle, img.ctypes.data, 8, 8, 4, 16, 75, -1, out.ctypes.data, out.size, C.byref(n)) == N.E_ARG  # stride < w*comp
    assert f(ctx.handle, None, 8, 8, 4, 32, 75, -1, out.ctypes.data, out.size, C.byref(n)) == N.E_ARG
    assert f(None, img.ctypes.data, 8, 8, 4, 32, 75, -1, out.ctypes.data, out.size, C.byref(n)) == N.E_ARG


def test_pillow_decodes_gpu_output(ljb, ctx):
    from PIL import Image

    crop = cases.og_crop()
    jpg = ljb.jfif.write_jpg(crop, 85, ctx=ctx)
    im = Image.open(io.BytesIO(jpg.tobytes()))
    im.load()
    assert im.size == (256, 100)
    mse = ((np.asarray(im, dtype=np.float64) - crop[:, :, :3]) ** 2).mean()
    assert 10 * np.log10(255.0 ** 2 / mse) > 30.0


def test_large_image_property(ljb, ctx, oracle):
    """4096 x 4096 noise (64 Ki MCUs, 13 k rounds): header, EOI, stuffing invariant, and equality with the oracle."""
    img = ljb.synth.random_image(4096, 4096, seed=11)
    jpg = ljb.jfif.write_jpg(img, 75, ctx=ctx)
    assert jpg[0] == 0xFF and jpg[1] == 0xD8 and jpg[-2] == 0xFF and jpg[-1] == 0xD9
    body = jpg[607:-2]
    ff = np.nonzero(body[:-1] == 0xFF)[0]
    assert (body[ff + 1] == 0).all() and body[-1] != 0xFF  # every 0xFF of the entropy-coded segment is stuffed
    assert np.array_equal(jpg, oracle.jfif_encode(img, 75, -1))
