"""GPU parity: the CUDA JPEG-like encoder called through the C ABI, checked by the oracle / reference vectors."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

VEC = np.load(os.path.join(cases.GOLDEN, "jpeg_ref_vectors.npz"))
ALL = dict(cases.jpeg_cases())


@pytest.fixture(scope="module")
def ljb():
    import lz4jpeg_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(ljb):
    c = ljb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", sorted(ALL))
def test_matches_reference_build_vectors(ljb, ctx, name):
    """Quantised coefficients, bit lengths, offsets and packed bit strings == the reference build's (committed)."""
    enc = ljb.jpeg.process(ALL[name], ctx=ctx)
    assert np.array_equal(enc.coefs, VEC[f"{name}__coefs"])
    assert np.array_equal(enc.group_bits, VEC[f"{name}__bits"])
    assert np.array_equal(enc.group_offsets, VEC[f"{name}__offsets"])
    assert np.array_equal(enc.stream, VEC[f"{name}__stream"])


def test_noise_1024_exact_vs_oracle(ljb, ctx, oracle):
    """1024x1024 uniform noise (the benchmark distribution): zero coefficient mismatches, identical stream.
    north_star allows +-1 on <= 0.01 % of coefficients; this implementation is exact, tolerance 0."""
    img = ljb.synth.random_image(1024, 1024, seed=42)
    enc = ljb.jpeg.process(img, ctx=ctx)
    ref = oracle.jpeg_encode(img)
    mism = int((enc.coefs != ref["coefs"]).sum())
    assert mism == 0, f"{mism} of {ref['coefs'].size} coefficients differ"
    assert np.array_equal(enc.group_bits, ref["bits"])
    assert np.array_equal(enc.stream, ref["stream"])


def test_sub_range_of_groups(ljb, ctx, oracle):
    """Shards pass sub-ranges of groups; the records equal the corresponding slice of the whole stream."""
    img = ljb.synth.random_image(256, 64, seed=5)
    whole = ljb.jpeg.process(img, ctx=ctx)
    part = ljb.jpeg.process(img, first_group=96, ngroups=100, ctx=ctx)
    o0, o1 = int(whole.group_offsets[96]), int(whole.group_offsets[196])
    assert np.array_equal(part.stream, whole.stream[o0:o1])
    assert np.array_equal(part.coefs, whole.coefs[96:196])
    assert np.array_equal(part.group_offsets, whole.group_offsets[96:197] - np.uint64(o0))


def test_bit_string_accessor(ljb, ctx):
    enc = ljb.jpeg.process(cases.appendix_c_block(), ctx=ctx)
    s = enc.bit_string(0, 0)
    assert len(s) == 277 and s.startswith("0111101011010001010110111011000100110100")


def test_odd_width_rejected(ljb, ctx):
    with pytest.raises(ljb.LjbError) as e:
        ljb.jpeg.process(cases.synth_image(1, 9, 8), ctx=ctx)
    assert e.value.code == -1


def test_full_size_properties(ljb, ctx):
    """Size-independent checks at 4096x4096: every record length equals ceil(sum(bits)/8); padding bits are zero;
    per-tile look-back produced a gap-free, strictly ordered offset table."""
    img = ljb.synth.random_image(4096, 4096, seed=7)
    enc = ljb.jpeg.process(img, want_coefs=False, ctx=ctx)
    d = np.diff(enc.group_offsets.astype(np.int64))
    bits = enc.group_bits.astype(np.int64).sum(1)
    assert np.array_equal(d, (bits + 7) // 8)
    assert int(enc.group_offsets[-1]) == enc.stream.size
    last = enc.stream[enc.group_offsets[1:].astype(np.int64) - 1].astype(np.int64)
    pad = (8 - bits % 8) % 8
    assert ((last & ((1 << pad) - 1)) == 0).all()


def _binary_noise(w, h, seed):
    rng = np.random.default_rng(seed)
    img = np.empty((h, w, 4), np.uint8)
    img[..., :3] = rng.integers(0, 2, size=(h, w, 3), dtype=np.uint8) * 255
    img[..., 3] = 255
    return img


def test_extreme_contrast_vs_oracle(ljb, ctx, oracle):
    """0/255 pixels give the widest coefficient spread (values outside int8, > 32 distinct symbols per channel):
    exercises the hand-over from the shared-memory fast path to the general routine."""
    img = _binary_noise(128, 64, 3)
    img[:16, :16, :3] = np.indices((16, 16)).sum(0)[..., None] % 2 * 255  # checkerboard: one huge coefficient
    img[16:32, :16, :3] = (np.arange(16)[None, :, None] >= 8) * 255       # vertical edge
    enc = ljb.jpeg.process(img, ctx=ctx)
    ref = oracle.jpeg_encode(img)
    assert np.array_equal(enc.coefs, ref["coefs"])
    assert np.array_equal(enc.group_bits, ref["bits"])
    assert np.array_equal(enc.stream, ref["stream"])


@pytest.mark.parametrize("name", ["og_crop", "noise_64x48", "noise_18x13", "appendix_c", "white_16x16"])
def test_general_routine_alone(ljb, ctx, name, monkeypatch):
    """LJB_JPEG_FORCE_SLOW routes every channel through the general (local-memory) routine: same bytes."""
    monkeypatch.setenv("LJB_JPEG_FORCE_SLOW", "1")
    enc = ljb.jpeg.process(ALL[name], ctx=ctx)
    assert np.array_equal(enc.coefs, VEC[f"{name}__coefs"])
    assert np.array_equal(enc.group_bits, VEC[f"{name}__bits"])
    assert np.array_equal(enc.stream, VEC[f"{name}__stream"])


def test_photo_like_vs_oracle(ljb, ctx, oracle):
    """Smooth gradients + texture (large DC, many zero runs): the typical photographic case."""
    yy, xx = np.mgrid[0:96, 0:160]
    rng = np.random.default_rng(12)
    base = 128 + 100 * np.sin(xx / 23.0) * np.cos(yy / 17.0)
    img = np.empty((96, 160, 4), np.uint8)
    for c in range(3):
        img[..., c] = np.clip(base + (c - 1) * 30 + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)
    img[..., 3] = 255
    enc = ljb.jpeg.process(img, ctx=ctx)
    ref = oracle.jpeg_encode(img)
    assert np.array_equal(enc.coefs, ref["coefs"])
    assert np.array_equal(enc.stream, ref["stream"])


@pytest.mark.parametrize("chunk", [1, 4096, 20000])
def test_host_pipeline_many_bands(ljb, ctx, oracle, monkeypatch, chunk):
    """Bands of group rows through the upload / kernel / download pipeline, including a sub-range of groups that
    starts and ends in the middle of a group row."""
    monkeypatch.setenv("LJB_PIPE_CHUNK_BYTES", str(chunk))
    img = ljb.synth.random_image(72, 52, seed=21)  # 9 groups per row, 7 group rows, last row partial (52 = 6*8 + 4)
    enc = ljb.jpeg.process(img, ctx=ctx)
    ref = oracle.jpeg_encode(img)
    assert np.array_equal(enc.coefs, ref["coefs"])
    assert np.array_equal(enc.group_bits, ref["bits"])
    assert np.array_equal(enc.group_offsets, ref["offsets"])
    assert np.array_equal(enc.stream, ref["stream"])
    part = ljb.jpeg.process(img, first_group=5, ngroups=40, ctx=ctx)
    o0, o1 = int(enc.group_offsets[5]), int(enc.group_offsets[45])
    assert np.array_equal(part.stream, enc.stream[o0:o1])
    assert np.array_equal(part.group_offsets, enc.group_offsets[5:46] - np.uint64(o0))
    assert np.array_equal(part.coefs, enc.coefs[5:45])


@pytest.mark.parametrize("name", sorted(ALL))
def test_decode_matches_oracle(ljb, ctx, oracle, name):
    """Decode half on the GPU (Inverse_quantize, IDCT, assemble_image) == oracle, bit for bit, for every case."""
    img = ALL[name]
    h, w, _ = img.shape
    enc = ljb.jpeg.process(img, ctx=ctx)
    rec = ljb.jpeg.assemble_image(enc.coefs, w, h, original=img, ctx=ctx)
    assert np.array_equal(rec, oracle.jpeg_decode(enc.coefs, w, h, img))


def test_decode_golden_and_noise(ljb, ctx, oracle):
    img = cases.og_crop()
    h, w, _ = img.shape
    rec = ljb.jpeg.assemble_image(VEC["og_crop__coefs"], w, h, original=img, ctx=ctx)
    assert np.array_equal(rec, VEC["og_crop__reconstructed"])  # produced by the reference build
    noise = ljb.synth.random_image(512, 256, seed=3)
    enc = ljb.jpeg.process(noise, ctx=ctx)
    rec = ljb.jpeg.assemble_image(enc.coefs, 512, 256, ctx=ctx)  # multiples of 8: no original needed
    assert np.array_equal(rec, oracle.jpeg_decode(enc.coefs, 512, 256))
    with pytest.raises(ljb.LjbError):
        ljb.jpeg.assemble_image(ljb.jpeg.process(ALL["noise_18x13"], ctx=ctx).coefs, 18, 13, ctx=ctx)


@pytest.mark.parametrize("name", sorted(ALL))
def test_three_byte_pixels_match_reference_vectors(ljb, ctx, name):
    """The same images as r g b (3 bytes per pixel, what stbi_load(..., 3) gives): the reference build's vectors again."""
    enc = ljb.jpeg.process(np.ascontiguousarray(ALL[name][:, :, :3]), ctx=ctx)
    assert np.array_equal(enc.coefs, VEC[f"{name}__coefs"])
    assert np.array_equal(enc.group_bits, VEC[f"{name}__bits"])
    assert np.array_equal(enc.group_offsets, VEC[f"{name}__offsets"])
    assert np.array_equal(enc.stream, VEC[f"{name}__stream"])


@pytest.mark.parametrize("w,h", [(1024, 512), (72, 52), (136, 40), (8, 8), (2, 2)])
def test_three_byte_pixels_equal_four_byte_pixels(ljb, ctx, w, h):
    """Widths whose rows are 8-byte aligned (the vector path) and not (the general path), sub-ranges, the device entry point."""
    import torch

    img = ljb.synth.random_image(w, h, seed=w + h)
    rgb = np.ascontiguousarray(img[:, :, :3])
    a, b = ljb.jpeg.process(img, ctx=ctx), ljb.jpeg.process(rgb, ctx=ctx)
    assert np.array_equal(a.stream, b.stream) and np.array_equal(a.coefs, b.coefs)
    assert np.array_equal(a.group_offsets, b.group_offsets) and np.array_equal(a.group_bits, b.group_bits)
    ng = ljb.jpeg.group_count(w, h)
    if ng > 8:
        pa = ljb.jpeg.process(img, first_group=3, ngroups=ng - 5, ctx=ctx)
        pb = ljb.jpeg.process(rgb, first_group=3, ngroups=ng - 5, ctx=ctx)
        assert np.array_equal(pa.stream, pb.stream) and np.array_equal(pa.coefs, pb.coefs)
    d_in = torch.from_numpy(rgb).cuda()
    d_out = torch.empty(ng * 96 + 4096, dtype=torch.uint8, device="cuda")
    d_offs = torch.empty(ng + 1, dtype=torch.int64, device="cuda")
    d_bits = torch.empty(ng * 3, dtype=torch.int16, device="cuda")
    d_res = torch.zeros(3, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ljb.jpeg.encode_device(d_in, w, h, d_out, d_offs, d_bits, d_res, ctx, bpp=3)
    torch.cuda.synchronize()
    n = int(d_res[0].item())
    assert n == a.stream.size and np.array_equal(d_out[:n].cpu().numpy(), a.stream)


def test_three_byte_pixels_general_routine(ljb, ctx, monkeypatch):
    monkeypatch.setenv("LJB_JPEG_FORCE_SLOW", "1")
    img = ljb.synth.random_image(72, 52, seed=9)
    a, b = ljb.jpeg.process(img, ctx=ctx), ljb.jpeg.process(np.ascontiguousarray(img[:, :, :3]), ctx=ctx)
    assert np.array_equal(a.stream, b.stream) and np.array_equal(a.coefs, b.coefs)
