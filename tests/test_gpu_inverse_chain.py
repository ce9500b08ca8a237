"""GPU parity of the inverse chain and of BASELINE.json configs[1]: the whole Assets/Images/og.png through both encoders, the
entropy decoder (decode_huffman -> inverse_RLE -> reverse_zigzag_pattern, JPEG.c:1009, :811, :729) on the PACKED STREAM, and
the reconstruction (JPEG.c:1408-1428) — against hashes of the reference build's own results (tests/golden/og_full_ref.json)."""
import hashlib
import json
import os

import numpy as np
import pytest
from PIL import Image

import cases

pytestmark = pytest.mark.gpu

REFV = json.load(open(os.path.join(cases.GOLDEN, "og_full_ref.json")))
ALL = dict(cases.jpeg_cases())


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ljb():
    import lz4jpeg_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(ljb):
    c = ljb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def og():
    return np.asarray(Image.open(os.path.join(cases.GOLDEN, "og.png")).convert("RGBA"))


@pytest.mark.parametrize("name", sorted(ALL))
def test_bit_stream_decodes_to_the_coefficients(ljb, ctx, name):
    """What the encoder packs is decodable: Huffman decode + inverse RLE + reverse zig-zag of the stream == its coefficients."""
    enc = ljb.jpeg.process(ALL[name], ctx=ctx)
    trees = ljb.jpeg.huffman_trees(enc.coefs, ctx=ctx)
    assert np.array_equal(ljb.jpeg.decode_huffman(enc, trees, ctx=ctx), enc.coefs)


def test_noise_stream_decodes(ljb, ctx):
    img = ljb.synth.random_image(512, 256, seed=7)
    enc = ljb.jpeg.process(img, ctx=ctx)
    trees = ljb.jpeg.huffman_trees(enc.coefs, ctx=ctx)
    assert np.array_equal(ljb.jpeg.decode_huffman(enc, trees, ctx=ctx), enc.coefs)
    bad = trees.copy()
    bad[5, 0] = 0  # a tree with no leaves
    with pytest.raises(ljb.LjbError) as e:
        ljb.jpeg.decode_huffman(enc, bad, ctx=ctx)
    assert e.value.code == ljb.LJB_E_FORMAT


def test_og_png_whole_image_encode(ljb, ctx, og):
    """configs[1]: the whole 1200 x 630 og.png; coefficients, bit lengths, offsets and stream == the reference build's."""
    enc = ljb.jpeg.process(og, ctx=ctx)
    assert enc.coefs.shape[0] == REFV["groups"] and enc.stream.size == REFV["stream_bytes"]
    assert _sha(enc.coefs) == REFV["coefs_sha256"]
    assert _sha(enc.group_bits) == REFV["bits_sha256"]
    assert _sha(enc.group_offsets) == REFV["offsets_sha256"]
    assert _sha(enc.stream) == REFV["stream_sha256"]


def test_og_png_whole_inverse_chain(ljb, ctx, og):
    """JPEG_seq.exe end to end on the GPU: encode, decode the bit stream, dequantise, IDCT, assemble == the reference's pixels."""
    enc = ljb.jpeg.process(og, ctx=ctx)
    trees = ljb.jpeg.huffman_trees(enc.coefs, ctx=ctx)
    coefs = ljb.jpeg.decode_huffman(enc, trees, ctx=ctx)
    assert np.array_equal(coefs, enc.coefs)
    h, w, _ = og.shape
    rec = ljb.jpeg.assemble_image(coefs, w, h, original=og, ctx=ctx)
    assert _sha(rec) == REFV["reconstructed_sha256"]


@pytest.mark.parametrize("name,sub", [("444", 0), ("420", -1)])
def test_og_png_baseline_jpeg_file(ljb, ctx, og, name, sub):
    """configs[1] as worded (quality 75, 4:4:4) and stb's own rule (4:2:0): the .jpg == the vendored stbi_write_jpg's file."""
    jpg = ljb.jfif.write_jpg(og, 75, sub, ctx=ctx)
    assert jpg.size == REFV[f"jfif_q75_{name}_bytes"]
    assert _sha(jpg) == REFV[f"jfif_q75_{name}_sha256"]
