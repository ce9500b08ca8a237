"""CPU: the baseline-JPEG restatement (oracle/jfif_oracle.c) against the reference's vendored stb_image_write.h.

Pins the oracle three ways: (1) byte for byte against stbi_write_jpg_to_func compiled from where it lies under
/root/reference (oracle/_ref/libref_jfif.so; skipped where that build is absent), (2) against the committed SHA-256
of that build's output for every case (runs everywhere), (3) against a committed real file.  Also checks that an
independent decoder (Pillow) accepts the files.
"""
import hashlib
import io
import json
import os

import numpy as np
import pytest

import cases

VEC = json.load(open(os.path.join(cases.GOLDEN, "jfif_ref_vectors.json")))
ALL = {name: (px, q, sub) for name, px, q, sub in cases.jfif_cases()}


@pytest.fixture(scope="module")
def ref_jfif():
    from oracle.pyoracle import Ref

    if not Ref.available("jfif"):
        pytest.skip("oracle/_ref/libref_jfif.so not built (needs /root/reference)")
    return Ref("jfif")


def test_every_case_has_a_vector():
    assert sorted(ALL) == sorted(VEC)


@pytest.mark.parametrize("name", sorted(ALL))
def test_oracle_matches_committed_hash_of_the_stb_build(oracle, name):
    px, q, sub = ALL[name]
    jpg = oracle.jfif_encode(px, q, sub)
    assert jpg.size == VEC[name]["size"]
    assert hashlib.sha256(jpg.tobytes()).hexdigest() == VEC[name]["sha256"]


@pytest.mark.parametrize("name", sorted(ALL))
def test_oracle_matches_stb_build(oracle, ref_jfif, name):
    px, q, sub = ALL[name]
    assert np.array_equal(oracle.jfif_encode(px, q, sub), ref_jfif.jfif_encode(px, q, sub))


def test_committed_file_is_what_stb_writes(oracle):
    want = np.fromfile(os.path.join(cases.GOLDEN, "og_crop_q75.jpg"), dtype=np.uint8)
    assert np.array_equal(oracle.jfif_encode(cases.og_crop(), 75), want)


@pytest.mark.parametrize("quality", [1, 25, 50, 75, 90, 91, 100])
def test_random_sizes_against_stb_build(oracle, ref_jfif, quality):
    rng = np.random.default_rng(quality)
    for _ in range(12):
        h, w, comp = int(rng.integers(1, 70)), int(rng.integers(1, 70)), int(rng.integers(1, 5))
        px = rng.integers(0, 256, size=(h, w, comp), dtype=np.uint8)
        if rng.integers(0, 2):
            px = (px // 64 * 64).astype(np.uint8)  # few levels: flat areas, zero runs
        for sub in (-1, 0, 1):
            assert np.array_equal(oracle.jfif_encode(px, quality, sub), ref_jfif.jfif_encode(px, quality, sub)), (h, w, comp, sub)


def test_huffman_codes_are_annex_k(oracle):
    """A few code words of ITU-T T.81 Tables K.3 - K.6 (the ones stb hard-codes, stb_image_write.h:1428-1465)."""
    code, ln = oracle.jfif_huffman(1)  # AC luminance
    assert (code[0x00], ln[0x00]) == (0b1010, 4)            # EOB
    assert (code[0xF0], ln[0xF0]) == (0b11111111001, 11)    # ZRL
    assert (code[0x01], ln[0x01]) == (0b00, 2)
    assert (code[0xFA], ln[0xFA]) == (0xFFFE, 16)
    code, ln = oracle.jfif_huffman(3)  # AC chrominance
    assert (code[0x00], ln[0x00]) == (0b00, 2)
    assert (code[0xF0], ln[0xF0]) == (0b1111111010, 10)
    code, ln = oracle.jfif_huffman(0)  # DC luminance
    assert [(int(code[i]), int(ln[i])) for i in range(12)] == [(0, 2), (2, 3), (3, 3), (4, 3), (5, 3), (6, 3), (14, 4), (30, 5), (62, 6),
                                                               (126, 7), (254, 8), (510, 9)]
    code, ln = oracle.jfif_huffman(2)  # DC chrominance
    assert (code[11], ln[11]) == (2046, 11)


def test_quantisers_follow_the_quality_rule(oracle):
    qy, quv, dy, duv, rank = oracle.jfif_tables(50)  # scale 100: the Annex K tables themselves
    assert qy[0] == 16 and qy[1] == 11 and qy[2] == 12 and qy[63] == 99  # zig-zag order
    assert quv[0] == 17 and quv[63] == 99
    assert list(rank[:8]) == [0, 1, 5, 6, 14, 15, 27, 28]
    qy100, _, _, _, _ = oracle.jfif_tables(100)
    assert (qy100 == 1).all()
    qy1, _, _, _, _ = oracle.jfif_tables(1)
    assert (qy1 == 255).all()
    assert np.isclose(dy[0], 1.0 / (16 * 8.0), rtol=1e-6)


@pytest.mark.parametrize("name", ["og_crop_q75", "og_crop_q95", "gradient_300x200_q75", "noise_33x70_q90", "grey_100x37_q80"])
def test_pillow_decodes_the_file(oracle, name):
    from PIL import Image

    px, q, sub = ALL[name]
    jpg = oracle.jfif_encode(px, q, sub)
    im = Image.open(io.BytesIO(jpg.tobytes()))
    im.load()
    h, w = px.shape[:2]
    assert im.size == (w, h) and im.mode == "RGB"
    if name.startswith(("og_crop", "gradient")):
        src = px[:, :, :3].astype(np.float64)
        mse = ((np.asarray(im, dtype=np.float64) - src) ** 2).mean()
        assert 10 * np.log10(255.0 ** 2 / mse) > 28.0
