"""Pins oracle/jpeg_oracle.c: the reference's colour-plane PNGs, reference-build vectors, Appendix C."""
import os

import numpy as np
import pytest

import cases

GOLDEN = cases.GOLDEN
VEC = np.load(os.path.join(GOLDEN, "jpeg_ref_vectors.npz"))
ALL = dict(cases.jpeg_cases())


def _png(name):
    from PIL import Image

    return np.array(Image.open(os.path.join(GOLDEN, name)).convert("RGBA"))


def test_colour_planes_against_reference_pngs(oracle):
    """Output-Input/Images/{bChrominance,rChrominance}.png are rendered from the Cb/Cr planes by
    S-JPG:254-300; both reproduce exactly.  luminance.png (written by the reference's 32-bit x87 build)
    differs by +-1 on ~1 % of pixels from any SSE2 evaluation."""
    rgba = cases.og_crop()
    Y, Cr, Cb = oracle.jpeg_planes(rgba)
    gcb, gcr, gl = _png("og_crop_cb.png"), _png("og_crop_cr.png"), _png("og_crop_lum.png")
    cb = Cb.astype(np.float64)
    cr = Cr.astype(np.float64)
    # S-JPG:290-292 / 266-268, double -> unsigned char truncation
    assert np.array_equal(gcb[..., 0], np.full_like(Cb, 128))
    assert np.array_equal(gcb[..., 1], (128 - 0.344 * (cb - 128) - 0.714 * 0).astype(np.int64).astype(np.uint8))
    assert np.array_equal(gcb[..., 2], (128 + 1.772 * (cb - 128)).astype(np.int64).astype(np.uint8))
    assert np.array_equal(gcr[..., 0], (128 + 1.402 * (cr - 128)).astype(np.int64).astype(np.uint8))
    assert np.array_equal(gcr[..., 1], (128 - 0.344 * 0 - 0.714 * (cr - 128)).astype(np.int64).astype(np.uint8))
    d = np.abs(gl[..., 0].astype(int) - Y.astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.03


@pytest.mark.parametrize("name", sorted(ALL))
def test_oracle_matches_reference_build_vectors(oracle, name):
    r = oracle.jpeg_encode(ALL[name])
    assert np.array_equal(r["coefs"], VEC[f"{name}__coefs"])
    assert np.array_equal(r["bits"], VEC[f"{name}__bits"])
    assert np.array_equal(r["offsets"], VEC[f"{name}__offsets"])
    assert np.array_equal(r["stream"], VEC[f"{name}__stream"])


@pytest.mark.parametrize("name", ["og_crop", "noise_64x48", "gradient_24x40", "noise_18x13"])
def test_oracle_equals_reference_build_live(oracle, ref_jpeg, name):
    rgba = ALL[name]
    a, b = oracle.jpeg_encode(rgba), ref_jpeg.jpeg_encode(rgba)
    for k in ("stream", "coefs", "bits", "offsets"):
        assert np.array_equal(a[k], b[k]), k
    for g in (0, oracle.jpeg_group_count(rgba.shape[1], rgba.shape[0]) - 1):
        sa, ca, ra = oracle.jpeg_group_stages(rgba, g)
        sb, cb, rb = ref_jpeg.jpeg_group_stages(rgba, g)
        assert np.array_equal(sa, sb)
        assert np.array_equal(ca, cb)  # bit-identical doubles
        assert all(np.array_equal(x, y) for x, y in zip(ra, rb))
    assert all(np.array_equal(x, y) for x, y in zip(oracle.jpeg_planes(rgba), ref_jpeg.jpeg_planes(rgba)))


def test_appendix_c_known_answer(oracle):
    """SURVEY.md Appendix C (probed from the reference build)."""
    rgba = cases.appendix_c_block()
    Y, Cr, Cb = oracle.jpeg_planes(rgba)
    assert list(Y[0]) == [60, 89, 117, 145, 173, 172, 50, 78]
    assert list(Cr[0]) == [91, 84, 77, 70, 63, 74, 162, 155]
    assert list(Cb[0]) == [101, 113, 125, 138, 150, 50, 137, 149]
    samples, coef, rles = oracle.jpeg_group_stages(rgba, 0)
    assert list(samples[64:64 + 8]) == [84, 70, 74, 155, 141, 145, 131, 117]
    assert np.allclose(coef[:8], [-37.0, -57.220249, -54.063850, 64.117888, -21.0, 5.320104, -24.690081, 17.369111], atol=1e-6)
    r = oracle.jpeg_encode(rgba)
    assert list(r["coefs"][0, :8]) == [-4, -9, -9, 8, -2, 0, -1, 0]
    assert list(r["coefs"][0, 64:76]) == [-1, 2, 0, 0, -1, -2, 2, 0, 0, -2, 0, 0]
    assert len(rles[0]) == 98 and list(rles[0][:8]) == [1, -4, 1, -9, 1, 0, 1, 4] and list(rles[0][-6:]) == [1, 1, 1, -1, 3, 0]
    assert r["bits"][0, 0] == 277
    bits = "".join(f"{b:08b}" for b in r["stream"])
    assert bits.startswith("0111101011010001010110111011000100110100")


def test_zigzag_orders(oracle):
    """the generic walk yields the standard 8x8 order and the listed 4-wide x 8-tall order."""
    # feed a block whose quantised coefficients are all distinct is impractical; check through stages on a ramp
    # instead: verify the permutation by construction in Python against the listed orders.
    def zz(width, height):
        order = []
        for s in range(width + height - 1):
            start_row = 0 if s < width else s - width + 1
            end_row = s if s < height else height - 1
            rows = range(end_row, start_row - 1, -1) if s % 2 == 0 else range(start_row, end_row + 1)
            for row in rows:
                col = s - row
                if col < width:
                    order.append(row * width + col)
        return order
    assert zz(8, 8)[:15] == [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4]
    assert zz(4, 8) == [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 16, 13, 10, 7, 11, 14, 17, 20, 24, 21, 18, 15, 19, 22, 25, 28, 29, 26, 23, 27, 30, 31]


def test_partial_tiles_and_group_count(oracle):
    """B.8: only ceil(w*h/64) groups are processed; missing pixels are 0."""
    assert oracle.jpeg_group_count(10, 6) == 1
    assert oracle.jpeg_group_count(18, 13) == 4
    assert oracle.jpeg_group_count(1200, 630) == 11813
    samples, _, _ = oracle.jpeg_group_stages(ALL["noise_10x6"], 0)
    assert samples[:64].reshape(8, 8)[6:].max() == 0


def test_odd_width_rejected(oracle):
    with pytest.raises(RuntimeError):
        oracle.jpeg_encode(cases.synth_image(1, 9, 8))


@pytest.mark.parametrize("name", ["og_crop", "noise_64x48", "noise_18x13", "appendix_c", "gradient_24x40", "noise_10x6", "red_24x16"])
def test_decode_equals_reference_build(oracle, ref_jpeg, name):
    """Decode half (Inverse_quantize, IDCT, assemble_image): restatement == the reference's own functions, including
    sizes whose last tiled groups the reference leaves unprocessed."""
    img = dict(cases.jpeg_cases())[name]
    h, w, _ = img.shape
    coefs = oracle.jpeg_encode(img)["coefs"]
    a = oracle.jpeg_decode(coefs, w, h, img)
    b = ref_jpeg.jpeg_decode(coefs, w, h, img)
    assert np.array_equal(a, b)
    assert (a[..., 3] == 255).all()


def test_decode_golden_fixture(oracle):
    """Committed reconstruction of the og.png crop by the reference build (tests/golden/make_golden.py)."""
    vec = np.load(os.path.join(cases.GOLDEN, "jpeg_ref_vectors.npz"))
    img = cases.og_crop()
    h, w, _ = img.shape
    out = oracle.jpeg_decode(vec["og_crop__coefs"], w, h, img)
    assert np.array_equal(out, vec["og_crop__reconstructed"])
