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
