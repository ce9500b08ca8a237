"""GPU parity: the CUDA LZ4 encoder called through the C ABI, checked by the oracle / reference hashes."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

VEC = json.load(open(os.path.join(cases.GOLDEN, "lz4_ref_vectors.json")))
ALL = {name: (data, bl) for name, data, bl in cases.lz4_cases()}


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


def test_reference_golden_vector(ljb, ctx):
    """input.txt -> compressed.bin of the reference repository, byte for byte."""
    inp = np.fromfile(os.path.join(cases.GOLDEN, "lz4_input.txt"), dtype=np.uint8)
    gold = np.fromfile(os.path.join(cases.GOLDEN, "lz4_compressed.bin"), dtype=np.uint8)
    f = ljb.lz4.lz4_encode(inp, ljb.lz4.DEFAULT_BLOCK_LENGTH, ctx=ctx)
    assert np.array_equal(f.stream, gold)
    assert list(f.block_offsets) == [1, 321, 377]


@pytest.mark.parametrize("name", sorted(ALL))
def test_stream_equals_reference_build(ljb, ctx, oracle, name):
    """Bit-exact with the streams of the reference's own block_encode (hashes committed from oracle/_ref)
    and with the oracle restatement on the same seeded input."""
    data, bl = ALL[name]
    if data.size < bl:  # the reference refuses inputs shorter than a block (LZ4.c:632); exercise via block_len = n
        bl_eff = data.size
        s, offs, ph = oracle.lz4_compress(data, bl_eff, 1)
        f = ljb.lz4.lz4_encode(data, bl_eff, ctx=ctx)
        assert np.array_equal(f.stream, s) and np.array_equal(f.block_offsets, offs) and f.phantom == ph
        return
    f = ljb.lz4.lz4_encode(data, bl, ctx=ctx)
    assert f.stream.size == VEC[name]["size"]
    assert _sha(f.stream) == VEC[name]["sha256"]
    assert _sha(f.block_offsets) == VEC[name]["offsets_sha256"]
    s, offs, ph = oracle.lz4_compress(data, bl, 1)
    assert np.array_equal(f.stream, s) and np.array_equal(f.block_offsets, offs) and f.phantom == ph


@pytest.mark.parametrize("name", ["golden_input", "extract_30000", "long_runs_b3000", "zero_containing", "same_byte_2500",
                                  "repeats_ge1024_b4096", "wrap_257", "periodic_text"])
def test_match_stage_equals_find_longest_match(ljb, ctx, oracle, name):
    """Per-position (length, distance) of the search stage == exhaustive find_longest_match (bounded)."""
    data, bl = ALL[name]
    blk = data[:bl]
    ln, ds = ljb.lz4.find_longest_match(blk, ctx=ctx)
    l0, d0 = oracle.lz4_matches(blk, 0 if blk.size <= 8192 else 1)
    assert np.array_equal(ln, l0)
    used = ln != 1024  # (uint8_t)1024 == 0 turns the match into a literal step (LZ4.c:317); its distance is never used
    assert np.array_equal(ds[used], d0[used])


def test_match_stage_64k_block(ljb, ctx, oracle):
    blk = cases.synth_text(65536, seed=5)
    ln, ds = ljb.lz4.find_longest_match(blk, ctx=ctx)
    l0, d0 = oracle.lz4_matches(blk, 1)
    used = ln != 1024
    assert np.array_equal(ln, l0) and np.array_equal(ds[used], d0[used])


def test_many_blocks_parity_and_lookback(ljb, ctx, oracle):
    """More blocks than CTAs: exercises the ticket counter and the decoupled look-back (600 blocks of 4 KiB + tail)."""
    data = cases.synth_text(600 * 4096 + 777, seed=9)
    f = ljb.lz4.lz4_encode(data, 4096, ctx=ctx)
    s, offs, ph = oracle.lz4_compress(data, 4096, 1)
    assert np.array_equal(f.block_offsets, offs)
    assert np.array_equal(f.stream, s) and f.phantom == ph


def test_full_size_blocks_parity(ljb, ctx, oracle):
    """40 blocks of 64 KiB random_extract-style text (the benchmark distribution), bit-exact vs the oracle."""
    data = cases.synth_text(40 * 65536, seed=42)
    f = ljb.lz4.lz4_encode(data, 65536, ctx=ctx)
    s, offs, ph = oracle.lz4_compress(data, 65536, 1)
    assert np.array_equal(f.block_offsets, offs)
    assert np.array_equal(f.stream, s) and f.phantom == ph


def test_pathological_blocks_terminate(ljb, ctx, oracle):
    """All-equal and two-symbol 64 KiB blocks: worst cases for the candidate search, still exact."""
    rng = np.random.default_rng(3)
    mixed = cases.synth_text(2 * 65536, seed=33)
    mixed[10000:50000] = 0
    mixed[70000:90000] = np.tile(np.array([7, 8, 9], np.uint8), 6667)[:20000]
    twice = np.tile(rng.integers(0, 256, 32768, dtype=np.uint8), 2)
    twice[32768 + 20::40] ^= 0xFF  # ... with a changed byte every 40: ~39-byte matches instead of capped ones
    for data in (np.zeros(65536, np.uint8), rng.integers(0, 2, 65536, dtype=np.uint8) + 65, twice,
                 np.tile(np.array([97, 98], np.uint8), 32768), np.tile(rng.integers(0, 256, 1000, dtype=np.uint8), 66)[:65536].copy(),
                 np.tile(np.frombuffer(b"abcdefg", dtype=np.uint8), 9363)[:65536].copy(), mixed,
                 # half a block of noise, twice: 32 Ki groups of two, the group directory no longer fits behind the sorted
                 # positions and neighbouring groups share buckets (the search's hashed-bucket mode)
                 np.tile(rng.integers(0, 256, 32768, dtype=np.uint8), 2),
                 np.tile(rng.integers(0, 256, 21846, dtype=np.uint8), 3)[:65536].copy()):
        f = ljb.lz4.lz4_encode(data, 65536, ctx=ctx)
        s, offs, ph = oracle.lz4_compress(data, 65536, 1)
        assert np.array_equal(f.stream, s) and f.phantom == ph


@pytest.mark.parametrize("tune", ["2", "4"])
def test_wide_walker_segments_forced_and_forbidden(ljb, ctx, oracle, monkeypatch, tune):
    """Low-entropy blocks are walked in segments of 528 bytes instead of 66 (lz4_lazy.cuh, `wide`): the same streams with that
    mode forced on every block (LJB_LZ4_TUNE=2) and forbidden on every block (4) — text, two-symbol data, a ragged last block."""
    monkeypatch.setenv("LJB_LZ4_TUNE", tune)
    rng = np.random.default_rng(11)
    for data in (cases.synth_text(3 * 65536 + 1234, seed=5), rng.integers(0, 2, 2 * 65536 + 700, dtype=np.uint8) + 48,
                 rng.integers(0, 4, 65536, dtype=np.uint8), np.zeros(70000, np.uint8), cases.synth_text(5000, seed=1)):
        bl = min(65536, data.size)
        f = ljb.lz4.lz4_encode(data, bl, ctx=ctx)
        s, offs, ph = oracle.lz4_compress(data, bl, 1)
        assert np.array_equal(f.stream, s) and np.array_equal(f.block_offsets, offs) and f.phantom == ph


def test_capacity_error(ljb, ctx):
    data = cases.synth_text(8192, seed=1)
    with pytest.raises(ljb.LjbError) as e:
        ljb.lz4.lz4_encode(data, 4096, ctx=ctx, out_cap=100)
    assert e.value.code == -3


def test_roundtrip_gpu_decoder(ljb, ctx):
    for name in ("golden_input", "extract_30000", "periodic_text", "repeats_ge1024_b4096", "metamorphosis_64k", "lit_271",
                 "lit_526", "random_65535", "long_runs_b3000", "same_byte_2500", "synth_64k_x3", "hex_65536", "base32_2x65536",
                 "two_symbol_65536"):
        data, bl = ALL[name]
        f = ljb.lz4.lz4_encode(data, min(bl, data.size), ctx=ctx)
        if name in ("hex_65536", "base32_2x65536", "synth_64k_x3"):
            assert f.phantom == 0  # (the decode below must not drop out silently for these)
        if f.phantom == 0:
            out = ljb.lz4.LZ4_decode(f, ctx=ctx)
            assert np.array_equal(out, data), name


def test_decoder_rejects_bad_tables(ljb, ctx):
    """A short intermediate block or a non-monotonic offset table is LJB_E_FORMAT, not a hole of stale device memory."""
    data, bl = ALL["synth_64k_x3"]
    f = ljb.lz4.lz4_encode(data, bl, ctx=ctx)
    bad = f.block_offsets.copy()
    bad[1], bad[2] = f.block_offsets[2], f.block_offsets[1]
    with pytest.raises(ljb.LjbError) as e:
        ljb.lz4.lz4_decompress_raw(f.stream, bad, bl, data.size, ctx=ctx)
    assert e.value.code == ljb.LJB_E_FORMAT
    f2 = ljb.lz4.lz4_encode(data[: 2 * bl - 100], bl, ctx=ctx)
    s3 = np.concatenate([f2.stream, f.stream[int(f.block_offsets[2]):]])
    offs3 = np.concatenate([f2.block_offsets, [f2.block_offsets[-1] + (f.block_offsets[3] - f.block_offsets[2])]]).astype(np.uint64)
    with pytest.raises(ljb.LjbError) as e:
        ljb.lz4.lz4_decompress_raw(s3, offs3, bl, 3 * bl, ctx=ctx)
    assert e.value.code == ljb.LJB_E_FORMAT


def test_roundtrip_full_size_property(ljb, ctx):
    """Size-independent property at benchmark scale: decode(encode(x)) == x on 256 MiB... scaled to 64 MiB here."""
    data = ljb.synth.random_extract(64 << 20, seed=123)
    f = ljb.lz4.lz4_encode(data, 65536, ctx=ctx)
    assert f.blocks == 1024
    # block offsets are strictly increasing and every block header's low bytes match the true size when no phantom
    d = np.diff(f.block_offsets.astype(np.int64))
    assert (d > 3).all()
    assert f.phantom == 0  # the benchmark distribution has no 257..259-byte matches: the decode below always runs
    out = ljb.lz4.LZ4_decode(f, ctx=ctx)
    assert np.array_equal(out, data)


@pytest.mark.parametrize("chunk", [65536, 3 * 65536, 1000])
def test_host_pipeline_many_chunks(ljb, ctx, oracle, monkeypatch, chunk):
    """The host-buffer entry point cuts the input into chunks that overlap upload, kernel and download; a tiny
    LJB_PIPE_CHUNK_BYTES drives many chunks (and the per-chunk offset fix-up) through a small input."""
    monkeypatch.setenv("LJB_PIPE_CHUNK_BYTES", str(chunk))
    data, bl = ALL["synth_64k_x3"]
    data = np.concatenate([data, cases.synth_text(4 * 65536 + 777, seed=99)])
    f = ljb.lz4.lz4_encode(data, bl, ctx=ctx)
    s, offs, ph = oracle.lz4_compress(data, bl, 1)
    assert np.array_equal(f.stream, s) and np.array_equal(f.block_offsets, offs) and f.phantom == ph
    f2 = ljb.lz4.lz4_encode(data[:30000], 300, ctx=ctx)  # 100 blocks of 300 B, chunks of 3 blocks at chunk = 1000
    s2, offs2, ph2 = oracle.lz4_compress(data[:30000], 300, 1)
    assert np.array_equal(f2.stream, s2) and np.array_equal(f2.block_offsets, offs2) and f2.phantom == ph2
