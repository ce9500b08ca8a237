"""Pins oracle/lz4_oracle.c: reference golden vector, reference-build streams, known answers."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

GOLDEN = cases.GOLDEN
VEC = json.load(open(os.path.join(GOLDEN, "lz4_ref_vectors.json")))
ALL = {name: (data, bl) for name, data, bl in cases.lz4_cases()}
SLOW_EXHAUSTIVE = {"metamorphosis_64k", "synth_64k_x3", "random_65536", "random_65535", "random_65536_plus", "hex_65536",
                   "base32_2x65536", "two_symbol_65536"}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_reference_golden_vector(oracle):
    """Output-Input/input/input.txt -> Output-Input/out/compressed.bin (the reference's only LZ4 known answer)."""
    inp = np.fromfile(os.path.join(GOLDEN, "lz4_input.txt"), dtype=np.uint8)
    gold = np.fromfile(os.path.join(GOLDEN, "lz4_compressed.bin"), dtype=np.uint8)
    for mode in (0, 1):
        s, offs, ph = oracle.lz4_compress(inp, 300, mode)
        assert np.array_equal(s, gold)
        assert list(offs) == [1, 321, 377] and ph == 0
    # block 0 = 13 sequences / byte_size 320, block 1 = 3 sequences / 56
    g = [int(x) for x in gold]
    assert g[0] == 2 and g[1] == 13 and g[2] | (g[3] << 8) == 320
    assert g[321] == 3 and g[322] | (g[323] << 8) == 56


@pytest.mark.parametrize("name", sorted(ALL))
def test_oracle_matches_reference_build_hashes(oracle, name):
    """Streams of the reference's own block_encode/write_block (oracle/_ref, bounded) — committed as hashes."""
    data, bl = ALL[name]
    modes = (1,) if name in SLOW_EXHAUSTIVE else (0, 1)
    for mode in modes:
        s, offs, _ = oracle.lz4_compress(data, bl, mode)
        assert s.size == VEC[name]["size"], (name, mode)
        assert _sha(s) == VEC[name]["sha256"], (name, mode)
        assert _sha(offs) == VEC[name]["offsets_sha256"], (name, mode)


@pytest.mark.parametrize("name", sorted(set(ALL) - SLOW_EXHAUSTIVE))
def test_oracle_equals_reference_build_live(oracle, ref_lz4, name):
    data, bl = ALL[name]
    a, ao, _ = oracle.lz4_compress(data, bl, 0)
    b, bo, _ = ref_lz4.lz4_compress(data, bl)
    assert np.array_equal(a, b) and np.array_equal(ao, bo)


def test_known_answers_metamorphosis(oracle):
    """B=300 -> 123143 B, frame byte 0x8B; B=65536 -> 91323 B, block 0 token 0x10, size 50099."""
    meta = cases.corpus()
    s, _, _ = oracle.lz4_compress(meta, 300, 0)
    assert s.size == 123143 and s[0] == 0x8B
    s, offs, _ = oracle.lz4_compress(meta, 65536, 1)
    assert s.size == 91323 and s[1] == 0x10 and (int(s[2]) | int(s[3]) << 8) == 50099


def test_wrap_cases_tokens(oracle):
    """256 -> literal step; 257..259 -> tokens FD/FE/FF and header = payload + 1; 260 normal."""
    for k, tok in ((257, 0xFD), (258, 0xFE), (259, 0xFF)):
        s, offs, ph = oracle.lz4_compress(cases._wrap_case(k), 2048, 0)
        assert ph == 1
        hdr = int(s[2]) | int(s[3]) << 8
        assert hdr == (int(offs[1]) - int(offs[0])) + 1
        toks = _tokens(s, offs)
        assert tok in toks, (k, [hex(t) for t in toks])
    s, offs, ph = oracle.lz4_compress(cases._wrap_case(256), 2048, 0)
    assert ph == 0
    s, offs, ph = oracle.lz4_compress(cases._wrap_case(260), 2048, 0)
    assert ph == 0


def _tokens(s, offs):
    """Walk block 0 structurally using seq_byte_size (phantom sequences are one byte shorter than they say)."""
    toks = []
    q, e = int(offs[0]) + 3, int(offs[1])
    while q < e:
        tok = int(s[q])
        size = int(s[q + 1]) | int(s[q + 2]) << 8
        toks.append(tok)
        if tok in (0xFD, 0xFE, 0xFF) and q + size > e:
            size -= 1
        q += size
    return toks


def test_match_stage_fast_equals_exhaustive(oracle):
    rng = np.random.default_rng(5)
    for data in (cases.synth_text(6000, seed=21), rng.integers(0, 3, 3000, dtype=np.uint8), np.full(1500, 7, np.uint8),
                 np.frombuffer(b"abcdabcdabcd" * 100, dtype=np.uint8).copy()):
        l0, d0 = oracle.lz4_matches(data, 0)
        l1, d1 = oracle.lz4_matches(data, 1)
        assert np.array_equal(l0, l1) and np.array_equal(d0, d1)


def test_zero_sequence_block(oracle):
    """A.3-c: a 65536-byte block without any match wraps the uint16 literal counter: no sequence at all."""
    data, bl = ALL["random_65536"]
    s, offs, _ = oracle.lz4_compress(data, bl, 1)
    l, _ = oracle.lz4_matches(data, 1)
    if int(l.max()) == 0:
        assert s.size == 4 and list(s) == [1, 0, 3, 0]


def test_roundtrip_format_decoder(oracle):
    for name in ("golden_input", "extract_30000", "periodic_text", "repeats_ge1024_b4096", "metamorphosis_64k",
                 "lit_271", "lit_526", "random_65535", "long_runs_b3000", "same_byte_2500", "tiny_1", "tiny_9", "hex_65536",
                 "base32_2x65536", "two_symbol_65536"):
        data, bl = ALL[name]
        s, offs, ph = oracle.lz4_compress(data, bl, 1)
        rc, out = oracle.lz4_decompress(s, offs, bl, data.size)
        if ph == 0:
            assert rc == 0 and np.array_equal(out, data), name


def test_expanding_blocks_roundtrip(oracle):
    """Blocks that expand past 64 KiB (random hex / base32 text): every short sequence then lies 64 KiB or more before the
    block's end, and the size field of a sequence wraps only when it really holds >= 65531 literals."""
    for name in ("hex_65536", "base32_2x65536"):
        data, bl = ALL[name]
        s, offs, ph = oracle.lz4_compress(data, bl, 1)
        assert ph == 0 and int(offs[1]) - int(offs[0]) >= 65536
        rc, out = oracle.lz4_decompress(s, offs, bl, data.size)
        assert rc == 0 and np.array_equal(out, data), name


def test_decoder_rejects_bad_tables(oracle):
    """A short intermediate block or a non-monotonic offset table is a format error, not a hole in the output."""
    data, bl = ALL["synth_64k_x3"]
    s, offs, _ = oracle.lz4_compress(data, bl, 1)
    bad = offs.copy()
    bad[1], bad[2] = offs[2], offs[1]
    rc, _ = oracle.lz4_decompress(s, bad, bl, data.size)
    assert rc != 0
    s2, offs2, _ = oracle.lz4_compress(data[: 2 * bl - 100], bl, 1)  # last block short: fine as the last block ...
    rc, out = oracle.lz4_decompress(s2, offs2, bl, 2 * bl)
    assert rc == 0 and out.size == 2 * bl - 100
    s3 = np.concatenate([s2, s[int(offs[2]):]])                      # ... but not with another block behind it
    offs3 = np.concatenate([offs2, [offs2[-1] + (offs[3] - offs[2])]]).astype(np.uint64)
    rc, _ = oracle.lz4_decompress(s3, offs3, bl, 3 * bl)
    assert rc != 0
