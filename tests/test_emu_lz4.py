"""CPU check of the CUDA LZ4 kernels' LOGIC: lz4-jpeg_b200/csrc/lz4_encode.cu (+ lz4_lazy.cuh, lz4_decode.cu) compiled
for the host under the lock-step emulator of tests/emu/cuda_emu.h and compared with the oracle.

This is test infrastructure (a debugging aid for a container without a GPU), not a product path: the emulator build is a
separate shared object under tests/emu/ that nothing in lz4-jpeg_b200/ loads.  The parity tests proper are the `-m gpu`
tests, which run the same source on the B200."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import cases

ALL = {name: (data, bl) for name, data, bl in cases.lz4_cases()}
LAZY, FULL, SMALL = 1, 0, 2


@pytest.fixture(scope="module")
def emu():
    import importlib.util

    spec = importlib.util.spec_from_file_location("ljb_emu_build", os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "build.py"))
    emu_build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emu_build)

    L = C.CDLL(emu_build.build())
    L.emu_lz4_compress.restype = C.c_int
    L.emu_lz4_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    L.emu_lz4_decompress.restype = C.c_int
    L.emu_lz4_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                     C.c_void_p]
    return L


def _compress(L, data, bl, mode):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    nb = (data.size + bl - 1) // bl
    cap = 1 + 3 * nb + 6 * data.size + 64
    out = np.zeros(cap, np.uint8)
    offs = np.zeros(nb + 1, np.uint64)
    res = np.zeros(3, np.uint64)
    rc = L.emu_lz4_compress(data.ctypes.data, data.size, bl, out.ctypes.data, cap, offs.ctypes.data, res.ctypes.data, mode, None, None, None)
    assert rc == 0, f"emulator rc {rc} (-7: a write past the kernel's shared-memory extent)"
    assert int(res[2]) == 0
    return out[: int(res[0])], offs, int(res[1])


def _decompress(L, stream, offs, bl, n):
    out = np.zeros(max(n, 1), np.uint8)
    lens = np.zeros(offs.size - 1, np.uint32)
    res = np.zeros(3, np.uint64)
    rc = L.emu_lz4_decompress(stream.ctypes.data, stream.size, offs.ctypes.data, offs.size - 1, bl, out.ctypes.data, out.size,
                              lens.ctypes.data, res.ctypes.data)
    assert rc == 0
    return out, lens, int(res[2])


@pytest.mark.parametrize("name", sorted(ALL))
def test_chain_search_kernel_equals_oracle(emu, oracle, name):
    """The product kernel (search along the greedy chain, lz4_lazy.cuh) on every seeded case."""
    data, bl = ALL[name]
    bl = min(bl, data.size)
    s, offs, ph = _compress(emu, data, bl, LAZY)
    s0, o0, p0 = oracle.lz4_compress(data, bl, 1)
    assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0


@pytest.mark.parametrize("name", ["golden_input", "repeats_ge1024_b4096", "wrap_257", "zeros_4096", "text_with_runs_16384", "lit_271"])
def test_full_search_kernel_equals_oracle(emu, oracle, name):
    """The every-position search (what ljb_lz4_block_matches runs) on a few cases."""
    data, bl = ALL[name]
    bl = min(bl, data.size)
    s, offs, ph = _compress(emu, data, bl, FULL)
    s0, o0, p0 = oracle.lz4_compress(data, bl, 1)
    assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0


def test_chain_search_pathological_blocks(emu, oracle):
    """64 KiB blocks of one byte, of period 2, of two random symbols, a block repeated with sparse changes: the candidate
    lists are the whole block; the walkers' capped-run marking and the pruned long compares keep the result exact."""
    rng = np.random.default_rng(3)
    twice = np.tile(rng.integers(0, 256, 32768, dtype=np.uint8), 2)
    twice[32768 + 20::40] ^= 0xFF
    mixed = cases.synth_text(65536, seed=33)
    mixed[10000:50000] = 0
    for data in (np.zeros(65536, np.uint8), np.tile(np.array([97, 98], np.uint8), 32768), twice, mixed,
                 np.tile(rng.integers(0, 256, 1000, dtype=np.uint8), 66)[:65536].copy()):
        s, offs, ph = _compress(emu, data, 65536, LAZY)
        s0, o0, p0 = oracle.lz4_compress(data, 65536, 1)
        assert np.array_equal(s, s0) and ph == p0


def test_chain_search_wide_walker_segments(emu, oracle, monkeypatch):
    """Two-symbol data takes the wide walker segments (8 x 66 bytes) by itself; LJB_LZ4_TUNE=2 forces them on text as well."""
    rng = np.random.default_rng(11)
    two = rng.integers(0, 2, 65536 + 900, dtype=np.uint8) + 48
    s, offs, ph = _compress(emu, two, 65536, LAZY)
    s0, o0, p0 = oracle.lz4_compress(two, 65536, 1)
    assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0
    monkeypatch.setenv("LJB_LZ4_TUNE", "2")
    for data in (cases.synth_text(65536 + 3000, seed=5), cases.synth_text(1000, seed=2), np.zeros(66000, np.uint8)):
        s, offs, ph = _compress(emu, data, 65536, LAZY)
        s0, o0, p0 = oracle.lz4_compress(data, 65536, 1)
        assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0


def test_chain_search_benchmark_blocks(emu, oracle):
    """12 blocks of the benchmark distribution through one persistent CTA (ticket counter, look-back, deferred placement)."""
    data = cases.synth_text(12 * 65536, seed=42)
    s, offs, ph = _compress(emu, data, 65536, LAZY)
    s0, o0, p0 = oracle.lz4_compress(data, 65536, 1)
    assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0


def test_decoder_kernel_roundtrip_and_rejects(emu, oracle):
    """The decoder kernel: round trips (including the expanding low-alphabet blocks whose short sequences lie 64 KiB before the
    block's end), and malformed tables flagged as format errors."""
    for name in ("golden_input", "hex_65536", "base32_2x65536", "lit_526", "random_65535", "same_byte_2500", "periodic_text"):
        data, bl = ALL[name]
        bl = min(bl, data.size)
        s, offs, ph = oracle.lz4_compress(data, bl, 1)
        assert ph == 0 or name in ("same_byte_2500", "periodic_text")
        if ph:  # 257..259-byte matches are not representable in the format
            continue
        out, lens, err = _decompress(emu, s, offs, bl, data.size)
        assert err == 0 and np.array_equal(out[: data.size], data), name
    data, bl = ALL["synth_64k_x3"]
    s, offs, _ = oracle.lz4_compress(data, bl, 1)
    bad = offs.copy()
    bad[1], bad[2] = offs[2], offs[1]
    assert _decompress(emu, s, bad, bl, data.size)[2] & 2
    s2, offs2, _ = oracle.lz4_compress(data[: 2 * bl - 100], bl, 1)
    s3 = np.concatenate([s2, s[int(offs[2]):]])
    offs3 = np.concatenate([offs2, [offs2[-1] + (offs[3] - offs[2])]]).astype(np.uint64)
    assert _decompress(emu, s3, offs3, bl, 3 * bl)[2] & 2


@pytest.mark.parametrize("name", sorted(n for n, (d, bl) in ALL.items() if min(bl, d.size) <= 4096))
def test_small_block_kernel_equals_oracle(emu, oracle, name):
    """The warp-per-block kernel (lz4_small.cuh: block lengths up to 4096, the reference's own 300 among them)."""
    data, bl = ALL[name]
    bl = min(bl, data.size)
    s, offs, ph = _compress(emu, data, bl, SMALL)
    s0, o0, p0 = oracle.lz4_compress(data, bl, 1)
    assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0


def test_small_block_kernel_many_blocks(emu, oracle):
    """More blocks than warps (ticket counter, look-back between warps), ragged last block, several block lengths; zeros at 4096
    (matches capped at 1024: the literal-run skip)."""
    for bl in (300, 1024, 4096, 5, 33):
        data = cases.synth_text(70 * bl + 17, seed=3)
        s, offs, ph = _compress(emu, data, bl, SMALL)
        s0, o0, p0 = oracle.lz4_compress(data, bl, 1)
        assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0, bl
    z = np.zeros(3 * 4096, np.uint8)
    for bl in (300, 2500, 4096):
        s, offs, ph = _compress(emu, z, bl, SMALL)
        s0, o0, p0 = oracle.lz4_compress(z, bl, 1)
        assert np.array_equal(s, s0) and np.array_equal(offs, o0) and ph == p0, bl
