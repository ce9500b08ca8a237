"""tests/cases.py — seeded inputs shared by the golden-fixture generator and the tests.

LZ4 cases follow the list SURVEY.md section 8(c) used to validate the format restatement (golden
input, extract, runs, random printable / binary, periodic, zero bytes, >=1024 repeats, crafted
256..260-byte matches), plus the integer-width edge cases of Appendix A.3 at the 64 KiB block size.
JPEG cases cover the reference's only real image input (a crop of Assets/Images/og.png whose height
is not a multiple of 8), seeded noise (Experiment/random_image.c), the Appendix C
block and degenerate blocks (single-symbol alphabets, flat colours, partial tiles).
"""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def corpus() -> np.ndarray:
    return np.fromfile(os.path.join(GOLDEN, "Metamorphosis.txt"), dtype=np.uint8)


def synth_text(n: int, seed: int = 42, passage: int = 30000) -> np.ndarray:
    """random_extract-style text (Experiment/random_extract.c:36,49-53) with the repo's splitmix64."""
    c = corpus()
    out = np.empty(n, dtype=np.uint8)
    o = 0
    M = (1 << 64) - 1
    state = int(seed)
    while o < n:
        state = (state + 0x9E3779B97F4A7C15) & M
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        start = z % (c.size - passage)
        take = min(passage, n - o)
        seg = c[start:start + take].copy()
        seg[(seg == 10) | (seg == 13)] = 32
        out[o:o + take] = seg
        o += take
    return out


def synth_image(seed: int, w: int, h: int) -> np.ndarray:
    """random_image-style noise (Experiment/random_image.c:58-77): 6 bytes of one splitmix64 draw -> 2 pixels."""
    npx = w * h
    ndraw = (npx + 1) // 2
    idx = np.arange(1, ndraw + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    b = np.stack([(z >> np.uint64(8 * k)) & np.uint64(0xFF) for k in range(6)], axis=1).astype(np.uint8)
    rgb = b.reshape(-1, 3)[:npx]
    out = np.empty((npx, 4), dtype=np.uint8)
    out[:, :3] = rgb
    out[:, 3] = 255
    return out.reshape(h, w, 4)


def _wrap_case(k: int) -> np.ndarray:
    """2048-byte block = R[0:600] + R[0:k] + one differing byte + filler (SURVEY.md A.5 wrap cases):
    position 600 has exactly one match, at position 0, of length exactly k."""
    rng = np.random.default_rng(1000 + k)
    R = rng.integers(32, 127, size=600, dtype=np.uint8)
    blk = np.empty(2048, dtype=np.uint8)
    blk[:600] = R
    blk[600:600 + k] = R[:k]
    blk[600 + k] = 200
    blk[601 + k:] = rng.integers(128, 255, size=2048 - 601 - k, dtype=np.uint8)
    return blk


def lz4_cases():
    """Yields (name, data uint8[n], block_len)."""
    rng = np.random.default_rng(42)
    out = []
    out.append(("golden_input", np.fromfile(os.path.join(GOLDEN, "lz4_input.txt"), dtype=np.uint8), 300))
    out.append(("extract_30000", synth_text(30000, seed=7), 300))
    runs = np.concatenate([np.full(700, 65, np.uint8), np.full(5, 66, np.uint8), np.full(1300, 65, np.uint8),
                           rng.integers(97, 123, size=40, dtype=np.uint8), np.full(955, 32, np.uint8)])
    out.append(("long_runs", runs, 300))
    out.append(("long_runs_b3000", runs, 3000))
    out.append(("random_printable_900", rng.integers(32, 127, size=900, dtype=np.uint8), 300))
    out.append(("random_binary_1500", rng.integers(0, 256, size=1500, dtype=np.uint8), 300))
    out.append(("periodic_text", np.frombuffer((b"the quick brown fox " * 90)[:1777], dtype=np.uint8).copy(), 300))
    z = rng.integers(0, 4, size=2000, dtype=np.uint8)
    out.append(("zero_containing", z, 300))
    base = synth_text(3000, seed=11)
    rep = np.concatenate([base, synth_text(2000, seed=12), base[:2500], synth_text(1700, seed=13), base[200:2900],
                          synth_text(388, seed=14)])
    out.append(("repeats_ge1024_b4096", rep[:12288], 4096))
    for k in (255, 256, 257, 258, 259, 260, 511, 512, 513, 515, 516):
        out.append((f"wrap_{k}", _wrap_case(k), 2048))
    out.append(("same_byte_2500", np.full(2500, 120, np.uint8), 2500))
    out.append(("tiny_1", np.array([65], np.uint8), 300))
    out.append(("tiny_3", np.array([65, 65, 65], np.uint8), 300))
    out.append(("tiny_4", np.array([65, 65, 65, 65], np.uint8), 300))
    out.append(("tiny_9", np.frombuffer(b"abcabcabc", dtype=np.uint8).copy(), 300))
    out.append(("exact_multiple", synth_text(1200, seed=3), 300))
    out.append(("last_block_1", synth_text(601, seed=4), 300))
    # literal-run length edge cases (A.3-c): 14/15/16, 269/270/271 literals before a match
    for lit in (14, 15, 16, 269, 270, 271, 526):
        pre = rng.permutation(np.arange(256, dtype=np.uint8).repeat(3))[:lit]
        # ensure no 4-gram repeats inside the random prefix: bytes from a permutation are near-unique
        tail = np.frombuffer(b"MATCHME-MATCHME-", dtype=np.uint8)
        blk = np.concatenate([tail, pre, tail])
        out.append((f"lit_{lit}", blk.astype(np.uint8), 1024))
    # 64 KiB block-size cases (Appendix A.3 header wraps: nseq u8, block byte_size u16)
    out.append(("metamorphosis_64k", corpus(), 65536))
    out.append(("synth_64k_x3", synth_text(3 * 65536, seed=42), 65536))
    out.append(("random_65536", rng.integers(0, 256, size=65536, dtype=np.uint8), 65536))
    out.append(("random_65535", rng.integers(0, 256, size=65535, dtype=np.uint8), 65536))
    out.append(("random_65536_plus", rng.integers(0, 256, size=65536 + 100, dtype=np.uint8), 65536))
    return out


def og_crop() -> np.ndarray:
    from PIL import Image

    return np.array(Image.open(os.path.join(GOLDEN, "og_crop.png")).convert("RGBA"))


def appendix_c_block() -> np.ndarray:
    i = np.arange(64)
    px = np.stack([(17 * i + 3) % 256, (29 * i + 101) % 256, (53 * i + 7) % 256, np.full(64, 255)], axis=1)
    return px.astype(np.uint8).reshape(8, 8, 4)


def jpeg_cases():
    """Yields (name, rgba uint8[h,w,4])."""
    out = []
    out.append(("og_crop", og_crop()))
    out.append(("noise_64x48", synth_image(42, 64, 48)))
    out.append(("noise_16x8", synth_image(7, 16, 8)))
    out.append(("appendix_c", appendix_c_block()))
    flat = np.zeros((16, 16, 4), np.uint8)
    flat[..., 3] = 255
    out.append(("black_16x16", flat.copy()))
    w = flat.copy()
    w[..., :3] = 255
    out.append(("white_16x16", w))
    red = flat.copy()[:16, :16]
    red[..., 0] = 255
    out.append(("red_24x16", np.concatenate([red, red[:, :8]], axis=1)))
    yy, xx = np.mgrid[0:40, 0:24]
    grad = np.stack([(xx * 10) % 256, (yy * 6) % 256, ((xx + yy) * 5) % 256, np.full_like(xx, 255)], axis=2)
    out.append(("gradient_24x40", grad.astype(np.uint8)))
    out.append(("noise_2x2", synth_image(9, 2, 2)))
    out.append(("noise_10x6", synth_image(10, 10, 6)))
    out.append(("noise_18x13", synth_image(11, 18, 13)))
    g128 = flat.copy()
    g128[..., :3] = 128
    out.append(("grey128_8x8", g128[:8, :8]))
    return out
