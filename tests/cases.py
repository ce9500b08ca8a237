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
    # huge 8-gram groups (> 512 occurrences of one 8-gram inside a block): the search's budgeted walks and group walks
    unit = np.frombuffer(b"ABCDEFGHI", dtype=np.uint8)
    big = np.concatenate([np.concatenate([unit, rng.integers(0, 256, size=5, dtype=np.uint8)]) for _ in range(1400)])
    out.append(("big_group_19600", big, big.size))
    out.append(("zeros_4096", np.zeros(4096, np.uint8), 4096))
    out.append(("period2_4096", np.tile(np.array([97, 98], np.uint8), 2048), 4096))
    out.append(("period7_6000", np.tile(np.frombuffer(b"abcdefg", dtype=np.uint8), 858)[:6000].copy(), 6000))
    # back-to-back 4-byte matches: 33 sequences per 132 bytes (more than a parse segment parks for the eight-lane passes)
    words = rng.integers(128, 256, size=(600, 4), dtype=np.uint8)
    sec1 = np.concatenate([np.concatenate([w, np.array([j % 128], np.uint8)]) for j, w in enumerate(words)])
    sec2 = words[rng.integers(0, 600, size=3000)].reshape(-1)
    dense = np.concatenate([sec1, sec2])
    out.append(("dense_ml4_15000", dense, dense.size))
    mixed = synth_text(16384, seed=21)
    mixed[3000:9000] = 0
    mixed[11000:13000] = np.tile(np.array([1, 2, 3], np.uint8), 667)[:2000]
    out.append(("text_with_runs_16384", mixed, 16384))
    # low-alphabet random text at 64 KiB: blocks that EXPAND (many 4-byte matches, 5 bytes of sequence overhead each), so an
    # ordinary short sequence sits 64 KiB or more before the block's end — what a size-field wrap heuristic keyed on the
    # block extent mistook for a wrapped field.  Own generator: the cases above keep their bytes.
    rng2 = np.random.default_rng(4242)
    out.append(("hex_65536", np.frombuffer(b"0123456789abcdef", dtype=np.uint8)[rng2.integers(0, 16, size=65536)].copy(), 65536))
    out.append(("base32_2x65536", np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ234567", dtype=np.uint8)[rng2.integers(0, 32, size=2 * 65536)].copy(), 65536))
    out.append(("two_symbol_65536", (rng2.integers(0, 2, size=65536) + 65).astype(np.uint8), 65536))
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


def jfif_cases():
    """Yields (name, pixels uint8[h,w(,comp)], quality, subsample) for the baseline-JPEG writer (stb_image_write.h:1398).
    subsample: -1 = stb's own rule (4:2:0 when quality <= 90), 0 / 1 = forced 4:4:4 / 4:2:0."""
    rng = np.random.default_rng(2024)
    out = []
    crop = og_crop()  # 256 x 100: the height is not a multiple of 8 or 16 (edge MCUs repeat the last row)
    out.append(("og_crop_q75", crop, 75, -1))
    out.append(("og_crop_q75_444", crop, 75, 0))
    out.append(("og_crop_q95", crop, 95, -1))
    out.append(("og_crop_rgb_q50", crop[:, :, :3].copy(), 50, -1))
    out.append(("og_crop_q0_default", crop, 0, -1))
    out.append(("noise_64x48_q75", synth_image(42, 64, 48), 75, -1))
    out.append(("noise_64x48_q75_444", synth_image(42, 64, 48), 75, 0))
    out.append(("noise_160x96_q100", synth_image(3, 160, 96), 100, -1))       # longest codes, 4:4:4 by stb's rule
    out.append(("noise_160x96_q100_420", synth_image(3, 160, 96), 100, 1))
    out.append(("noise_96x80_q1", synth_image(4, 96, 80), 1, -1))             # quantisers clamp to 255: runs of zeros, ZRL
    out.append(("noise_33x70_q90", synth_image(5, 33, 70), 90, -1))           # ragged in both directions
    out.append(("noise_17x23_rgb_q91", synth_image(6, 17, 23)[:, :, :3].copy(), 91, -1))
    out.append(("noise_1x1_q75", synth_image(8, 1, 1), 75, -1))
    out.append(("noise_1x1_q95", synth_image(8, 1, 1), 95, -1))
    out.append(("noise_8x8_q75_444", synth_image(9, 8, 8), 75, 0))
    out.append(("noise_16x16_q75", synth_image(10, 16, 16), 75, -1))
    out.append(("noise_400x8_q75", synth_image(11, 400, 8), 75, -1))          # one MCU row, several rounds
    out.append(("noise_8x400_q75_444", synth_image(12, 8, 400), 75, 0))      # one MCU column
    out.append(("grey_100x37_q80", rng.integers(0, 256, size=(37, 100), dtype=np.uint8), 80, -1))       # comp 1
    out.append(("greyalpha_64x48_q60", rng.integers(0, 256, size=(48, 64, 2), dtype=np.uint8), 60, -1))  # comp 2: alpha ignored
    yy, xx = np.mgrid[0:200, 0:300]
    grad = np.stack([(xx * 255) // 300, (yy * 255) // 200, (xx + yy) % 256, np.full_like(xx, 255)], axis=2).astype(np.uint8)
    out.append(("gradient_300x200_q30", grad, 30, -1))
    out.append(("gradient_300x200_q75", grad, 75, -1))
    out.append(("gradient_300x200_q95", grad, 95, -1))
    flat = np.zeros((64, 64, 4), np.uint8)
    flat[..., 3] = 255
    out.append(("black_64x64_q75", flat.copy(), 75, -1))
    w = flat.copy()
    w[..., :3] = 255
    out.append(("white_64x64_q75_444", w, 75, 0))
    sparse = flat.copy()
    sparse[::16, ::16, :3] = 255   # isolated bright pixels: long zero runs between a few large coefficients
    out.append(("sparse_64x64_q50", sparse, 50, -1))
    chk = flat.copy()
    chk[(np.indices((64, 64)).sum(0) & 1) == 1, :3] = 255  # checkerboard: extreme high-frequency coefficient
    out.append(("checker_64x64_q100", chk, 100, -1))
    out.append(("noise_640x360_q75", synth_image(13, 640, 360), 75, -1))      # many tiles, ragged bottom (360 = 22.5 MCUs)
    out.append(("noise_640x360_q75_444", synth_image(13, 640, 360), 75, 0))
    return out
