"""GPU: BASELINE.json configs[4] in miniature — a batch of 1920x1080 random_image-style frames through the JPEG encoder,
then an LZ4 round trip of every frame's bit stream (per-image seed = 42 + index)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batched_1080p_encode_and_lz4_round_trip(oracle):
    import lz4jpeg_b200 as ljb

    ctx = ljb.Context(0)
    try:
        for index in range(3):
            img = ljb.synth.random_image(1920, 1080, seed=42 + index)
            enc = ljb.jpeg.process(img, want_coefs=False, ctx=ctx)
            assert enc.group_offsets.size == 240 * 135 + 1
            # a slice of the frame against the oracle (the whole frame would take the CPU checker ~15 s)
            ref = oracle.jpeg_encode(img, 12000, 12600, want_coefs=False)
            o0, o1 = int(enc.group_offsets[12000]), int(enc.group_offsets[12600])
            assert np.array_equal(enc.stream[o0:o1], ref["stream"])
            assert np.array_equal(enc.group_bits[12000:12600], ref["bits"])
            # record framing properties over the whole frame
            d = np.diff(enc.group_offsets.astype(np.int64))
            assert np.array_equal(d, (enc.group_bits.astype(np.int64).sum(1) + 7) // 8)
            # LZ4 round trip of the bit stream at 64 KiB blocks
            frame = ljb.lz4.lz4_encode(enc.stream, 65536, ctx=ctx)
            assert frame.phantom == 0
            back = ljb.lz4.LZ4_decode(frame, ctx=ctx)
            assert np.array_equal(back, enc.stream)
    finally:
        ctx.close()
