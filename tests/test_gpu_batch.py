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


def test_batch_entry_point_equals_per_image_calls():
    """ljb_jpeg_encode_batch: N images in one call; image i's records are byte for byte what a call for that image alone gives.
    1920x1080 (sides multiples of 8: one launch over the whole batch) and 30x20 (not: tail groups differ per image)."""
    import lz4jpeg_b200 as ljb

    ctx = ljb.Context(0)
    try:
        for (w, h, n) in ((1920, 1080, 4), (30, 20, 5), (64, 48, 7)):
            imgs = np.stack([ljb.synth.random_image(w, h, seed=42 + i) for i in range(n)])
            enc = ljb.jpeg.process_batch(imgs, ctx=ctx)
            G = ljb.jpeg.group_count(w, h)
            assert enc.group_offsets.size == n * G + 1
            for i in range(n):
                one = ljb.jpeg.process(imgs[i], want_coefs=False, ctx=ctx)
                o0, o1 = int(enc.group_offsets[i * G]), int(enc.group_offsets[(i + 1) * G])
                assert np.array_equal(enc.stream[o0:o1], one.stream), (w, h, i)
                assert np.array_equal(enc.group_bits[i * G:(i + 1) * G], one.group_bits)
                assert np.array_equal(enc.group_offsets[i * G:(i + 1) * G + 1] - np.uint64(o0), one.group_offsets)
            rgb = ljb.jpeg.process_batch(np.ascontiguousarray(imgs[:, :, :, :3]), ctx=ctx)  # the same frames as r g b, three bytes per pixel
            assert np.array_equal(rgb.stream, enc.stream) and np.array_equal(rgb.group_offsets, enc.group_offsets), (w, h)
            assert np.array_equal(rgb.group_bits, enc.group_bits)
    finally:
        ctx.close()


def test_batch_device_pipeline_jpeg_then_lz4_round_trip():
    """configs[4] on one GPU, device-resident end to end: batch JPEG encode -> LZ4 compress of the concatenated bit streams ->
    LZ4 decompress == the bit streams.  Also a batch whose image sides are not multiples of 8 through the one-launch device path."""
    import torch

    import lz4jpeg_b200 as ljb

    ctx = ljb.Context(0)
    try:
        for (w, h, n) in ((1920, 1080, 6), (30, 20, 9)):
            imgs = np.stack([ljb.synth.random_image(w, h, seed=142 + i) for i in range(n)])
            G = ljb.jpeg.group_count(w, h)
            d_in = torch.from_numpy(imgs).cuda()
            cap = n * G * 96 + 4096
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            d_offs = torch.empty(n * G + 1, dtype=torch.int64, device="cuda")
            d_bits = torch.empty(n * G * 3, dtype=torch.int16, device="cuda")
            d_res = torch.zeros(3, dtype=torch.int64, device="cuda")
            ljb.jpeg.encode_batch_device(d_in, w, h, n, d_out, d_offs, d_bits, d_res, ctx)
            torch.cuda.synchronize()
            assert int(d_res[2].item()) == 0
            jlen = int(d_res[0].item())
            offs = d_offs.cpu().numpy().astype(np.int64)
            stream = d_out[:jlen].cpu().numpy()
            for i in range(n):
                one = ljb.jpeg.process(imgs[i], want_coefs=False, ctx=ctx)
                assert np.array_equal(stream[offs[i * G]:offs[(i + 1) * G]], one.stream), (w, h, i)
            if jlen < 65536:
                continue
            # LZ4 of the concatenated bit streams, 64 KiB blocks, all on the device
            nb = (jlen + 65535) // 65536
            d_lz = torch.empty(jlen + jlen // 4 + 16 * nb + 4096, dtype=torch.uint8, device="cuda")
            d_boffs = torch.empty(nb + 1, dtype=torch.int64, device="cuda")
            d_lres = torch.zeros(3, dtype=torch.int64, device="cuda")
            ljb.lz4.compress_device(d_out[:jlen], 65536, d_lz, d_boffs, d_lres, ctx)
            torch.cuda.synchronize()
            assert int(d_lres[2].item()) == 0 and int(d_lres[1].item()) == 0
            d_back = torch.empty(nb * 65536, dtype=torch.uint8, device="cuda")
            d_blen = torch.empty(nb, dtype=torch.int32, device="cuda")
            d_dres = torch.zeros(3, dtype=torch.int64, device="cuda")
            ljb.lz4.decompress_device(d_lz, int(d_lres[0].item()), d_boffs, nb, 65536, d_back, d_blen, d_dres, ctx)
            torch.cuda.synchronize()
            assert int(d_dres[2].item()) == 0 and int(d_dres[0].item()) == jlen
            assert torch.equal(d_back[:jlen], d_out[:jlen])
    finally:
        ctx.close()
