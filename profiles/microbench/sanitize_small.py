"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): few blocks /
groups, all entry points: LZ4 chain search and full search, LZ4 decoder, JPEG encoder (single, sub-range, batch, sample groups),
JPEG tree + entropy-decode + reconstruction kernels, JFIF encoder (4:4:4 and 4:2:0)."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
small = len(sys.argv) > 1 and sys.argv[1] == "small"   # racecheck of the 1024-thread LZ4 kernels is slow: fewer, smaller blocks
ctx = ljb.Context(0)
text = ljb.synth.random_extract((1 if small else 2) * 65536 + 777, seed=1)
f = ljb.lz4.lz4_encode(text, 65536, ctx=ctx)
if f.phantom == 0:
    assert np.array_equal(ljb.lz4.LZ4_decode(f, ctx=ctx), text)
f2 = ljb.lz4.lz4_encode(text[:5000], 300, ctx=ctx)
if f2.phantom == 0:
    assert np.array_equal(ljb.lz4.LZ4_decode(f2, ctx=ctx), text[:5000])
rnd = np.random.default_rng(0).integers(0, 256, 65536, dtype=np.uint8)
ljb.lz4.lz4_encode(rnd, 65536, ctx=ctx)
ljb.lz4.lz4_encode(np.full(70000, 65, np.uint8), 65536, ctx=ctx)
ljb.lz4.lz4_encode(np.tile(np.array([97, 98], np.uint8), 20000), 40000, ctx=ctx)
ln, ds = ljb.lz4.find_longest_match(text[:20000 if small else 65536], ctx=ctx)
img = ljb.synth.random_image(136, 52, seed=2)
enc = ljb.jpeg.process(img, ctx=ctx)
trees = ljb.jpeg.huffman_trees(enc.coefs, ctx=ctx)
assert np.array_equal(ljb.jpeg.decode_huffman(enc, trees, ctx=ctx), enc.coefs)
rec = ljb.jpeg.assemble_image(enc.coefs, 136, 52, original=img, ctx=ctx)
part = ljb.jpeg.process(img, first_group=3, ngroups=50, ctx=ctx)
batch = ljb.jpeg.process_batch(np.stack([ljb.synth.random_image(30, 20, seed=s) for s in range(3)]), ctx=ctx)
batch8 = ljb.jpeg.process_batch(np.stack([ljb.synth.random_image(64, 48, seed=s) for s in range(3)]), ctx=ctx)
smp, coefs = ljb.jpeg.process_groups(np.random.default_rng(1).integers(0, 256, (40, 128), dtype=np.uint8), ctx=ctx)
j444 = ljb.jfif.write_jpg(img, 75, 0, ctx=ctx)
j420 = ljb.jfif.write_jpg(img, 75, -1, ctx=ctx)
print("sanitize_small ok", f.stream.size, f2.stream.size, enc.stream.size, rec.shape, batch.stream.size, j444.size, j420.size)
ctx.close()
