"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): few blocks / groups, all entry points."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
ctx = ljb.Context(0)
text = ljb.synth.random_extract(2 * 65536 + 777, seed=1)
f = ljb.lz4.lz4_encode(text, 65536, ctx=ctx)
assert np.array_equal(ljb.lz4.LZ4_decode(f, ctx=ctx), text)
f2 = ljb.lz4.lz4_encode(text[:5000], 300, ctx=ctx)
assert np.array_equal(ljb.lz4.LZ4_decode(f2, ctx=ctx), text[:5000])
rnd = np.random.default_rng(0).integers(0, 256, 65536, dtype=np.uint8)
ljb.lz4.lz4_encode(rnd, 65536, ctx=ctx)
ljb.lz4.lz4_encode(np.full(70000, 65, np.uint8), 65536, ctx=ctx)
ln, ds = ljb.lz4.find_longest_match(text[:65536], ctx=ctx)
img = ljb.synth.random_image(136, 52, seed=2)
enc = ljb.jpeg.process(img, ctx=ctx)
rec = ljb.jpeg.assemble_image(enc.coefs, 136, 52, original=img, ctx=ctx)
part = ljb.jpeg.process(img, first_group=3, ngroups=50, ctx=ctx)
print("sanitize_small ok", f.stream.size, f2.stream.size, enc.stream.size, rec.shape)
ctx.close()
