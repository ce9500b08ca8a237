"""End-to-end (pinned host buffers) timing of ljb_jpeg_encode_rgb / _rgba; LJB_PIPE_CHUNK_BYTES sets the band size."""
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
from lz4jpeg_b200 import _native as N
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = ljb.Context(0); lib = N.lib()
ng = ljb.jpeg.group_count(dim, dim); jcap = ng * 96 + 4096
hj = torch.empty((dim, dim, 4), dtype=torch.uint8, pin_memory=True); ljb.synth.random_image(dim, dim, seed=42, out=hj.numpy())
h3 = torch.empty((dim, dim, 3), dtype=torch.uint8, pin_memory=True); h3.copy_(hj[:, :, :3])
hjo = torch.empty(jcap, dtype=torch.uint8, pin_memory=True); hjf = torch.empty(ng + 1, dtype=torch.int64, pin_memory=True)
ol = C.c_size_t(0)
for name, fn, buf, bpp in (("rgb", lib.ljb_jpeg_encode_rgb, h3, 3), ("rgba", lib.ljb_jpeg_encode_rgba, hj, 4)):
    ts = []
    for i in range(5):
        t0 = time.perf_counter()
        rc = fn(ctx.handle, buf.data_ptr(), dim, dim, bpp * dim, 0, ng, hjo.data_ptr(), jcap, hjf.data_ptr(), None, None, C.byref(ol))
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts[1:]))
    print(f"jpeg e2e {name} {dim}x{dim} rc={rc}: {dt*1e3:.2f} ms  {dim*dim/dt/1e6:.0f} MPix/s  out={ol.value}")
