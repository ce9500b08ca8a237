"""End-to-end (pinned host buffers, H2D + kernel + D2H) timing of the two host-buffer entry points."""
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
import lz4jpeg_b200 as ljb
from lz4jpeg_b200 import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
ctx = ljb.Context(0); lib = N.lib()
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); ljb.synth.random_extract(n, seed=42, out=h_in.numpy())
nb = n // 65536; cap = n + n // 8 + 16 * nb + 4096
h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True); h_offs = torch.empty(nb + 1, dtype=torch.int64, pin_memory=True)
ol = C.c_size_t(0); ph = C.c_uint64(0)
for i in range(3):
    t0 = time.perf_counter()
    rc = lib.ljb_lz4_compress(ctx.handle, h_in.data_ptr(), n, 65536, h_out.data_ptr(), cap, h_offs.data_ptr(), C.byref(ol), C.byref(ph))
    dt = time.perf_counter() - t0
    print(f"lz4 e2e {n>>20} MiB rc={rc}: {dt*1e3:.1f} ms  {n/dt/1e9:.2f} GB/s  out={ol.value}")
del h_in, h_out
ng = ljb.jpeg.group_count(dim, dim); jcap = ng * 96 + 4096
hj = torch.empty((dim, dim, 4), dtype=torch.uint8, pin_memory=True); ljb.synth.random_image(dim, dim, seed=42, out=hj.numpy())
hjo = torch.empty(jcap, dtype=torch.uint8, pin_memory=True); hjf = torch.empty(ng + 1, dtype=torch.int64, pin_memory=True)
for i in range(3):
    t0 = time.perf_counter()
    rc = lib.ljb_jpeg_encode_rgba(ctx.handle, hj.data_ptr(), dim, dim, 4 * dim, 0, ng, hjo.data_ptr(), jcap, hjf.data_ptr(), None, None, C.byref(ol))
    dt = time.perf_counter() - t0
    print(f"jpeg e2e {dim}x{dim} rc={rc}: {dt*1e3:.1f} ms  {dim*dim/dt/1e6:.0f} MPix/s  out={ol.value}")
