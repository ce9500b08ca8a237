#!/usr/bin/env python3
"""profiles/tools/jfif_time.py — time the device-resident baseline-JPEG path alone (development aid, not the bench).

    python profiles/tools/jfif_time.py [--dim 16384] [--quality 75] [--sub -1|0|1] [--iters 10] [--natural]

Prints whole-call ms (three kernels, CUDA events on the library stream) and the encode kernel's own ms.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import lz4jpeg_b200 as ljb

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=16384)
ap.add_argument("--quality", type=int, default=75)
ap.add_argument("--sub", type=int, default=-1)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--comp", type=int, default=4, help="bytes per pixel of the input: 4 (r g b a) or 3 (r g b)")
ap.add_argument("--e2e", action="store_true", help="also time ljb_jfif_encode from pinned host memory")
ap.add_argument("--natural", action="store_true", help="tile the og.png crop instead of noise")
a = ap.parse_args()
W = H = a.dim
ctx = ljb.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
if a.natural:
    from PIL import Image

    crop = np.array(Image.open(os.path.join(ROOT, "tests", "golden", "og_crop.png")).convert("RGBA"))
    img = np.tile(crop, (H // crop.shape[0] + 1, W // crop.shape[1] + 1, 1))[:H, :W].copy()
else:
    img = ljb.synth.random_image(W, H, seed=42)
img = np.ascontiguousarray(img[:, :, :a.comp])
d_in = torch.from_numpy(img).cuda()
cap = 607 + 2 + 3 * W * H + 4096
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_res = torch.zeros(3, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
tot, ker = [], []
for i in range(a.iters + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    ljb.jfif.encode_device(d_in, W, H, a.comp, a.quality, a.sub, d_out, d_res, ctx)
    e1.record(st)
    st.synchronize()
    if i >= 2:
        tot.append(e0.elapsed_time(e1))
        ker.append(ctx.last_kernel_ms())
n = int(d_res[0].item())
flags = int(d_res[2].item())
spilled = int(d_res[1].item())
t, k = float(np.median(tot)), float(np.median(ker))
print(f"dim {W} q{a.quality} sub {a.sub} comp {a.comp} {'natural' if a.natural else 'noise'}: file {n} B ({n / (W * H):.3f} B/px) flags {flags} spilled tiles {spilled} | "
      f"whole {t:.3f} ms ({W * H / t / 1e6:.1f} GPix/s) | encode kernel {k:.3f} ms | "
      f"algorithmic {(a.comp * W * H + n) / t / 1e6:.1f} GB/s")

if a.e2e:
    import ctypes as C, time
    h_in = torch.empty(img.shape, dtype=torch.uint8, pin_memory=True)
    h_in.numpy()[...] = img
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    ln = C.c_size_t(0)
    ts = []
    for i in range(a.iters + 1):
        t0 = time.perf_counter()
        rc = ljb._native.lib().ljb_jfif_encode(ctx.handle, h_in.data_ptr(), W, H, a.comp, a.comp * W, a.quality, a.sub, h_out.data_ptr(), cap, C.byref(ln))
        ts.append(time.perf_counter() - t0)
        assert rc == 0, rc
    t = float(np.median(ts[1:])) * 1e3
    print(f"  e2e ljb_jfif_encode comp {a.comp}: {t:.2f} ms ({W * H / t / 1e6:.2f} GPix/s), file {ln.value} B")
