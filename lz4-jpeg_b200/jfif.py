"""Host-side mirror of the baseline-JPEG writer the reference vendors (stb_image_write.h v1.16) over the C ABI.

  stbi_write_jpg_to_func / stbi_write_jpg_core   stb_image_write.h:1607 / :1398-1605   write_jpg(pixels, quality) -> bytes

All compute (colour conversion, AAN DCT, quantisation, Huffman coding, bit packing, byte stuffing) runs in
lz4-jpeg_b200/csrc/jfif_encode.cu; the output is byte-identical to stb's.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

HEADER_BYTES = 607


def bound(w: int, h: int) -> int:
    return int(N.lib().ljb_jfif_bound(w, h))


def _pixels(px) -> tuple[np.ndarray, int, int, int]:
    a = np.ascontiguousarray(px, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3 or not 1 <= a.shape[2] <= 4:
        raise ValueError("expect an H x W [x comp] uint8 array, comp = 1..4 (stb's `comp`)")
    return a, a.shape[0], a.shape[1], a.shape[2]


def write_jpg(pixels, quality: int = 90, subsample: int = -1, out_cap: int | None = None,
              ctx: N.Context | None = None) -> np.ndarray:
    """stbi_write_jpg_to_func(func, ctx, w, h, comp, data, quality): returns the .jpg file as a uint8 array.
    subsample: -1 = stb's rule (4:2:0 when quality <= 90), 0 = 4:4:4, 1 = 4:2:0."""
    a, h, w, comp = _pixels(pixels)
    ctx = ctx or N.default_context()
    cap = int(out_cap) if out_cap is not None else min(bound(w, h), HEADER_BYTES + 2 + 3 * w * h + 4096)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = N.lib().ljb_jfif_encode(ctx.handle, a.ctypes.data, w, h, comp, w * comp, quality, subsample, out.ctypes.data, cap, C.byref(n))
    if rc == N.E_CAPACITY and out_cap is None:
        cap = bound(w, h)
        out = np.empty(cap, dtype=np.uint8)
        rc = N.lib().ljb_jfif_encode(ctx.handle, a.ctypes.data, w, h, comp, w * comp, quality, subsample, out.ctypes.data, cap,
                                     C.byref(n))
    N.check(rc, "ljb_jfif_encode")
    return out[: n.value].copy()


def encode_device(d_pixels, w: int, h: int, comp: int, quality: int, subsample: int, d_out, d_result, ctx: N.Context,
                  stride: int | None = None, d_coefs=None) -> None:
    """Asynchronous on ctx.stream; d_* are torch CUDA tensors (pointers only)."""
    rc = N.lib().ljb_jfif_encode_dev(ctx.handle, d_pixels.data_ptr(), w, h, comp, stride if stride is not None else w * comp, quality,
                                     subsample, d_out.data_ptr(), d_out.numel(), d_result.data_ptr(),
                                     d_coefs.data_ptr() if d_coefs is not None else None)
    N.check(rc, "ljb_jfif_encode_dev")
