"""Host-side mirror of the reference's JPEG interface (Algorithms/sequential/JPEG/JPEG.c) over the C ABI.

reference stages fused in one kernel                                 here
  build_{luminance,rChrominance,bChrominance}_matrix  JPEG.c:114-185
  chroma_subsample / divide_image                     JPEG.c:302 / :496
  discrete_cosine_transform / Quantize                JPEG.c:451 / :621
  zigzag_pattern / RLE                                JPEG.c:693 / :767
  encode_huffman / generate_encoded_sequence          JPEG.c:1035 / :993    process(rgba) -> EncodedImage
(`process` is the reference's fused per-block function, Algorithms/parallel/JPEG/JPEG.c:1103.)
inverse chain, entropy half (JPEG.c:1253-1403)
  decode_huffman / inverse_RLE / reverse_zigzag_pattern   JPEG.c:1009 / :811 / :729
                                                                            huffman_trees(coefs) -> trees
                                                                            decode_huffman(enc, trees) -> coefs
decode half (JPEG.c:1408-1428)
  Inverse_quantize / inverse_discrete_cosine_transform / assemble_image   JPEG.c:631 / :399 / :552
                                                                            assemble_image(coefs, w, h) -> RGBA
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N

GROUP_SIZE = 8


@dataclass
class EncodedImage:
    stream: np.ndarray         # uint8: per group lum|Cr|Cb code bits, MSB first, group records byte aligned
    group_offsets: np.ndarray  # uint64[ngroups+1]
    group_bits: np.ndarray     # uint16[ngroups,3]: bit lengths of the lum, r, b strings
    coefs: np.ndarray | None   # int16[ngroups,128]: quantised lum[64], Cr[32], Cb[32] (row-major, before zig-zag)
    width: int
    height: int
    first_group: int

    def bit_string(self, g: int, channel: int) -> str:
        """The '0'/'1' string generate_encoded_sequence (JPEG.c:993) builds for group g, channel 0/1/2 = lum/r/b."""
        o = int(self.group_offsets[g])
        nb = [int(x) for x in self.group_bits[g]]
        rec = self.stream[o:int(self.group_offsets[g + 1])]
        bits = "".join(f"{b:08b}" for b in rec)
        start = sum(nb[:channel])
        return bits[start:start + nb[channel]]


def group_count(w: int, h: int) -> int:
    """Groups the reference processes: ceil(w*h/64) (JPEG.c:1131)."""
    return int(N.lib().ljb_jpeg_group_count(w, h))


def process(rgba, first_group: int = 0, ngroups: int | None = None, want_coefs: bool = True,
            ctx: N.Context | None = None) -> EncodedImage:
    """Encode the 8x8 groups [first_group, first_group+ngroups) of an H x W x 4 (r g b a) or H x W x 3 (r g b) uint8 image."""
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] not in (3, 4):
        raise ValueError("expect an H x W x 4 or H x W x 3 uint8 array (the reference's Pixel, JPEG.c:29-32)")
    h, w, bpp = a.shape
    total = group_count(w, h)
    if ngroups is None:
        ngroups = total - first_group
    ctx = ctx or N.default_context()
    cap = int(N.lib().ljb_jpeg_bound(ngroups))
    out = np.empty(cap, dtype=np.uint8)
    offs = np.zeros(ngroups + 1, dtype=np.uint64)
    bits = np.zeros((ngroups, 3), dtype=np.uint16)
    coefs = np.zeros((ngroups, 128), dtype=np.int16) if want_coefs else None
    out_len = C.c_size_t(0)
    fn = N.lib().ljb_jpeg_encode_rgba if bpp == 4 else N.lib().ljb_jpeg_encode_rgb
    rc = fn(ctx.handle, a.ctypes.data, w, h, bpp * w, first_group, ngroups, out.ctypes.data, cap, offs.ctypes.data, bits.ctypes.data,
            coefs.ctypes.data if want_coefs else None, C.byref(out_len))
    N.check(rc, "ljb_jpeg_encode_rgba" if bpp == 4 else "ljb_jpeg_encode_rgb")
    return EncodedImage(out[: out_len.value].copy(), offs, bits, coefs, w, h, first_group)


def encode_device(d_rgba, w: int, h: int, d_out, d_group_offsets, d_group_bits, d_result, ctx: N.Context,
                  first_group: int = 0, ngroups: int | None = None, d_coefs=None, bpp: int = 4) -> None:
    """Asynchronous on ctx.stream; d_* are torch CUDA tensors (pointers only).  bpp: bytes per pixel of d_rgba, 4 or 3."""
    if ngroups is None:
        ngroups = group_count(w, h) - first_group
    fn = N.lib().ljb_jpeg_encode_rgba_dev if bpp == 4 else N.lib().ljb_jpeg_encode_rgb_dev
    rc = fn(ctx.handle, d_rgba.data_ptr(), w, h, bpp * w, first_group, ngroups, d_out.data_ptr(), d_out.numel(), d_group_offsets.data_ptr(),
            d_group_bits.data_ptr() if d_group_bits is not None else None, d_coefs.data_ptr() if d_coefs is not None else None,
            d_result.data_ptr())
    N.check(rc, "ljb_jpeg_encode_rgba_dev" if bpp == 4 else "ljb_jpeg_encode_rgb_dev")


def assemble_image(coefs, w: int, h: int, original=None, ctx: N.Context | None = None) -> np.ndarray:
    """Quantised coefficients (EncodedImage.coefs of the WHOLE image) -> the reference's reconstructed.png pixels:
    Inverse_quantize, inverse_discrete_cosine_transform, assemble_image (JPEG.c:631, :399, :552).
    `original` (H x W x 4) is needed when w or h is not a multiple of 8: the reference leaves the last tiled
    groups unprocessed (JPEG.c:1131) and shows their colour-converted original samples."""
    c = np.ascontiguousarray(coefs, dtype=np.int16)
    if c.size != group_count(w, h) * 128:
        raise ValueError("coefs must hold 128 values for each of the ceil(w*h/64) groups")
    o = None
    if original is not None:
        o = np.ascontiguousarray(original, dtype=np.uint8)
        if o.shape != (h, w, 4):
            raise ValueError("original must be H x W x 4")
    out = np.empty((h, w, 4), dtype=np.uint8)
    ctx = ctx or N.default_context()
    rc = N.lib().ljb_jpeg_decode_coefs(ctx.handle, c.ctypes.data, w, h, o.ctypes.data if o is not None else None, 4 * w,
                                       out.ctypes.data, 4 * w)
    N.check(rc, "ljb_jpeg_decode_coefs")
    return out


TREE_BYTES = 1024  # LJB_JPEG_TREE_BYTES


def huffman_trees(coefs, ctx: N.Context | None = None) -> np.ndarray:
    """The Huffman tree of every (group, channel) — what calculate_frequency / build_heap / build_huffman_tree build
    (JPEG.c:864-961) and the reference keeps in memory for its decoder — serialised at TREE_BYTES per group."""
    c = np.ascontiguousarray(coefs, dtype=np.int16).reshape(-1, 128)
    ctx = ctx or N.default_context()
    trees = np.zeros((c.shape[0], TREE_BYTES), dtype=np.uint8)
    N.check(N.lib().ljb_jpeg_trees(ctx.handle, c.ctypes.data, c.shape[0], trees.ctypes.data), "ljb_jpeg_trees")
    return trees


def decode_huffman(enc: EncodedImage, trees, ctx: N.Context | None = None) -> np.ndarray:
    """decode_huffman -> inverse_RLE -> reverse_zigzag_pattern (JPEG.c:1009, :811, :729) of every group's bit strings:
    returns the quantised coefficients int16[ngroups, 128] recovered from the packed stream alone (+ trees)."""
    ctx = ctx or N.default_context()
    ng = enc.group_offsets.size - 1
    s = np.ascontiguousarray(enc.stream, dtype=np.uint8)
    offs = np.ascontiguousarray(enc.group_offsets, dtype=np.uint64)
    bits = np.ascontiguousarray(enc.group_bits, dtype=np.uint16)
    t = np.ascontiguousarray(trees, dtype=np.uint8)
    out = np.zeros((ng, 128), dtype=np.int16)
    N.check(N.lib().ljb_jpeg_entropy_decode(ctx.handle, s.ctypes.data, s.size, offs.ctypes.data, bits.ctypes.data, t.ctypes.data, ng,
                                            out.ctypes.data), "ljb_jpeg_entropy_decode")
    return out


def process_batch(images, ctx: N.Context | None = None) -> EncodedImage:
    """N equal-sized images (N x H x W x 4 or N x H x W x 3 uint8) in one call (ljb_jpeg_encode_batch / _rgb): image i holds groups
    [i*G, (i+1)*G)."""
    a = np.ascontiguousarray(images, dtype=np.uint8)
    if a.ndim != 4 or a.shape[3] not in (3, 4):
        raise ValueError("expect an N x H x W x 4 or N x H x W x 3 uint8 array")
    n, h, w, bpp = a.shape
    ctx = ctx or N.default_context()
    G = group_count(w, h)
    cap = int(N.lib().ljb_jpeg_bound(n * G))
    out = np.empty(cap, dtype=np.uint8)
    offs = np.zeros(n * G + 1, dtype=np.uint64)
    bits = np.zeros((n * G, 3), dtype=np.uint16)
    out_len = C.c_size_t(0)
    fn = N.lib().ljb_jpeg_encode_batch if bpp == 4 else N.lib().ljb_jpeg_encode_batch_rgb
    rc = fn(ctx.handle, a.ctypes.data, w, h, bpp * w, bpp * w * h, n, out.ctypes.data, cap, offs.ctypes.data, bits.ctypes.data, C.byref(out_len))
    N.check(rc, "ljb_jpeg_encode_batch" if bpp == 4 else "ljb_jpeg_encode_batch_rgb")
    return EncodedImage(out[: out_len.value].copy(), offs, bits, None, w, h, 0)


def encode_batch_device(d_rgba, w: int, h: int, nimages: int, d_out, d_group_offsets, d_group_bits, d_result, ctx: N.Context,
                        bpp: int = 4) -> None:
    """Asynchronous on ctx.stream: nimages back-to-back images in one launch (ljb_jpeg_encode_batch_dev / _rgb_dev)."""
    fn = N.lib().ljb_jpeg_encode_batch_dev if bpp == 4 else N.lib().ljb_jpeg_encode_batch_rgb_dev
    rc = fn(ctx.handle, d_rgba.data_ptr(), w, h, bpp * w, bpp * w * h, nimages, d_out.data_ptr(), d_out.numel(), d_group_offsets.data_ptr(),
            d_group_bits.data_ptr() if d_group_bits is not None else None, None, d_result.data_ptr())
    N.check(rc, "ljb_jpeg_encode_batch_dev" if bpp == 4 else "ljb_jpeg_encode_batch_rgb_dev")


def process_groups(samples, ctx: N.Context | None = None):
    """The reference's process() (Algorithms/parallel/JPEG/JPEG.c:1103) on groups given by their samples (uint8[n,128]: lum 64 | b 32
    | r 32): returns (reconstructed samples uint8[n,128], quantised coefficients int16[n,128])."""
    s = np.ascontiguousarray(samples, dtype=np.uint8).reshape(-1, 128).copy()
    ctx = ctx or N.default_context()
    coefs = np.zeros((s.shape[0], 128), dtype=np.int16)
    N.check(N.lib().ljb_jpeg_process_groups(ctx.handle, s.ctypes.data, s.shape[0], coefs.ctypes.data), "ljb_jpeg_process_groups")
    return s, coefs
