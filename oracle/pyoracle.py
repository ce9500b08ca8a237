"""oracle/pyoracle.py — ctypes bindings of the CPU checkers (TEST INFRASTRUCTURE).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (lz4-jpeg_b200/) never imports this module.

Two families with the same call shapes:
  * ``Oracle``  — this repo's C restatement (oracle/liboracle.so, built by oracle/build.py)
  * ``Ref``     — the reference's own sources compiled into oracle/_ref/ (present where the build ran
                  with /root/reference mounted; travels to the GPU box as prebuilt .so files)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)


def _p(a: np.ndarray, ty=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(ty))


def _as_u8(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a, dtype=np.uint8)


def lz4_bound(n: int, block_len: int) -> int:
    """Generous capacity for the reference dialect: worst case is 5 header bytes per input byte."""
    nblocks = (n + block_len - 1) // block_len
    return 1 + 3 * nblocks + 6 * n + 64


class _LZ4Mixin:
    _compress_name = ""

    def lz4_compress(self, data, block_len: int, mode: int = 0):
        """Returns (stream bytes as np.uint8, block_offsets np.uint64[nblocks+1], phantom_count|None)."""
        a = _as_u8(data)
        n = a.size
        nblocks = (n + block_len - 1) // block_len
        out = np.empty(lz4_bound(n, block_len), dtype=np.uint8)
        offs = np.zeros(nblocks + 1, dtype=np.uint64)
        out_len = C.c_size_t(0)
        fn = getattr(self.lib, self._compress_name)
        if self._compress_name == "oracle_lz4_compress":
            ph = C.c_uint64(0)
            rc = fn(_p(a), C.c_size_t(n), C.c_size_t(block_len), _p(out), C.c_size_t(out.size),
                    _p(offs, C.c_uint64), C.byref(out_len), C.byref(ph), C.c_int(mode))
            phantom = int(ph.value)
        else:
            rc = fn(_p(a), C.c_size_t(n), C.c_size_t(block_len), _p(out), C.c_size_t(out.size),
                    _p(offs, C.c_uint64), C.byref(out_len))
            phantom = None
        if rc != 0:
            raise RuntimeError(f"{self._compress_name} rc={rc}")
        return out[: out_len.value].copy(), offs, phantom


class Oracle(_LZ4Mixin):
    _compress_name = "oracle_lz4_compress"

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            from . import build  # type: ignore

            build.build_restatement()
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_lz4_compress.restype = C.c_int
        L.oracle_lz4_matches.restype = C.c_int
        L.oracle_lz4_decompress.restype = C.c_int
        L.oracle_jpeg_encode.restype = C.c_int
        L.oracle_jpeg_group_count.restype = C.c_size_t
        L.oracle_jpeg_group_stages.restype = C.c_int
        L.oracle_jpeg_decode.restype = C.c_int
        L.oracle_synth_text.restype = None
        L.oracle_synth_image.restype = None
        L.oracle_jpeg_planes.restype = None
        L.oracle_jpeg_basis.restype = None
        L.oracle_jfif_encode.restype = C.c_int
        L.oracle_jfif_unit_count.restype = C.c_size_t
        L.oracle_jfif_tables.restype = None
        L.oracle_jfif_huffman.restype = None

    # ---- LZ4 -------------------------------------------------------------------------------
    def lz4_matches(self, block, mode: int = 0):
        a = _as_u8(block)
        ln = np.zeros(a.size, dtype=np.uint16)
        ds = np.zeros(a.size, dtype=np.uint16)
        rc = self.lib.oracle_lz4_matches(_p(a), C.c_size_t(a.size), _p(ln, C.c_uint16), _p(ds, C.c_uint16), C.c_int(mode))
        if rc != 0:
            raise RuntimeError(f"oracle_lz4_matches rc={rc}")
        return ln, ds

    def lz4_decompress(self, comp, block_offsets, block_len: int, out_cap: int):
        c = _as_u8(comp)
        offs = np.ascontiguousarray(block_offsets, dtype=np.uint64)
        out = np.empty(max(out_cap, 1), dtype=np.uint8)
        out_len = C.c_size_t(0)
        rc = self.lib.oracle_lz4_decompress(_p(c), C.c_size_t(c.size), _p(offs, C.c_uint64), C.c_size_t(offs.size - 1),
                                            C.c_size_t(block_len), _p(out), C.c_size_t(out_cap), C.byref(out_len))
        return rc, out[: out_len.value].copy()

    def synth_text(self, corpus, seed: int, passage: int, n: int) -> np.ndarray:
        c = _as_u8(corpus)
        out = np.empty(n, dtype=np.uint8)
        self.lib.oracle_synth_text(_p(c), C.c_size_t(c.size), C.c_uint64(seed), C.c_size_t(passage), _p(out), C.c_size_t(n))
        return out

    # ---- JPEG ------------------------------------------------------------------------------
    def synth_image(self, seed: int, w: int, h: int) -> np.ndarray:
        out = np.empty((h, w, 4), dtype=np.uint8)
        self.lib.oracle_synth_image(C.c_uint64(seed), C.c_int(w), C.c_int(h), _p(out))
        return out

    def jpeg_group_count(self, w: int, h: int) -> int:
        return int(self.lib.oracle_jpeg_group_count(C.c_int(w), C.c_int(h)))

    def jpeg_planes(self, rgba: np.ndarray):
        return _jpeg_planes(self.lib.oracle_jpeg_planes, rgba)

    def jpeg_encode(self, rgba: np.ndarray, g0: int = 0, g1: int | None = None, want_coefs: bool = True):
        return _jpeg_encode(self.lib.oracle_jpeg_encode, rgba, g0, g1, want_coefs)

    def jpeg_group_stages(self, rgba: np.ndarray, g: int):
        return _jpeg_stages(self.lib.oracle_jpeg_group_stages, rgba, g)

    def jpeg_decode(self, coefs: np.ndarray, w: int, h: int, orig: np.ndarray | None = None):
        """Quantised coefficients -> reconstructed RGBA (Inverse_quantize, IDCT, assemble_image)."""
        return _jpeg_decode(self.lib.oracle_jpeg_decode, coefs, w, h, orig)

    # ---- baseline JFIF (stb_image_write's encoder) -------------------------------------------
    def jfif_encode(self, px: np.ndarray, quality: int, force_subsample: int = -1, want_coefs: bool = False):
        """px: H x W (grey), H x W x 2/3/4 uint8.  Returns the .jpg bytes (and the quantised data units, zig-zag order)."""
        a, h, w, comp = _check_pixels(px)
        cap = jfif_bound(w, h)
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        units = int(self.lib.oracle_jfif_unit_count(C.c_int(w), C.c_int(h), C.c_int(quality), C.c_int(force_subsample)))
        coefs = np.zeros((units, 64), dtype=np.int16) if want_coefs else None
        rc = self.lib.oracle_jfif_encode(_p(a), C.c_int(w), C.c_int(h), C.c_int(comp), C.c_size_t(w * comp), C.c_int(quality),
                                         C.c_int(force_subsample), _p(out), C.c_size_t(cap), C.byref(n),
                                         _p(coefs, C.c_int16) if want_coefs else None)
        if rc != 0:
            raise RuntimeError(f"oracle_jfif_encode rc={rc}")
        return (out[: n.value].copy(), coefs) if want_coefs else out[: n.value].copy()

    def jfif_tables(self, quality: int):
        qy = np.zeros(64, np.uint8)
        quv = np.zeros(64, np.uint8)
        dy = np.zeros(64, np.float32)
        duv = np.zeros(64, np.float32)
        rank = np.zeros(64, np.uint8)
        self.lib.oracle_jfif_tables(C.c_int(quality), _p(qy), _p(quv), _p(dy, C.c_float), _p(duv, C.c_float), _p(rank))
        return qy, quv, dy, duv, rank

    def jfif_huffman(self, which: int):
        code = np.zeros(256, np.uint16)
        ln = np.zeros(256, np.uint8)
        self.lib.oracle_jfif_huffman(C.c_int(which), _p(code, C.c_uint16), _p(ln))
        return code, ln

    def jpeg_basis(self):
        cos8 = np.zeros(64)
        cos4 = np.zeros(16)
        a8 = np.zeros(2)
        a4 = np.zeros(2)
        self.lib.oracle_jpeg_basis(_p(cos8, C.c_double), _p(cos4, C.c_double), _p(a8, C.c_double), _p(a4, C.c_double))
        return cos8, cos4, a8, a4


def jfif_bound(w: int, h: int) -> int:
    """Worst case of the baseline stream: 27 bits per coefficient, every byte stuffed, three full-size components."""
    units = ((w + 7) // 8) * ((h + 7) // 8) * 3
    return 607 + 2 + units * 64 * 27 // 8 * 2 + 64


def _check_pixels(px: np.ndarray):
    a = np.ascontiguousarray(px, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    assert a.ndim == 3 and 1 <= a.shape[2] <= 4, "expect H x W [x comp] uint8"
    return a, a.shape[0], a.shape[1], a.shape[2]


def _check_rgba(rgba: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 4, "expect H x W x 4 uint8"
    return a


def _jpeg_planes(fn, rgba):
    a = _check_rgba(rgba)
    h, w, _ = a.shape
    Y = np.empty((h, w), dtype=np.uint8)
    Cr = np.empty((h, w), dtype=np.uint8)
    Cb = np.empty((h, w), dtype=np.uint8)
    fn(_p(a), C.c_int(w), C.c_int(h), C.c_size_t(4 * w), _p(Y), _p(Cr), _p(Cb))
    return Y, Cr, Cb


def _jpeg_encode(fn, rgba, g0, g1, want_coefs):
    a = _check_rgba(rgba)
    h, w, _ = a.shape
    total = (w * h + 63) // 64
    if g1 is None:
        g1 = total
    ng = g1 - g0
    coefs = np.zeros((ng, 128), dtype=np.int16) if want_coefs else None
    cap = ng * 512 + 64
    out = np.zeros(cap, dtype=np.uint8)
    offs = np.zeros(ng + 1, dtype=np.uint64)
    bits = np.zeros((ng, 3), dtype=np.uint16)
    out_len = C.c_size_t(0)
    maxlen = C.c_int(0)
    rc = fn(_p(a), C.c_int(w), C.c_int(h), C.c_size_t(4 * w), C.c_size_t(g0), C.c_size_t(g1),
            _p(coefs, C.c_int16) if want_coefs else None, _p(out), C.c_size_t(cap), _p(offs, C.c_uint64),
            _p(bits, C.c_uint16), C.byref(out_len), C.byref(maxlen))
    if rc != 0:
        raise RuntimeError(f"jpeg encode rc={rc}")
    return {"stream": out[: out_len.value].copy(), "offsets": offs, "bits": bits, "coefs": coefs,
            "max_code_len": int(maxlen.value)}


def _jpeg_decode(fn, coefs, w, h, orig):
    c = np.ascontiguousarray(coefs, dtype=np.int16)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    o = _check_rgba(orig) if orig is not None else None
    rc = fn(_p(c, C.c_int16), C.c_int(w), C.c_int(h), _p(o) if o is not None else None, C.c_size_t(4 * w), _p(out),
            C.c_size_t(4 * w))
    if rc != 0:
        raise RuntimeError(f"jpeg decode rc={rc}")
    return out


def _jpeg_stages(fn, rgba, g):
    a = _check_rgba(rgba)
    h, w, _ = a.shape
    samples = np.zeros(128, dtype=np.uint8)
    coef = np.zeros(128, dtype=np.float64)
    rle = np.zeros(3 * 128, dtype=np.int32)
    rle_len = (C.c_size_t * 3)()
    rc = fn(_p(a), C.c_int(w), C.c_int(h), C.c_size_t(4 * w), C.c_size_t(g), _p(samples), _p(coef, C.c_double),
            _p(rle, C.c_int), rle_len)
    if rc != 0:
        raise RuntimeError(f"jpeg stages rc={rc}")
    rles = [rle[128 * c: 128 * c + rle_len[c]].copy() for c in range(3)]
    return samples, coef, rles


class Ref(_LZ4Mixin):
    """The reference's own code (oracle/_ref).  ``Ref.available()`` says whether the .so files exist."""

    _compress_name = "ref_lz4_compress"

    @staticmethod
    def paths() -> dict[str, str]:
        d = os.path.join(HERE, "_ref")
        return {k: os.path.join(d, f) for k, f in
                (("lz4", "libref_lz4.so"), ("lz4_verbatim", "libref_lz4_verbatim.so"), ("jpeg", "libref_jpeg.so"),
                 ("jfif", "libref_jfif.so"), ("lz4_par", "libref_lz4_par.so"), ("jpeg_par", "libref_jpeg_par.so"))}

    @staticmethod
    def available(which: str = "lz4") -> bool:
        return os.path.exists(Ref.paths()[which])

    def __init__(self, which: str = "lz4"):
        self.which = which
        self.lib = C.CDLL(Ref.paths()[which])
        if which == "lz4_par":    # the reference's parallel build (CPU baseline only)
            self.lib.ref_lz4par_time_blocks.restype = C.c_int
        elif which == "jpeg_par":
            self.lib.ref_jpegpar_time_groups.restype = C.c_int
        elif which.startswith("lz4"):
            self.lib.ref_lz4_compress.restype = C.c_int
            self.lib.ref_lz4_time_blocks.restype = C.c_int
        elif which == "jfif":
            self.lib.ref_jfif_encode.restype = C.c_int
            self.lib.ref_jfif_time_mt.restype = C.c_double
        else:
            self.lib.ref_jpeg_encode.restype = C.c_int
            self.lib.ref_jpeg_planes.restype = C.c_int
            self.lib.ref_jpeg_group_stages.restype = C.c_int
            self.lib.ref_jpeg_time_groups.restype = C.c_int
            self.lib.ref_jpeg_decode.restype = C.c_int

    def lz4_time_blocks(self, data, block_len: int, nthreads: int):
        a = _as_u8(data)
        sec = C.c_double(0)
        nb = C.c_uint64(0)
        self.lib.ref_lz4_time_blocks(_p(a), C.c_size_t(a.size), C.c_size_t(block_len), C.c_int(nthreads), C.byref(sec), C.byref(nb))
        return sec.value, int(nb.value)

    def lz4par_time_blocks(self, data, block_len: int, nthreads: int):
        """parallel_block_encode (with its global locks) over a pool of `nthreads` workers; (seconds, sum of block byte sizes)."""
        a = _as_u8(data)
        sec = C.c_double(0)
        nb = C.c_uint64(0)
        rc = self.lib.ref_lz4par_time_blocks(_p(a), C.c_size_t(a.size), C.c_size_t(block_len), C.c_int(nthreads), C.byref(sec), C.byref(nb))
        if rc != 0:
            raise RuntimeError(f"ref_lz4par_time_blocks rc={rc}")
        return sec.value, int(nb.value)

    def jpegpar_time_groups(self, rgba, nthreads: int):
        """process() (forward + inverse chain, results discarded as in the reference) over a pool of `nthreads` workers."""
        a = _check_rgba(rgba)
        h, w, _ = a.shape
        sec = C.c_double(0)
        self.lib.ref_jpegpar_time_groups(_p(a), C.c_int(w), C.c_int(h), C.c_size_t(4 * w), C.c_int(nthreads), C.byref(sec))
        return sec.value

    def jfif_encode(self, px: np.ndarray, quality: int, force_subsample: int = -1) -> np.ndarray:
        """stbi_write_jpg_to_func of the reference's vendored stb_image_write.h on tightly packed pixels."""
        a, h, w, comp = _check_pixels(px)
        cap = jfif_bound(w, h)
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self.lib.ref_jfif_encode(_p(a), C.c_int(w), C.c_int(h), C.c_int(comp), C.c_int(quality), C.c_int(force_subsample),
                                      _p(out), C.c_size_t(cap), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"ref_jfif_encode rc={rc}")
        return out[: n.value].copy()

    def jfif_time_mt(self, px: np.ndarray, quality: int, force_subsample: int, nthreads: int, reps: int = 1):
        a, h, w, comp = _check_pixels(px)
        nb = C.c_size_t(0)
        sec = self.lib.ref_jfif_time_mt(_p(a), C.c_int(w), C.c_int(h), C.c_int(comp), C.c_int(quality), C.c_int(force_subsample),
                                        C.c_int(nthreads), C.c_int(reps), C.byref(nb))
        return float(sec), int(nb.value)

    def jpeg_planes(self, rgba):
        return _jpeg_planes(self.lib.ref_jpeg_planes, rgba)

    def jpeg_encode(self, rgba, g0: int = 0, g1: int | None = None, want_coefs: bool = True):
        return _jpeg_encode(self.lib.ref_jpeg_encode, rgba, g0, g1, want_coefs)

    def jpeg_group_stages(self, rgba, g: int):
        return _jpeg_stages(self.lib.ref_jpeg_group_stages, rgba, g)

    def jpeg_decode(self, coefs, w: int, h: int, orig):
        return _jpeg_decode(self.lib.ref_jpeg_decode, coefs, w, h, orig)

    def jpeg_time_groups(self, rgba, nthreads: int):
        a = _check_rgba(rgba)
        h, w, _ = a.shape
        sec = C.c_double(0)
        nb = C.c_uint64(0)
        self.lib.ref_jpeg_time_groups(_p(a), C.c_int(w), C.c_int(h), C.c_size_t(4 * w), C.c_int(nthreads), C.byref(sec), C.byref(nb))
        return sec.value, int(nb.value)
