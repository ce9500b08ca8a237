"""ctypes binding of liblz4jpeg_b200.so (the C ABI in include/lz4jpeg_b200.h).

There is deliberately no fallback: if the shared object is missing or no B200 is usable, every call
raises.  Nothing in this package imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblz4jpeg_b200.so")

OK, E_ARG, E_CUDA, E_CAPACITY, E_FORMAT, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_i16p = C.POINTER(C.c_int16)
_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); also the list the symbol-export test checks against the header
SIGNATURES = {
    "ljb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ljb_ctx_destroy": (None, [C.c_void_p]),
    "ljb_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "ljb_strerror": (C.c_char_p, [C.c_int]),
    "ljb_last_cuda_error": (C.c_char_p, []),
    "ljb_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "ljb_ctx_last_kernel_ms": (C.c_float, [C.c_void_p]),
    "ljb_lz4_bound": (C.c_size_t, [C.c_size_t, C.c_size_t]),
    "ljb_lz4_block_count": (C.c_size_t, [C.c_size_t, C.c_size_t]),
    "ljb_lz4_compress": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, _szp, _u64p]),
    "ljb_lz4_compress_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_size_t]),
    "ljb_lz4_block_matches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ljb_lz4_decompress": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                     C.c_size_t, _szp]),
    "ljb_lz4_decompress_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_void_p]),
    "ljb_jpeg_group_count": (C.c_size_t, [C.c_int, C.c_int]),
    "ljb_jpeg_bound": (C.c_size_t, [C.c_size_t]),
    "ljb_jpeg_encode_rgba": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                       C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, _szp]),
    "ljb_jpeg_encode_rgba_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ljb_jpeg_encode_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, _szp]),
    "ljb_jpeg_encode_rgb_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t,
                                          C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ljb_jpeg_encode_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_void_p, _szp]),
    "ljb_jpeg_encode_batch_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_void_p, _szp]),
    "ljb_jpeg_encode_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ljb_jpeg_encode_batch_rgb_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ljb_jpeg_encode_groups_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    "ljb_jpeg_decode_groups_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ljb_jpeg_process_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ljb_jpeg_trees": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ljb_jpeg_trees_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ljb_jpeg_entropy_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ljb_jpeg_entropy_decode_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64,
                                              C.c_void_p, C.c_void_p]),
    "ljb_jpeg_decode_coefs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "ljb_jpeg_decode_coefs_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_void_p]),
    "ljb_jfif_bound": (C.c_size_t, [C.c_int, C.c_int]),
    "ljb_jfif_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                  C.c_size_t, _szp]),
    "ljb_jfif_encode_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                      C.c_size_t, C.c_void_p, C.c_void_p]),
    "ljb_synth_text": (None, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_size_t, C.c_void_p, C.c_size_t]),
    "ljb_synth_image": (None, [C.c_uint64, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


class LjbError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        lib = _lib
        msg = lib.ljb_strerror(code).decode() if lib is not None else str(code)
        if code == E_CUDA and lib is not None:
            msg += ": " + lib.ljb_last_cuda_error().decode()
        super().__init__(f"{what}: {msg} ({code})")


def lib() -> C.CDLL:
    """Load the shared object (built by lz4-jpeg_b200/build.py).  Raises if it does not exist."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` — there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int, what: str) -> None:
    if code != OK:
        raise LjbError(code, what)


class Context:
    """One per (process, GPU): owns the CUDA stream and the persistent device scratch."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self.device = device
        check(lib().ljb_ctx_create(device, C.byref(self._h)), "ljb_ctx_create")

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("context destroyed")
        return self._h

    @property
    def stream(self) -> int:
        return int(lib().ljb_ctx_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib().ljb_ctx_launch_count(self.handle))

    def last_kernel_ms(self) -> float:
        return float(lib().ljb_ctx_last_kernel_ms(self.handle))

    def close(self) -> None:
        if self._h:
            lib().ljb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
