"""Host-side mirror of the reference's LZ4 interface (Algorithms/sequential/LZ4/LZ4.c) over the C ABI.

reference function                         here
  lz4_encode()            LZ4.c:670        lz4_encode(data, block_len)            -> LZ4Frame
  parallel_LZ4_encode()   P-LZ4:680        (same call: blocks are always encoded in parallel on the GPU)
  find_longest_match()    LZ4.c:290        find_longest_match(block)              -> (len[], dist[]) for every position
  divide_input()          LZ4.c:123        divide_input(n, block_len)             -> block extents
  LZ4_decode()            LZ4.c:1038       LZ4_decode(frame)                      -> bytes
Constants keep the reference's names (LZ4.c:20-23).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N

MAX_MATCH_LENGTH = 1024
MIN_MATCH_LENGTH = 4
WINDOW_SIZE = 65535
DEFAULT_BLOCK_LENGTH = 300
MAX_BLOCK_LENGTH = 65536


@dataclass
class LZ4Frame:
    """What write_output() serialises (LZ4.c:427-441) plus the out-of-band framing the 8/16-bit headers cannot carry."""
    stream: np.ndarray          # uint8: u8 nblocks_lo8 | blocks
    block_offsets: np.ndarray   # uint64[nblocks+1]: byte offset of every block in `stream`, and the end
    block_length: int
    input_size: int
    phantom: int                # sequences the reference format cannot represent

    @property
    def blocks(self) -> int:
        return self.block_offsets.size - 1


def _u8(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a, dtype=np.uint8)


def divide_input(input_size: int, block_size: int):
    """LZ4.c:123-177: ceil(n/B) exact-size blocks; returns [(offset, length)]."""
    count = (input_size + block_size - 1) // block_size
    return [(i * block_size, min(block_size, input_size - i * block_size)) for i in range(count)]


def bound(n: int, block_len: int) -> int:
    return int(N.lib().ljb_lz4_bound(n, block_len))


def lz4_encode(data, block_len: int = DEFAULT_BLOCK_LENGTH, ctx: N.Context | None = None, out_cap: int | None = None) -> LZ4Frame:
    """Encode `data` like lz4_encode() (LZ4.c:670-742) does for input.txt; returns the compressed.bin bytes.

    Like extract_uncompressed_file (LZ4.c:632-637) the input must be at least one block long."""
    a = _u8(data)
    if a.size == 0 or not (1 <= block_len <= MAX_BLOCK_LENGTH):
        raise ValueError("empty input or block_len outside 1..65536")
    if a.size < block_len:
        raise ValueError("Error: default block length is too high, please reduce it before proceding.")  # LZ4.c:634
    ctx = ctx or N.default_context()
    nblocks = (a.size + block_len - 1) // block_len
    cap = out_cap if out_cap is not None else min(bound(a.size, block_len), 2 * a.size + 8 * nblocks + 4096)
    out = np.empty(cap, dtype=np.uint8)
    offs = np.zeros(nblocks + 1, dtype=np.uint64)
    out_len = C.c_size_t(0)
    ph = C.c_uint64(0)
    rc = N.lib().ljb_lz4_compress(ctx.handle, a.ctypes.data, a.size, block_len, out.ctypes.data, cap, offs.ctypes.data,
                                  C.byref(out_len), C.byref(ph))
    if rc == N.E_CAPACITY and out_cap is None:  # (on LJB_E_CAPACITY out_len is only a lower bound when the input went up in chunks)
        return lz4_encode(a, block_len, ctx, out_cap=bound(a.size, block_len))
    N.check(rc, "ljb_lz4_compress")
    return LZ4Frame(out[: out_len.value].copy(), offs, block_len, int(a.size), int(ph.value))


parallel_LZ4_encode = lz4_encode  # Algorithms/parallel/LZ4/LZ4.c:680 — same result, the GPU path is always block-parallel


def find_longest_match(block, ctx: N.Context | None = None):
    """LZ4.c:290-323 for every current_index of one block at once.

    Returns (length uint16[n], distance uint16[n]); length is the true longest match (0 or 4..1024) before the
    reference's (uint8_t) cast, distance = current_index - earliest best position."""
    a = _u8(block)
    if not (1 <= a.size <= MAX_BLOCK_LENGTH):
        raise ValueError("block must hold 1..65536 bytes")
    ctx = ctx or N.default_context()
    ln = np.zeros(a.size, dtype=np.uint16)
    ds = np.zeros(a.size, dtype=np.uint16)
    N.check(N.lib().ljb_lz4_block_matches(ctx.handle, a.ctypes.data, a.size, ln.ctypes.data, ds.ctypes.data), "ljb_lz4_block_matches")
    return ln, ds


def lz4_decompress_raw(stream, block_offsets, block_length: int, out_cap: int, ctx: N.Context | None = None) -> np.ndarray:
    """ljb_lz4_decompress on a stream and an offset table given separately (what a caller holding compressed.bin and its
    table passes); raises LjbError(LJB_E_FORMAT) on a malformed stream or table."""
    ctx = ctx or N.default_context()
    out = np.empty(max(int(out_cap), 1), dtype=np.uint8)
    out_len = C.c_size_t(0)
    offs = np.ascontiguousarray(block_offsets, dtype=np.uint64)
    s = _u8(stream)
    N.check(N.lib().ljb_lz4_decompress(ctx.handle, s.ctypes.data, s.size, offs.ctypes.data, offs.size - 1, block_length,
                                       out.ctypes.data, out.size, C.byref(out_len)), "ljb_lz4_decompress")
    return out[: out_len.value]


def LZ4_decode(frame: LZ4Frame, ctx: N.Context | None = None) -> np.ndarray:
    """Decode a frame produced by lz4_encode (compute of LZ4_decode, LZ4.c:1038-1121)."""
    return lz4_decompress_raw(frame.stream, frame.block_offsets, frame.block_length, frame.input_size, ctx)


parallel_LZ4_decode = LZ4_decode  # Algorithms/parallel/LZ4/LZ4.c:1105


# ---- device-resident entry point (torch tensors by pointer; used by bench.py and the sharded driver) ----
def compress_device(d_in, block_len: int, d_out, d_block_offsets, d_result, ctx: N.Context, first_block: int = 0,
                    frame_blocks: int | None = None) -> None:
    """Asynchronous on ctx.stream.  d_* are torch CUDA tensors (uint8 / int64 views are fine: only pointers are used)."""
    n = d_in.numel()
    nblocks = (n + block_len - 1) // block_len
    if d_block_offsets.numel() * d_block_offsets.element_size() < 8 * (nblocks + 1) or d_result.numel() * d_result.element_size() < 24:
        raise ValueError("d_block_offsets needs nblocks+1 and d_result 3 64-bit slots")
    rc = N.lib().ljb_lz4_compress_dev(ctx.handle, d_in.data_ptr(), n, block_len, d_out.data_ptr(), d_out.numel(),
                                      d_block_offsets.data_ptr(), d_result.data_ptr(), first_block,
                                      nblocks if frame_blocks is None else frame_blocks)
    N.check(rc, "ljb_lz4_compress_dev")


def decompress_device(d_comp, comp_len: int, d_block_offsets, nblocks: int, block_len: int, d_out, d_block_out_len, d_result,
                      ctx: N.Context) -> None:
    """Asynchronous on ctx.stream: decode a device-resident stream (ljb_lz4_decompress_dev)."""
    rc = N.lib().ljb_lz4_decompress_dev(ctx.handle, d_comp.data_ptr(), comp_len, d_block_offsets.data_ptr(), nblocks, block_len,
                                        d_out.data_ptr(), d_out.numel(), d_block_out_len.data_ptr(), d_result.data_ptr())
    N.check(rc, "ljb_lz4_decompress_dev")
