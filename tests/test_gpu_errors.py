"""GPU: error behaviour of the C ABI (the reference perror()+exit(1)s; the library returns codes, include/lz4jpeg_b200.h)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import lz4jpeg_b200 as ljb
    from lz4jpeg_b200 import _native as N

    ctx = ljb.Context(0)
    yield ljb, N, N.lib(), ctx
    ctx.close()


def test_lz4_argument_errors(env):
    ljb, N, lib, ctx = env
    a = np.zeros(1000, np.uint8)
    out = np.zeros(10000, np.uint8)
    ol = C.c_size_t(0)
    assert lib.ljb_lz4_compress(ctx.handle, a.ctypes.data, 0, 300, out.ctypes.data, out.size, None, C.byref(ol), None) == N.E_ARG
    assert lib.ljb_lz4_compress(ctx.handle, a.ctypes.data, 1000, 0, out.ctypes.data, out.size, None, C.byref(ol), None) == N.E_ARG
    assert lib.ljb_lz4_compress(ctx.handle, a.ctypes.data, 1000, 65537, out.ctypes.data, out.size, None, C.byref(ol), None) == N.E_ARG
    assert lib.ljb_lz4_compress(None, a.ctypes.data, 1000, 300, out.ctypes.data, out.size, None, C.byref(ol), None) == N.E_ARG
    with pytest.raises(ValueError):  # extract_uncompressed_file refuses inputs shorter than a block (LZ4.c:632-637)
        ljb.lz4.lz4_encode(a[:100], 300, ctx=ctx)


def test_lz4_capacity_error_reports_needed_size(env):
    ljb, N, lib, ctx = env
    data = ljb.synth.random_extract(200000, seed=3)
    full = ljb.lz4.lz4_encode(data, 65536, ctx=ctx)
    out = np.zeros(full.stream.size - 1000, np.uint8)
    ol = C.c_size_t(0)
    rc = lib.ljb_lz4_compress(ctx.handle, data.ctypes.data, data.size, 65536, out.ctypes.data, out.size, None, C.byref(ol), None)
    assert rc == N.E_CAPACITY
    again = ljb.lz4.lz4_encode(data, 65536, ctx=ctx)  # the context stays usable after an error
    assert np.array_equal(again.stream, full.stream)


def test_lz4_decoder_rejects_inconsistent_stream(env):
    ljb, N, lib, ctx = env
    data = ljb.synth.random_extract(70000, seed=5)
    f = ljb.lz4.lz4_encode(data, 65536, ctx=ctx)
    bad = f.stream.copy()
    bad[int(f.block_offsets[0]) + 4] ^= 0xFF  # corrupt the first sequence's size field
    broken = ljb.lz4.LZ4Frame(bad, f.block_offsets, f.block_length, f.input_size, 0)
    with pytest.raises(ljb.LjbError) as e:
        ljb.lz4.LZ4_decode(broken, ctx=ctx)
    assert e.value.code in (N.E_FORMAT, N.E_CAPACITY)


def test_jpeg_argument_errors(env):
    ljb, N, lib, ctx = env
    img = np.zeros((16, 16, 4), np.uint8)
    out = np.zeros(100000, np.uint8)
    ol = C.c_size_t(0)
    enc = lambda w, h, stride, g0, ng: lib.ljb_jpeg_encode_rgba(ctx.handle, img.ctypes.data, w, h, stride, g0, ng, out.ctypes.data,
                                                               out.size, None, None, None, C.byref(ol))
    assert enc(15, 16, 64, 0, 4) == N.E_ARG      # odd width: the reference reads past its subsampled rows (JPEG.c:543)
    assert enc(16, 16, 32, 0, 4) == N.E_ARG      # stride smaller than a row
    assert enc(16, 16, 64, 0, 5) == N.E_ARG      # more groups than ceil(w*h/64)
    assert enc(16, 16, 64, 0, 0) == N.E_ARG
    assert enc(16, 16, 64, 0, 4) == N.OK


def test_jpeg_capacity_error(env):
    ljb, N, lib, ctx = env
    img = ljb.synth.random_image(64, 64, seed=1)
    out = np.zeros(100, np.uint8)
    ol = C.c_size_t(0)
    rc = lib.ljb_jpeg_encode_rgba(ctx.handle, img.ctypes.data, 64, 64, 256, 0, 64, out.ctypes.data, out.size, None, None, None,
                                  C.byref(ol))
    assert rc == N.E_CAPACITY
    assert ljb.jpeg.process(img, ctx=ctx).stream.size > 100  # still usable


def test_context_rejects_bad_device(env):
    ljb, N, lib, ctx = env
    h = C.c_void_p()
    assert lib.ljb_ctx_create(99, C.byref(h)) == N.E_ARG
    assert lib.ljb_strerror(N.E_CAPACITY) == b"output buffer too small"
