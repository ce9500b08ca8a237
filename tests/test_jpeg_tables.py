"""The DCT basis constants embedded in the CUDA source are the doubles this libm's cos()/sqrt() return for the
reference's expressions (JPEG.c:481-482, :487-488)."""
import importlib.util
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "lz4-jpeg_b200", "csrc", "jpeg_tables.inc")


def _parse(name, text):
    body = re.search(name + r"\[\d+\] = \{(.*?)\};", text, flags=re.S).group(1)
    return np.array([float.fromhex(x) for x in re.findall(r"-?0x[0-9a-f.]+p[+-]?\d+", body)])


def test_embedded_tables_equal_libm(oracle):
    text = open(INC).read()
    cos8, cos4, a8, a4 = oracle.jpeg_basis()
    assert np.array_equal(_parse("kCos8", text), cos8)
    assert np.array_equal(_parse("kCos4", text), cos4)
    assert np.array_equal(_parse("kAlpha8", text), a8)
    assert np.array_equal(_parse("kAlpha4", text), a4)
    assert (a8[0] * a8[0]).hex() == "0x1.0000000000001p-3"  # alpha0^2 is not exactly 1/8


def test_generator_reproduces_the_committed_file():
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "lz4-jpeg_b200", "tools", "gen_jpeg_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert gen.render() == open(INC).read()
