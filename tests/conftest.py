"""tests/conftest.py — markers and shared fixtures.

``-m "not gpu"``: oracle vs golden vectors / reference build, host logic, C-ABI symbol check (CPU only).
``-m gpu``      : parity tests proper — the CUDA path called through the C ABI, checked by the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import build
    from oracle.pyoracle import Oracle

    build.build_restatement()
    return Oracle()


@pytest.fixture(scope="session")
def ref_lz4():
    from oracle.pyoracle import Ref

    if not Ref.available("lz4"):
        pytest.skip("oracle/_ref/libref_lz4.so not built (needs /root/reference)")
    return Ref("lz4")


@pytest.fixture(scope="session")
def ref_jpeg():
    from oracle.pyoracle import Ref

    if not Ref.available("jpeg"):
        pytest.skip("oracle/_ref/libref_jpeg.so not built (needs /root/reference)")
    return Ref("jpeg")
