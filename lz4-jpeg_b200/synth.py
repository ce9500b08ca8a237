"""Workload generators of the reference's harnesses (Experiment/random_extract.c, Experiment/random_image.c),
host code in csrc/synth.c with an explicit seed."""
from __future__ import annotations

import os

import numpy as np

from . import _native as N

CORPUS_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "Metamorphosis.txt")


def corpus() -> np.ndarray:
    """The text the reference samples (Output-Input/input/Metamorphosis.txt), committed as a fixture."""
    return np.fromfile(CORPUS_PATH, dtype=np.uint8)


def random_extract(n: int, seed: int = 42, passage: int = 30000, out: np.ndarray | None = None) -> np.ndarray:
    c = corpus()
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n and out.flags["C_CONTIGUOUS"]
    N.lib().ljb_synth_text(c.ctypes.data, c.size, seed, passage, out.ctypes.data, n)
    return out


def random_image(w: int, h: int, seed: int = 42, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.empty((h, w, 4), dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= 4 * w * h and out.flags["C_CONTIGUOUS"]
    N.lib().ljb_synth_image(seed, w, h, out.ctypes.data)
    return out
