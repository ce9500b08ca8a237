"""lz4jpeg_b200 — B200-native (sm_100a CUDA) replacement of the two per-block compression hot paths of
CyrilMorel42/LZ4-JPEG: the custom "LZ4" block encoder and the "JPEG-like" per-8x8-group encoder.

The package is a thin host-side mirror of the reference's own function names over the C ABI declared in
include/lz4jpeg_b200.h; all compute happens in lz4-jpeg_b200/csrc/*.cu.  There is no CPU fallback.
"""
from . import _native  # noqa: F401
from ._native import Context, LjbError, default_context  # noqa: F401
from ._native import OK as LJB_OK, E_ARG as LJB_E_ARG, E_CUDA as LJB_E_CUDA, E_CAPACITY as LJB_E_CAPACITY, E_FORMAT as LJB_E_FORMAT, E_UNSUPPORTED as LJB_E_UNSUPPORTED  # noqa: F401,E501
from . import lz4  # noqa: F401
from . import jpeg  # noqa: F401
from . import jfif  # noqa: F401
from . import synth  # noqa: F401
from . import sharding  # noqa: F401

__all__ = ["Context", "LjbError", "default_context", "lz4", "jpeg", "jfif", "synth", "sharding"]
