"""resvg_b200 — B200-native implementation of resvg's pixel hot path (rasteriser + filters).

The compute lives in ``libresvg_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/resvg_b200.h``).  This package is the thin host-side mirror used by tests and bench.py; it
raises at import time if the library is missing — there is no CPU fallback.
"""
from . import _ffi  # noqa: F401  (raises ImportError if the CUDA library is not built)
from .api import (Context, Layer, PinnedBuffer, ResvgB200Error, filters, make_light,  # noqa: F401
                  make_transfer)

__all__ = ["Context", "Layer", "PinnedBuffer", "ResvgB200Error", "filters", "make_light", "make_transfer"]
