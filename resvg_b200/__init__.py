"""resvg_b200 — B200-native implementation of resvg's pixel hot path (rasteriser + filters).

The compute lives in ``libresvg_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/resvg_b200.h``).  This package is the thin host-side mirror used by tests and bench.py; it
raises at import time if the library is missing — there is no CPU fallback.
"""
from . import _ffi  # noqa: F401  (raises ImportError if the CUDA library is not built)
from . import tree  # noqa: F401
from .api import (Batch, Context, Layer, Mask, PinnedBuffer, ResvgB200Error, apply_mask,  # noqa: F401
                  draw_layer, draw_layer_rects, fill_path, filters, make_light, make_paint, make_transfer, stroke_path, dash_path, hairline_blits)

__all__ = ["tree", "Batch", "Context", "Layer", "Mask", "PinnedBuffer", "ResvgB200Error", "apply_mask", "draw_layer", "draw_layer_rects",
           "fill_path", "filters", "make_light", "make_paint", "make_transfer", "stroke_path", "dash_path", "hairline_blits"]
