"""Path geometry hooks of the test-side front end and of the oracle back end: the stroker, the dasher and the
anti-aliased hairline walker.  ONE place decides which implementation the checker uses, so that neither arm of a
GPU-vs-oracle comparison shares geometry code with the product.

`impl` = "oracle": oracle/stroke.c, oracle/dash.c, oracle/hairline.c (independent C restatements, test infrastructure).
"""
import numpy as np

_IMPL = None


def _impl():
    global _IMPL
    if _IMPL is None:
        from tests import oracle_geom
        _IMPL = oracle_geom
    return _IMPL


def stroke_outline(verbs, pts, width, miter_limit, cap, join, res_scale):
    """tiny_skia_path::Path::stroke -> (verbs, pts) or None."""
    return _impl().stroke_path(np.asarray(verbs, np.uint8), np.asarray(pts, np.float32).reshape(-1, 2), width, miter_limit,
                               cap, join, res_scale)


def dash_path(verbs, pts, dash_array, dash_offset, res_scale):
    """tiny_skia_path::Path::dash -> (verbs, pts) or None."""
    return _impl().dash_path(np.asarray(verbs, np.uint8), np.asarray(pts, np.float32).reshape(-1, 2), dash_array, dash_offset,
                             res_scale)


def hairline_blits(verbs, dev_pts, cap, clip_w, clip_h):
    """scan/hairline_aa.rs over a device-space path -> int32 array of (x, y, alpha) blits in walker order."""
    return _impl().hairline_blits(np.asarray(verbs, np.uint8), np.asarray(dev_pts, np.float32).reshape(-1, 2), cap, clip_w, clip_h)
