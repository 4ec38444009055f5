"""ctypes binding of the raster half of oracle/liboracle.so (oracle/raster.h) — test infrastructure only."""
import ctypes as C

import numpy as np

from tests.oracle_ffi import lib

_vp, _i, _u32, _f = C.c_void_p, C.c_int32, C.c_uint32, C.c_float
f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)

MOVE, LINE, QUAD, CUBIC, CLOSE = 0, 1, 2, 3, 4

BLEND = {n: i for i, n in enumerate([
    "clear", "source", "destination", "source_over", "destination_over", "source_in", "destination_in",
    "source_out", "destination_out", "source_atop", "destination_atop", "xor", "plus", "modulate", "screen",
    "overlay", "darken", "lighten", "color_dodge", "color_burn", "hard_light", "soft_light", "difference",
    "exclusion", "multiply", "hue", "saturation", "color", "luminosity"])}
SPREAD = {"pad": 0, "reflect": 1, "repeat": 2}
QUALITY = {"nearest": 0, "bilinear": 1, "bicubic": 2}
IDENTITY = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)


class Paint(C.Structure):
    _fields_ = [
        ("shader", _i), ("color", _f * 4),
        ("x0", _f), ("y0", _f), ("r0", _f), ("x1", _f), ("y1", _f), ("r1", _f),
        ("n_stops", _i), ("stops", f32p), ("spread", _i), ("ts", _f * 6),
        ("pattern", _vp), ("pattern_w", _u32), ("pattern_h", _u32), ("quality", _i), ("opacity", _f),
        ("blend_mode", _i), ("anti_alias", _i), ("force_hq", _i),
    ]


def _sig(name, res, args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args


_sig("orc_fill_path", _i, [_vp, _u32, _u32, _vp, _i, _vp, _i, C.POINTER(Paint), _i, f32p])
_sig("orc_fill_paths", _i, [_vp, _u32, _u32, _i, _vp, _vp, _vp, _vp, _vp, _vp, f32p])
_sig("orc_blit_coverage", _i, [_vp, _u32, _u32, _i, _vp, C.POINTER(Paint), f32p])
_sig("orc_fill_rect", _i, [_vp, _u32, _u32, _f, _f, _f, _f, C.POINTER(Paint), f32p])
_sig("orc_draw_pixmap", _i, [_vp, _u32, _u32, _i, _i, _vp, _u32, _u32, _f, _i, _i, f32p])
_sig("orc_pixmap_fill", None, [_vp, _u32, _u32, _f, _f, _f, _f])
_sig("orc_mask_from_pixmap", None, [_vp, _u32, _u32, _i, _vp])
_sig("orc_mask_invert", None, [_vp, _u32, _u32])
_sig("orc_apply_mask", None, [_vp, _u32, _u32, _vp])
_sig("orc_mask_fill_path", _i, [_vp, _u32, _u32, _vp, _i, _vp, _i, _i, _i, f32p])
_sig("orc_path_coverage", _i, [_vp, _u32, _u32, _vp, _i, _vp, _i, _i, _i, f32p])


def ts_arr(ts):
    return (C.c_float * 6)(*[float(v) for v in ts])


def make_paint(spec, blend="source_over", anti_alias=True, keep=None):
    """spec: dict describing tiny_skia::Paint's shader (see tests/scene.py).  `keep` collects buffers that must
    outlive the returned struct."""
    p = Paint()
    p.blend_mode = BLEND[blend] if isinstance(blend, str) else int(blend)
    p.anti_alias = 1 if anti_alias else 0
    p.ts[:] = IDENTITY
    kind = spec["kind"]
    if kind == "solid":
        p.shader = 0
        p.color[:] = [float(np.float32(c)) for c in spec["color"]]
    elif kind in ("linear", "radial"):
        p.shader = 1 if kind == "linear" else 2
        p.x0, p.y0, p.x1, p.y1 = spec["x0"], spec["y0"], spec["x1"], spec["y1"]
        p.r0, p.r1 = spec.get("r0", 0.0), spec.get("r1", 0.0)
        stops = np.ascontiguousarray(spec["stops"], dtype=np.float32).reshape(-1, 5)
        if keep is not None:
            keep.append(stops)
        p.n_stops = stops.shape[0]
        p.stops = stops.ctypes.data_as(f32p)
        p.spread = SPREAD[spec.get("spread", "pad")]
        p.ts[:] = spec.get("ts", IDENTITY)
        p._stops_keepalive = stops
    elif kind == "pattern":
        p.shader = 3
        pix = np.ascontiguousarray(spec["pixmap"], dtype=np.uint8)
        if keep is not None:
            keep.append(pix)
        p.pattern = pix.ctypes.data
        p.pattern_h, p.pattern_w = pix.shape[0], pix.shape[1]
        p.spread = SPREAD[spec.get("spread", "repeat")]
        p.quality = QUALITY[spec.get("quality", "bicubic")]
        p.opacity = spec.get("opacity", 1.0)
        p.ts[:] = spec.get("ts", IDENTITY)
        p._pix_keepalive = pix
    else:
        raise ValueError(kind)
    return p


def _path_arrays(verbs, pts):
    v = np.ascontiguousarray(verbs, dtype=np.uint8)
    p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
    return v, p


def fill_path(px, verbs, pts, paint, rule="nonzero", ts=IDENTITY):
    """In place on px (h, w, 4) uint8."""
    v, p = _path_arrays(verbs, pts)
    h, w = px.shape[:2]
    return lib.orc_fill_path(px.ctypes.data, w, h, v.ctypes.data, len(v), p.ctypes.data, len(p), C.byref(paint),
                             1 if rule == "evenodd" else 0, ts_arr(ts))


def blit_coverage(px, blits, paint, ts=IDENTITY):
    """blits: (n, 3) int32 {x, y, alpha}, blended in order."""
    b = np.ascontiguousarray(blits, dtype=np.int32).reshape(-1, 3)
    h, w = px.shape[:2]
    return lib.orc_blit_coverage(px.ctypes.data, w, h, len(b), b.ctypes.data, C.byref(paint), ts_arr(ts))


def fill_rect(px, x, y, rw, rh, paint, ts=IDENTITY):
    h, w = px.shape[:2]
    return lib.orc_fill_rect(px.ctypes.data, w, h, x, y, rw, rh, C.byref(paint), ts_arr(ts))


def draw_pixmap(dst, x, y, src, opacity=1.0, blend="source_over", quality="nearest", ts=IDENTITY):
    src = np.ascontiguousarray(src)
    dh, dw = dst.shape[:2]
    sh, sw = src.shape[:2]
    b = BLEND[blend] if isinstance(blend, str) else int(blend)
    return lib.orc_draw_pixmap(dst.ctypes.data, dw, dh, int(x), int(y), src.ctypes.data, sw, sh, float(opacity), b,
                               QUALITY[quality], ts_arr(ts))


def pixmap_fill(px, r, g, b, a):
    h, w = px.shape[:2]
    lib.orc_pixmap_fill(px.ctypes.data, w, h, r, g, b, a)


def mask_from_pixmap(px, kind="alpha"):
    h, w = px.shape[:2]
    m = np.zeros((h, w), dtype=np.uint8)
    lib.orc_mask_from_pixmap(np.ascontiguousarray(px).ctypes.data, w, h, 1 if kind == "luminance" else 0, m.ctypes.data)
    return m


def mask_invert(m):
    h, w = m.shape
    lib.orc_mask_invert(m.ctypes.data, w, h)


def apply_mask(px, m):
    h, w = px.shape[:2]
    assert m.shape == (h, w)
    lib.orc_apply_mask(px.ctypes.data, w, h, np.ascontiguousarray(m).ctypes.data)


def mask_fill_path(m, verbs, pts, rule="nonzero", anti_alias=True, ts=IDENTITY):
    v, p = _path_arrays(verbs, pts)
    h, w = m.shape
    return lib.orc_mask_fill_path(m.ctypes.data, w, h, v.ctypes.data, len(v), p.ctypes.data, len(p),
                                  1 if rule == "evenodd" else 0, 1 if anti_alias else 0, ts_arr(ts))


def path_coverage(w, h, verbs, pts, rule="nonzero", anti_alias=True, ts=IDENTITY):
    v, p = _path_arrays(verbs, pts)
    cov = np.zeros((h, w), dtype=np.uint8)
    lib.orc_path_coverage(cov.ctypes.data, w, h, v.ctypes.data, len(v), p.ctypes.data, len(p),
                          1 if rule == "evenodd" else 0, 1 if anti_alias else 0, ts_arr(ts))
    return cov
