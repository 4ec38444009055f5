"""ctypes binding of the oracle's path geometry (oracle/stroke.c, dash.c, hairline.c) — test infrastructure only."""
import ctypes as C

import numpy as np

from tests.oracle_ffi import lib

CAPS = {"butt": 0, "round": 1, "square": 2}
JOINS = {"miter": 0, "miter-clip": 1, "round": 2, "bevel": 3}
_vp, _i, _f = C.c_void_p, C.c_int32, C.c_float
_vpp, _ip = C.POINTER(C.c_void_p), C.POINTER(C.c_int32)

lib.orc_path_stroke.restype = _i
lib.orc_path_stroke.argtypes = [_vp, _i, _vp, _i, _f, _f, _i, _i, _f, _vpp, _ip, _vpp, _ip]
lib.orc_geom_free.restype = None
lib.orc_geom_free.argtypes = [_vp]


def _take_path(ov, nv, op, npt):
    try:
        v = np.ctypeslib.as_array((C.c_uint8 * nv.value).from_address(ov.value)).copy()
        p = np.ctypeslib.as_array((C.c_float * (max(npt.value, 1) * 2)).from_address(op.value)).copy()[: npt.value * 2].reshape(-1, 2)
    finally:
        lib.orc_geom_free(ov)
        lib.orc_geom_free(op)
    return v, p


def stroke_path(verbs, pts, width, miter_limit=4.0, cap="butt", join="miter", res_scale=1.0):
    v = np.ascontiguousarray(verbs, np.uint8)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    ov, op, nv, npt = C.c_void_p(), C.c_void_p(), C.c_int32(), C.c_int32()
    ok = lib.orc_path_stroke(v.ctypes.data, len(v), p.ctypes.data, len(p), float(width), float(miter_limit),
                             CAPS[cap] if isinstance(cap, str) else int(cap), JOINS[join] if isinstance(join, str) else int(join),
                             float(res_scale), C.byref(ov), C.byref(nv), C.byref(op), C.byref(npt))
    if not ok:
        return None
    return _take_path(ov, nv, op, npt)


lib.orc_path_dash.restype = _i
lib.orc_path_dash.argtypes = [_vp, _i, _vp, _i, _vp, _i, _f, _f, _vpp, _ip, _vpp, _ip]


def dash_path(verbs, pts, dash_array, dash_offset=0.0, res_scale=1.0):
    v = np.ascontiguousarray(verbs, np.uint8)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    d = np.ascontiguousarray(dash_array, np.float32)
    ov, op, nv, npt = C.c_void_p(), C.c_void_p(), C.c_int32(), C.c_int32()
    ok = lib.orc_path_dash(v.ctypes.data, len(v), p.ctypes.data, len(p), d.ctypes.data, len(d), float(dash_offset), float(res_scale),
                           C.byref(ov), C.byref(nv), C.byref(op), C.byref(npt))
    if not ok:
        return None
    return _take_path(ov, nv, op, npt)


lib.orc_path_hairline.restype = _i
lib.orc_path_hairline.argtypes = [_vp, _i, _vp, _i, _i, _i, _i, _vpp]


def hairline_blits(verbs, pts, cap="butt", clip_w=1, clip_h=1):
    v = np.ascontiguousarray(verbs, np.uint8)
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = C.c_void_p()
    n = lib.orc_path_hairline(v.ctypes.data, len(v), p.ctypes.data, len(p), CAPS[cap] if isinstance(cap, str) else int(cap),
                              int(clip_w), int(clip_h), C.byref(out))
    if n <= 0 or not out.value:
        return np.zeros((0, 3), np.int32)
    try:
        arr = np.ctypeslib.as_array((C.c_int32 * (n * 3)).from_address(out.value)).copy()
    finally:
        lib.orc_geom_free(out)
    return arr.reshape(-1, 3)
