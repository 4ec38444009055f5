"""ctypes binding of oracle/liboracle.so — the CPU checker (test infrastructure only).

Each wrapper takes/returns numpy uint8 arrays of shape (h, w, 4) and never mutates its inputs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "liboracle.so")


def _ensure_built():
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ROOT, "oracle"], stdout=subprocess.DEVNULL)


_ensure_built()
lib = C.CDLL(LIB_PATH)

_vp, _i, _u32, _f, _d, _u8, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_double, C.c_uint8, C.c_size_t
f32p = C.POINTER(C.c_float)


class TransferFn(C.Structure):
    _fields_ = [("type", C.c_int32), ("n_values", C.c_int32), ("values", f32p), ("slope", _f), ("intercept", _f),
                ("amplitude", _f), ("exponent", _f), ("offset", _f)]


class LightSource(C.Structure):
    _fields_ = [("kind", C.c_int32), ("azimuth", _f), ("elevation", _f), ("x", _f), ("y", _f), ("z", _f),
                ("points_at_x", _f), ("points_at_y", _f), ("points_at_z", _f), ("specular_exponent", _f),
                ("has_cone", C.c_int32), ("limiting_cone_angle", _f)]


def _sig(name, res, args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
    return fn


_sig("orc_multiply_alpha", None, [_vp, _sz])
_sig("orc_demultiply_alpha", None, [_vp, _sz])
_sig("orc_into_linear_rgb", None, [_vp, _sz])
_sig("orc_into_srgb", None, [_vp, _sz])
_sig("orc_box_blur", None, [_d, _d, _vp, _u32, _u32])
_sig("orc_create_box_gauss", None, [_f, C.POINTER(C.c_int32)])
_sig("orc_iir_blur", None, [_d, _d, _vp, _u32, _u32])
_sig("orc_morphology", None, [_i, _f, _f, _vp, _u32, _u32])
_sig("orc_convolve_matrix", None, [f32p, _u32, _u32, _u32, _u32, _f, _f, _i, _i, _vp, _u32, _u32])
_sig("orc_color_matrix", None, [_i, f32p, _vp, _sz])
_sig("orc_component_transfer", None, [C.POINTER(TransferFn), _vp, _sz])
_sig("orc_composite_arithmetic", None, [_f, _f, _f, _f, _vp, _vp, _vp, _sz])
_sig("orc_displacement_map", None, [_i, _i, _f, _f, _f, _vp, _vp, _vp, _u32, _u32])
_sig("orc_diffuse_lighting", None, [_f, _f, _u8, _u8, _u8, C.POINTER(LightSource), _vp, _vp, _u32, _u32])
_sig("orc_specular_lighting", None, [_f, _f, _f, _u8, _u8, _u8, C.POINTER(LightSource), _vp, _vp, _u32, _u32])
_sig("orc_turbulence", None, [_d, _d, _d, _d, _d, _d, _u32, C.c_int32, _i, _i, _vp, _u32, _u32])


def _own(img):
    a = np.array(img, dtype=np.uint8, order="C", copy=True)
    assert a.ndim == 3 and a.shape[2] == 4
    return a


def _f32(values):
    arr = np.ascontiguousarray(values, dtype=np.float32)
    return arr, arr.ctypes.data_as(f32p)


def multiply_alpha(img):
    a = _own(img); lib.orc_multiply_alpha(a.ctypes.data, a.shape[0] * a.shape[1]); return a


def demultiply_alpha(img):
    a = _own(img); lib.orc_demultiply_alpha(a.ctypes.data, a.shape[0] * a.shape[1]); return a


def into_linear_rgb(img):
    a = _own(img); lib.orc_into_linear_rgb(a.ctypes.data, a.shape[0] * a.shape[1]); return a


def into_srgb(img):
    a = _own(img); lib.orc_into_srgb(a.ctypes.data, a.shape[0] * a.shape[1]); return a


def create_box_gauss(sigma):
    out = (C.c_int32 * 5)()
    lib.orc_create_box_gauss(sigma, out)
    return list(out)


def box_blur(sigma_x, sigma_y, img):
    a = _own(img); lib.orc_box_blur(sigma_x, sigma_y, a.ctypes.data, a.shape[1], a.shape[0]); return a


def iir_blur(sigma_x, sigma_y, img):
    a = _own(img); lib.orc_iir_blur(sigma_x, sigma_y, a.ctypes.data, a.shape[1], a.shape[0]); return a


def morphology(operator, rx, ry, img):
    a = _own(img)
    lib.orc_morphology({"erode": 0, "dilate": 1}[operator], rx, ry, a.ctypes.data, a.shape[1], a.shape[0])
    return a


def convolve_matrix(kernel, columns, rows, target_x, target_y, divisor, bias, edge_mode, preserve_alpha, img):
    a = _own(img)
    arr, ptr = _f32(kernel)
    lib.orc_convolve_matrix(ptr, columns, rows, target_x, target_y, divisor, bias,
                            {"none": 0, "duplicate": 1, "wrap": 2}[edge_mode], 1 if preserve_alpha else 0,
                            a.ctypes.data, a.shape[1], a.shape[0])
    return a


def color_matrix(kind, params, img):
    a = _own(img)
    arr, ptr = _f32(params if len(params) else [0.0])
    lib.orc_color_matrix({"matrix": 0, "saturate": 1, "hueRotate": 2, "luminanceToAlpha": 3}[kind], ptr,
                         a.ctypes.data, a.shape[0] * a.shape[1])
    return a


def make_transfer(kind="identity", values=(), slope=1.0, intercept=0.0, amplitude=1.0, exponent=1.0, offset=0.0):
    types = {"identity": 0, "table": 1, "discrete": 2, "linear": 3, "gamma": 4}
    arr, ptr = _f32(list(values))
    return TransferFn(types[kind], len(arr), ptr if len(arr) else None, slope, intercept, amplitude, exponent,
                      offset), arr


def component_transfer(funcs, img):
    a = _own(img)
    arr = (TransferFn * 4)(*[f[0] for f in funcs])
    lib.orc_component_transfer(arr, a.ctypes.data, a.shape[0] * a.shape[1])
    return a


def arithmetic(k1, k2, k3, k4, src1, src2, dest=None):
    s1, s2 = _own(src1), _own(src2)
    d = np.zeros_like(s1) if dest is None else _own(dest)
    lib.orc_composite_arithmetic(k1, k2, k3, k4, s1.ctypes.data, s2.ctypes.data, d.ctypes.data,
                                 s1.shape[0] * s1.shape[1])
    return d


def displacement_map(x_channel, y_channel, scale, sx, sy, src, map_):
    s, m = _own(src), _own(map_)
    d = np.zeros_like(s)
    lib.orc_displacement_map(x_channel, y_channel, scale, sx, sy, s.ctypes.data, m.ctypes.data, d.ctypes.data,
                             s.shape[1], s.shape[0])
    return d


def make_light(kind="distant", azimuth=0.0, elevation=0.0, x=0.0, y=0.0, z=0.0, points_at=(0.0, 0.0, 0.0),
               specular_exponent=1.0, limiting_cone_angle=None):
    kinds = {"distant": 0, "point": 1, "spot": 2}
    return LightSource(kinds[kind], azimuth, elevation, x, y, z, points_at[0], points_at[1], points_at[2],
                       specular_exponent, 0 if limiting_cone_angle is None else 1,
                       0.0 if limiting_cone_angle is None else limiting_cone_angle)


def diffuse_lighting(surface_scale, diffuse_constant, color, light, src):
    s = _own(src)
    d = np.zeros_like(s)
    lib.orc_diffuse_lighting(surface_scale, diffuse_constant, color[0], color[1], color[2], C.byref(light),
                             s.ctypes.data, d.ctypes.data, s.shape[1], s.shape[0])
    return d


def specular_lighting(surface_scale, specular_constant, specular_exponent, color, light, src):
    s = _own(src)
    d = np.zeros_like(s)
    lib.orc_specular_lighting(surface_scale, specular_constant, specular_exponent, color[0], color[1], color[2],
                              C.byref(light), s.ctypes.data, d.ctypes.data, s.shape[1], s.shape[0])
    return d


def turbulence(offset_x, offset_y, sx, sy, bfx, bfy, num_octaves, seed, stitch_tiles, fractal_noise, w, h):
    d = np.zeros((h, w, 4), dtype=np.uint8)
    lib.orc_turbulence(offset_x, offset_y, sx, sy, bfx, bfy, num_octaves, seed, 1 if stitch_tiles else 0,
                       1 if fractal_noise else 0, d.ctypes.data, w, h)
    return d
