"""ctypes binding of include/resvg_b200.h (the C ABI of the CUDA library).

There is deliberately no fallback: if ``libresvg_b200.so`` is missing, or exports fewer symbols
than the header declares, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libresvg_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make lib` (or __graft_entry__.build()). "
        "resvg_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

c_void_pp = C.POINTER(C.c_void_p)
u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)


class TransferFn(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("n_values", C.c_int32),
        ("values", f32p),
        ("slope", C.c_float),
        ("intercept", C.c_float),
        ("amplitude", C.c_float),
        ("exponent", C.c_float),
        ("offset", C.c_float),
    ]


class LightSource(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("azimuth", C.c_float),
        ("elevation", C.c_float),
        ("x", C.c_float),
        ("y", C.c_float),
        ("z", C.c_float),
        ("points_at_x", C.c_float),
        ("points_at_y", C.c_float),
        ("points_at_z", C.c_float),
        ("specular_exponent", C.c_float),
        ("has_cone", C.c_int32),
        ("limiting_cone_angle", C.c_float),
    ]


_vp, _i, _u32, _f, _d, _u8 = C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_double, C.c_uint8


class Paint(C.Structure):
    """rb_paint (include/resvg_b200.h)."""
    _fields_ = [
        ("shader", C.c_int32), ("color", C.c_float * 4),
        ("x0", _f), ("y0", _f), ("r0", _f), ("x1", _f), ("y1", _f), ("r1", _f),
        ("n_stops", C.c_int32), ("stops", f32p), ("spread", C.c_int32), ("ts", C.c_float * 6),
        ("pattern", _vp), ("quality", C.c_int32), ("opacity", _f),
        ("blend_mode", C.c_int32), ("anti_alias", C.c_int32), ("force_hq", C.c_int32),
    ]


class Stroke(C.Structure):
    """rb_stroke"""
    _fields_ = [("width", C.c_float), ("miter_limit", C.c_float), ("cap", C.c_int32), ("join", C.c_int32),
                ("dash_array", f32p), ("n_dash", C.c_int32), ("dash_offset", C.c_float)]


# name -> (restype, argtypes); mirrors include/resvg_b200.h one to one
SIGNATURES = {
    "rb_ctx_create": (_i, [_i, c_void_pp]),
    "rb_ctx_destroy": (None, [_vp]),
    "rb_ctx_synchronize": (_i, [_vp]),
    "rb_last_error": (C.c_char_p, [_vp]),
    "rb_ctx_stream": (_vp, [_vp]),
    "rb_ctx_device": (_i, [_vp]),
    "rb_timer_begin": (_i, [_vp]),
    "rb_timer_end": (_i, [_vp, f32p]),
    "rb_ctx_launch_count": (C.c_uint64, [_vp]),
    "rb_ctx_h2d_bytes": (C.c_uint64, [_vp]),
    "rb_host_alloc": (_i, [C.c_size_t, c_void_pp]),
    "rb_host_free": (None, [_vp]),
    "rb_layer_create": (_i, [_vp, _u32, _u32, c_void_pp]),
    "rb_layer_destroy": (None, [_vp]),
    "rb_layer_width": (_u32, [_vp]),
    "rb_layer_height": (_u32, [_vp]),
    "rb_layer_device_ptr": (_vp, [_vp]),
    "rb_layer_upload": (_i, [_vp, _vp]),
    "rb_layer_download": (_i, [_vp, _vp]),
    "rb_layer_fill": (_i, [_vp, _u8, _u8, _u8, _u8]),
    "rb_layer_copy": (_i, [_vp, _vp]),
    "rb_layer_multiply_alpha": (_i, [_vp]),
    "rb_layer_demultiply_alpha": (_i, [_vp]),
    "rb_layer_into_linear_rgb": (_i, [_vp]),
    "rb_layer_into_srgb": (_i, [_vp]),
    "rb_filter_box_blur": (_i, [_vp, _d, _d]),
    "rb_filter_iir_blur": (_i, [_vp, _d, _d]),
    "rb_filter_iir_blur_fast": (_i, [_vp, _d, _d]),
    "rb_filter_morphology": (_i, [_vp, _i, _f, _f]),
    "rb_filter_convolve_matrix": (_i, [_vp, f32p, _u32, _u32, _u32, _u32, _f, _f, _i, _i]),
    "rb_filter_color_matrix": (_i, [_vp, _i, f32p]),
    "rb_filter_component_transfer": (_i, [_vp, C.POINTER(TransferFn)]),
    "rb_filter_composite_arithmetic": (_i, [_vp, _vp, _vp, _f, _f, _f, _f]),
    "rb_filter_displacement_map": (_i, [_vp, _vp, _vp, _i, _i, _f, _f, _f]),
    "rb_filter_diffuse_lighting": (_i, [_vp, _vp, _f, _f, _u8, _u8, _u8, C.POINTER(LightSource)]),
    "rb_filter_specular_lighting": (_i, [_vp, _vp, _f, _f, _f, _u8, _u8, _u8, C.POINTER(LightSource)]),
    "rb_filter_turbulence": (_i, [_vp, _d, _d, _d, _d, _d, _d, _u32, C.c_int32, _i, _i]),
    "rb_fill_path": (_i, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.POINTER(Paint), C.c_int32, f32p]),
    "rb_batch_begin": (_i, [_vp, c_void_pp]),
    "rb_batch_fill_path": (_i, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.POINTER(Paint), C.c_int32, f32p]),
    "rb_batch_stroke_path": (_i, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.POINTER(Paint), _vp, f32p]),
    "rb_batch_draw_paths": (_i, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, f32p]),
    "rb_batch_fill_paths": (_i, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, f32p]),
    "rb_batch_submit": (_i, [_vp, C.c_int32]),
    "rb_batch_submit_download": (_i, [_vp, C.c_int32, _vp]),
    "rb_debug_banded_downloads": (C.c_uint64, []),
    "rb_batch_prepare": (_i, [_vp, C.c_int32]),
    "rb_batch_run": (_i, [_vp]),
    "rb_batch_run_counting": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "rb_batch_destroy": (None, [_vp]),
    "rb_batch_stats": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "rb_draw_layer": (_i, [_vp, _vp, C.c_int32, C.c_int32, _f, C.c_int32]),
    "rb_mask_create": (_i, [_vp, _u32, _u32, c_void_pp]),
    "rb_mask_destroy": (None, [_vp]),
    "rb_mask_download": (_i, [_vp, _vp]),
    "rb_mask_upload": (_i, [_vp, _vp]),
    "rb_mask_from_layer": (_i, [_vp, _vp, C.c_int32]),
    "rb_mask_invert": (_i, [_vp]),
    "rb_layer_apply_mask": (_i, [_vp, _vp]),
    "rb_mask_fill_path": (_i, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, f32p]),
    "rb_path_stroke": (_i, [_vp, C.c_int32, _vp, C.c_int32, _f, _f, C.c_int32, C.c_int32, _f, c_void_pp,
                            C.POINTER(C.c_int32), c_void_pp, C.POINTER(C.c_int32)]),
    "rb_path_free": (None, [_vp]),
    "rb_path_hairline": (_i, [_vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_void_pp, C.POINTER(C.c_int32)]),
    "rb_path_dash": (_i, [_vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, _f, _f, c_void_pp, C.POINTER(C.c_int32), c_void_pp,
                          C.POINTER(C.c_int32)]),
    "rb_debug_force_wide_kernel": (None, [_i]),
    "rb_debug_build_edges": (_i, [_vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, f32p, _vp, _vp,
                                  C.c_int32, _vp]),
    "rb_ctx_last_run_ms": (_i, [_vp, _vp]),
    "rb_batch_set_viewport": (_i, [_vp, C.c_int32, C.c_int32, _u32, _u32]),
    "rb_layer_download_begin": (_i, [_vp, _vp]),
    "rb_layer_download_end": (_i, [_vp]),
    "rb_batch_draw_documents": (_i, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, f32p]),
    "rb_draw_layer_rects": (_i, [_vp, _vp, C.c_int32, _vp, _vp, _vp, C.c_int32]),
    "rb_filter_box_blur_reach": (_i, [_d]),
    "rb_filter_box_blur_cells": (_i, [_vp, C.c_int32, _vp, _vp, _vp]),
    "rb_filter_flood_alpha": (_i, [_vp, _u8, _u8, _u8, _u8]),
    "rb_debug_host_expand": (None, [_i]),
    "rb_debug_geo_mode": (None, [_i]),
    "rb_debug_geo_counts": (None, [_vp]),
    "rb_debug_geo_host_stats": (_i, [_vp, _vp]),
    "rb_debug_stroke_dashed_in_units": (_i, [_vp, C.c_int32, _vp, C.c_int32, _vp, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                         C.c_float, _vp, _vp, _vp, _vp]),
    "rb_debug_batch_begin_host": (_i, [_u32, _u32, c_void_pp]),
    "rb_debug_batch_phases": (_i, [_vp, _vp]),
    "rb_debug_batch_block": (C.c_int64, [_vp, C.c_int32, _vp, C.c_uint64]),
    "rb_debug_profile": (None, [_i]),
    "rb_stroke_path": (_i, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.POINTER(Paint), _vp, f32p]),
    "rb_fill_rect": (_i, [_vp, _f, _f, _f, _f, C.POINTER(Paint), f32p]),
    "rb_layer_clone_rect": (_i, [_vp, C.c_int32, C.c_int32, _u32, _u32, c_void_pp]),
    "rb_layer_apply_mask_layer": (_i, [_vp, _vp, C.c_int32]),
    "rb_layer_apply_clip_layer": (_i, [_vp, _vp]),
    "rb_tree_parse": (_i, [C.c_char_p, C.c_size_t, c_void_pp]),
    "rb_tree_destroy": (None, [_vp]),
    "rb_tree_size": (_i, [_vp, f32p, f32p]),
    "rb_tree_node_bbox": (_i, [_vp, C.c_char_p, f32p]),
    "rb_render": (_i, [_vp, _vp, f32p, _vp]),
    "rb_render_node": (_i, [_vp, _vp, C.c_char_p, f32p, _vp]),
    "rb_render_strip": (_i, [_vp, _vp, f32p, _u32, _u32, C.c_int32, _vp]),
    "rb_submit": (_i, [_vp, C.c_char_p, C.c_size_t, f32p, _vp]),
    "rb_render_to_host": (_i, [_vp, _vp, f32p, _u32, _u32, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args
