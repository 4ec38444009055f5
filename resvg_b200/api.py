"""Thin Python mirror of the reference's operator interface over the C ABI.

Names follow the reference: ``Layer`` plays tiny_skia::Pixmap, the functions in ``filters`` carry the
names and argument order of crates/resvg/src/filter/*.rs (``box_blur.apply(sigma_x, sigma_y, src)`` →
``filters.box_blur(sigma_x, sigma_y, layer)``).  Everything executes on the GPU through
libresvg_b200.so; numpy is only the host-side container for pixels.
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import LightSource, Paint, TransferFn, lib


class ResvgB200Error(RuntimeError):
    pass


_STATUS = {1: "invalid argument", 2: "CUDA error", 3: "out of memory", 4: "unsupported"}


class Context:
    """One per GPU (rb_ctx): owns the CUDA stream every operation is enqueued on."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        st = lib.rb_ctx_create(device, C.byref(h))
        if st != 0:
            raise ResvgB200Error(
                f"rb_ctx_create(device={device}) failed: {_STATUS.get(st, st)} — resvg_b200 needs a CUDA GPU "
                "(there is no CPU fallback)"
            )
        self._h = h
        self.device = device

    def check(self, st: int, what: str = ""):
        if st != 0:
            msg = lib.rb_last_error(self._h)
            raise ResvgB200Error(f"{what}: {_STATUS.get(st, st)} ({msg.decode() if msg else ''})")

    def synchronize(self):
        self.check(lib.rb_ctx_synchronize(self._h), "synchronize")

    def timer_begin(self):
        self.check(lib.rb_timer_begin(self._h), "timer_begin")

    def timer_end(self) -> float:
        ms = C.c_float()
        self.check(lib.rb_timer_end(self._h, C.byref(ms)), "timer_end")
        return ms.value

    def last_run_ms(self):
        """(pre-pass ms, raster kernel ms) of the last batch run on this context, from CUDA events."""
        ms = (C.c_float * 2)()
        self.check(lib.rb_ctx_last_run_ms(self._h, ms), "last_run_ms")
        return float(ms[0]), float(ms[1])

    @property
    def launch_count(self) -> int:
        return int(lib.rb_ctx_launch_count(self._h))

    @property
    def h2d_bytes(self) -> int:
        return int(lib.rb_ctx_h2d_bytes(self._h))

    @property
    def stream(self) -> int:
        return int(lib.rb_ctx_stream(self._h) or 0)

    def layer(self, width: int, height: int) -> "Layer":
        return Layer(self, width, height)

    def layer_from(self, rgba: np.ndarray) -> "Layer":
        assert rgba.dtype == np.uint8 and rgba.ndim == 3 and rgba.shape[2] == 4
        l = Layer(self, rgba.shape[1], rgba.shape[0])
        l.upload(rgba)
        return l

    def close(self):
        if self._h:
            lib.rb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host staging buffer exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        if lib.rb_host_alloc(nbytes, C.byref(p)) != 0:
            raise ResvgB200Error("rb_host_alloc failed")
        self._p = p
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))

    def close(self):
        if self._p:
            self.array = None
            lib.rb_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Layer:
    """Device-resident premultiplied RGBA8 pixmap (tiny_skia::Pixmap)."""

    def __init__(self, ctx: Context, width: int, height: int):
        h = C.c_void_p()
        ctx.check(lib.rb_layer_create(ctx._h, width, height, C.byref(h)), "layer_create")
        self._h = h
        self.ctx = ctx
        self.width = width
        self.height = height

    def upload(self, rgba: np.ndarray):
        a = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert a.size == self.width * self.height * 4
        self.ctx.check(lib.rb_layer_upload(self._h, a.ctypes.data), "upload")
        self.ctx.synchronize()  # `a` may be a temporary

    def upload_ptr(self, ptr: int):
        self.ctx.check(lib.rb_layer_upload(self._h, ptr), "upload")

    def download(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self.ctx.check(lib.rb_layer_download(self._h, out.ctypes.data), "download")
        return out

    def download_ptr(self, ptr: int):
        self.ctx.check(lib.rb_layer_download(self._h, ptr), "download")

    def download_begin(self, ptr: int):
        """Asynchronous download into pinned host memory at `ptr` (rb_layer_download_begin); overlaps later renders."""
        self.ctx.check(lib.rb_layer_download_begin(self._h, C.c_void_p(ptr)), "layer_download_begin")

    def download_end(self):
        self.ctx.check(lib.rb_layer_download_end(self._h), "layer_download_end")

    def fill(self, r: int, g: int, b: int, a: int):
        self.ctx.check(lib.rb_layer_fill(self._h, r, g, b, a), "fill")

    def copy_from(self, other: "Layer"):
        self.ctx.check(lib.rb_layer_copy(self._h, other._h), "copy")

    def clone(self) -> "Layer":
        l = Layer(self.ctx, self.width, self.height)
        l.copy_from(self)
        return l

    def close(self):
        if self._h:
            lib.rb_layer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _f32(values):
    arr = np.ascontiguousarray(values, dtype=np.float32)
    return arr, arr.ctypes.data_as(_ffi.f32p)


def make_transfer(kind="identity", values=(), slope=1.0, intercept=0.0, amplitude=1.0, exponent=1.0, offset=0.0):
    """usvg::filter::TransferFunction → (rb_transfer_fn, keep-alive)."""
    types = {"identity": 0, "table": 1, "discrete": 2, "linear": 3, "gamma": 4}
    arr, ptr = _f32(list(values))
    t = TransferFn(types[kind], len(arr), ptr if len(arr) else None, slope, intercept, amplitude, exponent, offset)
    return t, arr


def make_light(kind="distant", azimuth=0.0, elevation=0.0, x=0.0, y=0.0, z=0.0, points_at=(0.0, 0.0, 0.0),
               specular_exponent=1.0, limiting_cone_angle=None):
    kinds = {"distant": 0, "point": 1, "spot": 2}
    return LightSource(kinds[kind], azimuth, elevation, x, y, z, points_at[0], points_at[1], points_at[2],
                       specular_exponent, 0 if limiting_cone_angle is None else 1,
                       0.0 if limiting_cone_angle is None else limiting_cone_angle)


class filters:
    """crates/resvg/src/filter/* on the GPU.  Static methods, reference names and argument order."""

    @staticmethod
    def multiply_alpha(l: Layer):
        l.ctx.check(lib.rb_layer_multiply_alpha(l._h), "multiply_alpha")

    @staticmethod
    def demultiply_alpha(l: Layer):
        l.ctx.check(lib.rb_layer_demultiply_alpha(l._h), "demultiply_alpha")

    @staticmethod
    def into_linear_rgb(l: Layer):
        l.ctx.check(lib.rb_layer_into_linear_rgb(l._h), "into_linear_rgb")

    @staticmethod
    def into_srgb(l: Layer):
        l.ctx.check(lib.rb_layer_into_srgb(l._h), "into_srgb")

    @staticmethod
    def box_blur(sigma_x: float, sigma_y: float, src: Layer):
        src.ctx.check(lib.rb_filter_box_blur(src._h, sigma_x, sigma_y), "box_blur")

    @staticmethod
    def iir_blur(sigma_x: float, sigma_y: float, src: Layer):
        """iir_blur::apply, sequential f64: bit-identical to the reference (what the renderer uses)."""
        src.ctx.check(lib.rb_filter_iir_blur(src._h, sigma_x, sigma_y), "iir_blur")

    @staticmethod
    def iir_blur_fast(sigma_x: float, sigma_y: float, src: Layer):
        """The f32 segment kernels: within 1/255 of iir_blur::apply on the stage output, ~10x faster (opt-in)."""
        src.ctx.check(lib.rb_filter_iir_blur_fast(src._h, sigma_x, sigma_y), "iir_blur_fast")

    @staticmethod
    def morphology(operator: str, rx: float, ry: float, src: Layer):
        src.ctx.check(lib.rb_filter_morphology(src._h, {"erode": 0, "dilate": 1}[operator], rx, ry), "morphology")

    @staticmethod
    def convolve_matrix(kernel, columns, rows, target_x, target_y, divisor, bias, edge_mode, preserve_alpha,
                        src: Layer):
        arr, ptr = _f32(kernel)
        assert arr.size == columns * rows
        em = {"none": 0, "duplicate": 1, "wrap": 2}[edge_mode]
        src.ctx.check(
            lib.rb_filter_convolve_matrix(src._h, ptr, columns, rows, target_x, target_y, divisor, bias, em,
                                          1 if preserve_alpha else 0),
            "convolve_matrix",
        )

    @staticmethod
    def color_matrix(kind: str, params, src: Layer):
        k = {"matrix": 0, "saturate": 1, "hueRotate": 2, "luminanceToAlpha": 3}[kind]
        arr, ptr = _f32(params if len(params) else [0.0])
        src.ctx.check(lib.rb_filter_color_matrix(src._h, k, ptr), "color_matrix")

    @staticmethod
    def component_transfer(funcs, src: Layer):
        """funcs: four (rb_transfer_fn, keepalive) pairs from make_transfer, order r,g,b,a."""
        arr = (TransferFn * 4)(*[f[0] for f in funcs])
        src.ctx.check(lib.rb_filter_component_transfer(src._h, arr), "component_transfer")

    @staticmethod
    def box_blur_reach(sigma: float) -> int:
        """Halo (rows / columns) a strip needs for its own pixels to equal the box blur of the whole image."""
        return int(lib.rb_filter_box_blur_reach(float(sigma)))

    @staticmethod
    def box_blur_cells(rects, sigma_x, sigma_y, src: Layer):
        """box_blur::apply on every rectangle (x, y, w, h) of the layer as a pixmap of its own (rb_filter_box_blur_cells)."""
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        sx = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma_x, np.float64), (len(r),)))
        sy = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma_y, np.float64), (len(r),)))
        src.ctx.check(lib.rb_filter_box_blur_cells(src._h, len(r), r.ctypes.data, sx.ctypes.data, sy.ctypes.data), "box_blur_cells")

    @staticmethod
    def flood_alpha(color, opacity_u8: int, src: Layer):
        """apply_drop_shadow's flood (filter/mod.rs:606-617): colour (r, g, b as u8) with opacity_u8 scaled by every pixel's alpha."""
        src.ctx.check(lib.rb_filter_flood_alpha(src._h, int(color[0]), int(color[1]), int(color[2]), int(opacity_u8)), "flood_alpha")

    @staticmethod
    def arithmetic(k1, k2, k3, k4, src1: Layer, src2: Layer, dest: Layer):
        dest.ctx.check(lib.rb_filter_composite_arithmetic(dest._h, src1._h, src2._h, k1, k2, k3, k4), "arithmetic")

    @staticmethod
    def displacement_map(x_channel: int, y_channel: int, scale: float, sx: float, sy: float, src: Layer, map_: Layer,
                         dest: Layer):
        dest.ctx.check(
            lib.rb_filter_displacement_map(dest._h, src._h, map_._h, x_channel, y_channel, scale, sx, sy),
            "displacement_map",
        )

    @staticmethod
    def diffuse_lighting(surface_scale, diffuse_constant, color, light: LightSource, src: Layer, dest: Layer):
        dest.ctx.check(
            lib.rb_filter_diffuse_lighting(dest._h, src._h, surface_scale, diffuse_constant, color[0], color[1],
                                           color[2], C.byref(light)),
            "diffuse_lighting",
        )

    @staticmethod
    def specular_lighting(surface_scale, specular_constant, specular_exponent, color, light: LightSource, src: Layer,
                          dest: Layer):
        dest.ctx.check(
            lib.rb_filter_specular_lighting(dest._h, src._h, surface_scale, specular_constant, specular_exponent,
                                            color[0], color[1], color[2], C.byref(light)),
            "specular_lighting",
        )

    @staticmethod
    def turbulence(offset_x, offset_y, sx, sy, base_frequency_x, base_frequency_y, num_octaves, seed, stitch_tiles,
                   fractal_noise, dest: Layer):
        dest.ctx.check(
            lib.rb_filter_turbulence(dest._h, offset_x, offset_y, sx, sy, base_frequency_x, base_frequency_y,
                                     num_octaves, seed, 1 if stitch_tiles else 0, 1 if fractal_noise else 0),
            "turbulence",
        )


# --------------------------------------------------------------------------------------------------
# rasteriser: tiny-skia Paint / fill_path / draw_pixmap / Mask over the C ABI
# --------------------------------------------------------------------------------------------------
BLEND = {n: i for i, n in enumerate([
    "clear", "source", "destination", "source_over", "destination_over", "source_in", "destination_in",
    "source_out", "destination_out", "source_atop", "destination_atop", "xor", "plus", "modulate", "screen",
    "overlay", "darken", "lighten", "color_dodge", "color_burn", "hard_light", "soft_light", "difference",
    "exclusion", "multiply", "hue", "saturation", "color", "luminosity"])}
SPREAD = {"pad": 0, "reflect": 1, "repeat": 2}
QUALITY = {"nearest": 0, "bilinear": 1, "bicubic": 2}
IDENTITY = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)


def _ts(ts):
    return (C.c_float * 6)(*[float(v) for v in ts])


def make_paint(spec, blend="source_over", anti_alias=True):
    """Build an rb_paint from a plain dict:
    {"kind": "solid", "color": (r,g,b,a)} | {"kind": "linear"|"radial", x0,y0,[r0],x1,y1,[r1], stops, spread, ts}
    | {"kind": "pattern", "layer": Layer, spread, quality, opacity, ts}."""
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
        p.n_stops = stops.shape[0]
        p.stops = stops.ctypes.data_as(_ffi.f32p)
        p.spread = SPREAD[spec.get("spread", "pad")]
        p.ts[:] = spec.get("ts", IDENTITY)
        p._keep = stops
    elif kind == "pattern":
        p.shader = 3
        p.pattern = spec["layer"]._h
        p.spread = SPREAD[spec.get("spread", "repeat")]
        p.quality = QUALITY[spec.get("quality", "bicubic")]
        p.opacity = spec.get("opacity", 1.0)
        p.ts[:] = spec.get("ts", IDENTITY)
        p._keep = spec["layer"]
    else:
        raise ValueError(kind)
    return p


def _path(verbs, pts):
    v = np.ascontiguousarray(verbs, dtype=np.uint8)
    p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
    return v, p


def fill_path(layer: Layer, verbs, pts, paint: Paint, rule="nonzero", ts=IDENTITY):
    """PixmapMut::fill_path(path, paint, rule, transform, None)."""
    v, p = _path(verbs, pts)
    layer.ctx.check(lib.rb_fill_path(layer._h, v.ctypes.data, len(v), p.ctypes.data, len(p), C.byref(paint),
                                     1 if rule == "evenodd" else 0, _ts(ts)), "fill_path")


class Batch:
    """Many fill_path calls against one layer, executed by one tile-binned kernel launch."""

    def __init__(self, layer: Layer):
        h = C.c_void_p()
        layer.ctx.check(lib.rb_batch_begin(layer._h, C.byref(h)), "batch_begin")
        self._h = h
        self.layer = layer
        self._keep = []

    def fill_path(self, verbs, pts, paint: Paint, rule="nonzero", ts=IDENTITY):
        v, p = _path(verbs, pts)
        self.layer.ctx.check(lib.rb_batch_fill_path(self._h, v.ctypes.data, len(v), p.ctypes.data, len(p),
                                                    C.byref(paint), 1 if rule == "evenodd" else 0, _ts(ts)),
                             "batch_fill_path")
        if getattr(paint, "_keep", None) is not None and paint.shader == 3:
            self._keep.append(paint._keep)

    def fill_paths(self, scene, ts=IDENTITY):
        """Bulk recording from packed arrays: scene has verb_off, pt_off (uint32, n+1), verbs (uint8), pts (float32
        (m, 2)), paints (ctypes array of rb_paint), rules (uint8).  Recorded by reference: the arrays are kept alive
        by this batch and must not be modified before submit() / prepare() has returned."""
        n = len(scene["rules"])
        strokes = scene.get("strokes")
        self._keep.append(dict(scene))
        self.layer.ctx.check(
            lib.rb_batch_draw_paths(self._h, n, scene["verb_off"].ctypes.data, scene["pt_off"].ctypes.data,
                                    scene["verbs"].ctypes.data, scene["pts"].ctypes.data,
                                    C.addressof(scene["paints"]), scene["rules"].ctypes.data,
                                    C.addressof(strokes) if strokes is not None else None, _ts(ts)),
            "batch_draw_paths")

    def stroke_path(self, verbs, pts, paint: Paint, width, miter_limit=4.0, cap="butt", join="miter", ts=IDENTITY,
                    dash=None, dash_offset=0.0):
        """PixmapMut::stroke_path; dash = the StrokeDash::new array (None: solid)."""
        v, p = _path(verbs, pts)
        st = _ffi.Stroke(float(width), float(miter_limit), CAPS[cap] if isinstance(cap, str) else int(cap),
                         JOINS[join] if isinstance(join, str) else int(join))
        if dash is not None and len(dash):
            arr, ptr = _f32(dash)
            st.dash_array = ptr
            st.n_dash = arr.size
            st.dash_offset = float(dash_offset)
        self.layer.ctx.check(lib.rb_batch_stroke_path(self._h, v.ctypes.data, len(v), p.ctypes.data, len(p),
                                                      C.byref(paint), C.byref(st), _ts(ts)), "batch_stroke_path")

    def draw_documents(self, scene, viewports, doc_first, doc_count, ts=IDENTITY):
        """One document per viewport (rb_batch_draw_documents): document k = doc_count[k] paths from doc_first[k] of the
        packed scene arrays (same layout and lifetime rules as fill_paths), rendered into viewports[k] = (x, y, w, h)."""
        vp = np.ascontiguousarray(viewports, np.int32).reshape(-1, 4)
        first = np.ascontiguousarray(doc_first, np.uint32)
        count = np.ascontiguousarray(doc_count, np.uint32)
        assert len(first) == len(vp) == len(count)
        strokes = scene.get("strokes")
        self._keep.append((dict(scene), vp, first, count))
        self.layer.ctx.check(
            lib.rb_batch_draw_documents(self._h, len(vp), vp.ctypes.data, first.ctypes.data, count.ctypes.data,
                                        scene["verb_off"].ctypes.data, scene["pt_off"].ctypes.data, scene["verbs"].ctypes.data,
                                        scene["pts"].ctypes.data, C.addressof(scene["paints"]), scene["rules"].ctypes.data,
                                        C.addressof(strokes) if strokes is not None else None, _ts(ts)),
            "batch_draw_documents")

    def set_viewport(self, x=0, y=0, w=0, h=0):
        """Render the draws recorded next into the rectangle (x, y, w, h) of the layer as if it were a pixmap of its own
        (w = h = 0: the whole layer again)."""
        self.layer.ctx.check(lib.rb_batch_set_viewport(self._h, int(x), int(y), int(w), int(h)), "batch_set_viewport")

    def submit(self, n_threads: int = 0):
        self.layer.ctx.check(lib.rb_batch_submit(self._h, n_threads), "batch_submit")

    def submit_download(self, ptr: int, n_threads: int = 0):
        """submit(), then the layer into host memory at `ptr` (w * h * 4 bytes, pinned for the overlap): the bands of the last
        raster launch are copied out while the next band is rendered (rb_batch_submit_download).  Returns once everything is
        enqueued; `layer.download_end()` waits for the pixels."""
        self.layer.ctx.check(lib.rb_batch_submit_download(self._h, n_threads, C.c_void_p(ptr)), "batch_submit_download")

    def prepare(self, n_threads: int = 0):
        self.layer.ctx.check(lib.rb_batch_prepare(self._h, n_threads), "batch_prepare")

    def run(self):
        self.layer.ctx.check(lib.rb_batch_run(self._h), "batch_run")

    def run_counting(self):
        """One run with the pixel counters on → (pixels blended read-modify-write, pixels stored write-only)."""
        out = (C.c_uint64 * 2)()
        self.layer.ctx.check(lib.rb_batch_run_counting(self._h, out), "batch_run_counting")
        return int(out[0]), int(out[1])

    def stats(self):
        s = (C.c_uint64 * 6)()
        lib.rb_batch_stats(self._h, s)
        return dict(draws=s[0], edges=s[1], pairs=s[2], tiles=s[3], upload_bytes=s[4], host_us=s[5])

    def close(self):
        if self._h:
            lib.rb_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def draw_layer(dst: Layer, src: Layer, x=0, y=0, opacity=1.0, blend="source_over"):
    """PixmapMut::draw_pixmap(x, y, src, PixmapPaint{opacity, blend_mode, Nearest}, identity, None)."""
    b = BLEND[blend] if isinstance(blend, str) else int(blend)
    dst.ctx.check(lib.rb_draw_layer(dst._h, src._h, int(x), int(y), float(opacity), b), "draw_layer")


def draw_layer_rects(dst: Layer, src: Layer, rects, opacity, blend="source_over", src_xy=None):
    """Region-wise draw_pixmap for atlases (rb_draw_layer_rects): rects (n, 4) = x, y, w, h in dst; src_xy (n, 2) = where
    the pixels come from in src (None: the same position); opacity (n,) or a scalar."""
    r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
    o = np.ascontiguousarray(np.broadcast_to(np.asarray(opacity, np.float32), (len(r),)))
    s = None if src_xy is None else np.ascontiguousarray(src_xy, np.int32).reshape(-1, 2)
    assert s is None or len(s) == len(r)
    b = BLEND[blend] if isinstance(blend, str) else int(blend)
    dst.ctx.check(lib.rb_draw_layer_rects(dst._h, src._h, len(r), r.ctypes.data, s.ctypes.data if s is not None else None,
                                          o.ctypes.data, b), "draw_layer_rects")


class Mask:
    """tiny_skia::Mask on the device."""

    def __init__(self, ctx: Context, width: int, height: int):
        h = C.c_void_p()
        ctx.check(lib.rb_mask_create(ctx._h, width, height, C.byref(h)), "mask_create")
        self._h, self.ctx, self.width, self.height = h, ctx, width, height

    @staticmethod
    def from_layer(layer: Layer, kind="alpha") -> "Mask":
        m = Mask(layer.ctx, layer.width, layer.height)
        layer.ctx.check(lib.rb_mask_from_layer(m._h, layer._h, 1 if kind == "luminance" else 0), "mask_from_layer")
        return m

    def invert(self):
        self.ctx.check(lib.rb_mask_invert(self._h), "mask_invert")

    def fill_path(self, verbs, pts, rule="nonzero", anti_alias=True, ts=IDENTITY):
        v, p = _path(verbs, pts)
        self.ctx.check(lib.rb_mask_fill_path(self._h, v.ctypes.data, len(v), p.ctypes.data, len(p),
                                             1 if rule == "evenodd" else 0, 1 if anti_alias else 0, _ts(ts)),
                       "mask_fill_path")

    def download(self) -> np.ndarray:
        out = np.empty((self.height, self.width), dtype=np.uint8)
        self.ctx.check(lib.rb_mask_download(self._h, out.ctypes.data), "mask_download")
        return out

    def upload(self, m: np.ndarray):
        a = np.ascontiguousarray(m, dtype=np.uint8)
        self.ctx.check(lib.rb_mask_upload(self._h, a.ctypes.data), "mask_upload")
        self.ctx.synchronize()

    def close(self):
        if self._h:
            lib.rb_mask_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def apply_mask(layer: Layer, mask: Mask):
    """Pixmap::apply_mask(mask)."""
    layer.ctx.check(lib.rb_layer_apply_mask(layer._h, mask._h), "apply_mask")


CAPS = {"butt": 0, "round": 1, "square": 2}
JOINS = {"miter": 0, "miter-clip": 1, "round": 2, "bevel": 3}


def hairline_blits(verbs, pts, cap="butt", clip_w=1, clip_h=1):
    """The ordered {x, y, alpha} blits of tiny-skia's anti-aliased hairline walker for a path in device space
    → int32 array (n, 3).  Host-only."""
    v, p = _path(verbs, pts)
    out, n = C.c_void_p(), C.c_int32()
    st = lib.rb_path_hairline(v.ctypes.data, len(v), p.ctypes.data, len(p), CAPS[cap] if isinstance(cap, str) else int(cap),
                              int(clip_w), int(clip_h), C.byref(out), C.byref(n))
    if st != 0:
        return np.zeros((0, 3), np.int32)
    try:
        arr = np.ctypeslib.as_array((C.c_int32 * (max(n.value, 1) * 3)).from_address(out.value)).copy()[: n.value * 3]
    finally:
        lib.rb_path_free(out)
    return arr.reshape(-1, 3)


def dash_path(verbs, pts, dash_array, dash_offset=0.0, res_scale=1.0):
    """tiny_skia_path::Path::dash(StrokeDash::new(dash_array, dash_offset)?, res_scale) → (verbs, points) or None when
    the specification is rejected or nothing is left.  Host-only."""
    v, p = _path(verbs, pts)
    arr, ptr = _f32(dash_array)
    ov, op = C.c_void_p(), C.c_void_p()
    nv, np_ = C.c_int32(), C.c_int32()
    st = lib.rb_path_dash(v.ctypes.data, len(v), p.ctypes.data, len(p), C.cast(ptr, C.c_void_p), arr.size, float(dash_offset),
                          float(res_scale), C.byref(ov), C.byref(nv), C.byref(op), C.byref(np_))
    if st != 0:
        return None
    try:
        out_v = np.ctypeslib.as_array((C.c_uint8 * nv.value).from_address(ov.value)).copy()
        out_p = np.ctypeslib.as_array((C.c_float * (np_.value * 2)).from_address(op.value)).copy().reshape(-1, 2)
    finally:
        lib.rb_path_free(ov)
        lib.rb_path_free(op)
    return out_v, out_p


def stroke_path(verbs, pts, width, miter_limit=4.0, cap="butt", join="miter", res_scale=1.0):
    """tiny_skia_path::Path::stroke → (verbs uint8, points float32 (n, 2)) or None when the stroke is empty.
    Host-only (no GPU needed)."""
    v, p = _path(verbs, pts)
    ov, op = C.c_void_p(), C.c_void_p()
    nv, np_ = C.c_int32(), C.c_int32()
    st = lib.rb_path_stroke(v.ctypes.data, len(v), p.ctypes.data, len(p), float(width), float(miter_limit),
                            CAPS[cap] if isinstance(cap, str) else int(cap),
                            JOINS[join] if isinstance(join, str) else int(join), float(res_scale),
                            C.byref(ov), C.byref(nv), C.byref(op), C.byref(np_))
    if st != 0:
        return None
    try:
        out_v = np.ctypeslib.as_array((C.c_uint8 * nv.value).from_address(ov.value)).copy()
        out_p = np.ctypeslib.as_array((C.c_float * (np_.value * 2)).from_address(op.value)).copy().reshape(-1, 2)
    finally:
        lib.rb_path_free(ov)
        lib.rb_path_free(op)
    return out_v, out_p


def stroke_dashed_in_units(verbs, pts, dash_array, dash_offset, width, miter_limit=4.0, cap="butt", join="miter", res_scale=1.0):
    """The outline of a dashed stroke built dash by dash, as the geometry kernels do it (rb_debug_stroke_dashed_in_units:
    the same code on the host) → (verbs, points) or None.  Equals stroke_path(*dash_path(...)) — that is what the test checks."""
    v, p = _path(verbs, pts)
    arr, ptr = _f32(dash_array)
    ov, op = C.c_void_p(), C.c_void_p()
    nv, np_ = C.c_int32(), C.c_int32()
    st = lib.rb_debug_stroke_dashed_in_units(v.ctypes.data, len(v), p.ctypes.data, len(p), C.cast(ptr, C.c_void_p), arr.size, float(dash_offset),
                                             float(width), float(miter_limit), CAPS[cap] if isinstance(cap, str) else int(cap),
                                             JOINS[join] if isinstance(join, str) else int(join), float(res_scale),
                                             C.byref(ov), C.byref(nv), C.byref(op), C.byref(np_))
    if st != 0:
        return None
    try:
        out_v = np.ctypeslib.as_array((C.c_uint8 * nv.value).from_address(ov.value)).copy()
        out_p = np.ctypeslib.as_array((C.c_float * (np_.value * 2)).from_address(op.value)).copy().reshape(-1, 2)
    finally:
        lib.rb_path_free(ov)
        lib.rb_path_free(op)
    return out_v, out_p
