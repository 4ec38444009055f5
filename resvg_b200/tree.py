"""Host-side mirror of the whole-tree entry points: ``resvg::render(&usvg::Tree, Transform, &mut PixmapMut)``
(crates/resvg/src/lib.rs:34), ``resvg::render_node`` (lib.rs:55) and the C API's ``resvg_render`` (c-api/lib.rs:875).

``serialize(scene)`` writes the "RBT1" stream documented in include/resvg_b200.h — the job the Rust shim does by walking
``usvg::Tree``'s accessors.  A *scene* is the plain-dict image of a usvg tree (groups, paths with
resolved paints, clip paths, masks, filters, bounding boxes) that the test-side parser produces; nothing here parses SVG.
"""
import ctypes as C
import struct

import numpy as np

from ._ffi import lib

VERBS = {"M": 0, "L": 1, "Q": 2, "C": 3, "Z": 4}
BLEND = {n: i for i, n in enumerate(["normal", "multiply", "screen", "overlay", "darken", "lighten", "color-dodge", "color-burn",
                                     "hard-light", "soft-light", "difference", "exclusion", "hue", "saturation", "color",
                                     "luminosity"])}
SPREAD = {"pad": 0, "reflect": 1, "repeat": 2}
CAP = {"butt": 0, "round": 1, "square": 2}
JOIN = {"miter": 0, "miter-clip": 1, "round": 2, "bevel": 3}
QUALITY = {"nearest": 0, "bilinear": 1, "bicubic": 2}
COMPOSITE = {"over": 0, "in": 1, "out": 2, "atop": 3, "xor": 4, "arithmetic": 5}
TRANSFER = {"identity": 0, "table": 1, "discrete": 2, "linear": 3, "gamma": 4}
COLOR_MATRIX = {"matrix": 0, "saturate": 1, "hueRotate": 2, "luminanceToAlpha": 3}
EDGE = {"none": 0, "duplicate": 1, "wrap": 2}
LIGHT = {"distant": 0, "point": 1, "spot": 2}
PRIM = {"blend": 0, "drop_shadow": 1, "flood": 2, "blur": 3, "offset": 4, "composite": 5, "merge": 6, "tile": 7, "image": 8,
        "component_transfer": 9, "color_matrix": 10, "convolve": 11, "morphology": 12, "displacement": 13, "turbulence": 14,
        "diffuse": 15, "specular": 16}
IDENT = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
DEFAULT_BBOX = (0.0, 0.0, 1.0, 1.0)


class _W:
    def __init__(self):
        self.parts = []

    def u32(self, *v):
        self.parts.append(struct.pack(f"<{len(v)}I", *[int(x) & 0xFFFFFFFF for x in v]))

    def f32(self, *v):
        self.parts.append(struct.pack(f"<{len(v)}f", *[float(x) for x in v]))

    def raw(self, b):
        self.parts.append(b)
        pad = (-len(b)) % 4
        if pad:
            self.parts.append(b"\0" * pad)

    def str(self, s):
        b = (s or "").encode("utf-8")
        self.u32(len(b))
        self.raw(b)

    def rgb(self, c):
        self.u32(int(c[0]) | int(c[1]) << 8 | int(c[2]) << 16)

    def bytes(self):
        return b"".join(self.parts)


def _group(w, g):
    w.str(g.get("id", ""))
    w.f32(*g.get("ts", IDENT))
    w.f32(g.get("opacity", 1.0))
    w.u32(BLEND[g.get("blend", "normal")], 1 if g.get("isolate_attr") else 0)
    w.f32(*(g.get("layer_bbox") or DEFAULT_BBOX))
    w.f32(*(g.get("abs_layer_bbox") or DEFAULT_BBOX))
    clip = g.get("clip")
    w.u32(1 if clip else 0)
    if clip:
        _clip(w, clip)
    mask = g.get("mask")
    w.u32(1 if mask else 0)
    if mask:
        _mask(w, mask)
    filters = g.get("filters") or []
    w.u32(len(filters))
    for f in filters:
        w.f32(*f["rect"])
        w.u32(len(f["primitives"]))
        for p in f["primitives"]:
            _primitive(w, p)
    kids = g.get("children", [])
    w.u32(len(kids))
    for n in kids:
        if n["t"] == "g":
            w.u32(0)
            _group(w, n)
        elif n["t"] == "path":
            w.u32(1)
            _path(w, n)
        else:
            w.u32(2)
            _image(w, n)


def _clip(w, c):
    w.f32(*c.get("ts", IDENT))
    w.u32(1 if c.get("clip") else 0)
    if c.get("clip"):
        _clip(w, c["clip"])
    _group(w, {"children": c["children"]})  # clip.root()


def _mask(w, m):
    w.f32(*m["rect"])
    w.u32(1 if m["kind"] == "alpha" else 0, 1 if m.get("mask") else 0)
    if m.get("mask"):
        _mask(w, m["mask"])
    _group(w, m["root"])


def _paint(w, p):
    k = p["kind"]
    if k == "color":
        w.u32(0)
        w.rgb(p["rgb"])
    elif k in ("linear", "radial"):
        if k == "linear":
            w.u32(1)
            w.f32(p["x0"], p["y0"], p["x1"], p["y1"])
        else:
            w.u32(2)
            w.f32(p["x1"], p["y1"], p["r1"], p["x0"], p["y0"], p["r0"])  # cx cy r fx fy fr
        w.u32(SPREAD[p.get("spread", "pad")])
        w.f32(*p.get("ts", IDENT))
        w.u32(len(p["stops"]))
        for o, r, g, b, so in p["stops"]:
            w.f32(o)
            w.rgb((r, g, b))
            w.f32(so)
    else:  # pattern_tree
        w.u32(3)
        w.f32(*p["rect"])
        w.f32(*p.get("ts", IDENT))
        _group(w, p["root"])


def _path(w, n):
    w.str(n.get("id", ""))
    ab = n.get("abs_bbox")
    w.u32(1 if n.get("visible", True) else 0, 1 if n.get("stroke_first") else 0, 1 if n.get("aa", True) else 0, 1 if ab else 0)
    w.f32(*(ab or (0.0, 0.0, 0.0, 0.0)))
    f = n.get("fill")
    w.u32(1 if f else 0)
    if f:
        _paint(w, f["paint"])
        w.f32(f.get("opacity", 1.0))
        w.u32(1 if f.get("rule") == "evenodd" else 0)
    s = n.get("stroke")
    w.u32(1 if s else 0)
    if s:
        _paint(w, s["paint"])
        w.f32(s.get("opacity", 1.0), s["width"], s.get("miter", 4.0))
        w.u32(CAP[s.get("cap", "butt")], JOIN[s.get("join", "miter")])
        dash = s.get("dash") or []
        w.u32(len(dash))
        if dash:
            w.f32(*dash)
        w.f32(s.get("dash_offset", 0.0))
    verbs = bytes(int(v) for v in n["verbs"])
    w.u32(len(verbs))
    w.raw(verbs)
    pts = np.asarray(n["pts"], np.float32).reshape(-1, 2)
    w.u32(len(pts))
    w.raw(pts.tobytes())


def _image(w, n):
    w.str(n.get("id", ""))
    ab = n.get("abs_bbox")
    w.u32(1 if n.get("visible", True) else 0, QUALITY[n.get("quality", "bicubic")], 1 if ab else 0)
    w.f32(*(ab or (0.0, 0.0, 0.0, 0.0)))
    if n["kind"] == "svg":
        w.u32(0)
        _tree(w, n["tree"])
    else:
        if "pixels_z" in n:  # deflate + base64 (the JSON form of a decoded pixmap)
            import base64
            import zlib
            px = np.frombuffer(zlib.decompress(base64.b64decode(n["pixels_z"])), np.uint8).reshape(n["h"], n["w"], 4)
        else:
            px = np.ascontiguousarray(n["pixels"], np.uint8).reshape(n["h"], n["w"], 4)
        w.u32(1, n["w"], n["h"])
        w.raw(px.tobytes())


def _input(w, i):
    if i[0] == "source":
        w.u32(0)
    elif i[0] == "alpha":
        w.u32(1)
    else:
        w.u32(2)
        w.str(i[1])


def _primitive(w, p):
    k = p["kind"]
    w.f32(*p["rect"])
    w.u32(0 if p.get("cs", "linearRGB") == "sRGB" else 1)
    w.str(p.get("result", ""))
    w.u32(PRIM[k])
    if k == "blend":
        w.u32(BLEND[p["mode"]])
        _input(w, p["in"])
        _input(w, p["in2"])
    elif k == "drop_shadow":
        _input(w, p["in"])
        w.f32(p["dx"], p["dy"], p["sx"], p["sy"])
        w.rgb(p["color"])
        w.f32(p["opacity"])
    elif k == "flood":
        w.rgb(p["color"])
        w.f32(p["opacity"])
    elif k == "blur":
        _input(w, p["in"])
        w.f32(p["sx"], p["sy"])
    elif k == "offset":
        _input(w, p["in"])
        w.f32(p["dx"], p["dy"])
    elif k == "composite":
        w.u32(COMPOSITE[p["op"]])
        w.f32(*p["k"])
        _input(w, p["in"])
        _input(w, p["in2"])
    elif k == "merge":
        w.u32(len(p["inputs"]))
        for i in p["inputs"]:
            _input(w, i)
    elif k == "tile":
        _input(w, p["in"])
    elif k == "image":
        _group(w, p["root"])
    elif k == "component_transfer":
        _input(w, p["in"])
        for f in p["funcs"]:
            vals = f.get("values", [])
            w.u32(TRANSFER[f["kind"]], len(vals))
            if vals:
                w.f32(*vals)
            w.f32(f.get("slope", 1.0), f.get("intercept", 0.0), f.get("amplitude", 1.0), f.get("exponent", 1.0), f.get("offset", 0.0))
    elif k == "color_matrix":
        _input(w, p["in"])
        w.u32(COLOR_MATRIX[p["cm_kind"]], len(p["params"]))
        if p["params"]:
            w.f32(*p["params"])
    elif k == "convolve":
        _input(w, p["in"])
        w.u32(p["cols"], p["rows"], p["tx"], p["ty"])
        w.f32(p["divisor"], p["bias"])
        w.u32(EDGE[p["edge"]], 1 if p["preserve_alpha"] else 0, len(p["matrix"]))
        w.f32(*p["matrix"])
    elif k == "morphology":
        _input(w, p["in"])
        w.u32(1 if p["op"] == "dilate" else 0)
        w.f32(p["rx"], p["ry"])
    elif k == "displacement":
        _input(w, p["in"])
        _input(w, p["in2"])
        w.f32(p["scale"])
        w.u32(p["xch"], p["ych"])
    elif k == "turbulence":
        w.f32(p["bfx"], p["bfy"])
        w.u32(p["octaves"], p["seed"], 1 if p["stitch"] else 0, 1 if p["fractal"] else 0)
    elif k in ("diffuse", "specular"):
        _input(w, p["in"])
        w.f32(p["surface_scale"], p["constant"], p.get("exponent", 1.0))
        w.rgb(p["color"])
        l = p["light"]
        pa = l.get("points_at", (0.0, 0.0, 0.0))
        cone = l.get("limiting_cone_angle")
        w.u32(LIGHT[l["kind"]])
        w.f32(l.get("azimuth", 0.0), l.get("elevation", 0.0), l.get("x", 0.0), l.get("y", 0.0), l.get("z", 0.0), pa[0], pa[1], pa[2],
              l.get("specular_exponent", 1.0))
        w.u32(0 if cone is None else 1)
        w.f32(0.0 if cone is None else cone)
    else:
        raise ValueError(f"unknown primitive {k}")


def _tree(w, scene):
    w.f32(scene["width"], scene["height"])
    # usvg's root group (identity transform); the scene's root is its child carrying the viewBox transform
    _group(w, {"children": [scene["root"]]})


def serialize(scene) -> bytes:
    w = _W()
    w.u32(0x31544252)
    _tree(w, scene)
    return w.bytes()


class Tree:
    """rb_tree: a parsed usvg tree, reusable across renders (resvg_render_tree of the C API)."""

    def __init__(self, scene_or_bytes):
        blob = scene_or_bytes if isinstance(scene_or_bytes, (bytes, bytearray)) else serialize(scene_or_bytes)
        self.blob = bytes(blob)
        h = C.c_void_p()
        st = lib.rb_tree_parse(self.blob, len(self.blob), C.byref(h))
        if st != 0:
            raise ValueError("rb_tree_parse: malformed tree stream")
        self._h = h

    @property
    def size(self):
        w, h = C.c_float(), C.c_float()
        lib.rb_tree_size(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def node_bbox(self, node_id):
        """node.abs_layer_bounding_box() -> (x, y, w, h) or None."""
        out = (C.c_float * 4)()
        st = lib.rb_tree_node_bbox(self._h, node_id.encode(), out)
        return None if st != 0 else tuple(out)

    def close(self):
        if self._h:
            lib.rb_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def _ts6(ts):
    return (C.c_float * 6)(*[float(v) for v in ts])


def render(tree: Tree, ts, layer):
    """resvg::render(tree, transform, pixmap)"""
    layer.ctx.check(lib.rb_render(layer.ctx._h, tree._h, _ts6(ts), layer._h), "rb_render")


def render_strip(tree: Tree, ts, canvas_w: int, canvas_h: int, y0: int, layer):
    """Rows [y0, y0 + layer.height) of the canvas_w x canvas_h render of `tree` (rb_render_strip: canvas-strip sharding of one
    document across GPUs, bit-identical to the whole-canvas render)."""
    layer.ctx.check(lib.rb_render_strip(layer.ctx._h, tree._h, _ts6(ts), canvas_w, canvas_h, y0, layer._h), "rb_render_strip")


def render_node(tree: Tree, node_id: str, ts, layer) -> bool:
    """resvg::render_node(node, transform, pixmap) -> False for the reference's None."""
    st = lib.rb_render_node(layer.ctx._h, tree._h, node_id.encode(), _ts6(ts), layer._h)
    if st == 1:
        return False
    layer.ctx.check(st, "rb_render_node")
    return True


def submit(blob: bytes, ts, layer):
    """rb_submit: parse + render in one call."""
    layer.ctx.check(lib.rb_submit(layer.ctx._h, blob, len(blob), _ts6(ts), layer._h), "rb_submit")


def render_to_host(ctx, tree: Tree, ts, pixmap: np.ndarray):
    """resvg_render over a host pixmap (premultiplied RGBA8, drawn over in place)."""
    assert pixmap.dtype == np.uint8 and pixmap.ndim == 3 and pixmap.shape[2] == 4 and pixmap.flags.c_contiguous
    ctx.check(lib.rb_render_to_host(ctx._h, tree._h, _ts6(ts), pixmap.shape[1], pixmap.shape[0], pixmap.ctypes.data), "rb_render_to_host")
