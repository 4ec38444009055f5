"""Test-side SVG front end: a small subset of usvg (parse -> resolved scene tree) plus a restatement of resvg's render
traversal (crates/resvg/src/{lib,render,path,clip,mask}.rs, filter/mod.rs) that drives a drawing back end
(tests/backends.py: CPU oracle or CUDA library).

It exists because the reference's host side is Rust and cannot be built here: to pin the oracle against the reference's
golden PNGs (crates/resvg/tests/tests/**.png) something has to turn the test SVGs into the call sequence resvg would
issue.  Only static, text-free, raster-image-free documents are handled; `parse` raises Unsupported for anything else.
The scene tree is plain JSON so fixtures can be committed under tests/golden/ without the SVG sources.
"""
import math
import re
import xml.etree.ElementTree as ET

import numpy as np

f32 = np.float32


class Unsupported(Exception):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# tiny_skia_path::Transform in f32 (sx, ky, kx, sy, tx, ty)
# ---------------------------------------------------------------------------------------------------------------------
IDENT = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)


def _f(v):
    return float(f32(v))


def ts_is_identity(t):
    return tuple(t) == IDENT


def ts_has_skew(t):
    return t[1] != 0 or t[2] != 0


def ts_concat(a, b):
    """b applied first (tiny-skia concat(a, b))."""
    if ts_is_identity(a):
        return tuple(b)
    if ts_is_identity(b):
        return tuple(a)
    asx, aky, akx, asy, atx, aty = [f32(v) for v in a]
    bsx, bky, bkx, bsy, btx, bty = [f32(v) for v in b]
    if not ts_has_skew(a) and not ts_has_skew(b):
        return (_f(asx * bsx), 0.0, 0.0, _f(asy * bsy), _f(asx * btx + atx), _f(asy * bty + aty))

    def mam(p, q, r, s):
        return f32(float(p) * float(q) + float(r) * float(s))

    return (_f(mam(asx, bsx, akx, bky)), _f(mam(aky, bsx, asy, bky)), _f(mam(asx, bkx, akx, bsy)),
            _f(mam(aky, bkx, asy, bsy)), _f(mam(asx, btx, akx, bty) + atx), _f(mam(aky, btx, asy, bty) + aty))


def ts_pre(a, b):  # a.pre_concat(b)
    return ts_concat(a, b)


def ts_post(a, b):  # a.post_concat(b)
    return ts_concat(b, a)


def ts_translate(x, y):
    return (1.0, 0.0, 0.0, 1.0, _f(x), _f(y))


def ts_scale(x, y):
    return (_f(x), 0.0, 0.0, _f(y), 0.0, 0.0)


def ts_get_scale(t):
    sx = f32(math.sqrt(float(f32(t[0]) * f32(t[0]) + f32(t[2]) * f32(t[2]))))
    sy = f32(math.sqrt(float(f32(t[1]) * f32(t[1]) + f32(t[3]) * f32(t[3]))))
    return _f(sx), _f(sy)


def ts_map(t, x, y):
    x, y = f32(x), f32(y)
    if ts_is_identity(t):
        return _f(x), _f(y)
    sx, ky, kx, sy, tx, ty = [f32(v) for v in t]
    if not ts_has_skew(t):
        if sx == 1 and sy == 1:
            return _f(x + tx), _f(y + ty)
        return _f(x * sx + tx), _f(y * sy + ty)
    return _f(x * sx + y * kx + tx), _f(x * ky + y * sy + ty)


def ts_from_bbox(b):
    return (_f(b[2]), 0.0, 0.0, _f(b[3]), _f(b[0]), _f(b[1]))


def rect_transform(r, t):
    """Rect::transform: bbox of the mapped corners; r = (x, y, w, h)."""
    x, y, w, h = r
    pts = [ts_map(t, x, y), ts_map(t, x + w, y), ts_map(t, x + w, y + h), ts_map(t, x, y + h)]
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    return (min(xs), min(ys), _f(f32(max(xs)) - f32(min(xs))), _f(f32(max(ys)) - f32(min(ys))))


def to_int_rect(r):
    """Rect::to_int_rect: floor x/y, ceil w/h (>= 1)."""
    return (math.floor(r[0]), math.floor(r[1]), max(1, math.ceil(r[2])), max(1, math.ceil(r[3])))


def fit_to_rect(r, b):
    l, t = max(r[0], b[0]), max(r[1], b[1])
    rr, bb = min(r[0] + r[2], b[0] + b[2]), min(r[1] + r[3], b[1] + b[3])
    if rr <= l or bb <= t:
        return None
    return (l, t, rr - l, bb - t)


# ---------------------------------------------------------------------------------------------------------------------
# parsing helpers
# ---------------------------------------------------------------------------------------------------------------------
NAMED = {"black": (0, 0, 0), "white": (255, 255, 255), "red": (255, 0, 0), "green": (0, 128, 0), "blue": (0, 0, 255),
         "yellow": (255, 255, 0), "gray": (128, 128, 128), "grey": (128, 128, 128), "lime": (0, 255, 0),
         "orange": (255, 165, 0), "purple": (128, 0, 128), "seagreen": (46, 139, 87), "none": None,
         "pink": (255, 192, 203), "cyan": (0, 255, 255), "aqua": (0, 255, 255), "magenta": (255, 0, 255),
         "fuchsia": (255, 0, 255), "silver": (192, 192, 192), "maroon": (128, 0, 0), "olive": (128, 128, 0),
         "navy": (0, 0, 128), "teal": (0, 128, 128), "darkblue": (0, 0, 139), "lightblue": (173, 216, 230),
         "steelblue": (70, 130, 180), "skyblue": (135, 206, 235), "darkgreen": (0, 100, 0), "brown": (165, 42, 42),
         "gold": (255, 215, 0), "indigo": (75, 0, 130), "violet": (238, 130, 238), "coral": (255, 127, 80),
         "tomato": (255, 99, 71), "crimson": (220, 20, 60), "khaki": (240, 230, 140), "salmon": (250, 128, 114),
         "lightgreen": (144, 238, 144), "lightgray": (211, 211, 211), "lightgrey": (211, 211, 211),
         "darkgray": (169, 169, 169), "darkgrey": (169, 169, 169), "dimgray": (105, 105, 105),
         "darkorange": (255, 140, 0), "greenyellow": (173, 255, 47), "forestgreen": (34, 139, 34),
         "royalblue": (65, 105, 225), "slategray": (112, 128, 144), "wheat": (245, 222, 179), "tan": (210, 180, 140),
         "plum": (221, 160, 221), "orchid": (218, 112, 214), "beige": (245, 245, 220), "azure": (240, 255, 255),
         "darkred": (139, 0, 0), "firebrick": (178, 34, 34), "chocolate": (210, 105, 30), "peru": (205, 133, 63),
         "turquoise": (64, 224, 208), "aquamarine": (127, 255, 212), "lavender": (230, 230, 250),
         "mediumseagreen": (60, 179, 113), "mediumblue": (0, 0, 205), "midnightblue": (25, 25, 112),
         "blueviolet": (138, 43, 226), "darkviolet": (148, 0, 211), "hotpink": (255, 105, 180),
         "deeppink": (255, 20, 147), "yellowgreen": (154, 205, 50), "olivedrab": (107, 142, 35),
         "cornflowerblue": (100, 149, 237), "dodgerblue": (30, 144, 255), "deepskyblue": (0, 191, 255),
         "cadetblue": (95, 158, 160), "darkcyan": (0, 139, 139), "lightyellow": (255, 255, 224),
         "lightcoral": (240, 128, 128), "indianred": (205, 92, 92), "sienna": (160, 82, 45),
         "darkkhaki": (189, 183, 107), "gainsboro": (220, 220, 220), "whitesmoke": (245, 245, 245),
         "ivory": (255, 255, 240), "linen": (250, 240, 230), "snow": (255, 250, 250), "mistyrose": (255, 228, 225)}

INHERITED = {"fill", "fill-rule", "fill-opacity", "stroke", "stroke-width", "stroke-linecap", "stroke-linejoin",
             "stroke-miterlimit", "stroke-dasharray", "stroke-dashoffset", "stroke-opacity", "clip-rule",
             "shape-rendering", "visibility", "paint-order", "color", "color-interpolation-filters", "marker-start",
             "marker-mid", "marker-end", "font-size"}
UNSUPPORTED_ELEMS = {"text", "marker", "switch", "style", "a", "tspan", "textPath", "symbol",
                     "foreignObject", "script", "animate", "set", "animateTransform", "filter-unsupported"}


def parse_color(s, current=(0, 0, 0)):
    """-> (r, g, b, a) with r,g,b in 0..255 ints, a float, or None for `none`."""
    s = s.strip()
    low = s.lower()
    if low in ("none",):
        return None
    if low == "currentcolor":
        return (*current, 1.0)
    if low == "transparent":
        return (0, 0, 0, 0.0)
    if low in NAMED:
        return (*NAMED[low], 1.0)
    if low.startswith("#"):
        h = low[1:]
        if len(h) == 3:
            return (int(h[0] * 2, 16), int(h[1] * 2, 16), int(h[2] * 2, 16), 1.0)
        if len(h) == 6:
            return (int(h[0:2], 16), int(h[2:4], 16), int(h[4:6], 16), 1.0)
        if len(h) == 8:
            return (int(h[0:2], 16), int(h[2:4], 16), int(h[4:6], 16), int(h[6:8], 16) / 255.0)
        if len(h) == 4:
            return (int(h[0] * 2, 16), int(h[1] * 2, 16), int(h[2] * 2, 16), int(h[3] * 2, 16) / 255.0)
    m = re.match(r"rgba?\(([^)]*)\)", low)
    if m:
        parts = [p.strip() for p in re.split(r"[,\s/]+", m.group(1).strip()) if p.strip()]
        vals = []
        for p in parts[:3]:
            if p.endswith("%"):
                vals.append(int(min(max(float(p[:-1]), 0), 100) * 255 / 100 + 0.5))
            else:
                vals.append(int(min(max(round(float(p)), 0), 255)))
        a = 1.0
        if len(parts) > 3:
            a = float(parts[3][:-1]) / 100 if parts[3].endswith("%") else float(parts[3])
        return (*vals, min(max(a, 0.0), 1.0))
    raise Unsupported(f"color {s!r}")


NUM = r"[-+]?(?:\d*\.\d+|\d+\.?)(?:[eE][-+]?\d+)?"


def parse_length(s, default=0.0, ref=None, font=12.0):
    if s is None:
        return default
    s = s.strip()
    m = re.match(rf"^({NUM})\s*([a-zA-Z%]*)$", s)
    if not m:
        raise Unsupported(f"length {s!r}")
    v, u = float(m.group(1)), m.group(2)
    if u in ("", "px"):
        return v
    if u == "%":
        if ref is None:
            raise Unsupported("percent length without reference")
        return v * ref / 100.0
    k = {"in": 96.0, "cm": 96.0 / 2.54, "mm": 96.0 / 25.4, "pt": 96.0 / 72.0, "pc": 16.0, "em": font, "ex": font / 2.0}
    if u in k:
        return v * k[u]
    raise Unsupported(f"unit {u!r}")


def parse_transform(s):
    t = IDENT
    if not s:
        return t
    for name, args in re.findall(r"(\w+)\s*\(([^)]*)\)", s):
        a = [float(x) for x in re.findall(NUM, args)]
        if name == "matrix" and len(a) == 6:
            m = tuple(_f(v) for v in a)
        elif name == "translate":
            m = ts_translate(a[0], a[1] if len(a) > 1 else 0.0)
        elif name == "scale":
            m = ts_scale(a[0], a[1] if len(a) > 1 else a[0])
        elif name == "rotate":
            r = math.radians(a[0])
            c, s_ = math.cos(r), math.sin(r)
            m = (_f(c), _f(s_), _f(-s_), _f(c), 0.0, 0.0)
            if len(a) == 3:
                m64 = _mat64_mul(_mat64_mul((1, 0, 0, 1, a[1], a[2]), (c, s_, -s_, c, 0, 0)), (1, 0, 0, 1, -a[1], -a[2]))
                m = tuple(_f(v) for v in m64)
        elif name == "skewX":
            m = (1.0, 0.0, _f(math.tan(math.radians(a[0]))), 1.0, 0.0, 0.0)
        elif name == "skewY":
            m = (1.0, _f(math.tan(math.radians(a[0]))), 0.0, 1.0, 0.0, 0.0)
        else:
            raise Unsupported(f"transform {name}")
        t = ts_pre(t, m)
    return t


def _mat64_mul(a, b):
    return (a[0] * b[0] + a[2] * b[1], a[1] * b[0] + a[3] * b[1], a[0] * b[2] + a[2] * b[3], a[1] * b[2] + a[3] * b[3],
            a[0] * b[4] + a[2] * b[5] + a[4], a[1] * b[4] + a[3] * b[5] + a[5])


# ---- path data (svgtypes SimplifyingPathParser + kurbo arcs) ----
M, L, Q, C_, Z = 0, 1, 2, 3, 4


def _arc_to_cubics(x0, y0, rx, ry, rot_deg, large, sweep, x, y):
    """kurbo::Arc::from_svg_arc + to_cubic_beziers(0.1) in f64."""
    if abs(rx) <= 1e-5 or abs(ry) <= 1e-5 or (x0 == x and y0 == y):
        return None
    rx, ry = abs(rx), abs(ry)
    xr = math.radians(rot_deg) % (2 * math.pi)
    sin_phi, cos_phi = math.sin(xr), math.cos(xr)
    hd_x, hd_y = (x0 - x) * 0.5, (y0 - y) * 0.5
    hs_x, hs_y = (x0 + x) * 0.5, (y0 + y) * 0.5
    px, py = cos_phi * hd_x + sin_phi * hd_y, -sin_phi * hd_x + cos_phi * hd_y
    rf = px * px / (rx * rx) + py * py / (ry * ry)
    if rf > 1.0:
        sc = math.sqrt(rf)
        rx *= sc
        ry *= sc
    rxry, rxpy, rypx = rx * ry, rx * py, ry * px
    sum_sq = rxpy * rxpy + rypx * rypx
    sign = -1.0 if large == sweep else 1.0
    coe = sign * math.sqrt(abs((rxry * rxry - sum_sq) / sum_sq))
    tcx, tcy = coe * rxpy / ry, -coe * rypx / rx
    cx = cos_phi * tcx - sin_phi * tcy + hs_x
    cy = sin_phi * tcx + cos_phi * tcy + hs_y
    sv = ((px - tcx) / rx, (py - tcy) / ry)
    ev = ((-px - tcx) / rx, (-py - tcy) / ry)
    start = math.atan2(sv[1], sv[0])
    sweep_a = math.fmod(math.atan2(ev[1], ev[0]) - start, 2 * math.pi)
    if sweep and sweep_a < 0:
        sweep_a += 2 * math.pi
    elif not sweep and sweep_a > 0:
        sweep_a -= 2 * math.pi
    x_rot = math.radians(rot_deg)
    sgn = math.copysign(1.0, sweep_a)
    scaled_err = max(rx, ry) / 0.1
    n_err = max((1.1163 * scaled_err) ** (1.0 / 6.0), 3.999999)
    n = math.ceil(n_err * abs(sweep_a) * (1.0 / (2 * math.pi)))
    step = sweep_a / n
    arm = (4.0 / 3.0) * math.tan(abs(0.25 * step)) * sgn

    def sample(a):
        u, v = rx * math.cos(a), ry * math.sin(a)
        s_, c_ = math.sin(x_rot), math.cos(x_rot)
        return (u * c_ - v * s_, u * s_ + v * c_)

    out = []
    a0 = start
    p0 = sample(a0)
    for _ in range(int(n)):
        a1 = a0 + step
        d0 = sample(a0 + math.pi / 2)
        p1 = (p0[0] + arm * d0[0], p0[1] + arm * d0[1])
        p3 = sample(a1)
        d1 = sample(a1 + math.pi / 2)
        p2 = (p3[0] - arm * d1[0], p3[1] - arm * d1[1])
        out.append((cx + p1[0], cy + p1[1], cx + p2[0], cy + p2[1], cx + p3[0], cy + p3[1]))
        a0, p0 = a1, p3
    return out


def parse_path(d):
    toks = re.findall(rf"[MmLlHhVvCcSsQqTtAaZz]|{NUM}", d)
    verbs, pts = [], []
    i = 0
    cx = cy = sx = sy = 0.0
    prev_cmd = None
    pcx = pcy = 0.0  # previous control point
    cmd = None
    first = True

    def num():
        nonlocal i
        v = float(toks[i])
        i += 1
        return v

    def flag():
        nonlocal i
        t = toks[i]
        # flags may be glued ("01") — the NUM regex splits "01" as one token; handle single chars
        if len(t) > 1 and t[0] in "01" and re.fullmatch(r"[01]+.*", t):
            toks[i] = t[1:]
            return int(t[0])
        i += 1
        return int(float(t))

    def ensure_move():
        # after ClosePath an implicit MoveTo to the sub-path start is needed
        if verbs and verbs[-1] == Z:
            verbs.append(M)
            pts.append((sx, sy))

    try:
        while i < len(toks):
            if re.fullmatch(r"[A-Za-z]", toks[i]):
                cmd = toks[i]
                i += 1
                if cmd in "Zz":
                    if verbs and verbs[-1] != Z:
                        verbs.append(Z)
                    cx, cy = sx, sy
                    prev_cmd = "Z"
                    continue
            elif cmd is None:
                break
            elif cmd in "Mm" and prev_cmd in ("M",):
                cmd = "L" if cmd == "M" else "l"
            if first and cmd not in "Mm":
                break
            first = False
            rel = cmd.islower()
            c = cmd.upper()
            if c == "M":
                x, y = num(), num()
                if rel:
                    x, y = cx + x, cy + y
                verbs.append(M)
                pts.append((x, y))
                cx, cy, sx, sy = x, y, x, y
                prev_cmd = "M"
                continue
            ensure_move()
            if c == "L":
                x, y = num(), num()
                if rel:
                    x, y = cx + x, cy + y
                verbs.append(L); pts.append((x, y)); cx, cy = x, y
            elif c == "H":
                x = num()
                if rel:
                    x = cx + x
                verbs.append(L); pts.append((x, cy)); cx = x
            elif c == "V":
                y = num()
                if rel:
                    y = cy + y
                verbs.append(L); pts.append((cx, y)); cy = y
            elif c == "C":
                v = [num() for _ in range(6)]
                if rel:
                    v = [v[0] + cx, v[1] + cy, v[2] + cx, v[3] + cy, v[4] + cx, v[5] + cy]
                verbs.append(C_); pts += [(v[0], v[1]), (v[2], v[3]), (v[4], v[5])]
                pcx, pcy, cx, cy = v[2], v[3], v[4], v[5]
            elif c == "S":
                v = [num() for _ in range(4)]
                if rel:
                    v = [v[0] + cx, v[1] + cy, v[2] + cx, v[3] + cy]
                x1, y1 = (2 * cx - pcx, 2 * cy - pcy) if prev_cmd in ("C", "S") else (cx, cy)
                verbs.append(C_); pts += [(x1, y1), (v[0], v[1]), (v[2], v[3])]
                pcx, pcy, cx, cy = v[0], v[1], v[2], v[3]
            elif c == "Q":
                v = [num() for _ in range(4)]
                if rel:
                    v = [v[0] + cx, v[1] + cy, v[2] + cx, v[3] + cy]
                verbs.append(Q); pts += [(v[0], v[1]), (v[2], v[3])]
                pcx, pcy, cx, cy = v[0], v[1], v[2], v[3]
            elif c == "T":
                x, y = num(), num()
                if rel:
                    x, y = cx + x, cy + y
                x1, y1 = (2 * cx - pcx, 2 * cy - pcy) if prev_cmd in ("Q", "T") else (cx, cy)
                verbs.append(Q); pts += [(x1, y1), (x, y)]
                pcx, pcy, cx, cy = x1, y1, x, y
            elif c == "A":
                rx, ry, rot = num(), num(), num()
                la, sw = flag(), flag()
                x, y = num(), num()
                if rel:
                    x, y = cx + x, cy + y
                cubs = _arc_to_cubics(cx, cy, rx, ry, rot, bool(la), bool(sw), x, y)
                if cubs is None:
                    verbs.append(L); pts.append((x, y))
                else:
                    for cb in cubs:
                        verbs.append(C_); pts += [(cb[0], cb[1]), (cb[2], cb[3]), (cb[4], cb[5])]
                cx, cy = x, y
            prev_cmd = c
    except (IndexError, ValueError):
        pass  # svgtypes stops at the first error and keeps what it has
    # PathBuilder::finish: a trailing lone MoveTo is dropped; fewer than 2 verbs -> None
    while verbs and verbs[-1] == M:
        verbs.pop(); pts.pop()
    if len(verbs) < 2:
        return None
    return verbs, [(_f(x), _f(y)) for x, y in pts]


def ellipse_path(cx, cy, rx, ry):
    cx, cy, rx, ry = f32(cx), f32(cy), f32(rx), f32(ry)
    verbs, pts = [M], [(float(cx + rx), float(cy))]
    cur = (float(cx + rx), float(cy))
    for (x, y) in [(cx, cy + ry), (cx - rx, cy), (cx, cy - ry), (cx + rx, cy)]:
        cubs = _arc_to_cubics(cur[0], cur[1], float(rx), float(ry), 0.0, False, True, float(x), float(y))
        if cubs is None:
            verbs.append(L); pts.append((float(x), float(y)))
        else:
            for cb in cubs:
                verbs.append(C_); pts += [(_f(cb[0]), _f(cb[1])), (_f(cb[2]), _f(cb[3])), (_f(cb[4]), _f(cb[5]))]
        cur = pts[-1]
    verbs.append(Z)
    return verbs, pts


def rect_path(x, y, w, h, rx, ry):
    x, y, w, h = f32(x), f32(y), f32(w), f32(h)
    if rx == 0:
        return [M, L, L, L, Z], [(float(x), float(y)), (float(x + w), float(y)), (float(x + w), float(y + h)), (float(x), float(y + h))]
    rx, ry = f32(rx), f32(ry)
    verbs, pts = [M], [(float(x + rx), float(y))]

    def line(px, py):
        verbs.append(L); pts.append((float(px), float(py)))

    def arc(px, py):
        cur = pts[-1]
        cubs = _arc_to_cubics(cur[0], cur[1], float(rx), float(ry), 0.0, False, True, float(px), float(py))
        if cubs is None:
            line(px, py)
        else:
            for cb in cubs:
                verbs.append(C_); pts.extend([(_f(cb[0]), _f(cb[1])), (_f(cb[2]), _f(cb[3])), (_f(cb[4]), _f(cb[5]))])

    line(x + w - rx, y); arc(x + w, y + ry)
    line(x + w, y + h - ry); arc(x + w - rx, y + h)
    line(x + rx, y + h); arc(x, y + h - ry)
    line(x, y + ry); arc(x + rx, y)
    verbs.append(Z)
    return verbs, pts


def tight_bounds(verbs, pts):
    """Path::compute_tight_bounds: extrema of the curves."""
    xs, ys = [], []
    pi = 0
    last = None
    for v in verbs:
        if v == M or v == L:
            last = pts[pi]; pi += 1
            xs.append(last[0]); ys.append(last[1])
        elif v == Q:
            p0, p1, p2 = last, pts[pi], pts[pi + 1]; pi += 2
            for axis, acc in ((0, xs), (1, ys)):
                a, b, c = p0[axis], p1[axis], p2[axis]
                acc += [a, c]
                d = a - 2 * b + c
                if d != 0:
                    t = (a - b) / d
                    if 0 < t < 1:
                        acc.append((1 - t) ** 2 * a + 2 * t * (1 - t) * b + t * t * c)
            last = p2
        elif v == C_:
            p0, p1, p2, p3 = last, pts[pi], pts[pi + 1], pts[pi + 2]; pi += 3
            for axis, acc in ((0, xs), (1, ys)):
                a, b, c, d = p0[axis], p1[axis], p2[axis], p3[axis]
                acc += [a, d]
                A, B, C2 = -a + 3 * b - 3 * c + d, 2 * (a - 2 * b + c), b - a
                roots = []
                if abs(A) < 1e-12:
                    if abs(B) > 1e-12:
                        roots.append(-C2 / B)
                else:
                    disc = B * B - 4 * A * C2
                    if disc >= 0:
                        sq = math.sqrt(disc)
                        roots += [(-B + sq) / (2 * A), (-B - sq) / (2 * A)]
                for t in roots:
                    if 0 < t < 1:
                        acc.append((1 - t) ** 3 * a + 3 * t * (1 - t) ** 2 * b + 3 * t * t * (1 - t) * c + t ** 3 * d)
            last = p3
    if not xs:
        return None
    return (_f(min(xs)), _f(min(ys)), _f(f32(max(xs)) - f32(min(xs))), _f(f32(max(ys)) - f32(min(ys))))


# ---------------------------------------------------------------------------------------------------------------------
# document -> scene tree
# ---------------------------------------------------------------------------------------------------------------------
def _strip(tag):
    return tag.split("}")[-1]


class Doc:
    def __init__(self, text, base_dir=None):
        self.base_dir = base_dir
        self.root = ET.fromstring(text)
        self.ids = {}
        self.parent = {}
        for el in self.root.iter():
            tag = _strip(el.tag)
            if tag in UNSUPPORTED_ELEMS:
                raise Unsupported(tag)
            for k in list(el.attrib):
                if "}" in k:
                    if _strip(k) == "href":
                        el.attrib["href"] = el.attrib[k]
                    elif _strip(k) in ("space", "lang"):
                        pass
            if "id" in el.attrib:
                self.ids.setdefault(el.attrib["id"], el)
            for ch in el:
                self.parent[ch] = el
            st = el.attrib.get("style")
            if st:
                for decl in st.split(";"):
                    if ":" in decl:
                        k, v = decl.split(":", 1)
                        el.attrib[k.strip()] = v.strip()
            for k in ("systemLanguage", "requiredExtensions", "requiredFeatures"):
                if k in el.attrib:
                    raise Unsupported(k)

    def attr(self, el, name, inherit=None):
        if inherit is None:
            inherit = name in INHERITED
        cur = el
        while cur is not None:
            v = cur.attrib.get(name)
            if v is not None and v.strip() != "inherit":
                return v.strip()
            if not inherit and not (v is not None and v.strip() == "inherit"):
                return None
            cur = self.parent.get(cur)
        return None

    def link(self, value):
        if value is None:
            return None
        m = re.match(r"url\(\s*['\"]?#([^)'\"]+)['\"]?\s*\)", value.strip())
        if m:
            return self.ids.get(m.group(1))
        if value.startswith("#"):
            return self.ids.get(value[1:])
        return None


def opacity_val(s, default=1.0):
    if s is None:
        return default
    s = s.strip()
    v = float(s[:-1]) / 100.0 if s.endswith("%") else float(s)
    return _f(min(max(v, 0.0), 1.0))


def to_u8_opacity(o):
    """strict_num NormalizedF32::to_u8: (v * 255).ceil()"""
    return int(math.ceil(float(f32(o) * f32(255.0))))


def viewbox_transform(vbr, par_str, w, h):
    """usvg ViewBox::to_transform(size)"""
    par = (par_str or "xMidYMid meet").split()
    align = par[0] if par else "xMidYMid"
    slice_ = len(par) > 1 and par[1] == "slice"
    sx, sy = f32(w) / f32(vbr[2]), f32(h) / f32(vbr[3])
    if align == "none":
        ssx, ssy = sx, sy
    else:
        s_ = max(sx, sy) if slice_ else min(sx, sy)
        ssx = ssy = s_
    x = f32(-vbr[0]) * ssx
    y = f32(-vbr[1]) * ssy
    ww, hh = f32(w) - f32(vbr[2]) * ssx, f32(h) - f32(vbr[3]) * ssy
    ax = {"xMin": 0.0, "xMid": 0.5, "xMax": 1.0}.get(align[:4], 0.5) if align != "none" else 0.0
    ay = {"YMin": 0.0, "YMid": 0.5, "YMax": 1.0}.get(align[4:], 0.5) if align != "none" else 0.0
    return (_f(ssx), 0.0, 0.0, _f(ssy), _f(x + ww * f32(ax)), _f(y + hh * f32(ay)))


class Converter:
    def __init__(self, doc):
        self.doc = doc
        self.vb_w = self.vb_h = 100.0
        self.depth = 0
        self.pattern_depth = 0
        self.nested = False

    # ---- root ----
    def convert(self):
        d, root = self.doc, self.doc.root
        if _strip(root.tag) != "svg":
            raise Unsupported("root")
        vb = root.attrib.get("viewBox")
        vbr = None
        if vb:
            n = [float(x) for x in re.findall(NUM, vb)]
            if len(n) == 4 and n[2] > 0 and n[3] > 0:
                vbr = n
        def_w, def_h = (vbr[2], vbr[3]) if vbr else (100.0, 100.0)
        self.vb_w, self.vb_h = def_w, def_h
        w = parse_length(root.attrib.get("width"), def_w, ref=def_w) if root.attrib.get("width", "100%") != "100%" else def_w
        h = parse_length(root.attrib.get("height"), def_h, ref=def_h) if root.attrib.get("height", "100%") != "100%" else def_h
        if vbr is None:
            self.vb_w, self.vb_h = w, h
        if not (w > 0 and h > 0):
            raise Unsupported("size")
        ts = viewbox_transform(vbr, root.attrib.get("preserveAspectRatio"), w, h) if vbr else IDENT
        children = []
        g = self.group_for(root, children_of=root, is_root=True)
        tree = {"t": "g", "ts": list(ts), "children": [g] if g else [], "opacity": 1.0, "isolate": False}
        return {"width": _f(w), "height": _f(h), "root": tree}

    # ---- groups ----
    def group_for(self, el, children_of=None, is_root=False, extra_ts=IDENT):
        d = self.doc
        if d.attr(el, "display", inherit=False) == "none":
            return None
        self.depth += 1
        if self.depth > 256:
            raise Unsupported("depth")
        try:
            ts = IDENT if is_root else ts_pre(self.resolve_transform(el, el.attrib.get("transform")), extra_ts)
            kids = []
            src = children_of if children_of is not None else el
            for ch in src:
                n = self.node_for(ch)
                if n is not None:
                    kids.append(n)
            g = {"t": "g", "id": "" if is_root else el.attrib.get("id", ""), "ts": list(ts), "children": kids}
            self.group_effects(el, g)
            if not kids and not g.get("filters"):
                return None
            return g
        finally:
            self.depth -= 1

    def group_effects(self, el, g):
        d = self.doc
        g["opacity"] = opacity_val(d.attr(el, "opacity", inherit=False))
        blend = d.attr(el, "mix-blend-mode", inherit=False) or "normal"
        if blend not in BLEND_MAP:
            blend = "normal"
        g["blend"] = blend
        iso = d.attr(el, "isolation", inherit=False) == "isolate"
        bbox = self.node_bbox(g)
        cp = d.attr(el, "clip-path", inherit=False)
        g["clip"] = None
        if cp and cp != "none":
            target = d.link(cp)
            if target is None or _strip(target.tag) != "clipPath":
                if cp.startswith("url"):
                    g["children"] = []  # a link to a missing clipPath: usvg removes the element
                else:
                    raise Unsupported("clip-path value")
            else:
                c = self.clip_for(target, bbox)
                if c is None:
                    g["children"] = []
                g["clip"] = c
        mk = d.attr(el, "mask", inherit=False)
        g["mask"] = None
        if mk and mk != "none":
            target = d.link(mk)
            if target is None or _strip(target.tag) != "mask":
                if mk.startswith("url"):
                    g["children"] = []
                else:
                    raise Unsupported("mask value")
            else:
                m = self.mask_for(target, bbox)
                if m is None:
                    g["children"] = []
                g["mask"] = m
        flt = d.attr(el, "filter", inherit=False)
        g["filters"] = []
        if flt and flt != "none":
            if not flt.startswith("url"):
                raise Unsupported("filter functions")
            target = d.link(flt)
            if target is None or _strip(target.tag) != "filter":
                g["children"] = []  # invalid filter link: element is not rendered
            else:
                from tests.svgfilters import convert_filter
                f = convert_filter(self, target, bbox)
                if f is None:
                    g["children"] = []
                else:
                    g["filters"] = [f]
        g["isolate_attr"] = bool(iso)
        g["isolate"] = bool(g["opacity"] != 1.0 or g["clip"] or g["mask"] or g["filters"] or blend != "normal" or iso)

    # ---- nodes ----
    def node_for(self, el):
        tag = _strip(el.tag)
        d = self.doc
        if tag in ("defs", "title", "desc", "metadata", "linearGradient", "radialGradient", "clipPath", "mask", "filter",
                   "stop", "pattern"):
            return None
        if tag == "g":
            return self.group_for(el)
        if tag == "svg":
            raise Unsupported("nested svg")
        if tag == "use":
            return self.use_for(el)
        if tag in ("path", "rect", "circle", "ellipse", "line", "polyline", "polygon"):
            return self.shape_for(el)
        if tag == "image":
            return self.image_for(el)
        if tag.startswith("fe"):
            return None
        raise Unsupported(f"element {tag}")

    def use_for(self, el):
        d = self.doc
        target = d.link(el.attrib.get("href"))
        if target is None:
            return None
        if _strip(target.tag) in ("svg", "symbol"):
            raise Unsupported("use svg/symbol")
        x = parse_length(el.attrib.get("x"), 0.0, ref=self.vb_w)
        y = parse_length(el.attrib.get("y"), 0.0, ref=self.vb_h)
        # the referenced element inherits from the <use>: temporarily re-parent it
        old_parent = d.parent.get(target)
        d.parent[target] = el
        self.depth += 1
        try:
            if self.depth > 32:
                raise Unsupported("use recursion")
            child = self.node_for(target)
        finally:
            self.depth -= 1
            if old_parent is not None:
                d.parent[target] = old_parent
        if child is None:
            return None
        ts = ts_pre(self.resolve_transform(el, el.attrib.get("transform")), ts_translate(x, y))
        g = {"t": "g", "id": el.attrib.get("id", ""), "ts": list(ts), "children": [child]}
        self.group_effects(el, g)
        return g

    def shape_for(self, el):
        d = self.doc
        tag = _strip(el.tag)
        if d.attr(el, "display", inherit=False) == "none":
            return None
        W, H = self.vb_w, self.vb_h
        diag = math.sqrt((W * W + H * H) / 2.0)
        geo = None
        if tag == "path":
            dd = el.attrib.get("d")
            geo = parse_path(dd) if dd else None
        elif tag == "rect":
            w = _f(parse_length(el.attrib.get("width"), 0.0, ref=W))
            h = _f(parse_length(el.attrib.get("height"), 0.0, ref=H))
            if not (w > 0 and h > 0):
                return None
            x = _f(parse_length(el.attrib.get("x"), 0.0, ref=W))
            y = _f(parse_length(el.attrib.get("y"), 0.0, ref=H))
            rx_a, ry_a = el.attrib.get("rx"), el.attrib.get("ry")
            rx = parse_length(rx_a, None, ref=W) if rx_a not in (None, "auto") else None
            ry = parse_length(ry_a, None, ref=H) if ry_a not in (None, "auto") else None
            if rx is not None and rx < 0:
                rx = None
            if ry is not None and ry < 0:
                ry = None
            if rx is None and ry is None:
                rx = ry = 0.0
            elif rx is None:
                rx = ry
            elif ry is None:
                ry = rx
            rx, ry = _f(min(rx, w / 2.0)), _f(min(ry, h / 2.0))
            geo = rect_path(x, y, w, h, rx, ry)
        elif tag == "circle":
            r = _f(parse_length(el.attrib.get("r"), 0.0, ref=diag))
            if not r > 0:
                return None
            geo = ellipse_path(_f(parse_length(el.attrib.get("cx"), 0.0, ref=W)), _f(parse_length(el.attrib.get("cy"), 0.0, ref=H)), r, r)
        elif tag == "ellipse":
            rx_a, ry_a = el.attrib.get("rx"), el.attrib.get("ry")
            rx = parse_length(rx_a, None, ref=W) if rx_a not in (None, "auto") else None
            ry = parse_length(ry_a, None, ref=H) if ry_a not in (None, "auto") else None
            if rx is not None and rx < 0:
                rx = None
            if ry is not None and ry < 0:
                ry = None
            if rx is None and ry is None:
                return None
            rx, ry = (rx if rx is not None else ry), (ry if ry is not None else rx)
            if not (rx > 0 and ry > 0):
                return None
            geo = ellipse_path(_f(parse_length(el.attrib.get("cx"), 0.0, ref=W)), _f(parse_length(el.attrib.get("cy"), 0.0, ref=H)), _f(rx), _f(ry))
        elif tag == "line":
            x1, y1 = _f(parse_length(el.attrib.get("x1"), 0.0, ref=W)), _f(parse_length(el.attrib.get("y1"), 0.0, ref=H))
            x2, y2 = _f(parse_length(el.attrib.get("x2"), 0.0, ref=W)), _f(parse_length(el.attrib.get("y2"), 0.0, ref=H))
            geo = ([M, L], [(x1, y1), (x2, y2)])
        else:
            nums = [float(v) for v in re.findall(NUM, el.attrib.get("points", ""))]
            n = len(nums) // 2
            if n < 2:
                return None
            pts = [(_f(nums[2 * i]), _f(nums[2 * i + 1])) for i in range(n)]
            geo = ([M] + [L] * (n - 1) + ([Z] if tag == "polygon" else []), pts)
        if geo is None:
            return None
        verbs, pts = geo
        for m in ("marker-start", "marker-mid", "marker-end"):
            mv = d.attr(el, m)
            if mv and mv != "none" and tag in ("path", "line", "polyline", "polygon"):
                raise Unsupported("marker")
        bbox = tight_bounds(verbs, pts)
        vis = d.attr(el, "visibility") or "visible"
        node = {"t": "path", "id": el.attrib.get("id", ""), "verbs": verbs, "pts": [list(p) for p in pts], "visible": vis == "visible", "bbox": bbox}
        sr = d.attr(el, "shape-rendering") or "geometricPrecision"
        node["aa"] = sr not in ("optimizeSpeed", "crispEdges")
        po = (d.attr(el, "paint-order") or "normal").split()
        node["stroke_first"] = bool(po) and po[0] == "stroke" or (len(po) > 1 and po[0] == "markers" and po[1] == "stroke")
        color = parse_color(d.attr(el, "color") or "black") or (0, 0, 0, 1.0)
        node["fill"] = self.paint_for(el, "fill", "black", color, bbox)
        if node["fill"]:
            node["fill"]["rule"] = "evenodd" if (d.attr(el, "fill-rule") or "nonzero") == "evenodd" else "nonzero"
        st = self.paint_for(el, "stroke", "none", color, bbox)
        if st:
            width = _f(parse_length(d.attr(el, "stroke-width"), 1.0, ref=diag))
            if not width > 0:
                st = None
            else:
                st["width"] = width
                st["cap"] = d.attr(el, "stroke-linecap") or "butt"
                if st["cap"] not in ("butt", "round", "square"):
                    st["cap"] = "butt"
                join = d.attr(el, "stroke-linejoin") or "miter"
                st["join"] = {"arcs": "miter", "miter-clip": "miter-clip"}.get(join, join)
                if st["join"] not in ("miter", "miter-clip", "round", "bevel"):
                    st["join"] = "miter"
                ml = float(d.attr(el, "stroke-miterlimit") or 4.0)
                if ml < 1:
                    raise Unsupported("miterlimit < 1")
                st["miter"] = _f(ml)
                da = d.attr(el, "stroke-dasharray")
                if da and da != "none":
                    # usvg style.rs conv_dasharray: a negative value or a zero sum means no dashing; an odd list is repeated
                    fs = d.attr(el, "font-size")
                    font = float(parse_length(fs, 12.0, ref=diag)) if fs and re.match(rf"^{NUM}", fs.strip()) else 12.0
                    vals = [_f(parse_length(x, 0.0, ref=diag, font=font)) for x in re.split(r"[\s,]+", da.strip()) if x]
                    if vals and not any(v < 0 or (v == 0 and math.copysign(1.0, v) < 0) for v in vals) and abs(sum(vals)) > 1e-12:
                        if len(vals) % 2:
                            vals = vals + vals
                        st["dash"] = vals
                        st["dash_offset"] = _f(parse_length(d.attr(el, "stroke-dashoffset"), 0.0, ref=diag, font=font))
        node["stroke"] = st
        ts = self.resolve_transform(el, el.attrib.get("transform"))
        g = {"t": "g", "ts": list(ts), "children": [node]}
        self.group_effects(el, g)
        if not g["children"]:
            return None
        if ts_is_identity(ts) and not g["isolate"]:
            return node
        return g

    # ---- images (usvg parser/image.rs) ----
    QUALITY = {"optimizeQuality": "bicubic", "auto": "bicubic", "optimizeSpeed": "nearest", "smooth": "bilinear",
               "high-quality": "bicubic", "crisp-edges": "nearest", "pixelated": "nearest"}

    def load_image(self, href):
        """get_href_data + decode (resvg image.rs:62-170): -> ("svg", scene) | ("raster", w, h, premultiplied RGBA8) | None"""
        import base64
        import gzip
        import io
        import os
        from urllib.parse import unquote_to_bytes
        if href is None:
            return None
        href = href.strip()
        mime = None
        if href.startswith("data:"):
            head, _, payload = href[5:].partition(",")
            mime = head.split(";")[0].strip().lower()
            try:
                data = base64.b64decode("".join(payload.split())) if ";base64" in head else unquote_to_bytes(payload)
            except Exception:
                return None
        else:
            if self.doc.base_dir is None:
                raise Unsupported("external image without a resources dir")
            path = os.path.join(self.doc.base_dir, href)
            if not os.path.exists(path):
                return None
            data = open(path, "rb").read()
        if data[:2] == b"\x1f\x8b":
            try:
                data = gzip.decompress(data)
            except Exception:
                return None
        is_raster = data[:8] == b"\x89PNG\r\n\x1a\n" or data[:3] == b"\xff\xd8\xff" or data[:4] == b"GIF8" or (data[:4] == b"RIFF" and data[8:12] == b"WEBP")
        if not is_raster:
            if self.nested:
                return None  # Tree::from_data_nested: images inside a sub-SVG image are not loaded
            try:
                text = data.decode("utf-8")
                if "<svg" not in text:
                    return None
                return ("svg", parse(text, self.doc.base_dir, nested=True))
            except Unsupported:
                raise
            except Exception:
                return None
        from PIL import Image as PILImage
        try:
            im = PILImage.open(io.BytesIO(data))
            im.seek(0)
            rgba = np.array(im.convert("RGBA"), dtype=np.uint8)
        except Exception:
            return None
        a = rgba[..., 3:4].astype(np.float64) / 255.0
        px = rgba.copy()
        px[..., :3] = (rgba[..., :3].astype(np.float64) * a + 0.5).astype(np.uint8)  # rgba_to_pixmap, image.rs:160-170
        return ("raster", int(rgba.shape[1]), int(rgba.shape[0]), px)

    def image_for(self, el, rect_override=None):
        d = self.doc
        if d.attr(el, "display", inherit=False) == "none":
            return None
        vis = (d.attr(el, "visibility") or "visible") == "visible"
        quality = self.QUALITY.get(d.attr(el, "image-rendering") or "optimizeQuality", "bicubic")
        kind = self.load_image(el.attrib.get("href"))
        if kind is None:
            return None
        if kind[0] == "svg":
            aw, ah = f32(kind[1]["width"]), f32(kind[1]["height"])
        else:
            aw, ah = f32(kind[1]), f32(kind[2])
        if rect_override is not None:
            x, y, w, h = [f32(v) for v in rect_override]
        else:
            x = f32(parse_length(el.attrib.get("x"), 0.0, ref=self.vb_w))
            y = f32(parse_length(el.attrib.get("y"), 0.0, ref=self.vb_h))
            wa, ha = el.attrib.get("width"), el.attrib.get("height")
            w = f32(parse_length(wa, float(aw), ref=self.vb_w))
            h = f32(parse_length(ha, float(ah), ref=self.vb_h))
            if wa is not None and ha is None:
                h = ah * (w / aw)
            elif wa is None and ha is not None:
                w = aw * (h / ah)
        if not (w > 0 and h > 0):
            return None
        par = (el.attrib.get("preserveAspectRatio") or "xMidYMid meet").split()
        if par and par[0] == "defer":
            par = par[1:]
        align = par[0] if par else "xMidYMid"
        slice_ = len(par) > 1 and par[1] == "slice"
        # fit_view_box + aligned_pos
        if align == "none":
            vw, vh = w, h
        else:
            rw = h * aw / ah
            with_h = (rw <= w) if slice_ else (rw >= w)
            if not with_h:
                vw, vh = rw, h
            else:
                vw, vh = w, w * ah / aw
        ax = {"xMin": 0.0, "xMid": 0.5, "xMax": 1.0}.get(align[:4], 0.5) if align != "none" else 0.0
        ay = {"YMin": 0.0, "YMid": 0.5, "YMax": 1.0}.get(align[4:], 0.5) if align != "none" else 0.0
        dx, dy = w - vw, h - vh
        vx = x + (dx / f32(2) if ax == 0.5 else (dx if ax == 1.0 else f32(0)))
        vy = y + (dy / f32(2) if ay == 0.5 else (dy if ay == 1.0 else f32(0)))
        image_ts = (_f(vw / aw), 0.0, 0.0, _f(vh / ah), _f(vx), _f(vy))
        node = {"t": "image", "id": "", "visible": vis, "quality": quality, "bbox": [0.0, 0.0, float(aw), float(ah)]}
        if kind[0] == "svg":
            node.update(kind="svg", tree=kind[1])
        else:
            import base64
            import zlib
            # JSON-friendly: the decoded premultiplied RGBA8 pixmap, deflated and base64-encoded
            node.update(kind="raster", w=kind[1], h=kind[2], pixels_z=base64.b64encode(zlib.compress(kind[3].tobytes(), 9)).decode())
        g = {"t": "g", "id": el.attrib.get("id", "") if rect_override is None else "", "ts": list(image_ts), "children": [node],
             "opacity": 1.0, "blend": "normal", "clip": None, "mask": None, "filters": [], "isolate_attr": False, "isolate": False}
        if slice_:
            # an image slice acts like a rectangular clip, unaffected by the image's own view box transform
            rp = [(float(x), float(y)), (float(x + w), float(y)), (float(x + w), float(y + h)), (float(x), float(y + h))]
            cp = {"t": "path", "id": "", "verbs": [M, L, L, L, Z], "pts": [list(q) for q in rp], "visible": True, "aa": True,
                  "stroke_first": False, "bbox": [float(x), float(y), float(w), float(h)], "stroke": None,
                  "fill": {"paint": {"kind": "color", "rgb": [0, 0, 0]}, "opacity": 1.0, "rule": "nonzero"}}
            g2 = {"t": "g", "id": g["id"], "ts": list(IDENT), "children": [g], "opacity": 1.0, "blend": "normal",
                  "clip": {"ts": list(IDENT), "children": [cp], "clip": None}, "mask": None, "filters": [], "isolate_attr": False,
                  "isolate": True}
            g["id"] = ""
            g = g2
        if rect_override is not None:
            return g
        # the <image> element itself goes through convert_group (opacity, clip-path, mask, filter, transform)
        ts = self.resolve_transform(el, el.attrib.get("transform"))
        outer = {"t": "g", "id": "", "ts": list(ts), "children": [g]}
        self.group_effects(el, outer)
        if not outer["children"]:
            return None
        if ts_is_identity(ts) and not outer["isolate"]:
            return g
        return outer

    def paint_for(self, el, prop, default, color, bbox):
        """usvg style.rs resolve_fill / resolve_stroke + convert_paint: -> {"paint": usvg::Paint, "opacity": Opacity}.
        A colour's alpha (and the single stop a degenerate gradient collapses to) is folded into the opacity."""
        d = self.doc
        v = d.attr(el, prop) or default
        op = opacity_val(d.attr(el, prop + "-opacity"))
        if v == "none":
            return None
        if v.startswith("url"):
            m = re.match(r"(url\([^)]*\))\s*(.*)", v)
            target = d.link(m.group(1))
            fallback = m.group(2).strip()
            if target is not None and _strip(target.tag) in ("linearGradient", "radialGradient"):
                p = self.gradient_for(target, bbox)
                if p == "none":
                    return None
                if p is not None:
                    if p["kind"] == "color":
                        sub = p.pop("sub_opacity")
                        return {"paint": p, "opacity": _f(min(max(f32(sub) * f32(op), f32(0)), f32(1)))}
                    return {"paint": p, "opacity": op}
            elif target is not None and _strip(target.tag) == "pattern":
                p = self.pattern_for(target, bbox)
                if p is not None:
                    return {"paint": p, "opacity": op}
            if fallback:
                if fallback == "none":
                    return None
                c = parse_color(fallback, color[:3])
            elif target is None:
                # a missing link: fill falls back to none (usvg: "fill" -> none, with a warning)
                return None
            else:
                return None
            v = None
        else:
            c = parse_color(v, color[:3])
        if c is None:
            return None
        a = _f(f32(op) * f32(c[3]))
        return {"paint": {"kind": "color", "rgb": [int(c[0]), int(c[1]), int(c[2])]}, "opacity": a}

    # ---- gradients (usvg paint_server.rs) ----
    def _grad_chain(self, el):
        chain, seen = [], set()
        cur = el
        while cur is not None and id(cur) not in seen:
            seen.add(id(cur))
            chain.append(cur)
            cur = self.doc.link(cur.attrib.get("href"))
            if cur is not None and _strip(cur.tag) not in ("linearGradient", "radialGradient"):
                break
        return chain

    def gradient_for(self, el, bbox):
        d = self.doc
        chain = self._grad_chain(el)
        tag = _strip(el.tag)

        def ga(name, same_kind_only=False):
            for c in chain:
                if same_kind_only and _strip(c.tag) != tag:
                    continue
                if name in c.attrib:
                    return c.attrib[name]
            return None

        stops_el = None
        for c in chain:
            if any(_strip(k.tag) == "stop" for k in c):
                stops_el = c
                break
        if stops_el is None:
            return "none"
        stops = []
        prev = 0.0
        for s in stops_el:
            if _strip(s.tag) != "stop":
                continue
            off = s.attrib.get("offset", "0")
            o = float(off[:-1]) / 100.0 if off.strip().endswith("%") else float(off)
            o = min(max(o, 0.0), 1.0)
            col = parse_color(s.attrib.get("stop-color", "black"), (0, 0, 0)) or (0, 0, 0, 1.0)
            so = opacity_val(s.attrib.get("stop-opacity"))
            stops.append([o, col, _f(f32(so) * f32(col[3]))])
        if not stops:
            return "none"
        # usvg: offsets must be monotonic; equal offsets are nudged by f32 epsilon
        out = []
        for i, (o, col, so) in enumerate(stops):
            o = f32(o)
            if i > 0:
                p = f32(out[-1][0])
                if o < p:
                    o = p
                if o == p and i > 0 and len(out) >= 1:
                    o = p + np.finfo(np.float32).eps if float(p) + float(np.finfo(np.float32).eps) <= 1.0 else p
                    if o == p and len(out) >= 1:
                        out[-1][0] = float(p - np.finfo(np.float32).eps)
            out.append([float(o), int(col[0]), int(col[1]), int(col[2]), so])  # usvg::Stop: offset, color, opacity
        if len(out) == 1:
            return {"kind": "color", "rgb": out[0][1:4], "sub_opacity": out[0][4]}
        units = ga("gradientUnits") or "objectBoundingBox"
        spread = ga("spreadMethod") or "pad"
        if spread not in ("pad", "reflect", "repeat"):
            spread = "pad"
        gts = self.resolve_transform(el, ga("gradientTransform"))
        obb = units == "objectBoundingBox"

        def coord(name, default, ref):
            v = ga(name, same_kind_only=True)
            if v is None:
                v = default
            if obb:
                v = v.strip()
                return _f(float(v[:-1]) / 100.0 if v.endswith("%") else float(v))
            return _f(parse_length(v, 0.0, ref=ref))

        W, H = self.vb_w, self.vb_h
        diag = math.sqrt((W * W + H * H) / 2.0)
        if obb:
            if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
                return None
            gts = ts_post(gts, ts_from_bbox(bbox)) if False else ts_concat(ts_from_bbox(bbox), gts)
        if tag == "linearGradient":
            return {"kind": "linear", "x0": coord("x1", "0%", W), "y0": coord("y1", "0%", H), "x1": coord("x2", "100%", W),
                    "y1": coord("y2", "0%", H), "stops": out, "spread": spread, "ts": list(gts)}
        cx, cy, r = coord("cx", "50%", W), coord("cy", "50%", H), coord("r", "50%", diag)
        fxv, fyv = ga("fx", True), ga("fy", True)
        fx = coord("fx", "50%", W) if fxv is not None else cx
        fy = coord("fy", "50%", H) if fyv is not None else cy
        fr = coord("fr", "0%", diag)
        if not r > 0:
            # 'A value of zero will cause the area to be painted as a single color using the last stop'
            return {"kind": "color", "rgb": out[-1][1:4], "sub_opacity": out[-1][4]}
        return {"kind": "radial", "x0": fx, "y0": fy, "r0": fr, "x1": cx, "y1": cy, "r1": r, "stops": out, "spread": spread,
                "ts": list(gts)}

    # ---- patterns (usvg paint_server.rs convert_pattern + to_user_coordinates) ----
    def pattern_for(self, el, bbox):
        d = self.doc
        chain, seen, cur = [], set(), el
        while cur is not None and id(cur) not in seen and _strip(cur.tag) == "pattern":
            seen.add(id(cur))
            chain.append(cur)
            cur = d.link(cur.attrib.get("href"))

        def ga(name):
            for c in chain:
                if name in c.attrib:
                    return c.attrib[name]
            return None

        with_children = next((c for c in chain if len(list(c)) > 0), None)
        if with_children is None:
            return None
        units = ga("patternUnits") or "objectBoundingBox"
        cunits = ga("patternContentUnits") or "userSpaceOnUse"
        pts = self.resolve_transform(el, ga("patternTransform"))
        W, H = self.vb_w, self.vb_h

        def num(name):
            v = (ga(name) or "0").strip()
            if units == "objectBoundingBox":
                return _f(float(v[:-1]) / 100.0 if v.endswith("%") else float(v))
            return _f(parse_length(v, 0.0, ref=W if name in ("x", "width") else H))

        rect = (num("x"), num("y"), num("width"), num("height"))
        if not (rect[2] > 0 and rect[3] > 0):
            return None
        vb = ga("viewBox")
        vbr = None
        if vb:
            n = [float(x) for x in re.findall(NUM, vb)]
            if len(n) == 4 and n[2] > 0 and n[3] > 0:
                vbr = n
        self.pattern_depth += 1  # nesting of patterns only (groups have their own guard)
        try:
            if self.pattern_depth > 16:
                raise Unsupported("pattern recursion")
            kids = [n for n in (self.node_for(ch) for ch in with_children) if n is not None]
        finally:
            self.pattern_depth -= 1
        if not kids:
            return None
        if units == "objectBoundingBox":
            if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
                return None
            rect = (_f(f32(rect[0]) * f32(bbox[2]) + f32(bbox[0])), _f(f32(rect[1]) * f32(bbox[3]) + f32(bbox[1])),
                    _f(f32(rect[2]) * f32(bbox[2])), _f(f32(rect[3]) * f32(bbox[3])))
        root = {"t": "g", "ts": list(IDENT), "children": kids, "opacity": 1.0, "isolate": False}
        if cunits == "objectBoundingBox" and vbr is None:
            if bbox is None:
                return None
            root = {"t": "g", "ts": list(IDENT), "opacity": 1.0, "isolate": False,
                    "children": [dict(root, ts=list(ts_scale(bbox[2], bbox[3])))]}
        if vbr is not None:
            vts = viewbox_transform(vbr, ga("preserveAspectRatio"), rect[2], rect[3])
            root = {"t": "g", "ts": list(IDENT), "opacity": 1.0, "isolate": False, "children": [dict(root, ts=list(vts))]}
        return {"kind": "pattern_tree", "rect": list(rect), "ts": list(pts), "root": root}

    # ---- bbox of a converted node (object bounding box, no stroke) ----
    def node_bbox(self, n):
        """usvg Group::calculate_object_bbox (tree/mod.rs:1850-1864): the union of the children's object bounding boxes
        in the group's OWN coordinate system — child groups contribute their bbox mapped by their transform, the
        group's own transform is not applied."""
        if n["t"] in ("path", "image"):
            return tuple(n["bbox"]) if n.get("bbox") else None
        boxes = []
        for c in n["children"]:
            b = self.node_bbox(c)
            if b is None:
                continue
            if c["t"] == "g" and not ts_is_identity(tuple(c["ts"])):
                b = rect_transform(b, tuple(c["ts"]))
            boxes.append(b)
        if not boxes:
            return None
        l, tp = min(b[0] for b in boxes), min(b[1] for b in boxes)
        r, bt = max(_f(f32(b[0]) + f32(b[2])) for b in boxes), max(_f(f32(b[1]) + f32(b[3])) for b in boxes)
        return (l, tp, _f(f32(r) - f32(l)), _f(f32(bt) - f32(tp)))

    def resolve_transform(self, el, value):
        """usvg converter.rs:1100-1127 resolve_transform: the transform attribute combined with `transform-origin`
        (lengths in user space, percentages of the viewBox): translate(o) * transform * translate(-o)."""
        ts = parse_transform(value)
        origin = el.attrib.get("transform-origin") if el is not None else None
        if not origin:
            return ts
        toks = origin.replace(",", " ").split()
        kw_x = {"left": "0%", "center": "50%", "right": "100%"}
        kw_y = {"top": "0%", "center": "50%", "bottom": "100%"}
        x, y = "50%", "50%"
        if len(toks) == 1:
            t = toks[0]
            if t in ("top", "bottom"):
                y = kw_y[t]
            else:
                x = kw_x.get(t, t)
        elif len(toks) >= 2:
            a, b = toks[0], toks[1]
            if a in ("top", "bottom") or b in ("left", "right"):
                a, b = b, a
            x, y = kw_x.get(a, a), kw_y.get(b, b)
        try:
            dx = _f(parse_length(x, 0.0, ref=self.vb_w))
            dy = _f(parse_length(y, 0.0, ref=self.vb_h))
        except Unsupported:
            return ts
        return ts_pre(ts_pre(ts_translate(dx, dy), ts), ts_translate(-dx, -dy))

    # ---- clipPath / mask ----
    def clip_for(self, el, bbox):
        d = self.doc
        units = el.attrib.get("clipPathUnits", "userSpaceOnUse")
        ts = self.resolve_transform(el, el.attrib.get("transform"))
        if units == "objectBoundingBox":
            if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
                return None
            ts = ts_pre(ts, ts_from_bbox(bbox))
        nested = None
        cp = d.attr(el, "clip-path", inherit=False)
        if cp and cp != "none":
            t = d.link(cp)
            if t is None or _strip(t.tag) != "clipPath":
                return None
            nested = self.clip_for(t, bbox)
            if nested is None:
                return None
        kids = []
        for ch in el:
            tag = _strip(ch.tag)
            if tag in ("path", "rect", "circle", "ellipse", "line", "polyline", "polygon", "use"):
                if tag == "use":
                    tgt = d.link(ch.attrib.get("href"))
                    if tgt is None or _strip(tgt.tag) not in ("path", "rect", "circle", "ellipse", "line", "polyline", "polygon"):
                        continue
                n = self.node_for(ch)
                if n is None:
                    continue
                self._clip_fixup(n, ch)
                kids.append(n)
            elif tag in ("text",):
                raise Unsupported("text in clipPath")
        if not kids:
            # 'A clip path without children is invalid': the element is not rendered — modelled as an empty clip
            return {"ts": list(ts), "children": [], "clip": nested}
        return {"ts": list(ts), "children": kids, "clip": nested}

    def _clip_fixup(self, n, el):
        """Inside a clipPath only geometry and clip-rule matter: fill = opaque black with rule = clip-rule."""
        if n["t"] == "path":
            rule = self.doc.attr(el, "clip-rule") or "nonzero"
            n["fill"] = {"paint": {"kind": "color", "rgb": [0, 0, 0]}, "opacity": 1.0,
                         "rule": "evenodd" if rule == "evenodd" else "nonzero"}
            n["stroke"] = None
        else:
            n["opacity"] = 1.0
            n["mask"] = None
            n["filters"] = []
            n["blend"] = "normal"
            n["isolate_attr"] = False
            n["isolate"] = bool(n.get("clip"))
            for c in n["children"]:
                self._clip_fixup(c, el if c["t"] == "path" and len(n["children"]) == 1 else el)

    def mask_for(self, el, bbox):
        d = self.doc
        units = el.attrib.get("maskUnits", "objectBoundingBox")
        cunits = el.attrib.get("maskContentUnits", "userSpaceOnUse")
        W, H = self.vb_w, self.vb_h

        def ln(name, default):
            v = el.attrib.get(name, default).strip()
            if units == "objectBoundingBox":
                return _f(float(v[:-1]) / 100.0 if v.endswith("%") else float(v))
            return _f(parse_length(v, 0.0, ref=W if name in ("x", "width") else H))

        rect = (ln("x", "-10%"), ln("y", "-10%"), ln("width", "120%"), ln("height", "120%"))
        if not (rect[2] > 0 and rect[3] > 0):
            return None
        if units == "objectBoundingBox":
            if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
                return {"rect": list(rect), "kind": "luminance", "mask": None, "root": {"t": "g", "ts": list(IDENT), "children": [], "opacity": 1.0, "isolate": False}}
            rect = (_f(f32(rect[0]) * f32(bbox[2]) + f32(bbox[0])), _f(f32(rect[1]) * f32(bbox[3]) + f32(bbox[1])),
                    _f(f32(rect[2]) * f32(bbox[2])), _f(f32(rect[3]) * f32(bbox[3])))
        nested = None
        mk = d.attr(el, "mask", inherit=False)
        if mk and mk != "none":
            t = d.link(mk)
            if t is None or _strip(t.tag) != "mask":
                return None
            nested = self.mask_for(t, bbox)
            if nested is None:
                return None
        kind = "alpha" if (el.attrib.get("mask-type") == "alpha") else "luminance"
        kids = [n for n in (self.node_for(ch) for ch in el) if n is not None]
        if not kids:
            return None
        root = {"t": "g", "ts": list(IDENT), "children": kids, "opacity": 1.0, "isolate": False}
        if cunits == "objectBoundingBox":
            if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
                return None
            root = {"t": "g", "ts": list(IDENT), "opacity": 1.0, "isolate": False,
                    "children": [{"t": "g", "ts": list(ts_from_bbox(bbox)), "children": kids, "opacity": 1.0, "isolate": False}]}
        return {"rect": list(rect), "kind": kind, "mask": nested, "root": root}


def parse(svg_text, base_dir=None, nested=False):
    """`base_dir`: where relative image hrefs are resolved (usvg Options::resources_dir)."""
    conv = Converter(Doc(svg_text, base_dir))
    conv.nested = nested
    return finalize_scene(conv.convert())


# ---------------------------------------------------------------------------------------------------------------------
# bounding boxes (usvg tree/mod.rs: Path::new :1311-1354, Group::calculate_bounding_boxes :1866-1929)
# ---------------------------------------------------------------------------------------------------------------------
def _union(boxes):
    """usvg BBox::expand over Rects given as (x, y, w, h); -> (l, t, r, b) or None."""
    acc = None
    for b in boxes:
        l, t, r, bt = f32(b[0]), f32(b[1]), f32(b[0]) + f32(b[2]), f32(b[1]) + f32(b[3])
        acc = (l, t, r, bt) if acc is None else (min(acc[0], l), min(acc[1], t), max(acc[2], r), max(acc[3], bt))
    return acc


def _non_zero(acc):
    if acc is None:
        return None
    w, h = acc[2] - acc[0], acc[3] - acc[1]
    if not (w > 0 and h > 0 and math.isfinite(float(w)) and math.isfinite(float(h))):
        return None
    return (float(acc[0]), float(acc[1]), float(w), float(h))


def _rect_transform_nz(r, t):
    b = rect_transform(r, t) if not ts_is_identity(t) else tuple(r)
    if not (b[2] > 0 and b[3] > 0 and all(math.isfinite(v) for v in b)):
        return None
    return b


def _stroke_bbox(verbs, pts, st):
    """Path::calculate_stroke_bbox: the tight bounds of the outline at res_scale 1 (dashes ignored)."""
    from tests import geom
    out = geom.stroke_outline(verbs, pts, st["width"], st["miter"], st["cap"], st["join"], 1.0)
    if out is None:
        return None
    v, q = out
    return tight_bounds([int(x) for x in v], [(float(a), float(b)) for a, b in q])


DEFAULT_LAYER_BBOX = (0.0, 0.0, 1.0, 1.0)


def finalize_scene(scene):
    """Fills in what usvg stores on every node once the tree is built: stroke / layer / absolute layer bounding boxes."""
    _finalize_group(scene["root"], IDENT)
    return scene


def _finalize_paint_roots(n):
    for key in ("fill", "stroke"):
        pd = n.get(key)
        if pd and pd["paint"]["kind"] == "pattern_tree":
            _finalize_group(pd["paint"]["root"], IDENT)


def _finalize_group(g, parent_abs):
    abs_ts = ts_pre(parent_abs, tuple(g["ts"]))
    layer = []
    for c in g["children"]:
        if c["t"] == "path":
            bbox = c.get("bbox")
            sb = (_stroke_bbox(c["verbs"], c["pts"], c["stroke"]) if c.get("stroke") else None) or bbox
            c["stroke_bbox"] = list(sb) if sb else None
            if sb is not None:
                if ts_has_skew(abs_ts):
                    tp = [ts_map(abs_ts, x, y) for x, y in c["pts"]]
                    ab = tight_bounds(c["verbs"], tp)
                    asb = (_stroke_bbox(c["verbs"], tp, c["stroke"]) if c.get("stroke") else None) or ab
                else:
                    asb = rect_transform(sb, abs_ts) if not ts_is_identity(abs_ts) else tuple(sb)
                c["abs_bbox"] = list(asb) if asb else None
                layer.append(sb)
            _finalize_paint_roots(c)
        elif c["t"] == "image":
            c["abs_bbox"] = list(rect_transform(tuple(c["bbox"]), abs_ts)) if not ts_is_identity(abs_ts) else list(c["bbox"])
            layer.append(tuple(c["bbox"]))
            if c["kind"] == "svg":
                finalize_scene(c["tree"])
        else:
            _finalize_group(c, abs_ts)
            r = _rect_transform_nz(tuple(c["layer_bbox"]), tuple(c["ts"]))
            if r is not None:
                layer.append(r)
    if g.get("clip"):
        _finalize_group(g["clip"], IDENT)  # a clip dict is group-like (ts + children); its own "clip" is reached recursively
    m = g.get("mask")
    while m:
        _finalize_group(m["root"], IDENT)
        m = m.get("mask")
    for f in g.get("filters") or []:
        for prim in f["primitives"]:
            if prim["kind"] == "image":
                _finalize_group(prim["root"], IDENT)
    lb = None
    if g.get("filters"):
        lb = _non_zero(_union([tuple(f["rect"]) for f in g["filters"]]))  # the filter region has priority
    if lb is None and not g.get("filters"):
        lb = _non_zero(_union(layer))
    if lb is None:
        g["layer_bbox"] = list(DEFAULT_LAYER_BBOX)
        g["abs_layer_bbox"] = list(DEFAULT_LAYER_BBOX)
    else:
        g["layer_bbox"] = list(lb)
        ab = _rect_transform_nz(lb, abs_ts)
        g["abs_layer_bbox"] = list(ab) if ab else list(DEFAULT_LAYER_BBOX)


# ---------------------------------------------------------------------------------------------------------------------
# render traversal (crates/resvg/src/lib.rs, render.rs, path.rs, clip.rs, mask.rs, image.rs) — the CHECKER's copy: it
# drives the CPU oracle.  The product's traversal is resvg_b200/csrc/render.cpp behind rb_render.
# ---------------------------------------------------------------------------------------------------------------------
BLEND_MAP = {"normal": "source_over", "multiply": "multiply", "screen": "screen", "overlay": "overlay", "darken": "darken",
             "lighten": "lighten", "color-dodge": "color_dodge", "color-burn": "color_burn", "hard-light": "hard_light",
             "soft-light": "soft_light", "difference": "difference", "exclusion": "exclusion", "hue": "hue",
             "saturation": "saturation", "color": "color", "luminosity": "luminosity"}


def max_filter_bbox(w, h):
    """lib.rs:86-97"""
    return (-2 * w, -2 * h, 5 * w, 5 * h)


def _trunc_i32(v):
    v = float(v)
    if v != v:
        return 0
    return int(max(min(math.trunc(v), 2147483647), -2147483648))


class Renderer:
    def __init__(self, backend):
        self.be = backend

    # lib.rs:34-43
    def render(self, scene, width, height, ts, layer=None):
        if layer is None:
            layer = self.be.new_layer(width, height)
        lw, lh = self.be.size(layer)
        self.max_bbox = max_filter_bbox(lw, lh)
        self.render_group(scene["root"], tuple(ts), layer)  # scene root = the child of usvg's root that carries the viewBox transform
        return layer

    # lib.rs:55-70
    def render_node_by_id(self, scene, node_id, ts, layer):
        n = find_node(scene["root"], node_id)
        if n is None:
            return False
        bbox = n.get("abs_layer_bbox") if n["t"] == "g" else n.get("abs_bbox")
        if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
            return False
        lw, lh = self.be.size(layer)
        self.max_bbox = max_filter_bbox(lw, lh)
        self.render_node(n, ts_pre(tuple(ts), ts_translate(-bbox[0], -bbox[1])), layer)
        return True

    def render_nodes(self, group, ts, layer, blend="source_over"):
        for n in group["children"]:
            self.render_node(n, ts, layer, blend)

    def render_node(self, n, ts, layer, blend="source_over"):
        if n["t"] == "path":
            self.render_path(n, ts, layer, blend)
        elif n["t"] == "image":
            self.render_image(n, ts, layer)
        else:
            self.render_group(n, ts, layer)

    # path.rs:6-24
    def render_path(self, n, ts, layer, blend):
        if not n.get("visible", True):
            return
        if n.get("stroke_first"):
            self.stroke(n, ts, layer, blend)
            self.fill_path(n, ts, layer, blend)
        else:
            self.fill_path(n, ts, layer, blend)
            self.stroke(n, ts, layer, blend)

    # path.rs:26-75
    def fill_path(self, n, ts, layer, blend):
        f = n.get("fill")
        if not f:
            return
        xs = [p[0] for p in n["pts"]]
        ys = [p[1] for p in n["pts"]]
        if max(xs) - min(xs) == 0.0 or max(ys) - min(ys) == 0.0:  # path.rs:36
            return
        paint = self.convert_paint(f["paint"], f.get("opacity", 1.0), ts)
        if paint is None:
            return
        self.be.fill_path(layer, n["verbs"], n["pts"], paint, f["rule"], ts, blend, n.get("aa", True))

    def convert_paint(self, paint, opacity, ts):
        """path.rs:45-71 / 118-177 (colour, gradients) and 179-205 (a pattern is pre-rendered into a tile pixmap)."""
        k = paint["kind"]
        if k == "color":
            r, g, b = paint["rgb"]
            return {"kind": "solid", "color": [_f(f32(r) / f32(255)), _f(f32(g) / f32(255)), _f(f32(b) / f32(255)),
                                               _f(f32(to_u8_opacity(opacity)) / f32(255))]}
        if k in ("linear", "radial"):
            stops = []
            for o, r, g, b, so in paint["stops"]:
                a = min(max(f32(so) * f32(opacity), f32(0)), f32(1))
                stops.append([o, _f(f32(r) / f32(255)), _f(f32(g) / f32(255)), _f(f32(b) / f32(255)), _f(f32(to_u8_opacity(a)) / f32(255))])
            return dict(paint, stops=stops)
        sx, sy = ts_get_scale(ts_pre(ts, tuple(paint["ts"])))
        rect = paint["rect"]
        iw = _trunc_i32(math.floor(abs(float(f32(rect[2]) * f32(sx))) + 0.5))  # f32::round of a non-negative value
        ih = _trunc_i32(math.floor(abs(float(f32(rect[3]) * f32(sy))) + 0.5))
        if iw <= 0 or ih <= 0:
            return None
        tile = self.be.new_layer(iw, ih)
        self.render_nodes(paint["root"], ts_scale(sx, sy), tile)
        pts = ts_pre(IDENT, tuple(paint["ts"]))
        pts = ts_pre(pts, ts_translate(rect[0], rect[1]))
        pts = ts_pre(pts, ts_scale(_f(f32(1.0) / f32(sx)), _f(f32(1.0) / f32(sy))))
        return {"kind": "pattern", "layer": tile, "spread": "repeat", "quality": "bicubic", "opacity": opacity, "ts": pts}

    # path.rs:77-116 + tiny-skia painter.rs stroke_path
    def stroke(self, n, ts, layer, blend):
        from tests import geom
        s = n.get("stroke")
        if not s:
            return
        paint = self.convert_paint(s["paint"], s.get("opacity", 1.0), ts)
        if paint is None:
            return
        # painter.rs stroke_path: res_scale = compute_resolution_scale(ts); thin strokes take the hairline path
        sx = float(np.sqrt(f32(ts[0]) * f32(ts[0]) + f32(ts[2]) * f32(ts[2])))  # Point::length in f32
        sy = float(np.sqrt(f32(ts[1]) * f32(ts[1]) + f32(ts[3]) * f32(ts[3])))
        res_scale = max(sx, sy) if (math.isfinite(sx) and math.isfinite(sy) and max(sx, sy) > 0) else 1.0
        src_verbs, src_pts = n["verbs"], n["pts"]
        if s.get("dash"):
            dashed = geom.dash_path(src_verbs, src_pts, s["dash"], s.get("dash_offset", 0.0), res_scale)
            if dashed is None:
                return  # StrokeDash::new accepted the list (the front end filtered the others) but nothing is left
            src_verbs, src_pts = dashed
        if n.get("aa", True):
            # painter.rs treat_as_hairline, in f32: map (w, 0) and (0, w) by the transform without its translation
            (p0x, p0y), (p1x, p1y) = ts_map((ts[0], ts[1], ts[2], ts[3], 0.0, 0.0), s["width"], 0.0), \
                ts_map((ts[0], ts[1], ts[2], ts[3], 0.0, 0.0), 0.0, s["width"])

            def fast_len(x, y):
                x, y = abs(f32(x)), abs(f32(y))
                if x < y:
                    x, y = y, x
                return f32(x + y * f32(0.5))

            if fast_len(p0x, p0y) <= 1.0 and fast_len(p1x, p1y) <= 1.0:
                self.be.stroke_hairline(layer, src_verbs, src_pts, paint, ts, blend, s["width"], s["cap"])
                return
        out = geom.stroke_outline(src_verbs, src_pts, s["width"], s["miter"], s["cap"], s["join"], res_scale)
        if out is None:
            return
        verbs, pts = out
        self.be.fill_path(layer, verbs, pts, paint, "nonzero", ts, blend, n.get("aa", True))

    # render.rs:49-143
    def render_group(self, g, ts, layer):
        ts = ts_pre(ts, tuple(g["ts"]))
        if not g.get("isolate"):
            self.render_nodes(g, ts, layer)
            return
        bbox = _rect_transform_nz(tuple(g["layer_bbox"]), ts)
        if bbox is None:
            return
        bx, by, bw, bh = [f32(v) for v in bbox]
        if not g.get("filters"):
            ib = (_trunc_i32(math.floor(float(bx))) - 2, _trunc_i32(math.floor(float(by))) - 2,
                  _trunc_i32(math.ceil(float(bw))) + 4, _trunc_i32(math.ceil(float(bh))) + 4)
        else:
            ib = (_trunc_i32(math.floor(float(bx))), _trunc_i32(math.floor(float(by))),
                  _trunc_i32(max(math.ceil(float(bw)), 1.0)), _trunc_i32(max(math.ceil(float(bh)), 1.0)))
        ib = fit_to_rect(ib, self.max_bbox)
        if ib is None:
            return
        dx = bx - (bx - f32(ib[0]))
        dy = by - (by - f32(ib[1]))
        lts = ts_pre(ts_translate(-dx, -dy), ts)
        sub = self.be.new_layer(ib[2], ib[3])
        self.render_nodes(g, lts, sub)
        for f in g.get("filters", []):
            from tests.svgfilters import apply_filter
            apply_filter(self, f, lts, sub)
        if g.get("clip"):
            self.apply_clip(g["clip"], lts, sub)
        if g.get("mask"):
            self.apply_mask(g["mask"], lts, sub)
        self.be.draw_layer(layer, sub, ib[0], ib[1], g.get("opacity", 1.0), BLEND_MAP[g.get("blend", "normal")])

    # ---- clip.rs ----
    def apply_clip(self, clip, ts, layer):
        w, h = self.be.size(layer)
        cl = self.be.new_layer(w, h)
        self.be.fill_color(cl, 0.0, 0.0, 0.0, 1.0)
        self.clip_children(clip, "clear", ts_pre(ts, tuple(clip["ts"])), cl)
        if clip.get("clip"):
            self.apply_clip(clip["clip"], ts, layer)
        m = self.be.mask_from_layer(cl, "alpha")
        self.be.mask_invert(m)
        self.be.apply_mask(layer, m)

    def clip_children(self, group, mode, ts, layer):
        for n in group["children"]:
            if n["t"] == "path":
                if n.get("visible", True):
                    self.fill_path(n, ts, layer, mode)
            elif n["t"] == "g":
                t = ts_pre(ts, tuple(n["ts"]))
                if n.get("clip"):
                    w, h = self.be.size(layer)
                    tmp = self.be.new_layer(w, h)
                    self.clip_children(n, "source_over", t, tmp)
                    self.apply_clip(n["clip"], t, tmp)
                    self.be.draw_layer(layer, tmp, 0, 0, 1.0, "xor")
                else:
                    self.clip_children(n, mode, t, layer)

    # ---- mask.rs ----
    def apply_mask(self, mask, ts, layer):
        w, h = self.be.size(layer)
        if not mask["root"]["children"]:
            self.be.fill_color(layer, 0.0, 0.0, 0.0, 0.0)
            return
        ml = self.be.new_layer(w, h)
        am = self.be.mask_new(w, h)
        x, y, rw, rh = mask["rect"]
        x, y, rw, rh = f32(x), f32(y), f32(rw), f32(rh)
        rect_pts = [(float(x), float(y)), (float(x + rw), float(y)), (float(x + rw), float(y + rh)), (float(x), float(y + rh))]
        self.be.mask_fill_path(am, [M, L, L, L, Z], rect_pts, "nonzero", True, ts)
        self.render_nodes(mask["root"], ts, ml)
        self.be.apply_mask(ml, am)
        if mask.get("mask"):
            self.apply_mask(mask["mask"], ts, layer)
        m = self.be.mask_from_layer(ml, mask["kind"])
        self.be.apply_mask(layer, m)

    # ---- image.rs ----
    def render_image(self, n, ts, layer):
        if not n.get("visible", True):
            return
        w, h = self.be.size(layer)
        if n["kind"] == "svg":  # render_vector, image.rs:37-54
            sub = self.be.new_layer(w, h)
            saved = self.max_bbox
            Renderer(self.be).render(n["tree"], w, h, ts, sub)
            self.max_bbox = saved
            self.be.draw_layer(layer, sub, 0, 0, 1.0, "source_over")
            return
        # render_raster, image.rs:173-206: a Pad pattern of the decoded pixmap filled into its own rectangle
        px = image_pixels(n)
        raster = self.be.new_layer(n["w"], n["h"])
        self.be.upload(raster, px)
        spec = {"kind": "pattern", "layer": raster, "spread": "pad", "quality": n.get("quality", "bicubic"), "opacity": 1.0,
                "ts": IDENT}
        self.be.fill_rect(layer, 0.0, 0.0, float(n["w"]), float(n["h"]), spec, ts)


def image_pixels(n):
    """The decoded pixmap of a raster image node: premultiplied RGBA8 (h, w, 4)."""
    if "pixels_z" in n:
        import base64
        import zlib
        return np.frombuffer(zlib.decompress(base64.b64decode(n["pixels_z"])), np.uint8).reshape(n["h"], n["w"], 4)
    return np.asarray(n["pixels"], np.uint8).reshape(n["h"], n["w"], 4)


def find_node(group, node_id):
    """usvg Tree::node_by_id: depth first."""
    if not node_id:
        return None
    for n in group["children"]:
        if n.get("id") == node_id:
            return n
        if n["t"] == "g":
            r = find_node(n, node_id)
            if r is not None:
                return r
    return None


def target_for(scene, target_width=300):
    """tests/integration/main.rs:75-87 (TestMode::Normal): scale to width 300 -> (pixmap w, h, render transform)."""
    w, h = scene["width"], scene["height"]
    iw, ih = max(1, int(round(w))), max(1, int(round(h)))  # Size::to_int_size
    pw = target_width
    ph = int(math.ceil(float(f32(pw) * f32(ih) / f32(iw))))
    return pw, ph, ts_scale(_f(f32(pw) / f32(w)), _f(f32(ph) / f32(h)))


def render_scene(scene, backend, target_width=300):
    """The checker's render: test-side traversal + the given back end (the CPU oracle); premultiplied RGBA8."""
    pw, ph, ts = target_for(scene, target_width)
    layer = Renderer(backend).render(scene, pw, ph, ts)
    return backend.to_numpy(layer)


def render_scene_gpu(scene, ctx, target_width=300, via="tree"):
    """The product's render: the scene is serialised (RBT1) and drawn by ONE rb_render / rb_submit call — traversal,
    filter graph and pixels all inside libresvg_b200.so."""
    import resvg_b200 as rb
    pw, ph, ts = target_for(scene, target_width)
    layer = ctx.layer(pw, ph)
    if via == "submit":
        rb.tree.submit(rb.tree.serialize(scene), ts, layer)
    else:
        tree = rb.tree.Tree(scene)
        rb.tree.render(tree, ts, layer)
        tree.close()
    return layer.download()


def demultiply_f64(px):
    """tests/integration/main.rs:204-211"""
    a = px[..., 3:4].astype(np.float64) / 255.0
    with np.errstate(divide="ignore", invalid="ignore"):
        c = px[..., :3].astype(np.float64) / a + 0.5
    c = np.nan_to_num(c, nan=0.0, posinf=255.0)
    out = px.copy()
    out[..., :3] = np.clip(c, 0, 255).astype(np.uint8)
    return out


def diff_pixels(actual_premul, golden_rgba, threshold=1):
    """get_diff (tests/integration/main.rs:151-226): number of pixels differing by more than `threshold` in any
    channel; pixels transparent in both images are ignored."""
    if actual_premul.shape != golden_rgba.shape:
        return max(actual_premul.shape[0], golden_rgba.shape[0]) * max(actual_premul.shape[1], golden_rgba.shape[1])
    a = demultiply_f64(actual_premul).astype(np.int16)
    g = golden_rgba.astype(np.int16)
    both0 = (a[..., 3] == 0) & (g[..., 3] == 0)
    d = np.abs(a - g).max(axis=-1) > threshold
    return int((d & ~both0).sum())
