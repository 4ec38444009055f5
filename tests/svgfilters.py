"""Filter support for the test-side front end: usvg's filter conversion (crates/usvg/src/parser/filter.rs) and a
restatement of resvg's filter executor (crates/resvg/src/filter/mod.rs:335-1147) driving a back end."""
import math
import re

import numpy as np

from tests.svgfront import (IDENT, NUM, Unsupported, _f, _strip, f32, fit_to_rect, opacity_val, parse_color,
                            parse_length, rect_transform, to_int_rect, to_u8_opacity, ts_get_scale, ts_map, M, L, Z)


def _nums(s):
    return [float(x) for x in re.findall(NUM, s or "")]


def _approx_zero(v):
    v = f32(v)
    if v == 0:
        return True
    if v < 0:
        return False
    return int(np.array(v, dtype=np.float32).view(np.uint32)) <= 4


# ---------------------------------------------------------------------------------------------------------------------
# conversion
# ---------------------------------------------------------------------------------------------------------------------
def convert_filter(conv, el, bbox):
    d = conv.doc
    units = el.attrib.get("filterUnits", "objectBoundingBox")
    punits = el.attrib.get("primitiveUnits", "userSpaceOnUse")
    W, H = conv.vb_w, conv.vb_h

    def num(node, name, default, un):
        v = node.attrib.get(name, default)
        if v is None:
            return None
        v = v.strip()
        if un == "objectBoundingBox":
            return _f(float(v[:-1]) / 100.0 if v.endswith("%") else float(v))
        return _f(parse_length(v, 0.0, ref=W if name in ("x", "width") else H))

    rect = (num(el, "x", "-10%", units), num(el, "y", "-10%", units), num(el, "width", "120%", units),
            num(el, "height", "120%", units))
    if not (rect[2] > 0 and rect[3] > 0):
        return None
    if units == "objectBoundingBox":
        if bbox is None or not (bbox[2] > 0 and bbox[3] > 0):
            return None
        rect = _bbox_transform(rect, bbox)
    # find_filter_with_primitives via href chain
    node = el
    seen = set()
    while node is not None and id(node) not in seen:
        seen.add(id(node))
        if _strip(node.tag) != "filter":
            return None
        if len(list(node)) > 0:
            break
        node = d.link(node.attrib.get("href"))
    if node is None or len(list(node)) == 0:
        return None
    if punits == "objectBoundingBox":
        if bbox is None:
            return None
        scale = (bbox[2], bbox[3])
    else:
        scale = (1.0, 1.0)
    prims = []
    names, idx = set(), [1]

    def gen_result(ch):
        r = ch.attrib.get("result")
        if r is not None:
            names.add(r)
            idx[0] += 1
            return r
        while True:
            n = f"result{idx[0]}"
            idx[0] += 1
            if n not in names:
                return n

    def resolve_input(ch, name):
        s = ch.attrib.get(name)
        if s is not None:
            if s == "SourceGraphic":
                return ("source",)
            if s == "SourceAlpha":
                return ("alpha",)
            if s in ("BackgroundImage", "BackgroundAlpha", "FillPaint", "StrokePaint"):
                return ("source",)
            if not any(p["result"] == s for p in prims):
                return ("ref", prims[-1]["result"]) if prims else ("source",)
            return ("ref", s)
        return ("ref", prims[-1]["result"]) if prims else ("source",)

    for ch in node:
        tag = _strip(ch.tag)
        if not tag.startswith("fe"):
            continue
        # resolve_primitive_region
        x = num(ch, "x", None, punits)
        y = num(ch, "y", None, punits)
        w = num(ch, "width", None, punits)
        h = num(ch, "height", None, punits)
        region = rect
        if tag in ("feFlood", "feImage") and punits == "objectBoundingBox":
            if bbox is None:
                break
            r = (x or 0.0, y or 0.0, 1.0 if w is None else w, 1.0 if h is None else h)
            if not (r[2] > 0 and r[3] > 0):
                break
            sub = _bbox_transform(r, bbox)
        elif punits == "objectBoundingBox":
            r = (x or 0.0, y or 0.0, 1.0 if w is None else w, 1.0 if h is None else h)
            if not (r[2] > 0 and r[3] > 0):
                break
            sub = _bbox_transform(region, r)
        else:
            sub = (region[0] if x is None else x, region[1] if y is None else y, region[2] if w is None else w,
                   region[3] if h is None else h)
            if not (sub[2] > 0 and sub[3] > 0):
                break
        cs = d.attr(ch, "color-interpolation-filters", inherit=True) or "linearRGB"
        cs = "sRGB" if cs == "sRGB" else "linearRGB"
        conv._fe_subregion = sub
        k = _convert_kind(conv, ch, tag, scale, resolve_input)
        if k is None:
            continue
        k.update(rect=list(sub), cs=cs, result=gen_result(ch))
        prims.append(k)
    if not prims:
        return None
    return {"rect": list(rect), "primitives": prims}


def _bbox_transform(r, b):
    """NonZeroRect::bbox_transform"""
    return (_f(f32(r[0]) * f32(b[2]) + f32(b[0])), _f(f32(r[1]) * f32(b[3]) + f32(b[1])), _f(f32(r[2]) * f32(b[2])),
            _f(f32(r[3]) * f32(b[3])))


def _flood_color(conv, ch):
    c = parse_color(conv.doc.attr(ch, "flood-color", inherit=False) or "black",
                    (parse_color(conv.doc.attr(ch, "color") or "black") or (0, 0, 0, 1.0))[:3]) or (0, 0, 0, 1.0)
    fo = opacity_val(conv.doc.attr(ch, "flood-opacity", inherit=False))
    return [c[0], c[1], c[2]], _f(f32(c[3]) * f32(fo))


def _std_dev(ch, scale, default):
    n = _nums(ch.attrib.get("stdDeviation", default))
    if len(n) == 2:
        sx, sy = n
    elif len(n) == 1:
        sx = sy = n[0]
    else:
        sx = sy = 0.0
    sx, sy = _f(f32(sx) * f32(scale[0])), _f(f32(sy) * f32(scale[1]))
    return (sx if sx > 0 and math.isfinite(sx) else 0.0), (sy if sy > 0 and math.isfinite(sy) else 0.0)


def _convert_kind(conv, ch, tag, scale, resolve_input):
    a = ch.attrib
    fget = lambda n, dflt: _f(float(a[n])) if n in a and re.fullmatch(rf"\s*{NUM}\s*", a[n]) else dflt
    dummy = {"kind": "flood", "color": [0, 0, 0], "opacity": 0.0}
    if tag == "feGaussianBlur":
        sx, sy = _std_dev(ch, scale, "0 0")
        return {"kind": "blur", "in": resolve_input(ch, "in"), "sx": sx, "sy": sy}
    if tag == "feDropShadow":
        sx, sy = _std_dev(ch, scale, "2 2")
        col, op = _flood_color(conv, ch)
        return {"kind": "drop_shadow", "in": resolve_input(ch, "in"), "sx": sx, "sy": sy, "color": col, "opacity": op,
                "dx": _f(f32(fget("dx", 2.0)) * f32(scale[0])), "dy": _f(f32(fget("dy", 2.0)) * f32(scale[1]))}
    if tag == "feOffset":
        return {"kind": "offset", "in": resolve_input(ch, "in"), "dx": _f(f32(fget("dx", 0.0)) * f32(scale[0])),
                "dy": _f(f32(fget("dy", 0.0)) * f32(scale[1]))}
    if tag == "feFlood":
        col, op = _flood_color(conv, ch)
        return {"kind": "flood", "color": col, "opacity": op}
    if tag == "feBlend":
        mode = a.get("mode", "normal")
        from tests.svgfront import BLEND_MAP
        if mode not in BLEND_MAP:
            mode = "normal"
        return {"kind": "blend", "mode": mode, "in": resolve_input(ch, "in"), "in2": resolve_input(ch, "in2")}
    if tag == "feComposite":
        op = a.get("operator", "over")
        if op not in ("over", "in", "out", "atop", "xor", "arithmetic"):
            op = "over"
        return {"kind": "composite", "op": op, "k": [fget("k1", 0.0), fget("k2", 0.0), fget("k3", 0.0), fget("k4", 0.0)],
                "in": resolve_input(ch, "in"), "in2": resolve_input(ch, "in2")}
    if tag == "feMerge":
        return {"kind": "merge", "inputs": [resolve_input(c, "in") for c in ch]}
    if tag == "feTile":
        return {"kind": "tile", "in": resolve_input(ch, "in")}
    if tag == "feImage":
        return _convert_fe_image(conv, ch)
    if tag == "feComponentTransfer":
        funcs = {k: {"kind": "identity"} for k in "rgba"}
        for c in ch:
            t = _strip(c.tag)
            if t not in ("feFuncR", "feFuncG", "feFuncB", "feFuncA"):
                continue
            ty = c.attrib.get("type")
            cg = lambda n, dv: _f(float(c.attrib[n])) if n in c.attrib else dv
            if ty == "identity":
                fn = {"kind": "identity"}
            elif ty in ("table", "discrete"):
                fn = {"kind": ty, "values": [_f(v) for v in _nums(c.attrib.get("tableValues", ""))]}
            elif ty == "linear":
                fn = {"kind": "linear", "slope": cg("slope", 1.0), "intercept": cg("intercept", 0.0)}
            elif ty == "gamma":
                fn = {"kind": "gamma", "amplitude": cg("amplitude", 1.0), "exponent": cg("exponent", 1.0), "offset": cg("offset", 0.0)}
            else:
                continue
            funcs[t[-1].lower()] = fn
        return {"kind": "component_transfer", "in": resolve_input(ch, "in"), "funcs": [funcs[k] for k in "rgba"]}
    if tag == "feColorMatrix":
        ty = a.get("type")
        vals = [_f(v) for v in _nums(a.get("values", ""))] if "values" in a else None
        kind, params = "matrix", [1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0]
        if ty == "saturate":
            if vals is not None:
                kind, params = "saturate", [min(max(vals[0], 0.0), 1.0) if vals else 1.0]
        elif ty == "hueRotate":
            if vals is not None:
                kind, params = "hueRotate", [vals[0] if vals else 0.0]
        elif ty == "luminanceToAlpha":
            kind, params = "luminanceToAlpha", []
        elif vals is not None and len(vals) == 20:
            kind, params = "matrix", vals
        return {"kind": "color_matrix", "in": resolve_input(ch, "in"), "cm_kind": kind, "params": [float(p) for p in params]}
    if tag == "feConvolveMatrix":
        ox = oy = 3
        if "order" in a:
            n = _nums(a["order"])
            x = int(n[0]) if n else 3
            y = int(n[1]) if len(n) > 1 else x
            if x > 0 and y > 0:
                ox, oy = x, y
        matrix = []
        lst = [_f(v) for v in _nums(a.get("kernelMatrix", ""))]
        if len(lst) == ox * oy:
            matrix = lst
        ks = f32(0)
        for v in matrix:
            ks = ks + f32(v)
        ks = f32(round(float(ks * f32(1000000.0)))) / f32(1000000.0)
        if _approx_zero(ks):
            ks = f32(1.0)
        divisor = fget("divisor", _f(ks))
        if _approx_zero(divisor):
            return dict(dummy)
        bias = fget("bias", 0.0)

        def target(name, order):
            dflt = math.floor(order / 2.0)
            t = int(fget(name, float(dflt)))
            return None if (t < 0 or t >= order) else t

        tx, ty_ = target("targetX", ox), target("targetY", oy)
        if tx is None or ty_ is None or len(matrix) != ox * oy:
            return dict(dummy)
        em = a.get("edgeMode", "duplicate")
        em = em if em in ("none", "wrap") else "duplicate"
        return {"kind": "convolve", "in": resolve_input(ch, "in"), "matrix": matrix, "cols": ox, "rows": oy, "tx": tx,
                "ty": ty_, "divisor": divisor, "bias": bias, "edge": em, "preserve_alpha": a.get("preserveAlpha", "false") == "true"}
    if tag == "feMorphology":
        op = "dilate" if a.get("operator") == "dilate" else "erode"
        rx, ry = scale[0], scale[1]
        if "radius" in a:
            n = [_f(v) for v in _nums(a["radius"])]
            x = y = 0.0
            if len(n) == 2:
                x, y = n
            elif len(n) == 1:
                x = y = n[0]
            if _approx_zero(x) and _approx_zero(y):
                x = y = 1.0
            if _approx_zero(x) and not _approx_zero(y):
                x = 1.0
            if not _approx_zero(x) and _approx_zero(y):
                y = 1.0
            if math.copysign(1, x) > 0 and math.copysign(1, y) > 0:
                rx, ry = _f(f32(x) * f32(scale[0])), _f(f32(y) * f32(scale[1]))
        return {"kind": "morphology", "in": resolve_input(ch, "in"), "op": op, "rx": rx, "ry": ry}
    if tag == "feDisplacementMap":
        chn = lambda n: {"R": 0, "G": 1, "B": 2}.get(a.get(n, "A"), 3)
        sc = _f((f32(scale[0]) + f32(scale[1])) / f32(2.0))
        return {"kind": "displacement", "in": resolve_input(ch, "in"), "in2": resolve_input(ch, "in2"),
                "scale": _f(f32(fget("scale", 0.0)) * f32(sc)), "xch": chn("xChannelSelector"), "ych": chn("yChannelSelector")}
    if tag == "feTurbulence":
        bx = by = 0.0
        if "baseFrequency" in a:
            n = [_f(v) for v in _nums(a["baseFrequency"])]
            x = y = 0.0
            if len(n) == 2:
                x, y = n
            elif len(n) == 1:
                x = y = n[0]
            if math.copysign(1, x) > 0 and math.copysign(1, y) > 0:
                bx, by = x, y
        no = fget("numOctaves", 1.0)
        if math.copysign(1, no) < 0:
            no = 0.0
        return {"kind": "turbulence", "bfx": bx, "bfy": by, "octaves": int(math.floor(no + 0.5)), "seed": int(math.trunc(fget("seed", 0.0))),
                "stitch": a.get("stitchTiles") == "stitch", "fractal": a.get("type") == "fractalNoise"}
    if tag in ("feDiffuseLighting", "feSpecularLighting"):
        light = None
        for c in ch:
            t = _strip(c.tag)
            cg = lambda n, dv=0.0: _f(float(c.attrib[n])) if n in c.attrib else dv
            if t == "feDistantLight":
                light = {"kind": "distant", "azimuth": cg("azimuth"), "elevation": cg("elevation")}
            elif t == "fePointLight":
                light = {"kind": "point", "x": cg("x"), "y": cg("y"), "z": cg("z")}
            elif t == "feSpotLight":
                se = cg("specularExponent", 1.0)
                if not (se > 0 and math.isfinite(se)):
                    se = 1.0
                light = {"kind": "spot", "x": cg("x"), "y": cg("y"), "z": cg("z"),
                         "points_at": [cg("pointsAtX"), cg("pointsAtY"), cg("pointsAtZ")], "specular_exponent": se,
                         "limiting_cone_angle": cg("limitingConeAngle", None) if "limitingConeAngle" in c.attrib else None}
            if light:
                break
        if light is None:
            return dict(dummy)
        lc = a.get("lighting-color")
        if lc is None:
            color = [255, 255, 255]
        elif lc.strip() == "currentColor":
            color = list((parse_color(conv.doc.attr(ch, "color") or "black") or (0, 0, 0, 1))[:3])
        else:
            try:
                color = list((parse_color(lc) or (255, 255, 255, 1))[:3])
            except Unsupported:
                color = [255, 255, 255]
        if tag == "feDiffuseLighting":
            return {"kind": "diffuse", "in": resolve_input(ch, "in"), "surface_scale": fget("surfaceScale", 1.0),
                    "constant": fget("diffuseConstant", 1.0), "color": color, "light": light}
        se = fget("specularExponent", 1.0)
        if not (1.0 <= se <= 128.0):
            return dict(dummy)
        return {"kind": "specular", "in": resolve_input(ch, "in"), "surface_scale": fget("surfaceScale", 1.0),
                "constant": fget("specularConstant", 1.0), "exponent": se, "color": color, "light": light}
    return None


def _convert_fe_image(conv, ch):
    """usvg parser/filter.rs:809-880: a link to an element becomes that element's subtree; anything else is loaded as an
    image placed in the primitive subregion moved to the origin."""
    dummy = {"kind": "flood", "color": [0, 0, 0], "opacity": 0.0}
    href = ch.attrib.get("href")
    if href is None:
        return dict(dummy)
    target = conv.doc.link(href) if href.startswith("#") else None
    if href.startswith("#"):
        if target is None:
            return dict(dummy)
        n = conv.node_for(target)
        if n is None:
            return dict(dummy)
        if n["t"] == "g" and n["children"]:
            n["id"] = n["children"][0].get("id", "")
            n["children"][0]["id"] = ""
        return {"kind": "image", "root": {"t": "g", "id": "", "ts": list(IDENT), "children": [n], "opacity": 1.0, "isolate": False}}
    sub = conv._fe_subregion
    g = conv.image_for(ch, rect_override=(0.0, 0.0, sub[2], sub[3]))
    if g is None:
        return dict(dummy)
    return {"kind": "image", "root": {"t": "g", "id": "", "ts": list(IDENT), "children": [g], "opacity": 1.0, "isolate": False}}


# ---------------------------------------------------------------------------------------------------------------------
# execution (filter/mod.rs)
# ---------------------------------------------------------------------------------------------------------------------
def filter_region(f, ts):
    """render.rs:73-86: the group's layer bbox when it has filters."""
    b = rect_transform(tuple(f["rect"]), ts)
    if not (b[2] > 0 and b[3] > 0):
        return None
    return (int(math.floor(b[0])), int(math.floor(b[1])), int(max(math.ceil(b[2]), 1.0)), int(max(math.ceil(b[3]), 1.0)))


class _Img:
    def __init__(self, layer, region, cs):
        self.layer, self.region, self.cs = layer, region, cs


def _into_cs(be, img, cs):
    if img.cs == cs:
        return img
    l = be.clone(img.layer)
    be.f("into_srgb" if cs == "sRGB" else "into_linear_rgb", l)
    return _Img(l, img.region, cs)


def _scale_coords(x, y, ts):
    sx, sy = ts_get_scale(ts)
    return _f(f32(x) * f32(sx)), _f(f32(y) * f32(sy))


def _resolve_std_dev(sx, sy, ts):
    sx, sy = _scale_coords(sx, sy, ts)
    if _approx_zero(sx) and _approx_zero(sy):
        return None
    if sx < 0.05:
        sx = 0.0
    if sy < 0.05:
        sy = 0.0
    return float(sx), float(sy), (sx >= 2.0 or sy >= 2.0)


def _trunc_i32(v):
    return int(math.trunc(float(f32(v))))


def apply_filter(rd, f, ts, source):
    be = rd.be
    w, h = be.size(source)
    try:
        res = _apply_inner(rd, f, ts, source, w, h)
    except _FilterError:
        res = None
    if res is None:
        be.fill_color(source, 0, 0, 0, 0)
        return
    res = _into_cs(be, res, "sRGB")
    be.fill_color(source, 0, 0, 0, 0)
    be.draw_layer(source, res.layer, 0, 0)


class _FilterError(Exception):
    pass


def _apply_inner(rd, f, ts, source, sw, sh):
    be = rd.be
    r = rect_transform(tuple(f["rect"]), ts)
    if not (r[2] > 0 and r[3] > 0):
        raise _FilterError()
    region = fit_to_rect(to_int_rect(r), (0, 0, sw, sh))
    if region is None:
        raise _FilterError()
    results = []

    def get_input(inp):
        if inp[0] == "source" or inp[0] == "alpha":
            l = be.clone(source)
            if inp[0] == "alpha":
                # zero RGB, keep alpha: luminanceToAlpha-free way = colour matrix with zero rows for rgb
                be.f("color_matrix", l, "matrix", [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0])
            return _Img(l, region, "sRGB")
        for res in reversed(results):
            if res[0] == inp[1]:
                return res[1]
        return get_input(("source",))

    for p in f["primitives"]:
        sr = rect_transform(tuple(p["rect"]), ts)
        if not (sr[2] > 0 and sr[3] > 0):
            raise _FilterError()
        subregion = to_int_rect(sr)
        k = p["kind"]
        if k == "offset" and p["in"][0] == "ref":
            for res in reversed(results):
                if res[0] == p["in"][1]:
                    subregion = res[1].region
                    break
        cs = p["cs"]
        out = _apply_primitive(rd, p, k, cs, ts, region, subregion, get_input, source)
        if region != subregion:
            if k == "offset":
                sub2 = (0, 0, region[2], region[3])
            else:
                sub2 = (subregion[0] - region[0], subregion[1] - region[1], subregion[2], subregion[3])
            l = be.clone(out.layer)
            lw, lh = be.size(l)
            black = {"kind": "solid", "color": [0.0, 0.0, 0.0, 1.0]}

            def clear(x, y, rw, rh):
                if rw > 0 and rh > 0 and math.isfinite(rw) and math.isfinite(rh):
                    # fill_rect(rect, Clear, identity): non-AA rect, rounded coordinates (sub2 is integral)
                    x0, y0 = int(round(x)), int(round(y))
                    x1, y1 = x0 + max(1, int(round(rw))), y0 + max(1, int(round(rh)))
                    pts = [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]
                    be.fill_path(l, [M, L, L, L, Z], pts, black, "nonzero", IDENT, "clear", False)

            clear(0.0, 0.0, float(lw), float(sub2[1]))
            clear(0.0, 0.0, float(sub2[0]), float(lh))
            clear(float(sub2[0] + sub2[2]), 0.0, float(lw), float(lh))
            clear(0.0, float(sub2[1] + sub2[3]), float(lw), float(lh))
            out = _Img(l, subregion, out.cs)
        results.append((p["result"], out))
    return results[-1][1] if results else None


def _crop_to_region(be, layer, region, sw, sh):
    """Inputs are full-source-sized in resvg; filter buffers are region sized only for freshly created ones.  resvg's
    primitives that create a new pixmap use region.width()/height() while draw_pixmap(0,0) copies the top-left part."""
    return layer


def _apply_primitive(rd, p, k, cs, ts, region, subregion, get_input, source):
    be = rd.be
    rw, rh = region[2], region[3]
    if k == "blur":
        inp = get_input(p["in"])
        sd = _resolve_std_dev(p["sx"], p["sy"], ts)
        if sd is None:
            return inp
        img = _into_cs(be, inp, cs)
        l = be.clone(img.layer)
        be.f("box_blur" if sd[2] else "iir_blur", l, sd[0], sd[1])
        return _Img(l, _full(be, l), cs)
    if k == "drop_shadow":
        inp = get_input(p["in"])
        dx, dy = _scale_coords(p["dx"], p["dy"], ts)
        iw, ih = be.size(inp.layer)
        out = be.new_layer(iw, ih)
        img = _into_cs(be, inp, cs)
        shadow = be.clone(img.layer)
        sd = _resolve_std_dev(p["sx"], p["sy"], ts)
        if sd is not None:
            be.f("box_blur" if sd[2] else "iir_blur", shadow, sd[0], sd[1])
        _recolor(be, shadow, p["color"], p["opacity"])
        be.f("into_srgb" if cs == "sRGB" else "into_linear_rgb", shadow)
        be.draw_layer(out, shadow, _trunc_i32(dx), _trunc_i32(dy))
        be.draw_layer(out, img.layer, 0, 0)
        return _Img(out, _full(be, out), cs)
    if k == "offset":
        inp = get_input(p["in"])
        dx, dy = _scale_coords(p["dx"], p["dy"], ts)
        if _approx_zero(dx) and _approx_zero(dy):
            return inp
        iw, ih = be.size(inp.layer)
        out = be.new_layer(iw, ih)
        be.draw_layer(out, inp.layer, _trunc_i32(dx), _trunc_i32(dy))
        return _Img(out, _full(be, out), inp.cs)
    if k == "flood":
        out = be.new_layer(rw, rh)
        c = p["color"]
        be.fill_color(out, c[0] / 255.0, c[1] / 255.0, c[2] / 255.0, to_u8_opacity(p["opacity"]) / 255.0)
        return _Img(out, _full(be, out), "sRGB")
    if k in ("blend", "composite"):
        i1 = _into_cs(be, get_input(p["in"]), cs)
        i2 = _into_cs(be, get_input(p["in2"]), cs)
        if k == "composite" and p["op"] == "arithmetic":
            a, b = _sized(be, i1.layer, rw, rh), _sized(be, i2.layer, rw, rh)
            out = be.arithmetic(p["k"], a, b)
            return _Img(out, _full(be, out), cs)
        out = be.new_layer(rw, rh)
        be.draw_layer(out, i2.layer, 0, 0)
        if k == "blend":
            from tests.svgfront import BLEND_MAP
            mode = BLEND_MAP[p["mode"]]
        else:
            mode = {"over": "source_over", "in": "source_in", "out": "source_out", "atop": "source_atop", "xor": "xor"}[p["op"]]
        be.draw_layer(out, i1.layer, 0, 0, 1.0, mode)
        return _Img(out, _full(be, out), cs)
    if k == "merge":
        out = be.new_layer(rw, rh)
        for inp in p["inputs"]:
            i = _into_cs(be, get_input(inp), cs)
            be.draw_layer(out, i.layer, 0, 0)
        return _Img(out, _full(be, out), cs)
    if k == "tile":
        inp = get_input(p["in"])
        sub = (inp.region[0] - region[0], inp.region[1] - region[1], inp.region[2], inp.region[3])
        iw, ih = be.size(inp.layer)
        x0, y0 = max(sub[0], 0), max(sub[1], 0)
        x1, y1 = min(sub[0] + sub[2], iw), min(sub[1] + sub[3], ih)
        if x1 <= x0 or y1 <= y0:
            raise _FilterError()
        # Pixmap::clone_rect: the crop is a draw of the input at a negative offset onto a transparent layer (exact copy)
        tl = be.new_layer(x1 - x0, y1 - y0)
        be.draw_layer(tl, inp.layer, -x0, -y0)
        out = be.new_layer(rw, rh)
        spec = {"kind": "pattern", "layer": tl, "spread": "repeat", "quality": "bicubic", "opacity": 1.0,
                "ts": (1.0, 0.0, 0.0, 1.0, float(sub[0]), float(sub[1]))}
        pts = [(0, 0), (rw, 0), (rw, rh), (0, rh)]
        be.fill_path(out, [M, L, L, L, Z], pts, spec, "nonzero", IDENT, "source_over", False)
        return _Img(out, _full(be, out), "sRGB")
    if k == "image":
        # apply_image, mod.rs:870-897: the subtree rendered at the subregion's origin with the transform's scale only
        out = be.new_layer(rw, rh)
        sx, sy = ts_get_scale(ts)
        saved = rd.max_bbox
        rd.max_bbox = (0, 0, rw, rh)
        rd.render_nodes(p["root"], (sx, 0.0, 0.0, sy, float(subregion[0]), float(subregion[1])), out)
        rd.max_bbox = saved
        return _Img(out, _full(be, out), "sRGB")
    if k in ("component_transfer", "color_matrix"):
        img = _into_cs(be, get_input(p["in"]), cs)
        l = be.clone(img.layer)
        be.f("demultiply_alpha", l)
        if k == "component_transfer":
            be.component_transfer(l, p["funcs"])
        else:
            be.f("color_matrix", l, p["cm_kind"], p["params"])
        be.f("multiply_alpha", l)
        return _Img(l, _full(be, l), cs)
    if k == "convolve":
        img = _into_cs(be, get_input(p["in"]), cs)
        l = be.clone(img.layer)
        if p["preserve_alpha"]:
            be.f("demultiply_alpha", l)
        be.f("convolve_matrix", l, p["matrix"], p["cols"], p["rows"], p["tx"], p["ty"], p["divisor"], p["bias"], p["edge"],
             p["preserve_alpha"])
        return _Img(l, _full(be, l), cs)
    if k == "morphology":
        img = _into_cs(be, get_input(p["in"]), cs)
        l = be.clone(img.layer)
        rx, ry = _scale_coords(p["rx"], p["ry"], ts)
        if not (rx > 0.0 and ry > 0.0):
            be.fill_color(l, 0, 0, 0, 0)
            return _Img(l, _full(be, l), cs)
        be.f("morphology", l, p["op"], rx, ry)
        return _Img(l, _full(be, l), cs)
    if k == "displacement":
        i1 = _into_cs(be, get_input(p["in"]), cs)
        i2 = _into_cs(be, get_input(p["in2"]), cs)
        sx, sy = _scale_coords(p["scale"], p["scale"], ts)
        out = be.displacement_map(p["xch"], p["ych"], p["scale"], sx, sy, _sized(be, i1.layer, rw, rh), _sized(be, i2.layer, rw, rh))
        return _Img(out, _full(be, out), cs)
    if k == "turbulence":
        sx, sy = ts_get_scale(ts)
        if _approx_zero(sx) or _approx_zero(sy):
            out = be.new_layer(rw, rh)
            return _Img(out, _full(be, out), cs)
        out = be.turbulence(rw, rh, float(region[0]) - float(f32(ts[4])), float(region[1]) - float(f32(ts[5])), float(sx),
                            float(sy), float(f32(p["bfx"])), float(f32(p["bfy"])), p["octaves"], p["seed"], p["stitch"],
                            p["fractal"])
        be.f("multiply_alpha", out)
        return _Img(out, _full(be, out), cs)
    if k in ("diffuse", "specular"):
        inp = get_input(p["in"])
        light = dict(p["light"])
        if light["kind"] != "distant":
            sz = _f(f32(math.sqrt(float(f32(ts[0]) * f32(ts[0]) + f32(ts[3]) * f32(ts[3])))) / f32(math.sqrt(2.0)))
            x, y = ts_map(ts, light["x"], light["y"])
            light["x"], light["y"] = _f(f32(x) - f32(region[0])), _f(f32(y) - f32(region[1]))
            light["z"] = _f(f32(light["z"]) * f32(sz)) if light["kind"] == "spot" else _f(
                f32(light["z"]) * f32(math.sqrt(float(f32(ts[0]) * f32(ts[0]) + f32(ts[3]) * f32(ts[3])))) / f32(math.sqrt(2.0)))
            if light["kind"] == "spot":
                px, py = ts_map(ts, light["points_at"][0], light["points_at"][1])
                light["points_at"] = [_f(f32(px) - f32(region[0])), _f(f32(py) - f32(region[1])), _f(f32(light["points_at"][2]) * f32(sz))]
        src = _sized(be, inp.layer, rw, rh)
        if k == "diffuse":
            out = be.diffuse_lighting(p["surface_scale"], p["constant"], p["color"], light, src)
        else:
            out = be.specular_lighting(p["surface_scale"], p["constant"], p["exponent"], p["color"], light, src)
        return _Img(out, _full(be, out), cs)
    raise Unsupported(k)


def _full(be, l):
    w, h = be.size(l)
    return (0, 0, w, h)


def _sized(be, layer, w, h):
    lw, lh = be.size(layer)
    if (lw, lh) == (w, h):
        return layer
    raise Unsupported("filter region differs from layer size")


def _recolor(be, layer, color, opacity):
    """filter/mod.rs:606-617: every pixel becomes flood colour * (opacity.to_u8()/255 * alpha/255), premultiplied."""
    a8 = to_u8_opacity(opacity)
    if be.name == "gpu":
        be.rb.filters.flood_alpha(color, a8, layer)
        return
    arr = be.to_numpy(layer)
    ca = f32(a8) / f32(255.0)
    al = np.clip(ca * (arr[..., 3].astype(np.float32) / f32(255.0)), 0, 1).astype(np.float32)
    out = np.empty_like(arr)
    for i in range(3):
        c = f32(color[i]) / f32(255.0)
        pm = np.where(al == 1.0, c, np.clip(c * al, 0, 1)).astype(np.float32)
        out[..., i] = (pm * f32(255.0) + f32(0.5)).astype(np.uint8)
    out[..., 3] = (al * f32(255.0) + f32(0.5)).astype(np.uint8)
    layer[...] = out
