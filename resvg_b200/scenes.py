"""Synthetic workloads of BASELINE.json, generated deterministically (SplitMix64, counter based so it vectorises).

``paths_scene`` is config C2 of SURVEY.md §8(d): N random closed paths of cubic / quadratic / line segments on a
W x H canvas, non-zero and even-odd fills, solid / linear / radial paints, 95 % anti-aliased.  The result is a plain
dict of packed numpy arrays (the layout rb_batch_fill_paths consumes); ``to_rb_paints`` / ``to_paint_array`` turn the
neutral paint table into a ctypes array of the caller's paint struct (the library's rb_paint, or the CPU checker's
struct in tests/bench).
"""
import ctypes as C
import math

import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64_uniform(seed: int, n: int) -> np.ndarray:
    """First n outputs of SplitMix64(seed) mapped to [0, 1) doubles."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.arange(1, n + 1, dtype=np.uint64) * _GAMMA)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


STREAM = 160  # uniforms reserved per path

MOVE, LINE, QUAD, CUBIC, CLOSE = 0, 1, 2, 3, 4


def _gather_ranges(off, idx):
    """Indices selecting the concatenation of ranges off[i]:off[i+1] for i in idx, plus the new offsets."""
    lens = (off[1:] - off[:-1]).astype(np.int64)[idx]
    new_off = np.zeros(len(idx) + 1, np.uint32)
    new_off[1:] = np.cumsum(lens)
    starts = off[:-1].astype(np.int64)[idx]
    pos = np.arange(int(new_off[-1]), dtype=np.int64) - np.repeat(new_off[:-1].astype(np.int64), lens)
    return np.repeat(starts, lens) + pos, new_off


def paths_scene(width=8192, height=8192, n_paths=100_000, seed=0x5EED0002, rmin=8.0, rmax=256.0, strokes=True,
                stroke_wmin=0.5, stroke_wmax=16.0, dashed=0.1, u=None, p_solid=0.5, p_linear=0.8):
    """C2 'paths8k': n_paths random closed paths; 70 % are filled, 20 % filled then stroked, 10 % stroked only
    (strokes=True), a fraction `dashed` of the strokes with a dash array of 2-4 intervals U[2, 32] (an odd list is
    repeated, as usvg does).  The returned arrays hold one entry per DRAW (a filled+stroked path is two entries).  Stroke
    width is log-uniform [stroke_wmin, stroke_wmax]; anti-aliased strokes of at most 1 px are hairlines for tiny-skia."""
    if u is None:
        u = splitmix64_uniform(seed, n_paths * STREAM).reshape(n_paths, STREAM)
    else:  # caller-provided uniforms, one row of STREAM values per path (icons_docs: one SplitMix64 stream per document)
        n_paths = len(u)
    cx = u[:, 0] * width
    cy = u[:, 1] * height
    radius = np.exp(np.log(rmin) + u[:, 2] * (np.log(rmax) - np.log(rmin)))
    n_seg = 3 + np.minimum((u[:, 3] * 10).astype(np.int64), 9)  # 3..12
    evenodd = (u[:, 4] < 0.5).astype(np.uint8)
    paint_kind = np.where(u[:, 5] < p_solid, 0, np.where(u[:, 5] < p_linear, 1, 2)).astype(np.int32)  # solid/linear/radial
    anti_alias = (u[:, 6] < 0.95).astype(np.int32)

    # segments: per segment one kind draw + 6 coordinate draws, starting at column 32
    seg_kind_u = u[:, 32:32 + 12]
    seg_kind = np.where(seg_kind_u < 0.5, CUBIC, np.where(seg_kind_u < 0.8, QUAD, LINE)).astype(np.uint8)
    coords = u[:, 48:48 + 12 * 6].reshape(n_paths, 12, 3, 2) * 2.0 - 1.0
    start = u[:, 44:46] * 2.0 - 1.0
    seg_valid = np.arange(12)[None, :] < n_seg[:, None]
    pts_per_seg = np.where(seg_kind == CUBIC, 3, np.where(seg_kind == QUAD, 2, 1)) * seg_valid
    n_pts = 1 + pts_per_seg.sum(axis=1)
    n_verbs = 2 + n_seg  # move + segments + close
    pt_off = np.zeros(n_paths + 1, np.uint32)
    verb_off = np.zeros(n_paths + 1, np.uint32)
    pt_off[1:] = np.cumsum(n_pts)
    verb_off[1:] = np.cumsum(n_verbs)

    verbs = np.empty(int(verb_off[-1]), np.uint8)
    pts = np.empty((int(pt_off[-1]), 2), np.float32)
    centre = np.stack([cx, cy], axis=1)
    # move
    verbs[verb_off[:-1]] = MOVE
    pts[pt_off[:-1]] = (centre + radius[:, None] * start).astype(np.float32)
    verbs[verb_off[1:] - 1] = CLOSE
    seg_pt_start = 1 + np.concatenate([np.zeros((n_paths, 1), np.int64), np.cumsum(pts_per_seg, axis=1)[:, :-1]], axis=1)
    for s in range(12):
        m = seg_valid[:, s]
        idx = np.nonzero(m)[0]
        verbs[verb_off[idx] + 1 + s] = seg_kind[idx, s]
        for k in range(3):
            mk = pts_per_seg[idx, s] > k
            ii = idx[mk]
            dst = pt_off[ii].astype(np.int64) + seg_pt_start[ii, s] + k
            pts[dst] = (centre[ii] + radius[ii, None] * coords[ii, s, k]).astype(np.float32)

    # paints
    color = np.empty((n_paths, 4), np.float32)
    color[:, 0] = np.floor(u[:, 8] * 256) / 255.0
    color[:, 1] = np.floor(u[:, 9] * 256) / 255.0
    color[:, 2] = np.floor(u[:, 10] * 256) / 255.0
    color[:, 3] = (32 + np.floor(u[:, 11] * 224)) / 255.0
    spread = np.minimum((u[:, 12] * 3).astype(np.int32), 2)
    n_stops = np.where(paint_kind == 0, 0, 2 + np.minimum((u[:, 13] * 7).astype(np.int64), 6)).astype(np.int32)  # 2..8
    stop_off = np.zeros(n_paths + 1, np.int64)
    stop_off[1:] = np.cumsum(n_stops)
    stops = np.zeros((int(stop_off[-1]), 5), np.float32)
    su = u[:, 120:160].reshape(n_paths, 8, 5)
    for k in range(8):
        m = n_stops > k
        idx = np.nonzero(m)[0]
        dst = stop_off[idx] + k
        ns = n_stops[idx].astype(np.float64)
        # increasing offsets: (k + jitter) / n
        stops[dst, 0] = ((k + su[idx, k, 0] * 0.999) / ns).astype(np.float32)
        stops[dst, 1] = np.floor(su[idx, k, 1] * 256) / 255.0
        stops[dst, 2] = np.floor(su[idx, k, 2] * 256) / 255.0
        stops[dst, 3] = np.floor(su[idx, k, 3] * 256) / 255.0
        stops[dst, 4] = (32 + np.floor(su[idx, k, 4] * 224)) / 255.0
    ang = u[:, 14] * 2 * np.pi
    foc = u[:, 15] * 0.8 * radius
    geom = np.zeros((n_paths, 6), np.float32)  # x0,y0,r0,x1,y1,r1
    lin = paint_kind == 1
    geom[lin, 0] = (cx - radius * np.cos(ang))[lin]
    geom[lin, 1] = (cy - radius * np.sin(ang))[lin]
    geom[lin, 3] = (cx + radius * np.cos(ang))[lin]
    geom[lin, 4] = (cy + radius * np.sin(ang))[lin]
    rad = paint_kind == 2
    geom[rad, 0] = (cx + foc * np.cos(ang))[rad]
    geom[rad, 1] = (cy + foc * np.sin(ang))[rad]
    geom[rad, 2] = 0.0
    geom[rad, 3] = cx[rad]
    geom[rad, 4] = cy[rad]
    geom[rad, 5] = radius[rad]
    sc = dict(width=width, height=height, n_paths=n_paths, n_source_paths=n_paths, verb_off=verb_off, pt_off=pt_off,
              verbs=verbs, pts=pts, rules=evenodd, paint_kind=paint_kind, color=color, geom=geom, spread=spread,
              n_stops=n_stops, stop_off=stop_off, stops=stops, anti_alias=anti_alias, radius=radius.astype(np.float32),
              stroke_width=np.zeros(n_paths, np.float32), stroke_miter=np.full(n_paths, 4.0, np.float32),
              stroke_cap=np.zeros(n_paths, np.int32), stroke_join=np.zeros(n_paths, np.int32))
    if not strokes:
        return sc
    # draw list: fill (kind < 0.9), then stroke (kind >= 0.7), in path order
    kind = u[:, 7]
    has_fill, has_stroke = kind < 0.9, kind >= 0.7
    n_entries = has_fill.astype(np.int64) + has_stroke.astype(np.int64)
    src = np.repeat(np.arange(n_paths), n_entries)
    first = np.zeros(len(src), bool)
    first[np.cumsum(n_entries) - n_entries] = True
    is_stroke = np.where(first, ~has_fill[src], True)
    vi, new_voff = _gather_ranges(verb_off, src)
    pi, new_poff = _gather_ranges(pt_off, src)
    out = dict(sc)
    out.update(n_paths=len(src), verb_off=new_voff, pt_off=new_poff, verbs=verbs[vi], pts=pts[pi], rules=evenodd[src],
               anti_alias=anti_alias[src], radius=sc["radius"][src])
    # stroke draws are painted with their own solid colour; fills keep the path's paint
    pk = paint_kind[src].copy()
    col = color[src].copy()
    scol = np.stack([np.floor(u[:, 16] * 256) / 255.0, np.floor(u[:, 17] * 256) / 255.0, np.floor(u[:, 18] * 256) / 255.0,
                     (32 + np.floor(u[:, 19] * 224)) / 255.0], axis=1).astype(np.float32)
    pk[is_stroke] = 0
    col[is_stroke] = scol[src][is_stroke]
    ns = n_stops[src].copy()
    ns[is_stroke] = 0
    so = stop_off[:-1][src].copy()  # offsets into the unchanged stop pool (not a prefix sum any more)
    sw = np.exp(np.log(stroke_wmin) + u[:, 20] * (np.log(stroke_wmax) - np.log(stroke_wmin))).astype(np.float32)
    out.update(paint_kind=pk, color=col, geom=geom[src], spread=spread[src], n_stops=ns, stop_start=so,
               stroke_width=np.where(is_stroke, sw[src], 0.0).astype(np.float32),
               stroke_miter=np.full(len(src), 4.0, np.float32),
               stroke_cap=np.minimum((u[:, 21] * 3).astype(np.int32), 2)[src],
               stroke_join=np.array([0, 2, 3], np.int32)[np.minimum((u[:, 22] * 3).astype(np.int64), 2)][src])
    # dash arrays: up to 6 floats per draw (3 intervals are repeated to 6); n_dash = 0 for solid strokes and fills
    cnt = 2 + np.minimum((u[:, 24] * 3).astype(np.int64), 2)
    vals = (2.0 + 30.0 * u[:, 25:29]).astype(np.float32)
    dash = np.zeros((n_paths, 6), np.float32)
    dash[:, :4] = vals
    three = cnt == 3
    dash[three, 3:6] = vals[three, :3]
    n_dash = np.where(u[:, 23] < dashed, np.where(three, 6, cnt), 0).astype(np.int32)
    out.update(dash=np.ascontiguousarray(dash[src]), n_dash=np.where(is_stroke, n_dash[src], 0).astype(np.int32))
    return out


def to_paint_array(scene, paint_struct, blend_mode=3):
    """ctypes array of `paint_struct` (fields: shader, color, x0..r1, n_stops, stops, spread, ts, blend_mode,
    anti_alias ...) filled from the neutral tables without a Python loop."""
    n = scene["n_paths"]
    arr = (paint_struct * n)()
    view = np.frombuffer(arr, dtype=np.uint8).reshape(n, C.sizeof(paint_struct))

    def put(field, values, dtype):
        f = getattr(paint_struct, field)
        raw = np.ascontiguousarray(values, dtype=dtype)
        raw = raw.reshape(n, -1).view(np.uint8)
        view[:, f.offset:f.offset + raw.shape[1]] = raw

    put("shader", scene["paint_kind"], np.int32)
    put("color", scene["color"], np.float32)
    g = scene["geom"]
    for i, name in enumerate(["x0", "y0", "r0", "x1", "y1", "r1"]):
        put(name, g[:, i], np.float32)
    put("n_stops", scene["n_stops"], np.int32)
    base = scene["stops"].ctypes.data
    start = scene["stop_start"] if "stop_start" in scene else scene["stop_off"][:-1]
    ptr = np.where(scene["n_stops"] > 0, base + start * 20, 0).astype(np.uint64)
    put("stops", ptr, np.uint64)
    put("spread", scene["spread"], np.int32)
    put("ts", np.tile(np.array([1, 0, 0, 1, 0, 0], np.float32), (n, 1)), np.float32)
    put("blend_mode", np.full(n, blend_mode, np.int32), np.int32)
    put("anti_alias", scene["anti_alias"], np.int32)
    if hasattr(paint_struct, "opacity"):
        put("opacity", np.ones(n, np.float32), np.float32)
    return arr


def to_stroke_array(scene, stroke_struct):
    """ctypes array of rb_stroke {width, miter_limit, cap, join, dash_array, n_dash, dash_offset}; width 0 marks a fill
    entry.  dash_array points into scene["dash"], which must outlive the array."""
    n = scene["n_paths"]
    arr = (stroke_struct * n)()
    view = np.frombuffer(arr, dtype=np.uint8).reshape(n, C.sizeof(stroke_struct))
    if "n_dash" in scene and hasattr(stroke_struct, "n_dash"):
        nd = np.ascontiguousarray(scene["n_dash"], np.int32)
        f = stroke_struct.n_dash
        view[:, f.offset:f.offset + 4] = nd.reshape(n, 1).view(np.uint8)
        ptr = np.where(nd > 0, scene["dash"].ctypes.data + np.arange(n, dtype=np.uint64) * 24, 0).astype(np.uint64)
        f = stroke_struct.dash_array
        view[:, f.offset:f.offset + 8] = ptr.reshape(n, 1).view(np.uint8)
    for name, dt in (("width", np.float32), ("miter_limit", np.float32), ("cap", np.int32), ("join", np.int32)):
        f = getattr(stroke_struct, name)
        key = {"width": "stroke_width", "miter_limit": "stroke_miter", "cap": "stroke_cap", "join": "stroke_join"}[name]
        view[:, f.offset:f.offset + 4] = np.ascontiguousarray(scene[key], dtype=dt).reshape(n, 1).view(np.uint8)
    return arr


def subset(scene, n):
    """First n paths of a scene (same canvas): the bounded sample the CPU baseline is timed on."""
    n = min(n, scene["n_paths"])
    out = dict(scene)
    out["n_paths"] = n
    out["verb_off"] = scene["verb_off"][: n + 1].copy()
    out["pt_off"] = scene["pt_off"][: n + 1].copy()
    for k in ("rules", "paint_kind", "color", "geom", "spread", "n_stops", "anti_alias", "radius", "stroke_width",
              "stroke_miter", "stroke_cap", "stroke_join", "stop_start", "dash", "n_dash"):
        if k in scene:
            out[k] = scene[k][:n].copy()
    out["stop_off"] = scene["stop_off"][: n + 1].copy()
    return out


ICON_SEED = 0x5EED0005
ICON_SIZE = 256
ICON_HEAD = 8  # per-document uniforms drawn before the first path's row


def icons_docs(first_doc, n_docs, size=ICON_SIZE):
    """C5 'icons' (SURVEY.md section 8(d)): documents first_doc .. first_doc + n_docs - 1, each size x size px and seeded
    ICON_SEED + i: 5-40 closed paths (radius log-U[4, 96], filled, 80 % solid / 10 % linear / 10 % radial, 95 % AA);
    10 % of the documents put the second half of their paths into one group with opacity U[0.3, 0.9]; another 5 % wrap
    everything into a group with a drop shadow (stdDeviation U[2, 4], dx = dy = 4, black at 50 %, sRGB).  Returns the
    packed paths of all documents in document-local coordinates (the paths_scene layout) plus per-document tables:
    doc_first (n_docs + 1 path offsets), group_first (first path of the opacity group or -1), group_opacity, shadow_sigma
    (0 = none)."""
    ids = np.arange(first_doc, first_doc + n_docs, dtype=np.uint64)
    with np.errstate(over="ignore"):
        seeds = np.uint64(ICON_SEED) + ids

    def rows(seed_arr, start, count):
        """count consecutive SplitMix64 uniforms per row, row r starting at output index start[r] of stream seed_arr[r]"""
        with np.errstate(over="ignore"):
            k = (start.astype(np.uint64)[:, None] + np.arange(1, count + 1, dtype=np.uint64)[None, :])
            z = seed_arr[:, None] + k * _GAMMA
            z = (z ^ (z >> np.uint64(30))) * _M1
            z = (z ^ (z >> np.uint64(27))) * _M2
            z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))

    head = rows(seeds, np.zeros(n_docs, np.uint64), ICON_HEAD)
    n_paths = 5 + np.minimum((head[:, 0] * 36).astype(np.int64), 35)
    doc_first = np.zeros(n_docs + 1, np.uint32)
    doc_first[1:] = np.cumsum(n_paths)
    doc_of = np.repeat(np.arange(n_docs), n_paths)
    j = np.arange(int(doc_first[-1])) - doc_first[:-1].astype(np.int64)[doc_of]
    u = rows(seeds[doc_of], (ICON_HEAD + j * STREAM).astype(np.uint64), STREAM)
    sc = paths_scene(size, size, rmin=4.0, rmax=96.0, strokes=False, u=u, p_solid=0.8, p_linear=0.9)
    grouped = head[:, 1] < 0.10
    shadow = (head[:, 1] >= 0.10) & (head[:, 1] < 0.15)
    sc.update(n_docs=n_docs, first_doc=first_doc, doc_first=doc_first, doc_of=doc_of.astype(np.int32), doc_size=size,
              group_first=np.where(grouped, doc_first[:-1].astype(np.int64) + n_paths // 2, -1),
              group_opacity=(0.3 + 0.6 * head[:, 2]).astype(np.float32),
              shadow_sigma=np.where(shadow, 2.0 + 2.0 * head[:, 3], 0.0))
    return sc


STACK_SEED = 0x5EED0004


def stack_svg(size=4096, levels=64, seed=STACK_SEED, inset=None, shapes=10):
    """C4 'stack4k'  as SVG text: `levels` nested groups on a size x size canvas; level k has opacity
    U[0.85, 0.99] and, by k mod 4: 0 a luminance mask (a linear-gradient rectangle), 1 a clip-path (a circle; every 8th level
    that circle is itself clipped by a nested clip-path), 2 a pattern-filled rectangle (tile 32-128 px holding 3 shapes),
    3 nothing more; every level draws `shapes` C2-style closed paths (cubic / quad / line segments, solid or linear-gradient
    fill, both fill rules) inside a box inset `inset` px per level (default size / 256, i.e. 16 px at 4096)."""
    inset = size / 256.0 if inset is None else inset
    per = 16 + shapes * 64
    u = splitmix64_uniform(seed, levels * per).reshape(levels, per)
    defs, body, tail = [], [], []
    fmt = lambda v: f"{v:.3f}"

    def shape(uu, x0, y0, x1, y1, ident):
        """one closed path from 64 uniforms inside the box"""
        w, h = x1 - x0, y1 - y0
        cx, cy = x0 + uu[0] * w, y0 + uu[1] * h
        r = math.exp(math.log(8.0) + uu[2] * (math.log(max(16.0, min(w, h) / 4)) - math.log(8.0)))
        nseg = 3 + min(int(uu[3] * 6), 5)
        pt = lambda a, b: f"{fmt(cx + r * (2 * a - 1))} {fmt(cy + r * (2 * b - 1))}"
        d = ["M " + pt(uu[4], uu[5])]
        k = 8
        for _ in range(nseg):
            kind = uu[k]
            if kind < 0.5:
                d.append(f"C {pt(uu[k + 1], uu[k + 2])} {pt(uu[k + 3], uu[k + 4])} {pt(uu[k + 5], uu[k + 6])}")
            elif kind < 0.8:
                d.append(f"Q {pt(uu[k + 1], uu[k + 2])} {pt(uu[k + 3], uu[k + 4])}")
            else:
                d.append(f"L {pt(uu[k + 1], uu[k + 2])}")
            k += 7
        d.append("Z")
        col = "#%02x%02x%02x" % (int(uu[56] * 256), int(uu[57] * 256), int(uu[58] * 256))
        rule = "evenodd" if uu[59] < 0.5 else "nonzero"
        op = fmt(0.3 + 0.7 * uu[60])
        if uu[61] < 0.25:
            col2 = "#%02x%02x%02x" % (int(uu[62] * 256), int(uu[63] * 256), int(uu[6] * 256))
            defs.append(f'<linearGradient id="g{ident}" gradientUnits="userSpaceOnUse" x1="{fmt(cx - r)}" y1="{fmt(cy - r)}" '
                        f'x2="{fmt(cx + r)}" y2="{fmt(cy + r)}"><stop offset="0" stop-color="{col}"/>'
                        f'<stop offset="1" stop-color="{col2}" stop-opacity="{op}"/></linearGradient>')
            fill = f'url(#g{ident})'
        else:
            fill = col
        return f'<path d="{" ".join(d)}" fill="{fill}" fill-opacity="{op}" fill-rule="{rule}"/>'

    for k in range(levels):
        x0 = y0 = inset * k
        x1 = y1 = size - inset * k
        w = x1 - x0
        uu = u[k]
        attrs = [f'opacity="{fmt(0.85 + 0.14 * uu[0])}"']
        if k % 4 == 0:
            defs.append(f'<linearGradient id="ml{k}" x1="0" y1="0" x2="1" y2="{fmt(uu[1])}"><stop offset="0" stop-color="white"/>'
                        f'<stop offset="1" stop-color="#404040"/></linearGradient>'
                        f'<mask id="m{k}" maskUnits="userSpaceOnUse" x="{fmt(x0)}" y="{fmt(y0)}" width="{fmt(w)}" height="{fmt(w)}">'
                        f'<rect x="{fmt(x0)}" y="{fmt(y0)}" width="{fmt(w)}" height="{fmt(w)}" fill="url(#ml{k})"/></mask>')
            attrs.append(f'mask="url(#m{k})"')
        elif k % 4 == 1:
            c, rad = (x0 + x1) / 2, w * (0.45 + 0.1 * uu[1])
            nested = ""
            if k % 8 == 1:
                defs.append(f'<clipPath id="cc{k}"><rect x="{fmt(x0 + w * 0.05)}" y="{fmt(y0)}" width="{fmt(w * 0.9)}" height="{fmt(w)}"/></clipPath>')
                nested = f' clip-path="url(#cc{k})"'
            defs.append(f'<clipPath id="c{k}"{nested}><circle cx="{fmt(c)}" cy="{fmt(c)}" r="{fmt(rad)}"/></clipPath>')
            attrs.append(f'clip-path="url(#c{k})"')
        body.append(f'<g {" ".join(attrs)}>')
        if k % 4 == 2:
            tile = 32 + 96 * uu[1]
            inner = "".join(shape(uu[16 + 64 * j:16 + 64 * (j + 1)] * 1.0, 0, 0, tile, tile, f"p{k}_{j}") for j in range(3))
            defs.append(f'<pattern id="p{k}" patternUnits="userSpaceOnUse" width="{fmt(tile)}" height="{fmt(tile)}">{inner}</pattern>')
            body.append(f'<rect x="{fmt(x0)}" y="{fmt(y0)}" width="{fmt(w)}" height="{fmt(w)}" fill="url(#p{k})" fill-opacity="0.5"/>')
        for j in range(shapes):
            body.append(shape(uu[16 + 64 * j:16 + 64 * (j + 1)], x0, y0, x1, y1, f"{k}_{j}"))
        tail.append("</g>")
    return (f'<svg xmlns="http://www.w3.org/2000/svg" width="{size}" height="{size}" viewBox="0 0 {size} {size}">'
            f'<defs>{"".join(defs)}</defs>{"".join(body)}{"".join(tail)}</svg>')
