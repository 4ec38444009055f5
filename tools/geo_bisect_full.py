"""Debug: finds the draws of the full-size scene whose pixels differ between the host builder and the device geometry."""
import ctypes as C
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resvg_b200 as rb
from resvg_b200 import _ffi, scenes

def sub_scene(scene, paints, strokes, idx):
    s = {}
    vo, po = scene["verb_off"], scene["pt_off"]
    vs, ps, voff, poff = [], [], [0], [0]
    for i in idx:
        vs.append(scene["verbs"][vo[i]:vo[i + 1]]); ps.append(scene["pts"][po[i]:po[i + 1]])
        voff.append(voff[-1] + len(vs[-1])); poff.append(poff[-1] + len(ps[-1]))
    s["verb_off"] = np.array(voff, np.uint32); s["pt_off"] = np.array(poff, np.uint32)
    s["verbs"] = np.ascontiguousarray(np.concatenate(vs)); s["pts"] = np.ascontiguousarray(np.concatenate(ps))
    s["rules"] = np.ascontiguousarray(scene["rules"][idx])
    P = (_ffi.Paint * len(idx))(); S = (_ffi.Stroke * len(idx))()
    for k, i in enumerate(idx):
        C.memmove(C.byref(P, k * C.sizeof(_ffi.Paint)), C.byref(paints, int(i) * C.sizeof(_ffi.Paint)), C.sizeof(_ffi.Paint))
        C.memmove(C.byref(S, k * C.sizeof(_ffi.Stroke)), C.byref(strokes, int(i) * C.sizeof(_ffi.Stroke)), C.sizeof(_ffi.Stroke))
    s["paints"] = P; s["strokes"] = S
    return s

def render(ctx, l, s, mode):
    _ffi.lib.rb_debug_geo_mode(mode)
    l.fill(0, 0, 0, 0)
    b = rb.Batch(l); b.fill_paths(s); b.submit(); b.close()
    out = l.download()
    _ffi.lib.rb_debug_geo_mode(0)
    return out

def main():
    w = h = 8192
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    scene = scenes.paths_scene(w, h, n, 0x5EED0002)
    paints = scenes.to_paint_array(scene, _ffi.Paint)
    strokes = scenes.to_stroke_array(scene, _ffi.Stroke)
    ctx = rb.Context(0)
    l = ctx.layer(w, h)
    nd = scene["n_paths"]
    full = dict(scene); full["paints"] = paints; full["strokes"] = strokes
    a = render(ctx, l, full, 2); b = render(ctx, l, full, 1)
    d = np.argwhere((a != b).any(axis=-1))
    print(len(d), "pixels differ", d[:10].tolist())
    vo, po = scene["verb_off"], scene["pt_off"]
    mins = np.minimum.reduceat(scene["pts"], po[:-1].astype(np.int64), axis=0)
    maxs = np.maximum.reduceat(scene["pts"], po[:-1].astype(np.int64), axis=0)
    seen = set()
    for (y, x) in d[:6]:
        reach = 0.5 * scene["stroke_width"] * 4.0 + 2
        cand = np.nonzero((mins[:, 0] - reach <= x) & (maxs[:, 0] + reach >= x) & (mins[:, 1] - reach <= y) & (maxs[:, 1] + reach >= y))[0]
        print("pixel", x, y, "candidates", len(cand))
        for i in cand:
            if int(i) in seen: continue
            s = sub_scene(scene, paints, strokes, [i])
            ra = render(ctx, l, s, 2); rb_ = render(ctx, l, s, 1)
            if not np.array_equal(ra, rb_):
                seen.add(int(i))
                dd = np.argwhere((ra != rb_).any(axis=-1))
                S = strokes[int(i)]
                print("  draw", i, "differs at", dd[:4].tolist(), "stroke w", scene["stroke_width"][i], "cap", S.cap, "join", S.join, "miter", S.miter_limit, "ndash", S.n_dash,
                      "aa", paints[int(i)].anti_alias, "rule", scene["rules"][i])
                print("   verbs", scene["verbs"][vo[i]:vo[i + 1]].tolist())
                print("   pts", scene["pts"][po[i]:po[i + 1]].tolist())
main()
