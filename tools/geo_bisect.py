"""Debug: renders every dashed stroke of a scene alone with the host builder and with the device geometry; prints the
draws whose pixels differ."""
import ctypes as C
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resvg_b200 as rb
from resvg_b200 import _ffi, scenes

def sub_scene(scene, i):
    s = {}
    vo, po = scene["verb_off"], scene["pt_off"]
    s["verb_off"] = np.array([0, vo[i + 1] - vo[i]], np.uint32)
    s["pt_off"] = np.array([0, po[i + 1] - po[i]], np.uint32)
    s["verbs"] = np.ascontiguousarray(scene["verbs"][vo[i]:vo[i + 1]])
    s["pts"] = np.ascontiguousarray(scene["pts"][po[i]:po[i + 1]])
    s["rules"] = np.ascontiguousarray(scene["rules"][i:i + 1])
    return s

def main():
    w, h, n, seed = 640, 480, 900, 0xC2
    scene = scenes.paths_scene(w, h, n, seed, rmin=6.0, rmax=120.0)
    paints = scenes.to_paint_array(scene, _ffi.Paint)
    strokes = scenes.to_stroke_array(scene, _ffi.Stroke)
    ctx = rb.Context(0)
    idx = [i for i in range(scene["n_paths"]) if scene["n_dash"][i] > 0]
    print(len(idx), "dashed draws")
    for i in idx:
        s = sub_scene(scene, i)
        P1 = (_ffi.Paint * 1)(); C.memmove(P1, C.byref(paints, i * C.sizeof(_ffi.Paint)), C.sizeof(_ffi.Paint))
        S1 = (_ffi.Stroke * 1)(); C.memmove(S1, C.byref(strokes, i * C.sizeof(_ffi.Stroke)), C.sizeof(_ffi.Stroke))
        s["paints"] = P1; s["strokes"] = S1
        out = []
        for mode in (2, 1):
            _ffi.lib.rb_debug_geo_mode(mode)
            l = ctx.layer(w, h)
            b = rb.Batch(l)
            b.fill_paths(s)
            b.submit()
            out.append(l.download())
            _ffi.lib.rb_debug_geo_mode(0)
        if not np.array_equal(out[0], out[1]):
            d = np.argwhere((out[0] != out[1]).any(axis=-1))
            print("draw", i, "differs at", d[:5].tolist(), "width", scene["stroke_width"][i], "cap/join", S1[0].cap, S1[0].join,
                  "dash", [S1[0].dash_array[k] for k in range(S1[0].n_dash)], S1[0].dash_offset, "aa", P1[0].anti_alias)
            print(" verbs", s["verbs"].tolist()); print(" pts", s["pts"].tolist())
main()
