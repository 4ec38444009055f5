"""Times the host half of the device geometry path (classify, cull, paints, task lists) without a GPU.
usage: python tools/geo_host_profile.py [paths8k] [repeats]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resvg_b200 import _ffi, scenes  # noqa: E402
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "paths8k"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    W, H, n_paths, seed = bench.WORKLOADS[wl]
    scene = scenes.paths_scene(W, H, n_paths, seed)
    paints = scenes.to_paint_array(scene, _ffi.Paint)
    strokes = scenes.to_stroke_array(scene, _ffi.Stroke)
    lib = _ffi.lib
    for r in range(reps):
        h = C.c_void_p()
        assert lib.rb_debug_batch_begin_host(W, H, C.byref(h)) == 0
        st = lib.rb_batch_draw_paths(h, scene["n_paths"], scene["verb_off"].ctypes.data, scene["pt_off"].ctypes.data,
                                     scene["verbs"].ctypes.data, scene["pts"].ctypes.data, C.addressof(paints),
                                     scene["rules"].ctypes.data, C.addressof(strokes), None)
        assert st == 0, st
        out = (C.c_uint64 * 8)()
        t1 = time.perf_counter()
        assert lib.rb_debug_geo_host_stats(h, out) == 0
        t2 = time.perf_counter()
        print(f"rep {r}: {1e3*(t2-t1):.2f} ms | tasks {out[0]} dash {out[1]} stroke {out[2]} hair {out[3]} fill {out[4]} "
              f"bytes {out[5]/2**20:.1f} MiB verbs {out[6]} pts {out[7]}")
        lib.rb_batch_destroy(h)


if __name__ == "__main__":
    main()
