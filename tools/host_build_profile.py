"""Times the host half of a batch (record + edge build + binning + layout) without a GPU.
usage: python tools/host_build_profile.py [paths8k|paths2k] [threads] [repeats]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resvg_b200 import _ffi, scenes  # noqa: E402
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "paths8k"
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    W, H, n_paths, seed = bench.WORKLOADS[wl]
    scene = scenes.paths_scene(W, H, n_paths, seed)
    paints = scenes.to_paint_array(scene, _ffi.Paint)
    strokes = scenes.to_stroke_array(scene, _ffi.Stroke)
    lib = _ffi.lib
    for r in range(reps):
        h = C.c_void_p()
        assert lib.rb_debug_batch_begin_host(W, H, C.byref(h)) == 0
        t0 = time.perf_counter()
        st = lib.rb_batch_draw_paths(h, scene["n_paths"], scene["verb_off"].ctypes.data, scene["pt_off"].ctypes.data,
                                     scene["verbs"].ctypes.data, scene["pts"].ctypes.data, C.addressof(paints),
                                     scene["rules"].ctypes.data, C.addressof(strokes), None)
        assert st == 0, st
        t1 = time.perf_counter()
        assert lib.rb_batch_prepare(h, nt) == 0
        t2 = time.perf_counter()
        ph = (C.c_uint64 * 6)()
        lib.rb_debug_batch_phases(h, ph)
        stats = (C.c_uint64 * 6)()
        lib.rb_batch_stats(h, stats)
        print(f"rep {r}: record {1e3*(t1-t0):.1f} ms, prepare {1e3*(t2-t1):.1f} ms | build {ph[0]/1e3:.1f} layout+count {ph[1]/1e3:.1f} "
              f"alloc {ph[4]/1e3:.1f} pack {ph[2]/1e3:.1f} lists {ph[3]/1e3:.1f} total {ph[5]/1e3:.1f} | draws {stats[0]} edges {stats[1]} "
              f"pairs {stats[2]} tiles {stats[3]} bytes {stats[4]/2**20:.0f} MiB")
        lib.rb_batch_destroy(h)


if __name__ == "__main__":
    main()
