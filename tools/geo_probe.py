"""One end-to-end submit of a paths scene through the device geometry path (for ncu launch lists / timing).
usage: python tools/geo_probe.py [paths8k] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resvg_b200 as rb  # noqa: E402
from resvg_b200 import _ffi, scenes  # noqa: E402
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "paths8k"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    W, H, n_paths, seed = bench.WORKLOADS[wl]
    seed += int(sys.argv[3]) if len(sys.argv) > 3 else 0
    threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    scene = scenes.paths_scene(W, H, n_paths, seed)
    scene["paints"] = scenes.to_paint_array(scene, _ffi.Paint)
    scene["strokes"] = scenes.to_stroke_array(scene, _ffi.Stroke)
    ctx = rb.Context(0)
    layer = ctx.layer(W, H)
    for s in range(steps):
        layer.fill(0, 0, 0, 0)
        t0 = time.perf_counter()
        b = rb.Batch(layer)
        b.fill_paths(scene)
        b.submit(threads)
        b.close()
        ctx.synchronize()
        print(f"step {s}: {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr)
    ctx.close()


if __name__ == "__main__":
    main()
