"""Host half of the device geometry path (batch_host.cpp rb_geo_host_build): classification, culling, DrawTiler tiles,
per-verb hairline tasks, unit bounds — no GPU needed (host-only batches)."""
import ctypes as C

import numpy as np

from resvg_b200 import _ffi, scenes


def _stats(w, h, scene, strokes=True, ts=None):
    lib = _ffi.lib
    paints = scenes.to_paint_array(scene, _ffi.Paint)
    st = scenes.to_stroke_array(scene, _ffi.Stroke) if strokes else None
    hnd = C.c_void_p()
    assert lib.rb_debug_batch_begin_host(w, h, C.byref(hnd)) == 0
    try:
        t = (C.c_float * 6)(*ts) if ts is not None else None
        assert lib.rb_batch_draw_paths(hnd, scene["n_paths"], scene["verb_off"].ctypes.data, scene["pt_off"].ctypes.data,
                                       scene["verbs"].ctypes.data, scene["pts"].ctypes.data, C.addressof(paints),
                                       scene["rules"].ctypes.data, C.addressof(st) if st is not None else None, t) == 0
        out = (C.c_uint64 * 8)()
        assert lib.rb_debug_geo_host_stats(hnd, out) == 0
        return dict(zip(["tasks", "dashed", "stroked", "hair", "fill", "bytes", "verbs", "pts"], [int(v) for v in out]))
    finally:
        lib.rb_batch_destroy(hnd)


def test_every_draw_becomes_a_task_and_strokes_are_classified():
    w, h = 640, 480
    scene = scenes.paths_scene(w, h, 900, 0xC2, rmin=6.0, rmax=120.0)
    s = _stats(w, h, scene)
    n_dashed = int(((scene["n_dash"] > 0) & (scene["stroke_width"] > 0)).sum())
    n_strokes = int((scene["stroke_width"] > 0).sum())
    assert s["dashed"] == n_dashed  # every dashed stroke is on the dash (or dash-unit) list
    # hairlines without dashes are split into one task per drawing verb: more tasks than draws
    assert s["tasks"] > scene["n_paths"] - 50 and s["hair"] > 0
    assert s["stroked"] + s["dashed"] <= n_strokes
    # every task that is filled: fills + stroke outlines (dashed outlines are filled by the unit kernels)
    assert s["fill"] >= scene["n_paths"] - n_strokes
    # raw path data is uploaded once per draw
    assert s["verbs"] <= int(scene["verb_off"][-1]) and s["pts"] <= int(scene["pt_off"][-1])


def test_draws_outside_the_target_are_culled():
    w, h = 256, 256
    scene = scenes.paths_scene(4096, 4096, 2000, 7, rmin=8.0, rmax=64.0, strokes=False)
    s = _stats(w, h, scene, strokes=False)
    inside = 0
    for i in range(scene["n_paths"]):
        p = scene["pts"][scene["pt_off"][i]:scene["pt_off"][i + 1]]
        if p[:, 0].max() >= -2 and p[:, 1].max() >= -2 and p[:, 0].min() <= w + 2 and p[:, 1].min() <= h + 2:
            inside += 1
    assert s["tasks"] == inside and 0 < inside < 200


def test_draw_tiler_tiles_get_their_own_tasks():
    """A canvas wider than 8191 px: a draw crossing the tile boundary is one task per tile, all others one."""
    w, h = 8300, 128
    scene = scenes.paths_scene(w, h, 400, 0x7117, rmin=8.0, rmax=100.0, strokes=False)
    s = _stats(w, h, scene, strokes=False)
    crossing = 0
    for i in range(scene["n_paths"]):
        p = scene["pts"][scene["pt_off"][i]:scene["pt_off"][i + 1]]
        if p[:, 0].min() <= 8191 + 2 and p[:, 0].max() >= 8191 - 2:
            crossing += 1
    assert s["tasks"] == scene["n_paths"] + crossing


def test_upload_is_a_fraction_of_the_host_builders():
    """The point of the device path: the raw paths are uploaded, not the edges."""
    W, H = 2048, 2048
    scene = scenes.paths_scene(W, H, 6000, 0x5EED0002)
    s = _stats(W, H, scene)
    lib = _ffi.lib
    paints = scenes.to_paint_array(scene, _ffi.Paint)
    st = scenes.to_stroke_array(scene, _ffi.Stroke)
    hnd = C.c_void_p()
    assert lib.rb_debug_batch_begin_host(W, H, C.byref(hnd)) == 0
    assert lib.rb_batch_draw_paths(hnd, scene["n_paths"], scene["verb_off"].ctypes.data, scene["pt_off"].ctypes.data,
                                   scene["verbs"].ctypes.data, scene["pts"].ctypes.data, C.addressof(paints),
                                   scene["rules"].ctypes.data, C.addressof(st), None) == 0
    assert lib.rb_batch_prepare(hnd, 0) == 0
    stats = (C.c_uint64 * 6)()
    lib.rb_batch_stats(hnd, stats)
    lib.rb_batch_destroy(hnd)
    assert s["bytes"] * 3 < int(stats[4])


def test_libm_compat_matches_the_hosts_libm(tmp_path):
    """libm_compat.h (acosf / cosf / cbrtf as glibc computes them, for the device) against the host's libm: every sampled
    argument gives the identical float (tools/libm_compat_check.cpp: 22 M + 40 M + 80 M arguments)."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "lmc")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", os.path.join(root, "tools", "libm_compat_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count(" 0 mismatches") == 3, out.stdout


def test_dashed_stroke_built_dash_by_dash_equals_dash_then_stroke():
    """The decomposition the geometry kernels use for dashed strokes (one thread per dash: cut it out, stroke it alone, a
    non-last dash followed by the next one's move_to) against dash-then-stroke of the whole path, over random open and
    closed paths, every cap and join: identical verbs and bit-identical points."""
    import resvg_b200 as rb
    from resvg_b200 import api
    from tests.pathgen import SplitMix64, random_path

    rng = SplitMix64(0xDA5E)
    caps, joins = ["butt", "round", "square"], ["miter", "miter-clip", "round", "bevel"]
    checked = 0
    for it in range(400):
        cx, cy, r = rng.uniform(0, 300), rng.uniform(0, 300), rng.log_uniform(8, 150)
        verbs, pts = random_path(rng, cx, cy, r)
        verbs = list(verbs)
        pts = [tuple(p) for p in pts]
        if it % 3 == 0 and verbs and verbs[-1] == 4:  # an open contour as well
            verbs = verbs[:-1]
        if it % 5 == 0:  # a second contour
            verbs += [0, 1, 2]
            pts += [(cx + 3.0, cy - 7.5), (cx + r, cy + 2.0), (cx + 0.5 * r, cy + r), (cx - r, cy + 0.25 * r)]
        n = 2 * int(rng.uniform(1, 3.99))
        dash = [rng.uniform(0.5, 30.0) for _ in range(n)]
        off = rng.uniform(-40.0, 80.0)
        width = rng.log_uniform(1.2, 20.0)
        cap, join = caps[it % 3], joins[(it // 3) % 4]
        res = rng.uniform(0.5, 3.0)
        d = rb.dash_path(verbs, pts, dash, off, res)
        want = rb.stroke_path(d[0], d[1], width, 4.0, cap, join, res) if d is not None else None
        got = api.stroke_dashed_in_units(verbs, pts, dash, off, width, 4.0, cap, join, res)
        if want is None:
            assert got is None
            continue
        assert got is not None, (it, cap, join)
        assert np.array_equal(got[0], want[0]), (it, cap, join, len(got[0]), len(want[0]))
        assert got[1].tobytes() == want[1].tobytes(), (it, cap, join)
        checked += 1
    assert checked > 350
