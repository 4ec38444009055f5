"""Device path geometry (resvg_b200/csrc/geo.cu: dash, stroke, hairline walk, fill front end as CUDA kernels) against the
host builder (the same cores compiled by g++, pinned against the independent checker by the CPU suite) and against the
checker itself.  Bit-exact: both sides run the same source without FMA contraction."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_raster as R
from tests.pathgen import SplitMix64, random_paint_spec, random_path
from tests.util import assert_exact, assert_within, random_premul

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import resvg_b200 as rb

    c = rb.Context(0)
    yield c
    c.close()


def _counts():
    from resvg_b200 import _ffi

    out = (C.c_uint64 * 6)()
    _ffi.lib.rb_debug_geo_counts(out)
    return [int(v) for v in out][:3]


def _render_scene(ctx, scene, w, h, mode, base=None, ts=None):
    import resvg_b200 as rb
    from resvg_b200 import _ffi

    _ffi.lib.rb_debug_geo_mode(mode)
    try:
        before = _counts()
        l = ctx.layer_from(base) if base is not None else ctx.layer(w, h)
        b = rb.Batch(l)
        if ts is None:
            b.fill_paths(scene)
        else:
            b.fill_paths(scene, ts)
        b.submit()
        got = l.download()
        after = _counts()
    finally:
        _ffi.lib.rb_debug_geo_mode(0)
    return got, [a - c for a, c in zip(after, before)]


def _scene(w, h, n, seed, **kw):
    from resvg_b200 import _ffi, scenes

    scene = scenes.paths_scene(w, h, n, seed, **kw)
    scene["paints"] = scenes.to_paint_array(scene, _ffi.Paint)
    scene["strokes"] = scenes.to_stroke_array(scene, _ffi.Stroke)
    return scene


@pytest.mark.parametrize("seed,w,h,n", [(0xC2, 640, 480, 900), (0xD1CE, 1500, 1100, 4000)])
def test_bench_scene_device_geometry_equals_host_builder(ctx, seed, w, h, n):
    """The C2 scene (fills, strokes with every cap / join, dashes, hairlines, gradients) built by the geometry kernels
    gives the pixels of the host builder, bit for bit, and the device path really ran."""
    scene = _scene(w, h, n, seed, rmin=6.0, rmax=120.0)
    assert (scene["n_dash"] > 0).sum() > 5 and (scene["stroke_width"] > 0).sum() > 100
    host, hc = _render_scene(ctx, scene, w, h, 2)
    dev, dc = _render_scene(ctx, scene, w, h, 1)
    assert hc[0] == 0 and dc[0] >= 1 and dc[1] == 0, (hc, dc)
    assert_exact(dev, host, "device geometry vs host builder")


@pytest.mark.parametrize("seed_offset", [0, 3])
def test_full_size_scene_device_geometry_equals_host_builder(ctx, seed_offset):
    """(seed + 3: a scene in which combine_vertical merges edges of two neighbouring dash outlines — that draw's edges are
    rebuilt by one thread.)  BASELINE.json's C2 scene at full size (8192 x 8192, 100 000 paths -> 120 087 draws, DrawTiler tiles, dashed strokes in
    units, hairlines verb by verb): every pixel of the device build equals the host build, also when the two share the batch
    (the default for large batches: the host threads build the first draws while the geometry kernels build the rest)."""
    import zlib

    import resvg_b200 as rb
    from resvg_b200 import _ffi

    w = h = 8192
    scene = _scene(w, h, 100_000, 0x5EED0002 + seed_offset)
    l = ctx.layer(w, h)
    pinned = rb.PinnedBuffer(w * h * 4)
    crcs = {}
    for name, mode in (("host", 2), ("device", 1), ("shared", 0)):
        _ffi.lib.rb_debug_geo_mode(mode)
        try:
            before = _counts()
            l.fill(0, 0, 0, 0)
            b = rb.Batch(l)
            b.fill_paths(scene)
            b.submit()
            b.close()
            l.download_ptr(pinned.array.ctypes.data)
            after = _counts()
        finally:
            _ffi.lib.rb_debug_geo_mode(0)
        crcs[name] = zlib.crc32(pinned.array)
        if name == "device":
            assert after[0] - before[0] >= 1 and after[1] == before[1], (before, after)
    pinned.close()
    l.close()
    assert crcs["device"] == crcs["host"] and crcs["shared"] == crcs["host"], crcs


def test_bench_scene_device_geometry_matches_checker(ctx):
    import bench

    w, h = 640, 480
    scene = _scene(w, h, 700, 0xC2, rmin=6.0, rmax=90.0)
    got, dc = _render_scene(ctx, scene, w, h, 1)
    assert dc[0] >= 1 and dc[1] == 0
    want = np.zeros((h, w, 4), np.uint8)
    bench.cpu_render_sample(bench.oracle_lib(), scene, scene["n_paths"], want)
    assert_within(got, want, 1, "bench scene, device geometry")  # two-point conical gradients run the f32 pipeline
    assert (got != want).any(axis=-1).mean() < 0.02


@pytest.mark.parametrize("ts", [(1.0, 0.0, 0.0, 1.0, 13.25, -7.5), (0.75, 0.0, 0.0, 1.5, 20.0, 10.0), (0.9, 0.25, -0.3, 0.8, 100.0, 40.0)])
def test_transformed_bulk_draws(ctx, ts):
    """path.transform(ts) of bulk fills and the stroke transform (translate / scale / affine map_points forms) on the device."""
    w, h = 700, 560
    scene = _scene(w, h, 600, 0xABCD, rmin=6.0, rmax=100.0)
    host, _ = _render_scene(ctx, scene, w, h, 2, ts=ts)
    dev, dc = _render_scene(ctx, scene, w, h, 1, ts=ts)
    assert dc[0] >= 1 and dc[1] == 0
    assert_exact(dev, host, f"ts={ts}")


def test_draw_tiler_canvas_wider_than_8191(ctx):
    """Canvases beyond 8191 px are cut into DrawTiler tiles: one geometry task per (draw, tile)."""
    w, h = 8300, 200
    scene = _scene(w, h, 500, 0x7117, rmin=8.0, rmax=150.0)
    host, _ = _render_scene(ctx, scene, w, h, 2)
    dev, dc = _render_scene(ctx, scene, w, h, 1)
    assert dc[0] >= 1 and dc[1] == 0
    assert_exact(dev, host, "draw tiler")
    assert dev[:, 8191:].any(), "something must land in the second tile"


def test_small_heap_is_retried(ctx, monkeypatch):
    """A heap hint that is far too small: the launch is repeated with a larger heap and the result is the same."""
    w, h = 640, 480
    scene = _scene(w, h, 900, 0xC2, rmin=6.0, rmax=120.0)
    host, _ = _render_scene(ctx, scene, w, h, 2)
    monkeypatch.setenv("RB_GEO_HEAP_BYTES", "200000")
    dev, dc = _render_scene(ctx, scene, w, h, 1)
    assert dc[0] >= 1 and dc[1] == 0 and dc[2] >= 1, dc
    assert_exact(dev, host, "after a heap retry")


def test_recorded_draws_viewports_and_strokes(ctx):
    """Individually recorded fills / strokes / hairlines / dashed strokes in atlas viewports (some reaching beyond the target),
    over an existing background: device geometry == host builder."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi

    cw, ch, gap, nx, ny = 61, 47, 3, 7, 5
    W, H = gap + nx * (cw + gap), gap + ny * (ch + gap)
    base = random_premul(W, H, 9)
    out = []
    for mode in (2, 1):
        rng = SplitMix64(2025)
        _ffi.lib.rb_debug_geo_mode(mode)
        try:
            l = ctx.layer_from(base)
            b = rb.Batch(l)
            for j in range(ny + 1):
                for i in range(nx + 1):
                    x0, y0 = gap + i * (cw + gap) - 20, gap + j * (ch + gap) - 15  # the last row / column leave the target
                    b.set_viewport(x0, y0, cw, ch)
                    for k in range(6):
                        cx, cy, r = rng.uniform(-10, cw + 10), rng.uniform(-10, ch + 10), rng.log_uniform(5, 70)
                        verbs, pts = random_path(rng, cx, cy, r)
                        spec = random_paint_spec(rng, cx, cy, r, solid=0.6, linear=0.4)
                        if k == 3:
                            b.stroke_path(verbs, pts, rb.make_paint(spec), rng.uniform(0.1, 0.8), 4.0, "square", "miter")
                        elif k == 4:
                            b.stroke_path(verbs, pts, rb.make_paint(spec), 3.5, 4.0, "round", "round")
                        elif k == 5:
                            b.stroke_path(verbs, pts, rb.make_paint(spec), rng.uniform(0.3, 6.0), 2.0, "butt", "bevel",
                                          dash=[rng.uniform(1, 9), rng.uniform(1, 5)], dash_offset=rng.uniform(-3, 12))
                        else:
                            b.fill_path(verbs, pts, rb.make_paint(spec, "source_over", k != 2), "evenodd" if k % 2 else "nonzero")
            b.set_viewport()
            b.fill_path([0, 1, 1, 4], [(5.0, 5.0), (W - 9.5, 20.25), (40.0, H - 7.75)], rb.make_paint({"kind": "solid", "color": (0.1, 0.9, 0.4, 0.35)}))
            before = _counts()
            b.submit()
            out.append(l.download())
            after = _counts()
            if mode == 1:
                assert after[0] - before[0] >= 1 and after[1] == before[1]
        finally:
            _ffi.lib.rb_debug_geo_mode(0)
    assert_exact(out[1], out[0], "recorded draws in viewports")


def test_many_overlapping_loops_fall_back_to_the_host_builder(ctx):
    """A path winding 140 times around one point exceeds the packed winding range: the geometry kernels detect it exactly
    (k_geo_wide) and hand the batch to the host builder, which routes it to the any-winding kernel."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi

    w, h = 96, 96
    verbs, pts = [], []
    for k in range(140):
        d = 0.05 * k
        verbs += [0, 1, 1, 1, 4]
        pts += [(10 + d, 10 + d), (86 - d, 12 + d), (84 - d, 86 - d), (12 + d, 84 - d)]
    spec = {"kind": "solid", "color": (0.2, 0.4, 0.9, 0.7)}
    want = np.zeros((h, w, 4), np.uint8)
    R.fill_path(want, verbs, pts, R.make_paint(spec), "nonzero")
    _ffi.lib.rb_debug_geo_mode(1)
    try:
        before = _counts()
        l = ctx.layer(w, h)
        b = rb.Batch(l)
        b.fill_path(verbs, pts, rb.make_paint(spec), "nonzero")
        b.submit()
        got = l.download()
        after = _counts()
    finally:
        _ffi.lib.rb_debug_geo_mode(0)
    assert after[1] - before[1] == 1, (before, after)
    assert_exact(got, want, "wide draw through the fallback")


@pytest.mark.parametrize("mode,w,h,n", [(0, 2048, 1536, 40000), (1, 1500, 1100, 6000), (2, 1500, 1100, 40000), (0, 300, 200, 20), (0, 640, 483, 900)])
def test_submit_download_equals_submit_then_download(ctx, mode, w, h, n):
    """rb_batch_submit_download (the last raster launch in bands of tile rows, every band copied out while the next one is
    rendered) against rb_batch_submit + rb_layer_download: the same bytes — on the hybrid path (host parts, then the device
    range), on the device-only and host-only paths, on a batch too small to be binned and on a height that is no multiple
    of the tile height; the canvas starts from random pixels so that an untouched band is noticed too."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi

    scene = _scene(w, h, n, 0xBA2D + mode)
    base = random_premul(w, h, 5)
    want, _ = _render_scene(ctx, scene, w, h, mode, base=base)
    _ffi.lib.rb_debug_geo_mode(mode)
    try:
        l = ctx.layer_from(base)
        b = rb.Batch(l)
        b.fill_paths(scene)
        pinned = rb.PinnedBuffer(w * h * 4)
        pinned.array[:] = 0x5A
        banded0 = int(_ffi.lib.rb_debug_banded_downloads())
        b.submit_download(pinned.array.ctypes.data)
        l.download_end()
        assert int(_ffi.lib.rb_debug_banded_downloads()) - banded0 == (1 if n > 32 else 0)  # <= 32 draws: no bins, plain download
        got = np.array(pinned.array, copy=True).reshape(h, w, 4)
        after = l.download()
    finally:
        _ffi.lib.rb_debug_geo_mode(0)
    assert_exact(got, want, "submit_download vs submit + download")
    assert_exact(after, want, "the layer after submit_download")
