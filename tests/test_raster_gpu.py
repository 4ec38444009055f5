"""GPU parity of the rasteriser: CUDA tile kernel vs the CPU oracle (oracle/raster.c) on the same seeded paths.

Bars: coverage, the u16 "lowp" pipeline, masks and the non-AA path are bit-exact; the f32 "highp" pipeline
(two-point-conical gradients, highp-only blend modes, layer composites, pattern sampling) is within 1/255.
"""
import math

import numpy as np
import pytest

from tests import oracle_raster as R
from tests.pathgen import SplitMix64, random_paint_spec, random_path, random_stops
from tests.util import assert_exact, assert_within, random_premul

pytestmark = pytest.mark.gpu


def _gpu_fill(ctx, base, verbs, pts, spec, rule="nonzero", ts=R.IDENTITY, blend="source_over", aa=True):
    import resvg_b200 as rb

    l = ctx.layer_from(base)
    rb.fill_path(l, verbs, pts, rb.make_paint(spec, blend, aa), rule, ts)
    return l.download()


def _cpu_fill(base, verbs, pts, spec, rule="nonzero", ts=R.IDENTITY, blend="source_over", aa=True):
    px = base.copy()
    R.fill_path(px, verbs, pts, R.make_paint(spec, blend, aa), rule, ts)
    return px


SOLID = {"kind": "solid", "color": (0.2, 0.6, 0.9, 0.7)}
OPAQUE = {"kind": "solid", "color": (1.0, 0.5, 0.0, 1.0)}


@pytest.mark.parametrize("w,h", [(64, 16), (300, 300), (257, 131), (70, 500), (1, 1), (5, 3)])
@pytest.mark.parametrize("rule", ["nonzero", "evenodd"])
@pytest.mark.parametrize("aa", [True, False])
def test_coverage_random_paths(ctx, w, h, rule, aa):
    """Coverage through an opaque solid on a transparent layer: alpha == coverage, so this pins scan conversion."""
    rng = SplitMix64(1000 + w * 7 + h)
    base = np.zeros((h, w, 4), np.uint8)
    for i in range(12):
        r = rng.log_uniform(2, max(w, h))
        verbs, pts = random_path(rng, rng.uniform(0, w), rng.uniform(0, h), r)
        got = _gpu_fill(ctx, base, verbs, pts, OPAQUE, rule, aa=aa)
        want = _cpu_fill(base, verbs, pts, OPAQUE, rule, aa=aa)
        assert_exact(got, want, f"path {i} {w}x{h} {rule} aa={aa}")


def test_coverage_structured_shapes(ctx):
    """Axis-aligned rectangles, shared vertical edges, thin slivers and self-retracing contours: the cases where
    spans abut inside one pixel (the 63-vs-64 rule) and where edges tie."""
    w, h = 200, 120
    base = np.zeros((h, w, 4), np.uint8)
    M, L, Z = 0, 1, 4

    def rect(x, y, rw, rh, ccw=False):
        p = [(x, y), (x + rw, y), (x + rw, y + rh), (x, y + rh)]
        if ccw:
            p = p[::-1]
        return [M, L, L, L, Z], p

    shapes = []
    for (x, y, rw, rh) in [(10.25, 10.25, 50.5, 30.5), (0.75, 0.75, 198.5, 118.5), (20, 20, 0.1, 60), (30.5, 7.3, 90.2, 0.2)]:
        shapes.append(rect(x, y, rw, rh))
    # two rectangles sharing an edge inside a pixel, same and opposite orientation
    for ccw in (False, True):
        v1, p1 = rect(10.3, 40.6, 40.33, 30.1)
        v2, p2 = rect(50.63, 35.2, 33.3, 50.7, ccw)
        shapes.append((v1 + v2, p1 + p2))
    # stroke-like outline: outer rect + inner rect reversed + excursions through the corner
    v1, p1 = rect(100.75, 20.75, 80.5, 60.5)
    v2, p2 = rect(102.25, 22.25, 77.5, 57.5, True)
    shapes.append((v1 + v2, p1 + p2))
    # retraced spike (zero-area excursion) inside a filled region
    shapes.append(([M, L, L, L, L, L, L, Z], [(20.2, 80.1), (90.7, 80.1), (90.7, 110.9), (55.3, 110.9), (55.3, 85.0), (55.3, 110.9), (20.2, 110.9)]))
    for i, (v, p) in enumerate(shapes):
        for rule in ("nonzero", "evenodd"):
            for ts in (R.IDENTITY, (1.5, 0, 0, 1.5, 0.37, 0.21), (0.8, 0.3, -0.2, 1.1, 20, 5)):
                got = _gpu_fill(ctx, base, v, p, OPAQUE, rule, ts)
                want = _cpu_fill(base, v, p, OPAQUE, rule, ts)
                assert_exact(got, want, f"shape {i} {rule} {ts}")


def test_paths_leaving_the_canvas_are_clipped_identically(ctx):
    w, h = 160, 96
    rng = SplitMix64(77)
    base = random_premul(w, h, 5)
    for i in range(40):
        cx, cy = rng.uniform(-40, w + 40), rng.uniform(-40, h + 40)
        verbs, pts = random_path(rng, cx, cy, rng.log_uniform(20, 300))
        rule = "evenodd" if i % 2 else "nonzero"
        got = _gpu_fill(ctx, base, verbs, pts, SOLID, rule, aa=(i % 5 != 0))
        want = _cpu_fill(base, verbs, pts, SOLID, rule, aa=(i % 5 != 0))
        assert_exact(got, want, f"clipped path {i}")


LOWP_BLENDS = ["clear", "source", "source_over", "destination_over", "source_in", "destination_in", "source_out",
               "destination_out", "source_atop", "destination_atop", "xor", "plus", "modulate", "screen", "overlay",
               "darken", "lighten", "hard_light", "difference", "exclusion", "multiply"]
HIGHP_BLENDS = ["color_dodge", "color_burn", "soft_light", "hue", "saturation", "color", "luminosity"]


@pytest.mark.parametrize("blend", LOWP_BLENDS)
def test_lowp_blend_modes_exact(ctx, blend):
    w, h = 96, 64
    rng = SplitMix64(31)
    base = random_premul(w, h, 6)
    for i, spec in enumerate([SOLID, OPAQUE, {"kind": "linear", "x0": 5, "y0": 5, "x1": 90, "y1": 50,
                                              "stops": random_stops(rng, 4), "spread": "reflect"}]):
        verbs, pts = random_path(rng, 48, 32, 40)
        got = _gpu_fill(ctx, base, verbs, pts, spec, blend=blend)
        want = _cpu_fill(base, verbs, pts, spec, blend=blend)
        assert_exact(got, want, f"{blend} paint {i}")


@pytest.mark.parametrize("blend", HIGHP_BLENDS)
def test_highp_blend_modes_within_one(ctx, blend):
    w, h = 96, 64
    rng = SplitMix64(32)
    base = random_premul(w, h, 7)
    verbs, pts = random_path(rng, 48, 32, 40)
    for spec in (SOLID, OPAQUE):
        got = _gpu_fill(ctx, base, verbs, pts, spec, blend=blend)
        want = _cpu_fill(base, verbs, pts, spec, blend=blend)
        assert_within(got, want, 1, blend)  # f32 pipeline: tolerance 1/255 (north star)


def test_gradients(ctx):
    w, h = 200, 150
    rng = SplitMix64(33)
    base = random_premul(w, h, 8)
    for i in range(40):
        cx, cy, r = rng.uniform(20, 180), rng.uniform(20, 130), rng.log_uniform(10, 120)
        verbs, pts = random_path(rng, cx, cy, r)
        spec = random_paint_spec(rng, cx, cy, r, solid=0.0, linear=0.5)
        if i % 3 == 0:
            spec["ts"] = (0.9, 0.2, -0.3, 1.2, 3.0, -2.0)
        if i % 4 == 0:
            for s in spec["stops"]:
                s[4] = 1.0
        ts = (1.0, 0, 0, 1.0, 0, 0) if i % 2 else (1.3, 0.1, -0.1, 0.9, 4, 2)
        got = _gpu_fill(ctx, base, verbs, pts, spec, ts=ts)
        want = _cpu_fill(base, verbs, pts, spec, ts=ts)
        if spec["kind"] == "linear" or (abs(spec["x0"] - spec["x1"]) < 1e-9 and abs(spec["y0"] - spec["y1"]) < 1e-9):
            assert_exact(got, want, f"gradient {i} {spec['kind']}")  # lowp pipeline
        else:
            assert_within(got, want, 1, f"gradient {i} two-point conical")  # highp f32


def test_simple_radial_and_focal_variants(ctx):
    w, h = 128, 128
    base = np.zeros((h, w, 4), np.uint8)
    verbs, pts = [0, 1, 1, 1, 4], [(4, 4), (124, 4), (124, 124), (4, 124)]
    stops = [[0, 1, 1, 1, 1], [0.5, 0.2, 0.8, 0.3, 0.6], [1, 0, 0, 0, 1]]
    cases = [
        dict(x0=64, y0=64, r0=0, x1=64, y1=64, r1=50),      # simple radial: lowp, exact
        dict(x0=64, y0=64, r0=10, x1=64, y1=64, r1=50),     # concentric with fr
        dict(x0=50, y0=60, r0=0, x1=64, y1=64, r1=50),      # focal inside
        dict(x0=64, y0=14, r0=0, x1=64, y1=64, r1=50),      # focal on circle
        dict(x0=20, y0=20, r0=5, x1=90, y1=90, r1=30),      # general two point
        dict(x0=20, y0=64, r0=20, x1=100, y1=64, r1=20),    # strip
    ]
    for i, c in enumerate(cases):
        for spread in ("pad", "reflect", "repeat"):
            spec = dict(kind="radial", stops=stops, spread=spread, **c)
            got = _gpu_fill(ctx, base, verbs, pts, spec)
            want = _cpu_fill(base, verbs, pts, spec)
            if i == 0:
                assert_exact(got, want, f"radial {i} {spread}")
            else:
                assert_within(got, want, 1, f"radial {i} {spread}")


def test_batch_painters_order_matches_sequential_oracle(ctx):
    """200 overlapping translucent paths in one batch (one kernel launch) == 200 sequential oracle fills."""
    import resvg_b200 as rb

    w, h = 333, 217
    rng = SplitMix64(34)
    want = np.zeros((h, w, 4), np.uint8)
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    blends = ["source_over", "source_over", "source_over", "multiply", "screen", "xor", "plus", "darken"]
    for i in range(200):
        cx, cy, r = rng.uniform(0, w), rng.uniform(0, h), rng.log_uniform(4, 150)
        verbs, pts = random_path(rng, cx, cy, r)
        spec = random_paint_spec(rng, cx, cy, r, solid=0.6, linear=0.4)
        rule = "evenodd" if rng.u() < 0.5 else "nonzero"
        aa = rng.u() < 0.9
        blend = blends[rng.randint(0, len(blends) - 1)]
        b.fill_path(verbs, pts, rb.make_paint(spec, blend, aa), rule)
        R.fill_path(want, verbs, pts, R.make_paint(spec, blend, aa), rule)
    b.submit()
    st = b.stats()
    assert st["draws"] >= 190 and st["edges"] > 1000
    assert_exact(l.download(), want, "batch of 200")
    assert ctx.launch_count > 0


def test_large_canvas_draw_tiler_split(ctx):
    """A canvas wider than 8191 px is drawn as DrawTiler tiles (8191 + rest); paths across the seam must match."""
    import resvg_b200 as rb

    w, h = 8300, 40
    rng = SplitMix64(35)
    want = np.zeros((h, w, 4), np.uint8)
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    for i in range(30):
        cx = rng.uniform(8100, 8290) if i % 2 else rng.uniform(0, w)
        verbs, pts = random_path(rng, cx, rng.uniform(0, h), rng.log_uniform(5, 120))
        spec = random_paint_spec(rng, cx, 20, 60, solid=0.7, linear=0.3)
        b.fill_path(verbs, pts, rb.make_paint(spec), "nonzero")
        R.fill_path(want, verbs, pts, R.make_paint(spec), "nonzero")
    b.submit()
    assert_exact(l.download(), want, "draw tiler seam")


@pytest.mark.parametrize("blend", LOWP_BLENDS + HIGHP_BLENDS)
def test_draw_layer_all_blend_modes(ctx, blend):
    import resvg_b200 as rb

    dst, src = random_premul(150, 90, 9, sparse=True), random_premul(100, 70, 10, sparse=True)
    for (x, y, op) in [(0, 0, 1.0), (20, 10, 0.5), (-30, -20, 0.85), (100, 60, 1.0)]:
        ld, ls = ctx.layer_from(dst), ctx.layer_from(src)
        rb.draw_layer(ld, ls, x, y, op, blend)
        want = dst.copy()
        R.draw_pixmap(want, x, y, src, op, blend)
        # layer composites run the f32 pipeline; identical operation order gives identical bytes except where
        # division/sqrt sequences differ (dodge/burn/soft-light/non-separable): tolerance 1/255 there
        if blend in HIGHP_BLENDS:
            assert_within(ld.download(), want, 1, f"draw_layer {blend} {x},{y},{op}")
        else:
            assert_exact(ld.download(), want, f"draw_layer {blend} {x},{y},{op}")


def test_masks(ctx):
    import resvg_b200 as rb

    w, h = 131, 77
    px = random_premul(w, h, 11, sparse=True)
    l = ctx.layer_from(px)
    for kind in ("alpha", "luminance"):
        m = rb.Mask.from_layer(l, kind)
        assert_exact(m.download(), R.mask_from_pixmap(px, kind), f"mask_from_pixmap {kind}")
    m = rb.Mask.from_layer(l, "luminance")
    m.invert()
    want_m = R.mask_from_pixmap(px, "luminance")
    R.mask_invert(want_m)
    assert_exact(m.download(), want_m, "invert")
    other = random_premul(w, h, 12)
    lo = ctx.layer_from(other)
    rb.apply_mask(lo, m)
    want = other.copy()
    R.apply_mask(want, want_m)
    assert_exact(lo.download(), want, "apply_mask")
    # Mask::fill_path
    rng = SplitMix64(36)
    gm = rb.Mask(ctx, w, h)
    cm = np.zeros((h, w), np.uint8)
    for i in range(6):
        verbs, pts = random_path(rng, rng.uniform(0, w), rng.uniform(0, h), rng.log_uniform(10, 80))
        gm.fill_path(verbs, pts, "evenodd" if i % 2 else "nonzero", i % 3 != 0, (1.2, 0, 0, 1.2, 1.5, 0.5))
        R.mask_fill_path(cm, verbs, pts, "evenodd" if i % 2 else "nonzero", i % 3 != 0, (1.2, 0, 0, 1.2, 1.5, 0.5))
    assert_exact(gm.download(), cm, "mask fill_path")


def test_pattern_fill(ctx):
    import resvg_b200 as rb

    w, h = 120, 90
    tile = random_premul(17, 13, 13)
    lt = ctx.layer_from(tile)
    base = random_premul(w, h, 14)
    rng = SplitMix64(37)
    verbs, pts = random_path(rng, 60, 45, 50)
    for quality in ("nearest", "bilinear", "bicubic"):
        for spread in ("repeat", "pad", "reflect"):
            for ts in ((1, 0, 0, 1, 3, 4), (1.7, 0.2, -0.1, 1.4, 5.5, 2.25)):
                gspec = dict(kind="pattern", layer=lt, spread=spread, quality=quality, opacity=0.8, ts=ts)
                cspec = dict(kind="pattern", pixmap=tile, spread=spread, quality=quality, opacity=0.8, ts=ts)
                l = ctx.layer_from(base)
                rb.fill_path(l, verbs, pts, rb.make_paint(gspec))
                want = base.copy()
                R.fill_path(want, verbs, pts, R.make_paint(cspec))
                assert_within(l.download(), want, 1, f"pattern {quality} {spread} {ts}")


def test_wide_fallback_kernel_matches_too(ctx):
    """The any-winding fallback kernel (selected by the host when |winding| could exceed 127) gives the same bytes."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi

    w, h = 300, 180
    rng = SplitMix64(41)
    want = np.zeros((h, w, 4), np.uint8)
    recs = []
    for i in range(60):
        cx, cy, r = rng.uniform(0, w), rng.uniform(0, h), rng.log_uniform(4, 150)
        verbs, pts = random_path(rng, cx, cy, r)
        spec = random_paint_spec(rng, cx, cy, r, solid=0.6, linear=0.4)
        rule = "evenodd" if rng.u() < 0.5 else "nonzero"
        recs.append((verbs, pts, spec, rule, rng.u() < 0.9))
        R.fill_path(want, verbs, pts, R.make_paint(spec, "source_over", recs[-1][4]), rule)
    _ffi.lib.rb_debug_force_wide_kernel(1)
    try:
        l = ctx.layer(w, h)
        b = rb.Batch(l)
        for verbs, pts, spec, rule, aa in recs:
            b.fill_path(verbs, pts, rb.make_paint(spec, "source_over", aa), rule)
        b.submit()
        got = l.download()
    finally:
        _ffi.lib.rb_debug_force_wide_kernel(0)
    assert_exact(got, want, "wide kernel")


@pytest.mark.parametrize("split", [False, True])
def test_bench_scene_bulk_api_matches_oracle(ctx, split, monkeypatch):
    """The C2 bench scene (fills, strokes, dashes, gradients) at a small size, recorded through the by-reference bulk
    API exactly as bench.py does — once as a single submit, once forced through the multi-part pipelined submit — must
    equal the oracle rendering of the same draws."""
    import bench
    import resvg_b200 as rb
    from resvg_b200 import _ffi, scenes

    if split:
        monkeypatch.setenv("RB_SUBMIT_SPLIT_FROM", "64")
        monkeypatch.setenv("RB_SUBMIT_PARTS", "5")
    w, h = 640, 480
    scene = scenes.paths_scene(w, h, 700, 0xC2, rmin=6.0, rmax=90.0)
    scene["paints"] = scenes.to_paint_array(scene, _ffi.Paint)
    scene["strokes"] = scenes.to_stroke_array(scene, _ffi.Stroke)
    assert (scene["n_dash"] > 0).sum() > 5 and (scene["stroke_width"] > 0).sum() > 100
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    b.fill_paths(scene)
    b.submit()
    got = l.download()
    want = np.zeros((h, w, 4), np.uint8)
    bench.cpu_render_sample(bench.oracle_lib(), scene, scene["n_paths"], want)
    assert_within(got, want, 1, "bench scene")  # two-point conical gradients run the f32 pipeline
    assert (got != want).any(axis=-1).mean() < 0.02


def test_full_size_scene_is_independent_of_how_the_batch_is_cut(ctx, monkeypatch):
    """BASELINE.json's full-size C2 scene (8192 x 8192, 120 k draws incl. strokes, dashes, hairlines): the image must not
    depend on how rb_batch_submit cuts the batch into pipelined parts, nor on a resident prepare + run, and a second
    run over the cleared layer must reproduce it bit for bit (the tile kernel has no order-dependent arithmetic)."""
    import zlib

    import resvg_b200 as rb
    from resvg_b200 import _ffi, scenes

    w = h = 8192
    scene = scenes.paths_scene(w, h, 100_000, 0x5EED0002)
    scene["paints"] = scenes.to_paint_array(scene, _ffi.Paint)
    scene["strokes"] = scenes.to_stroke_array(scene, _ffi.Stroke)
    l = ctx.layer(w, h)
    pinned = rb.PinnedBuffer(w * h * 4)

    def render(parts):
        l.fill(0, 0, 0, 0)
        b = rb.Batch(l)
        b.fill_paths(scene)
        if parts == 0:
            b.prepare()
            b.run()
        else:
            monkeypatch.setenv("RB_SUBMIT_PARTS", str(parts))
            b.submit()
        b.close()
        l.download_ptr(pinned.array.ctypes.data)
        return zlib.crc32(pinned.array), int(pinned.array[3::4][:: 4097].astype(np.int64).sum())

    ref = render(1)
    assert ref[1] > 0
    for parts in (3, 8, 0, 1):
        assert render(parts) == ref, f"parts={parts}"
    pinned.close()
    l.close()


@pytest.mark.parametrize("mode", ["inline", "wide-kernel"])
def test_hairline_strokes_in_painters_order(ctx, mode):
    """Strokes tiny-skia treats as hairlines (anti-aliased, transformed width <= 1 px) interleaved with ordinary fills,
    painter's order kept.  "inline": their blits are applied by the tile kernel as a draw kind of its own;
    "wide-kernel": the fallback kernel does not know them, so the batch is cut into fill runs and hairline runs.
    Checked against the oracle back end (same host walker, oracle blending), exactly."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi
    from tests.backends import OracleBackend

    w, h = 260, 200
    rng = SplitMix64(4711)
    ob = OracleBackend()
    want = np.zeros((h, w, 4), np.uint8)
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    caps = ["butt", "round", "square"]
    n_hair = 0
    for i in range(90):
        cx, cy, r = rng.uniform(-10, w + 10), rng.uniform(-10, h + 10), rng.log_uniform(6, 120)
        verbs, pts = random_path(rng, cx, cy, r)
        if i % 2:
            verbs = verbs[:-1]  # open contour
        spec = random_paint_spec(rng, cx, cy, r, solid=0.7, linear=0.3)
        if i % 3 == 0:  # an ordinary fill between the hairlines
            b.fill_path(verbs, pts, rb.make_paint(spec), "nonzero")
            R.fill_path(want, verbs, pts, R.make_paint(spec), "nonzero")
            continue
        ts = [(1.0, 0.0, 0.0, 1.0, 0.0, 0.0), (0.6, 0.1, -0.2, 0.7, 4.0, 3.0), (1.4, 0.0, 0.0, 0.5, -3.0, 9.0)][i % 3]
        width = rng.uniform(0.05, 0.6)
        cap = caps[i % 3]
        dash, off = ([7.0, 4.0], 2.0) if i % 5 == 0 else (None, 0.0)
        b.stroke_path(verbs, pts, rb.make_paint(spec), width, 4.0, cap, "miter", ts, dash, off)
        ob.stroke_hairline(want, verbs, pts, spec, ts, "source_over", width, cap, dash, off)
        n_hair += 1
    assert n_hair > 40
    _ffi.lib.rb_debug_force_wide_kernel(1 if mode == "wide-kernel" else 0)
    try:
        b.submit()
    finally:
        _ffi.lib.rb_debug_force_wide_kernel(0)
    assert_exact(l.download(), want, f"hairlines ({mode})")


@pytest.mark.parametrize("mode", ["items", "wide-kernel"])
def test_viewports_render_many_documents_into_one_atlas(ctx, mode):
    """rb_batch_set_viewport: 35 small 'documents' (fills, strokes, hairlines, gradients leaving their cell) rendered by ONE
    batch into the cells of an atlas layer == the oracle rendering every document into a pixmap of its own; the gutter
    between the cells stays untouched."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi
    from tests.backends import OracleBackend

    cw, ch, gap, nx, ny = 61, 47, 3, 7, 5
    W, H = gap + nx * (cw + gap), gap + ny * (ch + gap)
    rng = SplitMix64(2025)
    ob = OracleBackend()
    base = random_premul(W, H, 9)
    want = base.copy()
    l = ctx.layer_from(base)
    b = rb.Batch(l)
    for j in range(ny):
        for i in range(nx):
            x0, y0 = gap + i * (cw + gap), gap + j * (ch + gap)
            doc = np.ascontiguousarray(want[y0:y0 + ch, x0:x0 + cw])
            b.set_viewport(x0, y0, cw, ch)
            for k in range(5):
                cx, cy, r = rng.uniform(-10, cw + 10), rng.uniform(-10, ch + 10), rng.log_uniform(5, 70)
                verbs, pts = random_path(rng, cx, cy, r)
                spec = random_paint_spec(rng, cx, cy, r, solid=0.6, linear=0.4)
                if k == 3:
                    width = rng.uniform(0.1, 0.8)
                    b.stroke_path(verbs, pts, rb.make_paint(spec), width, 4.0, "square", "miter")
                    ob.stroke_hairline(doc, verbs, pts, spec, R.IDENTITY, "source_over", width, "square")
                elif k == 4:
                    b.stroke_path(verbs, pts, rb.make_paint(spec), 3.5, 4.0, "round", "round")
                    out = rb.stroke_path(verbs, pts, 3.5, 4.0, "round", "round", 1.0)
                    if out is not None:
                        R.fill_path(doc, out[0], out[1], R.make_paint(spec), "nonzero")
                else:
                    rule = "evenodd" if k % 2 else "nonzero"
                    b.fill_path(verbs, pts, rb.make_paint(spec), rule)
                    R.fill_path(doc, verbs, pts, R.make_paint(spec), rule)
            want[y0:y0 + ch, x0:x0 + cw] = doc
    b.set_viewport()  # back to the whole layer: one shape across everything
    big_v, big_p = [0, 1, 1, 4], [(5.0, 5.0), (W - 9.5, 20.25), (40.0, H - 7.75)]
    big = {"kind": "solid", "color": (0.1, 0.9, 0.4, 0.35)}
    b.fill_path(big_v, big_p, rb.make_paint(big), "nonzero")
    R.fill_path(want, big_v, big_p, R.make_paint(big), "nonzero")
    _ffi.lib.rb_debug_force_wide_kernel(1 if mode == "wide-kernel" else 0)
    try:
        b.submit()
    finally:
        _ffi.lib.rb_debug_force_wide_kernel(0)
    assert_exact(l.download(), want, f"atlas ({mode})")


def test_batch_dashed_strokes(ctx):
    """stroke_path with a dash array: dash (tiny_skia_path::Path::dash) -> stroke -> fill inside the batch builder must
    equal the same host steps done one by one and filled by the oracle; rejected dash lists leave the stroke solid."""
    import resvg_b200 as rb

    w, h = 300, 220
    rng = SplitMix64(77)
    want = np.zeros((h, w, 4), np.uint8)
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    dashes = [([6.0, 3.0], 0.0), ([10.0, 2.0, 1.0, 2.0], 4.5), ([3.0, 3.0], -7.0), ([0.0, 5.0], 0.0), ([4.0], 0.0),
              ([5.0, -1.0], 0.0), ([0.0, 0.0], 0.0), ([12.5, 7.25], 100.0)]
    caps = ["butt", "round", "square"]
    for i in range(48):
        cx, cy, r = rng.uniform(20, w - 20), rng.uniform(20, h - 20), rng.log_uniform(10, 90)
        verbs, pts = random_path(rng, cx, cy, r)
        if i % 2 == 0:
            verbs = verbs[:-1]
        spec = random_paint_spec(rng, cx, cy, r, solid=1.0, linear=0.0)
        dash, off = dashes[i % len(dashes)]
        width, cap = rng.log_uniform(1.5, 8), caps[i % 3]
        b.stroke_path(verbs, pts, rb.make_paint(spec), width, 4.0, cap, "round", dash=dash, dash_offset=off)
        src = (np.asarray(verbs, np.uint8), np.asarray(pts, np.float32))
        valid = len(dash) >= 2 and len(dash) % 2 == 0 and min(dash) >= 0 and sum(dash) > 0
        if valid:
            src = rb.dash_path(verbs, pts, dash, off, 1.0)
            if src is None:
                continue
        out = rb.stroke_path(src[0], src[1], width, 4.0, cap, "round", 1.0)
        if out is not None:
            R.fill_path(want, out[0], out[1], R.make_paint(spec), "nonzero")
    b.submit()
    assert_exact(l.download(), want, "dashed strokes")


def test_device_curve_expansion_matches_host_expansion(ctx):
    """Curves are forward-differenced on the device by default; expanding them on the host (the fallback builder) must
    give the same pixels, and both must equal the oracle."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi

    w, h = 300, 260
    rng = SplitMix64(991)
    want = np.zeros((h, w, 4), np.uint8)
    recs = []
    for _ in range(60):
        cx, cy, r = rng.uniform(-20, w + 20), rng.uniform(-20, h + 20), rng.log_uniform(5, 200)
        verbs, pts = random_path(rng, cx, cy, r)
        spec = random_paint_spec(rng, cx, cy, r, solid=0.7, linear=0.3)
        rule = "evenodd" if rng.u() < 0.5 else "nonzero"
        recs.append((verbs, pts, spec, rule, rng.u() < 0.9))
        R.fill_path(want, verbs, pts, R.make_paint(spec, "source_over", recs[-1][4]), rule)
    got = []
    for host_expand in (0, 1):
        _ffi.lib.rb_debug_host_expand(host_expand)
        try:
            l = ctx.layer(w, h)
            b = rb.Batch(l)
            for verbs, pts, spec, rule, aa in recs:
                b.fill_path(verbs, pts, rb.make_paint(spec, "source_over", aa), rule)
            b.submit()
            got.append(l.download())
        finally:
            _ffi.lib.rb_debug_host_expand(0)
    assert_exact(got[0], want, "device expansion")
    assert_exact(got[1], want, "host expansion")


def test_many_overlapping_loops_select_wide_kernel(ctx):
    """A path winding around the same point 140 times exceeds the packed kernel's +-127 range: the host must route
    it to the fallback kernel and the result must still be exact."""
    import resvg_b200 as rb

    w, h = 96, 96
    verbs, pts = [], []
    for k in range(140):
        d = 0.05 * k
        verbs += [0, 1, 1, 1, 4]
        pts += [(10 + d, 10 + d), (80 - d, 12 + d), (78 - d, 82 - d), (12 + d, 80 - d)]
    spec = {"kind": "solid", "color": (0.1, 0.7, 0.3, 0.8)}
    l = ctx.layer(w, h)
    rb.fill_path(l, verbs, pts, rb.make_paint(spec), "nonzero")
    want = np.zeros((h, w, 4), np.uint8)
    R.fill_path(want, verbs, pts, R.make_paint(spec), "nonzero")
    assert_exact(l.download(), want, "140 nested loops")


def test_batch_stroke_path_matches_host_stroker_plus_oracle_fill(ctx):
    """PixmapMut::stroke_path through the batch API (host stroker inside prepare, device fill) == the same outline
    filled by the oracle."""
    import resvg_b200 as rb

    w, h = 320, 240
    rng = SplitMix64(51)
    want = np.zeros((h, w, 4), np.uint8)
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    caps, joins = ["butt", "round", "square"], ["miter", "round", "bevel", "miter-clip"]
    for i in range(60):
        cx, cy, r = rng.uniform(0, w), rng.uniform(0, h), rng.log_uniform(8, 120)
        verbs, pts = random_path(rng, cx, cy, r)
        if i % 3 == 0:
            verbs = verbs[:-1]  # open contour: caps
        spec = random_paint_spec(rng, cx, cy, r, solid=0.7, linear=0.3)
        width, cap, join = rng.log_uniform(1.2, 14), caps[i % 3], joins[i % 4]
        ts = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0) if i % 2 else (1.3, 0.2, -0.1, 0.8, 3.0, -2.0)
        b.stroke_path(verbs, pts, rb.make_paint(spec), width, 4.0, cap, join, ts)
        res = math.hypot(ts[0], ts[2]), math.hypot(ts[1], ts[3])
        out = rb.stroke_path(verbs, pts, width, 4.0, cap, join, max(res))
        if out is not None:
            R.fill_path(want, out[0], out[1], R.make_paint(spec), "nonzero", ts)
    b.submit()
    assert_exact(l.download(), want, "stroke batch")


def test_canvas_strips_equal_the_whole_canvas(ctx):
    """Canvas-strip sharding (shard.strip_for_rank / strip_viewport): the bench scene (fills, strokes, dashes, hairlines,
    gradients) rendered strip by strip, the document's pixmap placed above each strip layer with a negative viewport
    origin, is bit-identical to the whole-canvas render.  The DrawTiler-style alternative (translate the draws, let the
    strip be the pixmap) is NOT: tiny-skia clips curves to the pixmap before flattening them, so paths crossing a strip
    boundary change slightly — shown here so the difference stays documented."""
    import resvg_b200 as rb
    from resvg_b200 import _ffi, scenes, shard

    W, H = 1024, 1000
    scene = scenes.paths_scene(W, H, 1500, 77, rmin=8.0, rmax=160.0)
    scene["paints"] = scenes.to_paint_array(scene, _ffi.Paint)
    scene["strokes"] = scenes.to_stroke_array(scene, _ffi.Stroke)
    whole = ctx.layer(W, H)
    b = rb.Batch(whole)
    b.fill_paths(scene)
    b.submit()
    b.close()
    want = whole.download()
    for world in (2, 3):
        got, tiled = np.zeros_like(want), np.zeros_like(want)
        for r in range(world):
            y0, rows = shard.strip_for_rank(H, r, world)
            strip = ctx.layer(W, rows)
            b = rb.Batch(strip)
            b.set_viewport(*shard.strip_viewport(W, H, y0))
            b.fill_paths(scene)
            b.submit()
            b.close()
            got[y0:y0 + rows] = strip.download()
            strip.fill(0, 0, 0, 0)
            b = rb.Batch(strip)
            b.fill_paths(scene, ts=shard.strip_transform(y0))
            b.submit()
            b.close()
            tiled[y0:y0 + rows] = strip.download()
        assert_exact(got, want, f"{world} strips")
        differing = (np.abs(tiled.astype(np.int16) - want.astype(np.int16)).max(axis=2) > 0).mean()
        assert 0 < differing < 0.05, differing
    # a viewport reaching beyond the target on every side: the middle of the document
    mid = ctx.layer(500, 400)
    b = rb.Batch(mid)
    b.set_viewport(-300, -250, W, H)
    b.fill_paths(scene)
    b.submit()
    b.close()
    assert_exact(mid.download(), want[250:650, 300:800], "window in the middle of the document")
