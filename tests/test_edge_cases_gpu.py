"""GPU parity on the inputs where a rasteriser usually goes wrong: degenerate and non-finite geometry, huge
coordinates, shapes far outside the layer, the DrawTiler seam at 8191 px, tiny layers, empty batches, error returns.
Everything is compared with the CPU oracle on the same input (bit-exact: solid colours through the u16 pipeline)."""
import numpy as np
import pytest

from tests import oracle_raster as R
from tests.pathgen import SplitMix64, random_path
from tests.util import assert_exact

pytestmark = pytest.mark.gpu

M, L, Q, C, Z = 0, 1, 2, 3, 4
PAINT = {"kind": "solid", "color": (0.9, 0.3, 0.1, 0.8)}


def _both(ctx, w, h, draws, aa=True):
    """draws: (verbs, pts, rule, ts) — rendered in order by one GPU batch and by the oracle."""
    import resvg_b200 as rb

    want = np.zeros((h, w, 4), np.uint8)
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    for verbs, pts, rule, ts in draws:
        b.fill_path(verbs, pts, rb.make_paint(PAINT, "source_over", aa), rule, ts)
        R.fill_path(want, verbs, pts, R.make_paint(PAINT, "source_over", aa), rule, ts)
    b.submit()
    got = l.download()
    b.close()
    l.close()
    return got, want


DEGENERATE = [
    ("single move", [M], [(5, 5)]),
    ("move + close", [M, Z], [(5, 5)]),
    ("zero-length line", [M, L, Z], [(5, 5), (5, 5)]),
    ("two-point line (no area)", [M, L, Z], [(2, 2), (30, 20)]),
    ("horizontal sliver", [M, L, L, Z], [(2, 10), (40, 10), (20, 10.0005)]),
    ("vertical sliver", [M, L, L, Z], [(10, 2), (10, 40), (10.0005, 20)]),
    ("all points equal (curves)", [M, Q, C, Z], [(7, 7)] * 6),
    ("quad with coincident control", [M, Q, L, Z], [(3, 3), (3, 3), (40, 30), (3, 30)]),
    ("cubic cusp", [M, C, Z], [(5, 40), (60, 0), (0, 0), (55, 40)]),
    ("cubic loop", [M, C, Z], [(10, 10), (80, 60), (-20, 60), (50, 10)]),
    ("retraced contour", [M, L, L, L, L, Z], [(5, 5), (50, 5), (50, 40), (50, 5), (5, 5)]),
    ("two moves in a row", [M, M, L, L, Z], [(1, 1), (10, 10), (40, 10), (25, 40)]),
    ("unclosed contours", [M, L, L, M, L, L], [(5, 5), (30, 8), (12, 30), (35, 20), (60, 25), (40, 45)]),
]


@pytest.mark.parametrize("name,verbs,pts", DEGENERATE, ids=[d[0] for d in DEGENERATE])
@pytest.mark.parametrize("aa", [True, False])
def test_degenerate_geometry(ctx, name, verbs, pts, aa):
    got, want = _both(ctx, 64, 48, [(verbs, pts, "nonzero", R.IDENTITY), (verbs, pts, "evenodd", (1.5, 0.2, -0.3, 1.1, 4, 2))], aa)
    assert_exact(got, want, name)


@pytest.mark.parametrize("bad", [float("nan"), float("inf"), -float("inf"), 3.0e38, -3.0e38, 1.0e9, 2147483648.0])
def test_non_finite_and_huge_coordinates(ctx, bad):
    """tiny-skia drops paths with non-finite bounds; huge finite ones are clipped.  A normal shape after the bad one
    must still be drawn (the batch is not poisoned)."""
    tri = ([M, L, L, Z], [(5, 5), (60, 10), (20, 40)])
    for k in range(3):
        pts = [(5, 5), (60, 10), (20, 40)]
        pts[k] = (bad, pts[k][1]) if k % 2 == 0 else (pts[k][0], bad)
        got, want = _both(ctx, 64, 48, [([M, L, L, Z], pts, "nonzero", R.IDENTITY), (tri[0], tri[1], "evenodd", R.IDENTITY)])
        assert_exact(got, want, f"bad={bad} at point {k}")
        assert got[..., 3].any()


def test_shapes_far_outside_and_covering_everything(ctx):
    w, h = 96, 64
    big = 1.0e6
    draws = [
        ([M, L, L, L, Z], [(-big, -big), (big, -big), (big, big), (-big, big)], "nonzero", R.IDENTITY),   # covers everything
        ([M, L, L, Z], [(-500, -500), (-100, -480), (-300, -90)], "nonzero", R.IDENTITY),                 # entirely outside
        ([M, L, L, Z], [(w + 10, 5), (w + 200, 30), (w + 50, 60)], "evenodd", R.IDENTITY),                # right of the layer
        ([M, L, L, Z], [(-40, 20), (w + 40, 22), (30, h + 50)], "nonzero", R.IDENTITY),                   # crosses three sides
        ([M, C, Z], [(-1000, 30), (50, -4000), (60, 5000), (1200, 20)], "evenodd", R.IDENTITY),           # huge curve through the layer
    ]
    for d in draws:
        got, want = _both(ctx, w, h, [d])
        assert_exact(got, want, str(d[1][:2]))
    got, want = _both(ctx, w, h, draws)
    assert_exact(got, want, "all together")


def test_non_invertible_and_extreme_transforms(ctx):
    tri = ([M, L, L, Z], [(5, 5), (60, 10), (20, 40)])
    for ts in [(0, 0, 0, 0, 10, 10), (1, 2, 2, 4, 0, 0), (1e-6, 0, 0, 1e-6, 20, 20), (1e4, 0, 0, 1e4, -2e5, -2e5), (-1, 0, 0, -1, 64, 48),
               (0, 1, 1, 0, 0, 0)]:
        got, want = _both(ctx, 64, 48, [(tri[0], tri[1], "nonzero", ts)])
        assert_exact(got, want, f"ts={ts}")


@pytest.mark.parametrize("w,h", [(1, 1), (1, 37), (37, 1), (2, 2), (31, 7), (33, 9), (8191, 3), (3, 8191)])
def test_tiny_and_extreme_layer_shapes(ctx, w, h):
    rng = SplitMix64(w * 31 + h)
    draws = []
    for _ in range(6):
        verbs, pts = random_path(rng, rng.uniform(0, min(w, 400)), rng.uniform(0, min(h, 400)), rng.log_uniform(2, 300))
        draws.append((verbs, pts, "evenodd" if rng.u() < 0.5 else "nonzero", R.IDENTITY))
    got, want = _both(ctx, w, h, draws)
    assert_exact(got, want, f"{w}x{h}")


@pytest.mark.parametrize("w,h", [(8200, 24), (24, 8200), (8193, 17)])
def test_draw_tiler_seam(ctx, w, h):
    """Layers wider / taller than 8191 px are drawn as DrawTiler tiles (8191 + rest): shapes and hairlines across the
    seam must come out as the oracle's tiled painter draws them."""
    import resvg_b200 as rb
    from tests.backends import OracleBackend

    rng = SplitMix64(w + h)
    ob = OracleBackend()
    want = np.zeros((h, w, 4), np.uint8)
    l = ctx.layer(w, h)
    b = rb.Batch(l)
    horiz = w > h
    for i in range(10):
        c = 8191 + rng.uniform(-12, 12)
        cx, cy = (c, rng.uniform(0, h)) if horiz else (rng.uniform(0, w), c)
        verbs, pts = random_path(rng, cx, cy, rng.log_uniform(4, 30))
        if i % 3 == 2:
            width = rng.uniform(0.2, 0.9)
            b.stroke_path(verbs, pts, rb.make_paint(PAINT), width, 4.0, "round", "miter", R.IDENTITY)
            ob.stroke_hairline(want, verbs, pts, PAINT, R.IDENTITY, "source_over", width, "round")
        else:
            b.fill_path(verbs, pts, rb.make_paint(PAINT), "nonzero")
            R.fill_path(want, verbs, pts, R.make_paint(PAINT), "nonzero")
    b.submit()
    got = l.download()
    sl = (slice(None), slice(8150, None)) if horiz else (slice(8150, None), slice(None))
    assert_exact(got[sl], want[sl], "around the seam")
    assert_exact(got, want, "whole layer")


def test_empty_batches_and_error_returns(ctx):
    import ctypes as C

    import resvg_b200 as rb
    from resvg_b200 import _ffi
    from resvg_b200.api import ResvgB200Error

    l = ctx.layer(40, 30)
    b = rb.Batch(l)
    b.submit()  # nothing recorded: no-op
    b.prepare()
    b.run()
    assert not l.download().any()
    paint = rb.make_paint(PAINT)
    with pytest.raises(ResvgB200Error):
        b.fill_path([L, L], [(1, 1), (2, 2)], paint)  # does not start with a move
    with pytest.raises(ResvgB200Error):
        b.fill_path([M, L, 9], [(1, 1), (2, 2)], paint)  # unknown verb
    with pytest.raises(ResvgB200Error):
        b.fill_path([M, Q], [(1, 1), (2, 2)], paint)  # too few points for the verbs
    bad = rb.make_paint(PAINT)
    bad.blend_mode = 77
    with pytest.raises(ResvgB200Error):
        b.fill_path([M, L, L, Z], [(1, 1), (20, 2), (5, 20)], bad)
    # the failed records left nothing behind
    b.fill_path([M, L, L, Z], [(1, 1), (30, 2), (5, 25)], paint)
    b.submit()
    want = np.zeros((30, 40, 4), np.uint8)
    R.fill_path(want, [M, L, L, Z], [(1, 1), (30, 2), (5, 25)], R.make_paint(PAINT))
    assert_exact(l.download(), want, "after rejected records")
    # zero-sized layers cannot exist (tiny-skia Pixmap::new -> None)
    h = C.c_void_p()
    assert _ffi.lib.rb_layer_create(ctx._h, 0, 10, C.byref(h)) != 0
    assert _ffi.lib.rb_layer_create(ctx._h, 10, 0, C.byref(h)) != 0
    # negative stroke widths draw nothing and are not an error (painter.rs: width < 0 -> return)
    b2 = rb.Batch(l)
    b2.stroke_path([M, L], [(1, 1), (30, 20)], paint, -1.0)
    b2.submit()
    assert_exact(l.download(), want, "negative stroke width")


def test_immediate_fills_are_collected_and_flushed_in_call_order(ctx):
    """rb_fill_path records lazily: consecutive immediate draws on a layer run as one batch the next time the layer is
    read or written.  Interleavings with explicit batches, composites, pattern paints (executed at once), masks, copies
    and layer destruction must give the CPU checker's sequential result, with far fewer launches than draws."""
    import resvg_b200 as rb

    W, H = 150, 120
    rng = SplitMix64(99)
    shapes = []
    for k in range(60):
        cx, cy, r = rng.uniform(0, W), rng.uniform(0, H), rng.log_uniform(6, 50)
        verbs, pts = random_path(rng, cx, cy, r)
        spec = {"kind": "solid", "color": (rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0.3, 1.0))}
        shapes.append((verbs, pts, spec, "evenodd" if k % 3 == 0 else "nonzero"))
    want = np.zeros((H, W, 4), np.uint8)
    l = ctx.layer(W, H)
    ctx.synchronize()
    launches0 = ctx.launch_count
    for verbs, pts, spec, rule in shapes[:40]:                       # 40 immediate draws ...
        rb.fill_path(l, verbs, pts, rb.make_paint(spec), rule)
        R.fill_path(want, verbs, pts, R.make_paint(spec), rule)
    assert ctx.launch_count == launches0                              # ... nothing has run yet
    b = rb.Batch(l)                                                   # an explicit batch recorded now, submitted later
    for verbs, pts, spec, rule in shapes[40:50]:
        b.fill_path(verbs, pts, rb.make_paint(spec), rule)
    for verbs, pts, spec, rule in shapes[50:55]:                      # more immediate draws BEFORE the batch is submitted:
        rb.fill_path(l, verbs, pts, rb.make_paint(spec), rule)        # they come first (call order of the executions)
        R.fill_path(want, verbs, pts, R.make_paint(spec), rule)
    b.submit()
    for verbs, pts, spec, rule in shapes[40:50]:
        R.fill_path(want, verbs, pts, R.make_paint(spec), rule)
    b.close()
    assert ctx.launch_count - launches0 < 40                          # two batches, not 55 one-draw batches
    # a composite reads a layer with pending draws; the destination has pending draws of its own
    top = ctx.layer(W, H)
    top_want = np.zeros((H, W, 4), np.uint8)
    for verbs, pts, spec, rule in shapes[55:]:
        rb.fill_path(top, verbs, pts, rb.make_paint(spec), rule)
        R.fill_path(top_want, verbs, pts, R.make_paint(spec), rule)
    verbs, pts, spec, rule = shapes[0]
    rb.fill_path(l, verbs, pts, rb.make_paint(spec), rule)
    R.fill_path(want, verbs, pts, R.make_paint(spec), rule)
    rb.draw_layer(l, top, 0, 0, 1.0, "source_over")
    R.draw_pixmap(want, 0, 0, top_want, 1.0, "source_over")
    # a pattern paint whose source layer has pending draws and is destroyed right after the call
    tile = ctx.layer(16, 12)
    tile_want = np.zeros((12, 16, 4), np.uint8)
    tv, tp = [M, L, L, Z], [(1, 1), (15, 3), (6, 11)]
    rb.fill_path(tile, tv, tp, rb.make_paint(PAINT), "nonzero")
    R.fill_path(tile_want, tv, tp, R.make_paint(PAINT), "nonzero")
    pspec = {"kind": "pattern", "layer": tile, "spread": "repeat", "quality": "nearest", "opacity": 1.0, "ts": (1, 0, 0, 1, 0, 0)}
    verbs, pts, _, rule = shapes[1]
    rb.fill_path(l, verbs, pts, rb.make_paint(pspec), rule)
    tile.close()
    R.fill_path(want, verbs, pts, R.make_paint(dict(pspec, pixmap=tile_want)), rule)
    # clone, mask and download all see the pending draws
    verbs, pts, spec, rule = shapes[2]
    rb.fill_path(l, verbs, pts, rb.make_paint(spec), rule)
    R.fill_path(want, verbs, pts, R.make_paint(spec), rule)
    c = l.clone()
    m = rb.Mask.from_layer(l, "alpha")
    assert_exact(c.download(), want, "clone after immediate draws")
    assert_exact(m.download(), want[..., 3], "mask after immediate draws")
    assert_exact(l.download(), want, "download after immediate draws")
    # a layer destroyed with pending draws: nothing runs, nothing breaks
    dead = ctx.layer(64, 64)
    rb.fill_path(dead, *shapes[3][:2], rb.make_paint(shapes[3][2]), "nonzero")
    dead.close()
    ctx.synchronize()
    assert_exact(l.download(), want, "still intact")
