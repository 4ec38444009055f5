"""The product's host geometry (resvg_b200/csrc/{stroker,dasher,hairline}.cpp) against the oracle's independent
restatements (oracle/{stroke,dash,hairline}.c).  The two were written separately from the published tiny-skia-path /
tiny-skia algorithm (the crate's source is not under /root/reference); nothing is shared, so agreement is evidence, and the
GPU parity tests feed each arm with ITS OWN geometry (tests/geom.py).

Float note: outlines agree bit for bit except where two mathematically equal expressions round differently (a handful of
1-ulp points per thousand paths); the bound below is in pixels.  Dashed paths and hairline blit lists must be identical.
"""
import numpy as np
import pytest

import resvg_b200 as rb
from tests import oracle_geom as G
from tests.pathgen import SplitMix64, random_path

CAPS = ["butt", "round", "square"]
JOINS = ["miter", "miter-clip", "round", "bevel"]


def _cases(seed, n, lo=0, hi=300):
    rng = SplitMix64(seed)
    for _ in range(n):
        cx, cy, r = rng.uniform(lo, hi), rng.uniform(lo, hi), rng.log_uniform(2, 150)
        yield rng, random_path(rng, cx, cy, r)


def test_stroker_outlines_agree():
    exact = total = 0
    worst = 0.0
    for rng, (verbs, pts) in _cases(7, 2500):
        w, ml = rng.log_uniform(0.3, 30), rng.uniform(1, 8)
        cap, join = CAPS[int(rng.u() * 3)], JOINS[int(rng.u() * 4)]
        rs = rng.log_uniform(0.3, 10)
        a = rb.stroke_path(verbs, pts, w, ml, cap, join, rs)
        b = G.stroke_path(verbs, pts, w, ml, cap, join, rs)
        assert (a is None) == (b is None)
        if a is None:
            continue
        total += 1
        assert np.array_equal(a[0], b[0]), "different verb sequence"
        d = float(np.abs(a[1] - b[1]).max())
        worst = max(worst, d)
        exact += d == 0.0
    assert worst <= 1e-3, worst
    assert exact >= 0.99 * total, (exact, total)


@pytest.mark.parametrize("verbs,pts", [
    ([0, 4], [(10, 10)]),                                   # move + close: a dot (caps only)
    ([0, 1], [(10, 10), (10, 10)]),                         # zero-length line
    ([0, 1, 1, 4], [(0, 0), (50, 0), (0, 0)]),              # 180-degree turn
    ([0, 2], [(0, 0), (50, 0), (100, 0)]),                  # quad on a line
    ([0, 2], [(0, 0), (100, 0), (50, 0)]),                  # quad folded back on itself
    ([0, 3], [(0, 0), (100, 0), (-50, 0), (50, 0)]),        # cubic folded on a line
    ([0, 3], [(0, 0), (100, 100), (0, 100), (100, 0)]),     # cubic with a loop
    ([0, 3], [(0, 0), (0, 0), (100, 0), (100, 0)]),         # coincident control points
    ([0, 3], [(0, 0), (60, 80), (60, 80), (0, 0)]),         # cusp
    ([0, 1, 0, 1, 1, 4], [(0, 0), (30, 5), (50, 50), (90, 50), (70, 90)]),  # two contours, one closed
])
def test_stroker_degenerate_inputs_agree(verbs, pts):
    for cap in CAPS:
        for join in JOINS:
            for width in (0.7, 5.0, 40.0):
                a = rb.stroke_path(verbs, pts, width, 4.0, cap, join, 1.0)
                b = G.stroke_path(verbs, pts, width, 4.0, cap, join, 1.0)
                assert (a is None) == (b is None), (cap, join, width)
                if a is not None:
                    assert np.array_equal(a[0], b[0]), (cap, join, width)
                    assert np.abs(a[1] - b[1]).max() <= 1e-3, (cap, join, width)


def test_dasher_paths_identical():
    n = 0
    for rng, (verbs, pts) in _cases(11, 2500):
        k = 2 * (1 + int(rng.u() * 3))
        dash = [rng.log_uniform(0.2, 40) if rng.u() > 0.1 else 0.0 for _ in range(k)]
        off = rng.uniform(-100, 100) if rng.u() > 0.3 else 0.0
        rs = rng.log_uniform(0.3, 10)
        a = rb.dash_path(verbs, pts, dash, off, rs)
        b = G.dash_path(verbs, pts, dash, off, rs)
        assert (a is None) == (b is None)
        if a is not None:
            n += 1
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert n > 2000


@pytest.mark.parametrize("dash,off", [([], 0.0), ([5.0], 0.0), ([5.0, 3.0, 2.0], 0.0), ([5.0, -1.0], 0.0), ([0.0, 0.0], 0.0),
                                      ([4.0, 4.0], float("inf")), ([4.0, 4.0], -3.0), ([4.0, 4.0], 1e9)])
def test_dash_specifications_stroke_dash_new_rejects(dash, off):
    verbs, pts = [0, 1, 1], [(0, 0), (100, 0), (100, 100)]
    a = rb.dash_path(verbs, pts, dash, off, 1.0)
    b = G.dash_path(verbs, pts, dash, off, 1.0)
    assert (a is None) == (b is None)
    if a is not None:
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_hairline_blit_lists_identical():
    """Includes paths that cross and hug the canvas border, where tiny-skia's unsigned pixel coordinates bend the
    walker's output (pairs shifted inwards, ordinates clamped at 0) and the sub-clip cuts it again."""
    same = total = 0
    rng = SplitMix64(21)
    for i in range(4000):
        W, H = [(200, 150), (64, 64), (9000, 30)][i % 3]
        cx, cy, r = rng.uniform(-20, min(W, 300) + 20), rng.uniform(-20, H + 20), rng.log_uniform(2, 150)
        verbs, pts = random_path(rng, cx, cy, r)
        cap = CAPS[int(rng.u() * 3)]
        a = rb.hairline_blits(verbs, pts, cap, W, H)
        b = G.hairline_blits(verbs, pts, cap, W, H)
        total += 1
        same += a.shape == b.shape and np.array_equal(a, b)
    assert same >= total - 2, (same, total)  # a curve whose subdivision count sits on a float tie may differ


def test_hairline_border_cases():
    for p0, p1 in [((2.5, 66.0), (-1.5, 73.0)), ((2.5, 66.0), (-0.5, 73.0)), ((10.5, 3.0), (50.25, -0.7)), ((-3.0, -3.0), (40.0, 25.0)),
                   ((0.0, 0.0), (199.9, 149.9)), ((-5.0, 10.0), (300.0, 10.0)), ((10.0, -5.0), (10.0, 300.0)), ((0.2, 0.2), (0.7, 0.6))]:
        for cap in CAPS:
            a = rb.hairline_blits([0, 1], [p0, p1], cap, 200, 150)
            b = G.hairline_blits([0, 1], [p0, p1], cap, 200, 150)
            assert np.array_equal(a, b), (p0, p1, cap)
