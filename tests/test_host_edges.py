"""CPU tests of the product's host geometry (resvg_b200/csrc/raster_host.cpp, via rb_debug_build_edges) and of the
device coverage formulation (tests/device_model.py) against the sequential oracle walker (oracle/raster.c)."""
import numpy as np
import pytest

from tests import device_model as D
from tests import oracle_raster as R
from tests.pathgen import SplitMix64, random_path


def test_rect_edges_are_two_vertical_lines():
    verbs, pts = [0, 1, 1, 1, 4], [(10, 5), (30, 5), (30, 25), (10, 25)]
    edges, meta, geom = D.build_edges(verbs, pts, True, 64, 64)
    assert len(edges) == 2
    assert list(geom[:5]) == [10, 5, 20, 20, 2]
    # supersampled FDot16: x = 10*4 and 30*4 sub-pixels, y rows 20..99
    assert sorted(int(e[0]) >> 16 for e in edges) == [40, 120]
    assert all(int(e[1]) == 0 and int(e[2]) == 20 and int(e[3]) == 99 for e in edges)
    assert sorted(int(e[4]) for e in edges) == [-1, 1]


def test_non_aa_uses_pixel_grid():
    verbs, pts = [0, 1, 1, 4], [(2.2, 1.1), (40.7, 3.3), (20.5, 30.9)]
    edges, meta, geom = D.build_edges(verbs, pts, False, 64, 64)
    assert geom[4] == 0 and len(edges) >= 2
    assert min(int(e[2]) for e in edges) >= 1 and max(int(e[3]) for e in edges) <= 31


def test_curves_expand_to_monotone_line_edges():
    rng = SplitMix64(5)
    verbs, pts = random_path(rng, 100, 100, 80, n_seg=6, kinds=(1.0, 0.0, 0.0))
    edges, meta, geom = D.build_edges(verbs, pts, True, 256, 256)
    assert len(edges) > 12
    assert all(int(e[2]) <= int(e[3]) for e in edges)
    fy = [int(e[2]) for e in edges]
    assert fy == sorted(fy)  # sorted by first_y
    # continuation links point at the segment that ends right above
    for i, m in enumerate(meta):
        if m[0] >= 0:
            assert int(edges[m[0]][3]) + 1 == int(edges[i][2])


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("rule", ["nonzero", "evenodd"])
def test_device_model_matches_oracle_coverage(seed, rule):
    w, h = 72, 40
    rng = SplitMix64(9000 + seed)
    for i in range(14):
        verbs, pts = random_path(rng, rng.uniform(-5, w + 5), rng.uniform(-5, h + 5), rng.log_uniform(3, 90))
        aa = i % 4 != 3
        want = R.path_coverage(w, h, verbs, pts, rule, aa)
        got = D.coverage(verbs, pts, w, h, rule, aa)
        assert np.array_equal(got, want), (seed, i, rule, aa, np.argwhere(got != want)[:4].tolist())


def test_device_model_exact_tie_from_curve_continuation():
    """Regression: a curve's continuation segment and a new edge with the identical FDot16 x on the 4th sub-row
    of a fully covered pixel — the walker keeps the curve first, so the span does not break (63, not 64)."""
    w, h = 257, 131
    rng = SplitMix64(1000 + w * 7 + h)
    for i in range(8):
        r = rng.log_uniform(2, max(w, h))
        verbs, pts = random_path(rng, rng.uniform(0, w), rng.uniform(0, h), r)
    want = R.path_coverage(w, h, verbs, pts, "nonzero", True)
    got = D.coverage(verbs, pts, w, h, "nonzero", True)
    assert np.array_equal(got, want), np.argwhere(got != want)[:4].tolist()


def test_device_model_structured_shapes():
    """Coincident vertical edges of abutting rectangles, frames with corner excursions, retraced spikes."""
    w, h = 96, 64
    M, L, Z = 0, 1, 4

    def rect(x, y, rw, rh, ccw=False):
        p = [(x, y), (x + rw, y), (x + rw, y + rh), (x, y + rh)]
        return [M, L, L, L, Z], (p[::-1] if ccw else p)

    shapes = []
    for ccw in (False, True):
        v1, p1 = rect(10.3, 20.6, 20.33, 30.1)
        v2, p2 = rect(30.63, 15.2, 23.3, 40.7, ccw)
        shapes.append((v1 + v2, p1 + p2))
        shapes.append((v2 + v1, p2 + p1))
    v1, p1 = rect(40.75, 10.75, 40.5, 40.5)
    v2, p2 = rect(42.25, 12.25, 37.5, 37.5, True)
    shapes.append((v1 + v2, p1 + p2))
    shapes.append(([M, L, L, L, L, L, L, Z], [(5.2, 40.1), (60.7, 40.1), (60.7, 60.9), (35.3, 60.9), (35.3, 45.0), (35.3, 60.9), (5.2, 60.9)]))
    for i, (v, p) in enumerate(shapes):
        for rule in ("nonzero", "evenodd"):
            for ts in (D.IDENTITY, (1.5, 0, 0, 1.5, 0.37, 0.21)):
                want = R.path_coverage(w, h, v, p, rule, True, ts)
                got = D.coverage(v, p, w, h, rule, True, ts)
                assert np.array_equal(got, want), (i, rule, ts, np.argwhere(got != want)[:4].tolist())


def test_dasher_lengths_and_rejections():
    """tiny_skia_path::Path::dash on a straight line: the on-intervals come out where the pattern says; StrokeDash::new
    rejections (odd count, negative entry, zero sum) yield no path."""
    import resvg_b200 as rb

    verbs, pts = [0, 1], [(0.0, 0.0), (100.0, 0.0)]
    v, p = rb.dash_path(verbs, pts, [10.0, 5.0], 0.0)
    assert list(v[:4]) == [0, 1, 0, 1]
    xs = p[:, 0].reshape(-1, 2)
    assert np.allclose(xs[:, 0], np.arange(0, 100, 15.0)) and np.allclose(xs[:-1, 1], np.arange(10, 100, 15.0))
    assert xs[-1, 1] == 100.0  # the last dash is cut at the end of the contour
    v2, p2 = rb.dash_path(verbs, pts, [10.0, 5.0], 12.0)  # starts inside the gap
    assert np.allclose(p2[:2, 0], [3.0, 13.0])
    v3, p3 = rb.dash_path(verbs, pts, [10.0, 5.0], -5.0)  # negative offsets wrap
    assert np.allclose(p3[:2, 0], [0.0, 5.0]) or np.allclose(p3[:2, 0], [5.0, 15.0])
    for bad in ([4.0], [4.0, 4.0, 4.0], [5.0, -1.0], [0.0, 0.0]):
        assert rb.dash_path(verbs, pts, bad, 0.0) is None
    # closed contour: the first dash joins up with the last one (no move_to in between)
    sq_v, sq_p = [0, 1, 1, 1, 4], [(0.0, 0.0), (40.0, 0.0), (40.0, 40.0), (0.0, 40.0)]
    v4, p4 = rb.dash_path(sq_v, sq_p, [30.0, 10.0], 0.0)
    assert int((v4 == 0).sum()) == 4 and len(v4) > 8
