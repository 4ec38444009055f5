"""CPU tests of the filter oracle (oracle/filters.c): known-answer values read off the reference source and
independent numpy formulations of the same arithmetic.  Golden-PNG pinning lives in test_golden.py."""
import numpy as np

from tests import oracle_ffi as o
from tests.util import random_premul, random_rgba


def test_create_box_gauss_known_values():
    # sigma=2 -> [3;5]; sigma=6 -> [9,9,9,9,11]; sigma=64 -> [99;5]... from box_blur.rs:37-71
    assert o.create_box_gauss(2.0) == [3, 3, 3, 3, 3]
    assert o.create_box_gauss(6.0) == [9, 9, 9, 9, 11]
    assert o.create_box_gauss(0.0) == [1, 1, 1, 1, 1]
    assert o.create_box_gauss(-1.0) == [1, 1, 1, 1, 1]
    s = o.create_box_gauss(64.0)
    assert all(v % 2 == 1 for v in s) and 97 <= s[0] <= 101


def _box_axis_numpy(img, r, axis):
    """Zero-padded window sum, round-half-even of sum * f32(1/(2r+1))."""
    if r == 0:
        return img.copy()
    x = img.astype(np.int64)
    pad = [(0, 0)] * 3
    pad[axis] = (r + 1, r)
    c = np.cumsum(np.pad(x, pad), axis=axis)
    n = img.shape[axis]
    hi = np.take(c, np.arange(2 * r + 1, 2 * r + 1 + n), axis=axis)
    lo = np.take(c, np.arange(0, n), axis=axis)
    s = (hi - lo).astype(np.float32)
    iarr = np.float32(1.0) / np.float32(2 * r + 1)
    v = s * iarr
    v = (v + np.float32(12582912.0)) - np.float32(12582912.0)
    return np.clip(v, 0, 255).astype(np.uint8)


def test_box_quantisation_equals_an_exact_integer_quotient():
    """The identity the fast CUDA box-blur passes rely on (filters.cu, BOX2): for every radius r <= 127 and every
    possible window sum, round(sum as f32 * (1.0 / d as f32)) with the reference's add/subtract-1.5*2^23 rounding
    (box_blur.rs:148, 327-331) equals ((sum + r) * ceil(2^24 / d)) >> 24, d = 2r + 1."""
    magic = np.float32(12582912.0)
    for r in range(1, 128):
        d = 2 * r + 1
        sums = np.arange(0, 255 * d + 1, dtype=np.int64)
        v = sums.astype(np.float32) * (np.float32(1.0) / np.float32(d))
        ref = ((v + magic) - magic).astype(np.int64)
        m = ((1 << 24) + d - 1) // d
        prod = (sums + r) * m
        assert prod.max() < 2**32 and 255 * d + r < 65536
        assert np.array_equal(prod >> 24, ref), r


def test_box_blur_matches_window_sum_formulation():
    for (w, h, sx, sy) in [(37, 23, 2.0, 2.0), (64, 9, 6.0, 3.0), (15, 40, 0.0, 4.0), (8, 8, 30.0, 30.0)]:
        img = random_premul(w, h, 1, sparse=True)
        bh, bv = o.create_box_gauss(sx), o.create_box_gauss(sy)
        want = img.copy()
        for i in range(5):
            want = _box_axis_numpy(want, (bv[i] - 1) // 2, 0)
            want = _box_axis_numpy(want, (bh[i] - 1) // 2, 1)
        got = o.box_blur(sx, sy, img)
        assert np.array_equal(got, want), (w, h, sx, sy)


def test_premultiply_demultiply_formulas():
    img = random_rgba(33, 17, 2)
    a = img[..., 3:4].astype(np.float32) / np.float32(255.0)
    want = img.copy()
    want[..., :3] = (img[..., :3].astype(np.float32) * a + np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(o.multiply_alpha(img), want)
    pm = random_premul(33, 17, 3)
    a = pm[..., 3:4].astype(np.float32) / np.float32(255.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = pm[..., :3].astype(np.float32) / a + np.float32(0.5)
    v = np.nan_to_num(v, nan=0.0, posinf=255.0)
    want = pm.copy()
    want[..., :3] = np.clip(v, 0, 255).astype(np.uint8)
    assert np.array_equal(o.demultiply_alpha(pm), want)


def test_srgb_tables_are_monotone_and_inverse_at_ends():
    import ctypes as C

    o.lib.orc_srgb_to_linear_table.restype = C.POINTER(C.c_uint8)
    o.lib.orc_linear_to_srgb_table.restype = C.POINTER(C.c_uint8)
    a = np.ctypeslib.as_array(o.lib.orc_srgb_to_linear_table(), (256,))
    b = np.ctypeslib.as_array(o.lib.orc_linear_to_srgb_table(), (256,))
    assert a[0] == 0 and a[255] == 255 and b[0] == 0 and b[255] == 255
    assert np.all(np.diff(a.astype(int)) >= 0) and np.all(np.diff(b.astype(int)) >= 0)
    # the formula in the doc comment (filter/mod.rs:150-159) reproduces the table
    c = np.arange(256) / 255.0
    lin = np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    assert np.abs(np.round(lin * 255) - a).max() <= 1


def test_morphology_separable_equals_2d():
    img = random_premul(31, 19, 4, sparse=True)
    for op in ("erode", "dilate"):
        full = o.morphology(op, 3.0, 2.0, img)
        # asymmetric even window [x-3, x+2] x [y-2, y+1]
        want = np.zeros_like(img)
        h, w = img.shape[:2]
        for y in range(h):
            for x in range(w):
                win = img[max(0, y - 2):y + 2, max(0, x - 3):x + 3].reshape(-1, 4)
                want[y, x] = win.min(axis=0) if op == "erode" else win.max(axis=0)
        assert np.array_equal(full, want), op


def test_iir_blur_preserves_flat_and_decays():
    img = np.full((32, 32, 4), 255, dtype=np.uint8)
    out = o.iir_blur(1.5, 1.5, img)
    # truncating store makes the interior decay to 254 at most
    assert out[16, 16, 3] in (254, 255)
    # no boundary renormalisation in iir_blur.rs:84-101 (unlike Getreuer's original): edges darken
    assert out[0, 0, 3] < out[16, 16, 3]


def test_turbulence_is_deterministic_and_seed_sensitive():
    a = o.turbulence(0, 0, 1, 1, 0.05, 0.05, 2, 1, False, True, 32, 16)
    b = o.turbulence(0, 0, 1, 1, 0.05, 0.05, 2, 1, False, True, 32, 16)
    c = o.turbulence(0, 0, 1, 1, 0.05, 0.05, 2, 2, False, True, 32, 16)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_composite_arithmetic_skips_transparent():
    a, b = random_premul(16, 16, 5), random_premul(16, 16, 6)
    out = o.arithmetic(0.0, 0.0, 0.0, 0.0, a, b)
    assert not out.any()
    out = o.arithmetic(0.0, 1.0, 0.0, 0.0, a, b)
    # k2 = 1: result = src1 (premultiplied, clamp to alpha, truncating store)
    assert np.abs(out.astype(int) - a.astype(int)).max() <= 1
