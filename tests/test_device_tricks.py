"""CPU checks of two pieces of integer arithmetic the tile kernel relies on (resvg_b200/csrc/raster_warp.cuh), restated in
numpy with the same operations on 32-bit words:

* two channels per multiply: ``(x * k + 0x00ff00ff) >> 8 & 0x00ff00ff`` on ``R | B << 16`` is tiny-skia's lowp
  ``div255(v * k) = (v * k + 255) >> 8`` on both 16-bit lanes at once;
* the coverage bytes of blend_tile_gradient: four pixels' sample counts (0..16, one per byte) and their
  63-instead-of-64 flags (bit 4k) turned into ``min(16 * count - flag, 255)`` without a loop over the pixels.
"""
import numpy as np

U32 = np.uint64  # arithmetic in 64 bits, results masked to 32: what the GPU's 32-bit registers hold
M32 = np.uint64(0xFFFFFFFF)


def test_two_lanes_per_multiply_is_div255_on_each():
    v = np.arange(256, dtype=np.uint64)
    r, b, k = np.meshgrid(v, v[::5], v, indexing="ij")  # every r and k, every 5th b
    packed = (((r | (b << U32(16))) * k + U32(0x00FF00FF)) & M32) >> U32(8) & U32(0x00FF00FF)
    assert np.array_equal(packed & U32(0xFFFF), (r * k + U32(255)) >> U32(8))
    assert np.array_equal(packed >> U32(16), (b * k + U32(255)) >> U32(8))
    # the Source program adds two such products before the shift; their sum stays below 2^16 per lane when the factors sum to 255
    d, s, c = np.meshgrid(v, v[::3], v, indexing="ij")
    both = ((((d | (d << U32(16))) * (U32(255) - c) + (s | (s << U32(16))) * c + U32(0x00FF00FF)) & M32) >> U32(8)) & U32(0x00FF00FF)
    assert np.array_equal(both & U32(0xFFFF), ((d * (U32(255) - c) + s * c + U32(255)) >> U32(8)) & U32(0xFF))


def cov4(cnt, dnib):
    """blend_tile_gradient's lambda, operation by operation."""
    full = (cnt >> U32(4)) & U32(0x01010101)
    d = dnib & U32(0x1111)
    d = (d | (d << U32(8))) & U32(0x00FF00FF)
    d = (d | (d << U32(4))) & U32(0x01010101)
    return ((((cnt & U32(0x0F0F0F0F)) << U32(4)) - (d & (~full & M32))) | (full * U32(255))) & M32


def test_coverage_bytes_of_four_pixels_at_once():
    rng = np.random.default_rng(7)
    n = 200000
    counts = rng.integers(0, 17, size=(n, 4)).astype(np.uint64)
    # a pixel can only count 63 on its last sub-row if that sub-row is full: at least 4 of its 16 samples are inside
    flags = (rng.integers(0, 2, size=(n, 4)).astype(np.uint64)) * (counts >= 4)
    # every combination of the extreme values as well
    ext = np.array(np.meshgrid(*[[0, 1, 4, 15, 16]] * 4)).reshape(4, -1).T.astype(np.uint64)
    for f in (0, 1):
        counts = np.vstack([counts, ext])
        flags = np.vstack([flags, (ext >= 4) * np.uint64(f)])
    cnt = counts[:, 0] | counts[:, 1] << U32(8) | counts[:, 2] << U32(16) | counts[:, 3] << U32(24)
    noise = rng.integers(0, 1 << 16, size=len(cnt)).astype(np.uint64) & U32(0xEEEE)  # the other bits of the flag word are ignored
    dnib = (flags[:, 0] | flags[:, 1] << U32(4) | flags[:, 2] << U32(8) | flags[:, 3] << U32(12)) | noise
    got = cov4(cnt, dnib)
    for k in range(4):
        want = np.minimum(U32(16) * counts[:, k] - flags[:, k], U32(255))
        assert np.array_equal((got >> U32(8 * k)) & U32(0xFF), want), k
