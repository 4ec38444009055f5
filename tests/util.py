import numpy as np


def random_premul(w, h, seed, sparse=False):
    """Random premultiplied RGBA8: a ~ U{0..255}, c ~ U{0..a}  (SURVEY.md §8(d) C3 standalone buffers)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(h, w), dtype=np.int64)
    if sparse:  # large transparent / opaque areas, like real layers
        m = rng.integers(0, 4, size=(h, w))
        a = np.where(m == 0, 0, np.where(m == 1, 255, a))
    c = (rng.random((h, w, 3)) * (a[..., None] + 1)).astype(np.int64)
    c = np.minimum(c, a[..., None])
    return np.concatenate([c, a[..., None]], axis=2).astype(np.uint8)


def random_rgba(w, h, seed):
    """Unpremultiplied random RGBA8."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(h, w, 4), dtype=np.int64).astype(np.uint8)


def smooth_alpha(w, h, seed):
    """Premultiplied image with a smooth alpha bump field (for lighting normals)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    a = np.zeros((h, w))
    for _ in range(6):
        cx, cy = rng.random() * w, rng.random() * h
        s = (0.1 + rng.random() * 0.3) * max(w, h)
        a += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    a = np.clip(a / a.max() * 255.0, 0, 255).astype(np.uint8)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    img[..., 3] = a
    img[..., 0] = a // 2
    return img


def max_abs_diff(a, b):
    return int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max()) if a.size else 0


def assert_exact(got, want, what=""):
    if not np.array_equal(got, want):
        d = np.abs(got.astype(np.int16) - want.astype(np.int16))
        n = int((d.max(axis=-1) > 0).sum()) if d.ndim == 3 else int((d > 0).sum())
        idx = np.argwhere(d > 0)[:5]
        raise AssertionError(f"{what}: {n} pixels differ, max |diff| = {d.max()}, first at {idx.tolist()}")


def assert_within(got, want, tol, what=""):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    if d.size and d.max() > tol:
        n = int((d > tol).sum())
        idx = np.argwhere(d > tol)[:5]
        raise AssertionError(f"{what}: {n} channel values differ by more than {tol} (max {d.max()}), first at {idx.tolist()}")
