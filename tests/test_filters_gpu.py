"""GPU parity: every CUDA filter kernel against the CPU oracle (oracle/filters.c) on the same inputs.

Bars (BASELINE.json north_star): bit-exact for integer / u8 / LUT stages and for every f32/f64 stage
whose operation order is reproduced exactly; <= 1/255 per channel only where the device evaluates
`powf` (spot-light cone exponent, specular exponent) differently from glibc.
"""
import numpy as np
import pytest

from tests.util import assert_exact, assert_within, random_premul, random_rgba, smooth_alpha

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (3, 2), (7, 5), (64, 48), (257, 131), (300, 300), (1029, 67)]
# layers of >= 65536 px take the table-driven demultiply / colour-space kernels
SIZES_HELPERS = SIZES + [(256, 256), (517, 301)]


def _run(ctx, img, fn):
    import resvg_b200 as rb

    l = ctx.layer_from(img)
    fn(rb.filters, l)
    out = l.download()
    l.close()
    return out


@pytest.mark.parametrize("w,h", SIZES_HELPERS)
def test_alpha_and_colorspace_helpers(ctx, oracle, w, h):
    img = random_premul(w, h, 1)
    assert_exact(_run(ctx, img, lambda f, l: f.demultiply_alpha(l)), oracle.demultiply_alpha(img), "demultiply")
    un = random_rgba(w, h, 2)
    assert_exact(_run(ctx, un, lambda f, l: f.multiply_alpha(l)), oracle.multiply_alpha(un), "multiply")
    assert_exact(_run(ctx, img, lambda f, l: f.into_linear_rgb(l)), oracle.into_linear_rgb(img), "into_linear")
    assert_exact(_run(ctx, img, lambda f, l: f.into_srgb(l)), oracle.into_srgb(img), "into_srgb")
    # non-premultiplied garbage must follow the same saturating arithmetic
    assert_exact(_run(ctx, un, lambda f, l: f.into_linear_rgb(l)), oracle.into_linear_rgb(un), "into_linear(unpremul)")
    assert_exact(_run(ctx, un, lambda f, l: f.demultiply_alpha(l)), oracle.demultiply_alpha(un), "demultiply(unpremul)")


def test_every_alpha_channel_pair_through_the_tables(ctx, oracle):
    """All 65536 (alpha, channel) pairs, premultiplied or not, on a layer large enough for the table kernels."""
    a, c = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    img = np.zeros((256, 512, 4), dtype=np.uint8)
    img[:, :256, 3] = a
    img[:, :256, 0] = c
    img[:, :256, 1] = 255 - c
    img[:, :256, 2] = (c * 7) % 256
    img[:, 256:] = img[:, :256][:, ::-1]
    assert_exact(_run(ctx, img, lambda f, l: f.demultiply_alpha(l)), oracle.demultiply_alpha(img), "demultiply")
    assert_exact(_run(ctx, img, lambda f, l: f.into_linear_rgb(l)), oracle.into_linear_rgb(img), "into_linear")
    assert_exact(_run(ctx, img, lambda f, l: f.into_srgb(l)), oracle.into_srgb(img), "into_srgb")


def test_colorspace_round_trip_full_range(ctx, oracle):
    # every (c, a) pair with c <= a
    a, c = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    img = np.zeros((256, 256, 4), dtype=np.uint8)
    img[..., 3] = a
    img[..., 0] = np.minimum(c, a)
    img[..., 1] = np.minimum(255 - c, a)
    img[..., 2] = np.minimum((c * 7) % 256, a)
    assert_exact(_run(ctx, img, lambda f, l: f.into_linear_rgb(l)), oracle.into_linear_rgb(img))
    assert_exact(_run(ctx, img, lambda f, l: f.into_srgb(l)), oracle.into_srgb(img))


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("sx,sy", [(2.0, 2.0), (4.0, 3.0), (8.0, 0.0), (0.0, 16.0), (64.0, 64.0), (2.5, 40.0)])
def test_box_blur(ctx, oracle, w, h, sx, sy):
    img = random_premul(w, h, 3, sparse=True)
    assert_exact(_run(ctx, img, lambda f, l: f.box_blur(sx, sy, l)), oracle.box_blur(sx, sy, img), f"box {sx},{sy}")


def test_box_blur_radius_larger_than_image(ctx, oracle):
    img = random_premul(40, 9, 4)
    assert_exact(_run(ctx, img, lambda f, l: f.box_blur(30.0, 30.0, l)), oracle.box_blur(30.0, 30.0, img))
    # radius too large for the tiled horizontal kernel -> fallback kernel
    img = random_premul(700, 16, 5)
    assert_exact(_run(ctx, img, lambda f, l: f.box_blur(300.0, 2.0, l)), oracle.box_blur(300.0, 2.0, img))


def test_box_blur_radius_sweep(ctx, oracle):
    """Every box radius the integer-quotient passes can select (and the first ones beyond, which take the float kernels),
    on dense random pixels, odd and even widths."""
    for w, h in ((260, 70), (131, 33)):
        img = random_premul(w, h, 11)
        for sigma in [2.0 + 3.1 * k for k in range(0, 56)]:
            assert_exact(_run(ctx, img, lambda f, l: f.box_blur(sigma, sigma * 0.7, l)), oracle.box_blur(sigma, sigma * 0.7, img),
                         f"box sweep {sigma} {w}x{h}")


def test_box_blur_large(ctx, oracle):
    img = random_premul(2048, 1024, 6, sparse=True)
    for s in (2.0, 8.0, 64.0):
        assert_exact(_run(ctx, img, lambda f, l: f.box_blur(s, s, l)), oracle.box_blur(s, s, img), f"box {s}")


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("sx,sy", [(0.5, 0.5), (1.0, 1.9), (1.9, 0.0), (0.0, 1.2)])
def test_iir_blur(ctx, oracle, w, h, sx, sy):
    img = random_premul(w, h, 7, sparse=True)
    # Sequential recurrences in the reference order: bit-exact.
    want = oracle.iir_blur(sx, sy, img)
    assert_exact(_run(ctx, img, lambda f, l: f.iir_blur(sx, sy, l)), want, f"iir exact {sx},{sy}")
    # the opt-in f32 kernel (segments with halos): BASELINE.json grants the IIR blur 1/255 on the stage output
    got = _run(ctx, img, lambda f, l: f.iir_blur_fast(sx, sy, l))
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1, f"iir fast {sx},{sy}: max diff {d.max()} at {np.argwhere(d > 1)[:3].tolist()}"


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(1, 1), (5, 300), (300, 5), (223, 225), (449, 1000), (2048, 700)])
@pytest.mark.parametrize("sx,sy", [(0.05, 1.99), (1.99, 1.99), (1.0, 0.0), (0.0, 0.7), (0.3, 1.2)])
def test_iir_blur_fast_segments_and_borders(ctx, oracle, w, h, sx, sy):
    """Segment seams (multiples of the interior length), sizes around the 224-sample line, single rows / columns, one axis off."""
    img = random_premul(w, h, w * 7 + h, sparse=False)
    want = oracle.iir_blur(sx, sy, img)
    got = _run(ctx, img, lambda f, l: f.iir_blur_fast(sx, sy, l))
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1, f"{w}x{h} sigma {sx},{sy}: max diff {d.max()} at {np.argwhere(d > 1)[:3].tolist()}"
    assert (d > 0).mean() < 0.05


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("op", ["erode", "dilate"])
@pytest.mark.parametrize("rx,ry", [(1.0, 1.0), (3.0, 2.0), (0.4, 7.5), (32.0, 8.0)])
def test_morphology(ctx, oracle, w, h, op, rx, ry):
    if w * h > 100000 and rx * ry > 100:
        pytest.skip("oracle is O(rx*ry) per pixel")
    img = random_premul(w, h, 8, sparse=True)
    assert_exact(_run(ctx, img, lambda f, l: f.morphology(op, rx, ry, l)), oracle.morphology(op, rx, ry, img),
                 f"morph {op} {rx},{ry}")


def test_morphology_wide_windows_as_a_chain_of_tile_passes(ctx, oracle):
    """Windows wider than the tile kernel's 16 taps are applied as a chain of narrower windows (exact for min / max, also at
    the image borders); windows that reach the image size, odd sizes and different radii per axis included."""
    for (w, h, rx, ry) in [(200, 120, 32.0, 32.0), (97, 61, 20.0, 9.0), (150, 40, 8.5, 60.0), (64, 64, 100.0, 3.0), (300, 9, 127.0, 1.0)]:
        img = random_premul(w, h, 31, sparse=True)
        for op in ("erode", "dilate"):
            assert_exact(_run(ctx, img, lambda f, l: f.morphology(op, rx, ry, l)), oracle.morphology(op, rx, ry, img),
                         f"morph {op} {rx},{ry} on {w}x{h}")


KERNELS = {
    "sharpen3": ([0, -1, 0, -1, 5, -1, 0, -1, 0], 3, 3, 1, 1, 1.0, 0.0),
    "emboss3": ([-2, -1, 0, -1, 1, 1, 0, 1, 2], 3, 3, 1, 1, 1.0, 0.5),
    "blur5": ([1] * 25, 5, 5, 2, 2, 25.0, 0.0),
    "asym": ([1, 2, 3, 4, 5, 6], 3, 2, 0, 1, 3.0, 0.1),
    "wide": ([0.5, -1, 2, 0.25, 1, 1, -0.5], 7, 1, 6, 0, 2.0, 0.0),
}


@pytest.mark.parametrize("w,h", [(3, 2), (64, 48), (257, 131)])
@pytest.mark.parametrize("kname", list(KERNELS))
@pytest.mark.parametrize("edge", ["none", "duplicate", "wrap"])
@pytest.mark.parametrize("preserve", [False, True])
def test_convolve_matrix(ctx, oracle, w, h, kname, edge, preserve):
    k, cols, rows, tx, ty, div, bias = KERNELS[kname]
    img = random_premul(w, h, 9)
    got = _run(ctx, img, lambda f, l: f.convolve_matrix(k, cols, rows, tx, ty, div, bias, edge, preserve, l))
    assert_exact(got, oracle.convolve_matrix(k, cols, rows, tx, ty, div, bias, edge, preserve, img),
                 f"convolve {kname} {edge} {preserve}")


@pytest.mark.parametrize("w,h", [(7, 5), (257, 131)])
def test_color_matrix(ctx, oracle, w, h):
    img = random_rgba(w, h, 10)
    rng = np.random.default_rng(11)
    m = (rng.random(20) * 2 - 0.7).astype(np.float32)
    cases = [("matrix", m), ("saturate", [0.3]), ("saturate", [1.7]), ("hueRotate", [90.0]), ("hueRotate", [-37.5]),
             ("luminanceToAlpha", [])]
    for kind, params in cases:
        got = _run(ctx, img, lambda f, l: f.color_matrix(kind, params, l))
        assert_exact(got, oracle.color_matrix(kind, params, img), f"color_matrix {kind}")


@pytest.mark.parametrize("w,h", [(7, 5), (257, 131)])
def test_component_transfer(ctx, oracle, w, h):
    import resvg_b200 as rb

    img = random_rgba(w, h, 12)
    specs = [
        dict(kind="table", values=[0.0, 1.0, 0.2, 0.9]),
        dict(kind="discrete", values=[0.1, 0.5, 0.9]),
        dict(kind="linear", slope=1.5, intercept=-0.2),
        dict(kind="gamma", amplitude=0.9, exponent=2.2, offset=0.05),
        dict(kind="identity"),
        dict(kind="table", values=[]),
        dict(kind="table", values=[0.7]),
    ]
    for i in range(len(specs)):
        four = [specs[(i + j) % len(specs)] for j in range(4)]
        got = _run(ctx, img, lambda f, l: f.component_transfer([rb.make_transfer(**s) for s in four], l))
        want = oracle.component_transfer([oracle.make_transfer(**s) for s in four], img)
        assert_exact(got, want, f"component_transfer {i}")


@pytest.mark.parametrize("w,h", [(7, 5), (257, 131), (1029, 67)])
@pytest.mark.parametrize("k", [(0.5, 0.5, 0.5, 0.0), (1.0, 0.0, 0.0, 0.0), (0.0, 1.0, -1.0, 0.1), (0.0, 0.0, 0.0, 0.0),
                               (-0.3, 0.2, 1.4, -0.05)])
def test_composite_arithmetic(ctx, oracle, w, h, k):
    import resvg_b200 as rb

    a, b = random_premul(w, h, 13, sparse=True), random_premul(w, h, 14, sparse=True)
    la, lb, ld = ctx.layer_from(a), ctx.layer_from(b), ctx.layer(w, h)
    rb.filters.arithmetic(*k, la, lb, ld)
    assert_exact(ld.download(), oracle.arithmetic(*k, a, b), f"arithmetic {k}")


@pytest.mark.parametrize("w,h", [(7, 5), (257, 131)])
@pytest.mark.parametrize("xch,ych,scale,s", [(0, 1, 20.0, 1.0), (3, 3, 50.0, 1.5), (2, 0, -7.3, 0.5)])
def test_displacement_map(ctx, oracle, w, h, xch, ych, scale, s):
    import resvg_b200 as rb

    src, mp = random_premul(w, h, 15), random_rgba(w, h, 16)
    ls, lm, ld = ctx.layer_from(src), ctx.layer_from(mp), ctx.layer(w, h)
    rb.filters.displacement_map(xch, ych, scale, s * scale, s * scale, ls, lm, ld)
    assert_exact(ld.download(), oracle.displacement_map(xch, ych, scale, s * scale, s * scale, src, mp), "displacement")


LIGHTS = {
    "distant": dict(kind="distant", azimuth=45.0, elevation=60.0),
    "distant_flat": dict(kind="distant", azimuth=200.0, elevation=5.0),
    "point": dict(kind="point", x=40.0, y=30.0, z=25.0),
    "spot": dict(kind="spot", x=10.0, y=10.0, z=40.0, points_at=(60.0, 50.0, 0.0), specular_exponent=8.0),
    "spot_cone": dict(kind="spot", x=10.0, y=10.0, z=40.0, points_at=(60.0, 50.0, 0.0), specular_exponent=1.0,
                      limiting_cone_angle=25.0),
}


@pytest.mark.parametrize("w,h", [(2, 9), (3, 3), (64, 48), (257, 131)])
@pytest.mark.parametrize("lname", list(LIGHTS))
def test_diffuse_lighting(ctx, oracle, w, h, lname):
    import resvg_b200 as rb

    src = smooth_alpha(w, h, 17)
    ls, ld = ctx.layer_from(src), ctx.layer(w, h)
    rb.filters.diffuse_lighting(5.0, 1.2, (255, 200, 90), rb.make_light(**LIGHTS[lname]), ls, ld)
    want = oracle.diffuse_lighting(5.0, 1.2, (255, 200, 90), oracle.make_light(**LIGHTS[lname]), src)
    if lname == "spot":  # device powf for the cone exponent: <= 1/255
        assert_within(ld.download(), want, 1, f"diffuse {lname}")
    else:
        assert_exact(ld.download(), want, f"diffuse {lname}")


@pytest.mark.parametrize("w,h", [(3, 3), (64, 48), (257, 131)])
@pytest.mark.parametrize("lname", list(LIGHTS))
@pytest.mark.parametrize("exponent", [1.0, 20.0])
def test_specular_lighting(ctx, oracle, w, h, lname, exponent):
    import resvg_b200 as rb

    src = smooth_alpha(w, h, 18)
    ls, ld = ctx.layer_from(src), ctx.layer(w, h)
    rb.filters.specular_lighting(5.0, 1.1, exponent, (255, 255, 255), rb.make_light(**LIGHTS[lname]), ls, ld)
    want = oracle.specular_lighting(5.0, 1.1, exponent, (255, 255, 255), oracle.make_light(**LIGHTS[lname]), src)
    if exponent != 1.0 or lname == "spot":  # powf on the device: tolerance 1/255 (north star, highp float stages)
        assert_within(ld.download(), want, 1, f"specular {lname} {exponent}")
    else:
        assert_exact(ld.download(), want, f"specular {lname} {exponent}")


@pytest.mark.parametrize("w,h", [(7, 5), (200, 120)])
@pytest.mark.parametrize("bf,octaves,fractal,stitch", [
    ((0.05, 0.05), 1, False, False), ((0.01, 0.03), 4, True, False), ((0.05, 0.02), 3, False, True),
    ((0.1, 0.1), 2, True, True), ((0.0, 0.04), 2, True, True)])
def test_turbulence(ctx, oracle, w, h, bf, octaves, fractal, stitch):
    import resvg_b200 as rb

    ld = ctx.layer(w, h)
    args = (6.0 - 1.5, 6.0 - 0.25, 1.5, 1.5, bf[0], bf[1], octaves, 7, stitch, fractal)
    rb.filters.turbulence(*args, ld)
    assert_exact(ld.download(), oracle.turbulence(*args, w, h), f"turbulence {bf} {octaves} {fractal} {stitch}")
    for seed in (0, -5, 12345):
        args2 = (0.0, 0.0, 1.0, 2.0, bf[0], bf[1], octaves, seed, stitch, fractal)
        rb.filters.turbulence(*args2, ld)
        assert_exact(ld.download(), oracle.turbulence(*args2, w, h), f"turbulence seed {seed}")


def test_filter_chain_matches_oracle(ctx, oracle):
    """C3 chain at a size the oracle finishes in seconds."""
    import resvg_b200 as rb

    f = rb.filters
    w, h = 512, 384
    img = random_premul(w, h, 19, sparse=True)
    l = ctx.layer_from(img)
    f.into_linear_rgb(l)
    f.box_blur(8.0, 8.0, l)
    f.morphology("dilate", 3.0, 3.0, l)
    sharpen = [0, -1, 0, -1, 5, -1, 0, -1, 0]
    f.convolve_matrix(sharpen, 3, 3, 1, 1, 1.0, 0.0, "duplicate", False, l)
    t = ctx.layer(w, h)
    f.turbulence(0.0, 0.0, 1.0, 1.0, 0.02, 0.02, 3, 7, False, True, t)
    f.multiply_alpha(t)
    c = ctx.layer(w, h)
    f.arithmetic(0.5, 0.5, 0.5, 0.0, l, t, c)
    lit = ctx.layer(w, h)
    f.diffuse_lighting(5.0, 1.0, (255, 255, 255), rb.make_light(kind="distant", azimuth=45.0, elevation=60.0), c, lit)
    f.box_blur(64.0, 64.0, lit)
    f.into_srgb(lit)
    got = lit.download()

    o = oracle
    x = o.into_linear_rgb(img)
    x = o.box_blur(8.0, 8.0, x)
    x = o.morphology("dilate", 3.0, 3.0, x)
    x = o.convolve_matrix(sharpen, 3, 3, 1, 1, 1.0, 0.0, "duplicate", False, x)
    tt = o.multiply_alpha(o.turbulence(0.0, 0.0, 1.0, 1.0, 0.02, 0.02, 3, 7, False, True, w, h))
    cc = o.arithmetic(0.5, 0.5, 0.5, 0.0, x, tt)
    ll = o.diffuse_lighting(5.0, 1.0, (255, 255, 255), o.make_light(kind="distant", azimuth=45.0, elevation=60.0), cc)
    ll = o.into_srgb(o.box_blur(64.0, 64.0, ll))
    assert_exact(got, ll, "filter chain")


@pytest.mark.gpu
def test_filter_chain_on_strips_with_halo_equals_the_whole_layer(ctx, oracle):
    """Canvas-strip sharding of a filter chain (shard.strip_with_halo): every strip is filtered together with the halo rows
    the chain reads (box blur: sum of its radii; dilate 3: 3 + 3; convolve 3x3: 2; lighting: 1) and its own rows are
    bit-identical to the chain over the whole layer — redundant halo work instead of an exchange."""
    import resvg_b200 as rb
    from resvg_b200 import shard
    F = rb.filters
    W, H = 320, 1000
    img = random_premul(W, H, 21, sparse=True)
    light = rb.make_light("distant", azimuth=45.0, elevation=60.0)

    def chain(l):
        F.box_blur(3.0, 3.0, l)
        F.morphology("dilate", 3.0, 3.0, l)
        F.convolve_matrix([0, -1, 0, -1, 5, -1, 0, -1, 0], 3, 3, 1, 1, 1.0, 0.0, "duplicate", False, l)
        out = ctx.layer(l.width, l.height)
        F.diffuse_lighting(5.0, 1.0, (255, 255, 255), light, l, out)
        F.box_blur(9.0, 9.0, out)
        return out

    want = chain(ctx.layer_from(img)).download()
    halo = F.box_blur_reach(3.0) + 6 + 2 + 1 + F.box_blur_reach(9.0)
    assert F.box_blur_reach(3.0) == sum((b - 1) // 2 for b in oracle.create_box_gauss(3.0))
    for world in (2, 3):
        got = np.zeros_like(want)
        for r in range(world):
            lo, n, off, rows = shard.strip_with_halo(H, r, world, halo)
            part = chain(ctx.layer_from(np.ascontiguousarray(img[lo:lo + n]))).download()
            got[lo + off:lo + off + rows] = part[off:off + rows]
        assert_exact(got, want, f"{world} strips with a {halo}-row halo")
