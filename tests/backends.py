"""Two interchangeable drawing back ends for the test-side renderer (tests/svgfront.py): the CPU oracle and the CUDA
library.  Both expose the tiny-skia / filter operations resvg's render traversal issues, with identical signatures,
so one traversal can drive either and the results can be compared byte for byte."""
import numpy as np


class OracleBackend:
    name = "oracle"

    def __init__(self):
        from tests import oracle_ffi as O
        from tests import oracle_raster as R
        self.O, self.R = O, R

    # ---- layers (numpy arrays) ----
    def new_layer(self, w, h):
        return np.zeros((h, w, 4), np.uint8)

    def clone(self, l):
        return l.copy()

    def to_numpy(self, l):
        return l

    def size(self, l):
        return l.shape[1], l.shape[0]

    def fill_color(self, l, r, g, b, a):  # non-premultiplied floats (Pixmap::fill)
        self.R.pixmap_fill(l, r, g, b, a)

    def fill_path(self, l, verbs, pts, spec, rule, ts, blend="source_over", aa=True):
        if spec["kind"] == "pattern":
            spec = dict(spec, pixmap=spec["layer"])
        self.R.fill_path(l, verbs, pts, self.R.make_paint(spec, blend, aa), rule, ts)

    def draw_layer(self, dst, src, x, y, opacity=1.0, blend="source_over"):
        self.R.draw_pixmap(dst, x, y, src, opacity, blend)

    # ---- masks ----
    def mask_from_layer(self, l, kind):
        return self.R.mask_from_pixmap(l, kind)

    def mask_new(self, w, h):
        return np.zeros((h, w), np.uint8)

    def mask_fill_path(self, m, verbs, pts, rule, aa, ts):
        self.R.mask_fill_path(m, verbs, pts, rule, aa, ts)

    def mask_invert(self, m):
        self.R.mask_invert(m)

    def apply_mask(self, l, m):
        self.R.apply_mask(l, m)

    # ---- filters: in place unless they return a new layer ----
    def f(self, name, l, *args):
        O = self.O
        out = getattr(O, name)(*args, l) if name not in ("into_linear_rgb", "into_srgb", "multiply_alpha", "demultiply_alpha") else getattr(O, name)(l)
        l[...] = out

    def arithmetic(self, k, a, b):
        return self.O.arithmetic(*k, a, b)

    def displacement_map(self, xch, ych, scale, sx, sy, src, mp):
        return self.O.displacement_map(xch, ych, scale, sx, sy, src, mp)

    def diffuse_lighting(self, ss, kd, color, light, src):
        return self.O.diffuse_lighting(ss, kd, color, self.O.make_light(**light), src)

    def specular_lighting(self, ss, ks, exp, color, light, src):
        return self.O.specular_lighting(ss, ks, exp, color, self.O.make_light(**light), src)

    def turbulence(self, w, h, *args):
        return self.O.turbulence(*args, w, h)

    def component_transfer(self, l, funcs):
        l[...] = self.O.component_transfer([self.O.make_transfer(**f) for f in funcs], l)


class GpuBackend:
    name = "gpu"

    def __init__(self, ctx):
        import resvg_b200 as rb
        self.rb, self.ctx = rb, ctx

    def new_layer(self, w, h):
        return self.ctx.layer(w, h)

    def clone(self, l):
        return l.clone()

    def to_numpy(self, l):
        return l.download()

    def size(self, l):
        return l.width, l.height

    def fill_color(self, l, r, g, b, a):
        # Pixmap::fill(color): premultiply in f32, (c * 255 + 0.5) as u8 — host constants
        f = np.float32
        if a == 1.0:
            pr, pg, pb = f(r), f(g), f(b)
        else:
            pr, pg, pb = [min(max(f(c) * f(a), f(0)), f(1)) for c in (r, g, b)]
        q = [int(f(c) * f(255.0) + f(0.5)) for c in (pr, pg, pb, f(a))]
        l.fill(*q)

    def fill_path(self, l, verbs, pts, spec, rule, ts, blend="source_over", aa=True):
        self.rb.fill_path(l, verbs, pts, self.rb.make_paint(spec, blend, aa), rule, ts)

    def draw_layer(self, dst, src, x, y, opacity=1.0, blend="source_over"):
        self.rb.draw_layer(dst, src, x, y, opacity, blend)

    def mask_from_layer(self, l, kind):
        return self.rb.Mask.from_layer(l, kind)

    def mask_new(self, w, h):
        return self.rb.Mask(self.ctx, w, h)

    def mask_fill_path(self, m, verbs, pts, rule, aa, ts):
        m.fill_path(verbs, pts, rule, aa, ts)

    def mask_invert(self, m):
        m.invert()

    def apply_mask(self, l, m):
        self.rb.apply_mask(l, m)

    def f(self, name, l, *args):
        fn = getattr(self.rb.filters, name)
        if name in ("into_linear_rgb", "into_srgb", "multiply_alpha", "demultiply_alpha"):
            fn(l)
        else:
            fn(*args, l)

    def arithmetic(self, k, a, b):
        d = self.ctx.layer(a.width, a.height)
        self.rb.filters.arithmetic(*k, a, b, d)
        return d

    def displacement_map(self, xch, ych, scale, sx, sy, src, mp):
        d = self.ctx.layer(src.width, src.height)
        self.rb.filters.displacement_map(xch, ych, scale, sx, sy, src, mp, d)
        return d

    def diffuse_lighting(self, ss, kd, color, light, src):
        d = self.ctx.layer(src.width, src.height)
        self.rb.filters.diffuse_lighting(ss, kd, color, self.rb.make_light(**light), src, d)
        return d

    def specular_lighting(self, ss, ks, exp, color, light, src):
        d = self.ctx.layer(src.width, src.height)
        self.rb.filters.specular_lighting(ss, ks, exp, color, self.rb.make_light(**light), src, d)
        return d

    def turbulence(self, w, h, *args):
        d = self.ctx.layer(w, h)
        self.rb.filters.turbulence(*args, d)
        return d

    def component_transfer(self, l, funcs):
        self.rb.filters.component_transfer([self.rb.make_transfer(**f) for f in funcs], l)
