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

    def stroke_hairline(self, l, verbs, pts, spec, ts, blend, width, cap, dash=None, dash_offset=0.0):
        """tiny-skia painter.rs stroke_path for a stroke treat_as_hairline accepts (anti-aliased, transformed width <= 1
        px): dash (unless the caller already did), transform into device space, fold the hairline coverage into the paint's alpha,
        walk the segments (oracle/hairline.c) and blend every blit with the oracle's pipeline."""
        from tests import geom
        f = np.float32

        def fast_len(x, y):
            x, y = abs(f(x)), abs(f(y))
            if x < y:
                x, y = y, x
            return f(x + y * f(0.5))

        w = f(width)
        len0, len1 = fast_len(f(ts[0]) * w, f(ts[1]) * w), fast_len(f(ts[2]) * w, f(ts[3]) * w)
        coverage = f((len0 + len1) * f(0.5))
        v, p = np.asarray(verbs, np.uint8), np.asarray(pts, np.float32).reshape(-1, 2)
        if dash:
            import math
            sx, sy = math.hypot(ts[0], ts[2]), math.hypot(ts[1], ts[3])
            res = max(sx, sy) if (math.isfinite(sx) and math.isfinite(sy) and max(sx, sy) > 0) else 1.0
            out = geom.dash_path(v, p, dash, dash_offset, res)
            if out is None:
                return
            v, p = out
        sx_, ky, kx, sy_, tx, ty = [f(t) for t in ts]
        dev = np.stack([p[:, 0] * sx_ + p[:, 1] * kx + tx, p[:, 0] * ky + p[:, 1] * sy_ + ty], axis=1).astype(np.float32) \
            if tuple(ts) != (1.0, 0.0, 0.0, 1.0, 0.0, 0.0) else p
        if spec["kind"] == "pattern":
            spec = dict(spec, pixmap=spec["layer"])
        pre_scales = blend in ("destination", "destination_over", "plus", "destination_out", "source_atop", "source_over", "xor")
        if coverage != 1.0 and pre_scales:
            scale = int(f(coverage) * f(256.0))
            opacity = f(f((255 * scale) >> 8) / f(255.0))
            clamp = lambda a: float(min(max(f(a) * opacity, f(0.0)), f(1.0)))
            if spec["kind"] == "solid":
                c = list(spec["color"])
                c[3] = clamp(c[3])
                spec = dict(spec, color=c)
            elif spec["kind"] == "pattern":
                spec = dict(spec, opacity=clamp(spec.get("opacity", 1.0)))
            else:
                st = np.array(spec["stops"], dtype=np.float32).reshape(-1, 5).copy()
                st[:, 4] = [clamp(a) for a in st[:, 4]]
                spec = dict(spec, stops=st)
        h, wpx = l.shape[:2]
        paint = self.R.make_paint(spec, blend, True)
        for ty in range(0, h, 8191):  # DrawTiler: the painter draws layers larger than 8191 px tile by tile
            for tx in range(0, wpx, 8191):
                blits = geom.hairline_blits(v, dev - np.float32([tx, ty]), cap, min(wpx - tx, 8191), min(h - ty, 8191))
                if len(blits):
                    blits[:, 0] += tx
                    blits[:, 1] += ty
                    self.R.blit_coverage(l, blits, paint, ts)

    def upload(self, l, px):
        l[...] = px

    def fill_rect(self, l, x, y, w, h, spec, ts, blend="source_over", aa=True):
        if spec["kind"] == "pattern":
            spec = dict(spec, pixmap=spec["layer"])
        self.R.fill_rect(l, x, y, w, h, self.R.make_paint(spec, blend, aa), ts)

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

    def stroke_hairline(self, l, verbs, pts, spec, ts, blend, width, cap):
        b = self.rb.Batch(l)
        b.stroke_path(verbs, pts, self.rb.make_paint(spec, blend, True), width, 4.0, cap, "miter", ts)
        b.submit()
        b.close()

    def upload(self, l, px):
        l.upload(px)

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
