// fill_core.h — the front end of fill_path up to (not including) the scanline walk: bounds and clip decisions, path
// transform, y-monotone chopping, clipping against the pixmap, fixed-point line edges and curve records.
//
// Semantics follow tiny-skia 0.12.0 (painter.rs fill_path, scan/path.rs, scan/path_aa.rs, edge_builder.rs,
// edge_clipper.rs, line_clipper.rs, edge.rs; tiny-skia-path path_geometry.rs — a Rust port of Skia's
// SkEdgeBuilder / SkEdgeClipper / SkEdge), reached from crates/resvg/src/path.rs:73 (`pixmap.fill_path`).
//
// One source for both sides (geom_common.h): raster_host.cpp instantiates it with std::vector for the host builder,
// geo.cu with DVec for the device geometry kernels.  `Pts` is anything indexable that yields the path's points in
// device space, relative to the DrawTiler tile (a pointer, or a functor that maps them on the fly).
#pragma once

#include "batch.h"
#include "edge_math.h"
#include "geom_common.h"
#include "raster_host.h"

namespace geo {
namespace fl {

using rbh::CurveRec;
using rbh::DrawGeom;
using rbh::Edge;
using rbh::IRect;
using rbe::fdot6_round;
using rbe::shl;

constexpr float kNearlyZero = 1.0f / 4096.0f;
GEO_HDI inline bool nearly_zero(float v, float tol = kNearlyZero) { return fabsf(v) <= tol; }

// ---------------------------------------------------------------------------------------------------
// curve chopping (tiny-skia-path path_geometry.rs)
// ---------------------------------------------------------------------------------------------------
GEO_HDI inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }
GEO_HDI inline P lerp(P a, P b, float t) { return P{lerpf(a.x, b.x, t), lerpf(a.y, b.y, t)}; }

GEO_HDI inline bool unit_divide(float numer, float denom, float *ratio)
{
    if (numer < 0) { numer = -numer; denom = -denom; }
    if (denom == 0 || numer == 0 || numer >= denom) return false;
    float r = numer / denom;
    if (!(r > 0.0f && r < 1.0f)) return false;
    *ratio = r;
    return true;
}

GEO_HD inline int unit_quad_roots(float a, float b, float c, float roots[2])
{
    if (a == 0) return unit_divide(-c, b, roots) ? 1 : 0;
    double dr = (double)b * b - 4.0 * (double)a * c;
    if (dr < 0) return 0;
    float r = (float)sqrt(dr);
    if (!gfinite(r)) return 0;
    float q = (b < 0) ? -(b - r) / 2 : -(b + r) / 2;
    int n = 0;
    if (unit_divide(q, a, roots + n)) n++;
    if (unit_divide(c, q, roots + n)) n++;
    if (n == 2) {
        if (roots[0] > roots[1]) gswap(roots[0], roots[1]);
        else if (roots[0] == roots[1]) n = 1;
    }
    return n;
}

GEO_HD inline void split_quad(const P s[3], float t, P d[5])
{
    P p01 = lerp(s[0], s[1], t), p12 = lerp(s[1], s[2], t);
    d[0] = s[0]; d[1] = p01; d[2] = lerp(p01, p12, t); d[3] = p12; d[4] = s[2];
}
GEO_HD inline void split_cubic(const P s[4], float t, P d[7])
{
    P ab = lerp(s[0], s[1], t), bc = lerp(s[1], s[2], t), cd = lerp(s[2], s[3], t);
    P abc = lerp(ab, bc, t), bcd = lerp(bc, cd, t);
    d[0] = s[0]; d[1] = ab; d[2] = abc; d[3] = lerp(abc, bcd, t); d[4] = bcd; d[5] = cd; d[6] = s[3];
}

template <int AXIS> GEO_HDI inline float &ax(P &p) { return AXIS ? p.y : p.x; }
template <int AXIS> GEO_HDI inline float axv(const P &p) { return AXIS ? p.y : p.x; }

template <int AXIS> GEO_HD int quad_extrema(const P s[3], P d[5])
{
    float a = axv<AXIS>(s[0]), b = axv<AXIS>(s[1]), c = axv<AXIS>(s[2]);
    float ab = a - b, bc = b - c;
    if (ab < 0) bc = -bc;
    if (ab == 0 || bc < 0) { // not monotonic
        float t;
        if (unit_divide(a - b, a - b - b + c, &t)) {
            split_quad(s, t, d);
            ax<AXIS>(d[1]) = axv<AXIS>(d[2]);
            ax<AXIS>(d[3]) = axv<AXIS>(d[2]);
            return 1;
        }
        b = fabsf(a - b) < fabsf(b - c) ? a : c;
    }
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
    ax<AXIS>(d[1]) = b;
    return 0;
}

template <int AXIS> GEO_HD int cubic_extrema(const P s[4], P d[10])
{
    float a = axv<AXIS>(s[0]), b = axv<AXIS>(s[1]), c = axv<AXIS>(s[2]), e = axv<AXIS>(s[3]);
    float tv[2];
    int roots = unit_quad_roots(e - a + 3 * (b - c), 2 * (a - b - b + c), b - a, tv);
    if (roots == 0) { memcpy(d, s, 4 * sizeof(P)); return 0; }
    P src[4];
    memcpy(src, s, sizeof(src));
    P *dst = d;
    float t = tv[0];
    for (int i = 0; i < roots; i++) {
        split_cubic(src, t, dst);
        if (i == roots - 1) break;
        dst += 3;
        memcpy(src, dst, sizeof(src));
        if (!unit_divide(tv[i + 1] - tv[i], 1.0f - tv[i], &t)) {
            dst[4] = dst[5] = dst[6] = src[3];
            break;
        }
    }
    ax<AXIS>(d[2]) = axv<AXIS>(d[3]);
    ax<AXIS>(d[4]) = axv<AXIS>(d[3]);
    if (roots == 2) {
        ax<AXIS>(d[5]) = axv<AXIS>(d[6]);
        ax<AXIS>(d[7]) = axv<AXIS>(d[6]);
    }
    return roots;
}

// ---------------------------------------------------------------------------------------------------
// edge emission (tiny-skia edge.rs) — curves are expanded to their line edges immediately
// ---------------------------------------------------------------------------------------------------
template <template <class> class Vec> struct Sink {
    Vec<Edge> *out;
    size_t base;  // first edge of this draw
    int shift;
    Vec<uint8_t> kinds; // per emitted edge: 0 = from a line, 1 = from a curve (combine_vertical only
                                // ever looks at a preceding *line* edge)
    // item mode: curves are recorded (FDot6 control points + subdivision count) instead of being expanded; `order`
    // of a line / `item` of a curve is then the emission index shared by both kinds
    Vec<CurveRec> *curves = nullptr;
    uint32_t n_items = 0;

    GEO_HDI uint32_t next_order() { return curves ? n_items++ : (uint32_t)(out->size() - base); }

    // LineEdge::new / update tail: FDot6 end points, y0 <= y1
    GEO_HD bool emit(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int winding, Edge *e)
    {
        rbe::RawEdge r;
        if (!rbe::line_edge(x0, y0, x1, y1, &r)) return false;
        e->x = r.x;
        e->dx = r.dx;
        e->first_y = r.first_y;
        e->last_y = r.last_y;
        e->winding = winding;
        e->prev = -1;
        e->before = 0;
        e->order = 0;
        return true;
    }

    GEO_HD void line(P p0, P p1)
    {
        float scale = (float)(1 << (shift + 6));
        int32_t x0 = f2i(p0.x * scale), y0 = f2i(p0.y * scale), x1 = f2i(p1.x * scale), y1 = f2i(p1.y * scale);
        int w = 1;
        if (y0 > y1) { gswap(x0, x1); gswap(y0, y1); w = -1; }
        Edge e;
        if (!emit(x0, y0, x1, y1, w, &e)) return;
        if (e.dx == 0 && !kinds.empty() && kinds.back() == 0) {
            int c = combine_vertical(e, out->back());
            if (c == 2) { out->pop_back(); kinds.pop_back(); return; }
            if (c == 1) return;
        }
        e.order = next_order();
        out->push_back(e);
        kinds.push_back(0);
    }

    // edge_builder.rs combine_vertical: 0 no, 1 partial, 2 total
    GEO_HD static int combine_vertical(const Edge &edge, Edge &last)
    {
        if (last.dx != 0 || edge.x != last.x) return 0;
        if (edge.winding == last.winding) {
            if (edge.last_y + 1 == last.first_y) { last.first_y = edge.first_y; return 1; }
            if (edge.first_y == last.last_y + 1) { last.last_y = edge.last_y; return 1; }
            return 0;
        }
        if (edge.first_y == last.first_y) {
            if (edge.last_y == last.last_y) return 2;
            if (edge.last_y < last.last_y) { last.first_y = edge.last_y + 1; return 1; }
            last.first_y = last.last_y + 1;
            last.last_y = edge.last_y;
            last.winding = edge.winding;
            return 1;
        }
        if (edge.last_y == last.last_y) {
            if (edge.first_y > last.first_y) last.last_y = edge.first_y - 1;
            else {
                last.last_y = last.first_y - 1;
                last.first_y = edge.first_y;
                last.winding = edge.winding;
            }
            return 1;
        }
        return 0;
    }

    GEO_HD void push_segment(const rbe::RawEdge &r, int w, int32_t *link)
    {
        Edge e;
        e.x = r.x; e.dx = r.dx; e.first_y = r.first_y; e.last_y = r.last_y;
        e.winding = w;
        e.prev = *link;
        e.before = 0;
        e.order = (uint32_t)(out->size() - base);
        *link = (int32_t)e.order;
        out->push_back(e);
        kinds.push_back(1);
    }

    // QuadraticEdge::new + every update(): emits one line edge per non-degenerate segment.
    GEO_HD void quad(const P p[3])
    {
        float scale = (float)(1 << (shift + 6));
        int32_t x0 = f2i(p[0].x * scale), y0 = f2i(p[0].y * scale), x1 = f2i(p[1].x * scale), y1 = f2i(p[1].y * scale);
        int32_t x2 = f2i(p[2].x * scale), y2 = f2i(p[2].y * scale);
        int w = 1;
        if (y0 > y2) { gswap(x0, x2); gswap(y0, y2); w = -1; }
        if (fdot6_round(y0) == fdot6_round(y2)) return;
        const int sh = rbe::quad_shift(x0, y0, x1, y1, x2, y2, shift);
        if (curves) {
            CurveRec c;
            c.p[0] = x0; c.p[1] = y0; c.p[2] = x1; c.p[3] = y1; c.p[4] = x2; c.p[5] = y2; c.p[6] = 0; c.p[7] = 0;
            c.info = 0u | ((uint32_t)sh << 4) | (w < 0 ? 0x100u : 0u);
            c.item = n_items++;
            curves->push_back(c);
            kinds.push_back(1);
            return;
        }
        int32_t link = -1;
        rbe::quad_expand(x0, y0, x1, y1, x2, y2, sh, [&](const rbe::RawEdge &r) { push_segment(r, w, &link); });
    }

    // CubicEdge::new + every update()
    GEO_HD void cubic(const P p[4])
    {
        float scale = (float)(1 << (shift + 6));
        int32_t x0 = f2i(p[0].x * scale), y0 = f2i(p[0].y * scale), x1 = f2i(p[1].x * scale), y1 = f2i(p[1].y * scale);
        int32_t x2 = f2i(p[2].x * scale), y2 = f2i(p[2].y * scale), x3 = f2i(p[3].x * scale), y3 = f2i(p[3].y * scale);
        int w = 1;
        if (y0 > y3) { gswap(x0, x3); gswap(x1, x2); gswap(y0, y3); gswap(y1, y2); w = -1; }
        if (fdot6_round(y0) == fdot6_round(y3)) return;
        const int sh = rbe::cubic_shift(x0, y0, x1, y1, x2, y2, x3, y3);
        if (curves) {
            CurveRec c;
            c.p[0] = x0; c.p[1] = y0; c.p[2] = x1; c.p[3] = y1; c.p[4] = x2; c.p[5] = y2; c.p[6] = x3; c.p[7] = y3;
            c.info = 1u | ((uint32_t)sh << 4) | (w < 0 ? 0x100u : 0u);
            c.item = n_items++;
            curves->push_back(c);
            kinds.push_back(1);
            return;
        }
        int32_t link = -1;
        rbe::cubic_expand(x0, y0, x1, y1, x2, y2, x3, y3, sh, [&](const rbe::RawEdge &r) { push_segment(r, w, &link); });
    }
};

// ---------------------------------------------------------------------------------------------------
// clipping against the tile rectangle (tiny-skia edge_clipper.rs / line_clipper.rs)
// ---------------------------------------------------------------------------------------------------
struct Clip { float l, t, r, b; };

GEO_HDI inline float pin(double v, double a, double b)
{
    if (a > b) gswap(a, b);
    return (float)gmin(gmax(v, a), b);
}
GEO_HD inline float cut_h(const P s[2], float y)
{
    float dy = s[1].y - s[0].y;
    if (nearly_zero(dy)) return (s[0].x + s[1].x) * 0.5f;
    double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
    return pin(x0 + ((double)y - y0) * (x1 - x0) / (y1 - y0), x0, x1);
}
GEO_HD inline float cut_v(const P s[2], float x)
{
    float dx = s[1].x - s[0].x;
    float y;
    if (nearly_zero(dx)) y = (s[0].y + s[1].y) * 0.5f;
    else {
        double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
        y = (float)(y0 + ((double)x - x0) * (y1 - y0) / (x1 - x0));
    }
    float a = s[0].y, b = s[1].y;
    if (a > b) gswap(a, b);
    return gmin(gmax(y, a), b);
}

template <class Sink> struct Clipper {
    Sink *sink;
    Clip c;

    GEO_HD void line(P p0, P p1)
    {
        const P pts[2] = {p0, p1};
        int i0 = pts[0].y < pts[1].y ? 0 : 1, i1 = 1 - i0;
        if (pts[i1].y <= c.t || pts[i0].y >= c.b) return;
        P tmp[2] = {p0, p1};
        if (pts[i0].y < c.t) tmp[i0] = P{cut_h(pts, c.t), c.t};
        if (tmp[i1].y > c.b) tmp[i1] = P{cut_h(pts, c.b), c.b};
        P res[4];
        int n = 1;
        bool rev;
        if (pts[0].x < pts[1].x) { i0 = 0; i1 = 1; rev = false; } else { i0 = 1; i1 = 0; rev = true; }
        if (tmp[i1].x <= c.l) {
            res[0] = P{c.l, tmp[0].y}; res[1] = P{c.l, tmp[1].y}; rev = false;
        } else if (tmp[i0].x >= c.r) {
            res[0] = P{c.r, tmp[0].y}; res[1] = P{c.r, tmp[1].y}; rev = false;
        } else {
            P *r = res;
            if (tmp[i0].x < c.l) {
                *r++ = P{c.l, tmp[i0].y};
                *r = P{c.l, cut_v(tmp, c.l)};
            } else *r = tmp[i0];
            r++;
            if (tmp[i1].x > c.r) {
                *r++ = P{c.r, cut_v(tmp, c.r)};
                *r = P{c.r, tmp[i1].y};
            } else *r = tmp[i1];
            n = (int)(r - res);
        }
        if (rev) for (int i = n; i > 0; i--) sink->line(res[i], res[i - 1]);
        else for (int i = 0; i < n; i++) sink->line(res[i], res[i + 1]);
    }
    GEO_HD void vline(float x, float y0, float y1, bool rev)
    {
        if (rev) gswap(y0, y1);
        sink->line(P{x, y0}, P{x, y1});
    }
    GEO_HD void put_quad(const P p[3], bool rev)
    {
        if (rev) { P r[3] = {p[2], p[1], p[0]}; sink->quad(r); } else sink->quad(p);
    }
    GEO_HD void put_cubic(const P p[4], bool rev)
    {
        if (rev) { P r[4] = {p[3], p[2], p[1], p[0]}; sink->cubic(r); } else sink->cubic(p);
    }

    GEO_HD bool mono_quad_t(float c0, float c1, float c2, float target, float *t)
    {
        float roots[2];
        if (unit_quad_roots(c0 - c1 - c1 + c2, 2 * (c1 - c0), c0 - target, roots)) { *t = roots[0]; return true; }
        return false;
    }

    GEO_HD void mono_quad(const P src[3])
    {
        P p[3];
        bool rev = src[0].y > src[2].y;
        if (rev) { p[0] = src[2]; p[1] = src[1]; p[2] = src[0]; } else { p[0] = src[0]; p[1] = src[1]; p[2] = src[2]; }
        if (p[2].y <= c.t || p[0].y >= c.b) return;
        float t;
        P tmp[5];
        if (p[0].y < c.t) {
            if (mono_quad_t(p[0].y, p[1].y, p[2].y, c.t, &t)) {
                split_quad(p, t, tmp);
                tmp[2].y = c.t;
                tmp[3].y = gmax(tmp[3].y, c.t);
                p[0] = tmp[2]; p[1] = tmp[3];
            } else for (auto &q : p) if (q.y < c.t) q.y = c.t;
        }
        if (p[2].y > c.b) {
            if (mono_quad_t(p[0].y, p[1].y, p[2].y, c.b, &t)) {
                split_quad(p, t, tmp);
                tmp[1].y = gmin(tmp[1].y, c.b);
                tmp[2].y = c.b;
                p[1] = tmp[1]; p[2] = tmp[2];
            } else for (auto &q : p) if (q.y > c.b) q.y = c.b;
        }
        if (p[0].x > p[2].x) { gswap(p[0], p[2]); rev = !rev; }
        if (p[2].x <= c.l) { vline(c.l, p[0].y, p[2].y, rev); return; }
        if (p[0].x >= c.r) { vline(c.r, p[0].y, p[2].y, rev); return; }
        if (p[0].x < c.l) {
            if (mono_quad_t(p[0].x, p[1].x, p[2].x, c.l, &t)) {
                split_quad(p, t, tmp);
                vline(c.l, tmp[0].y, tmp[2].y, rev);
                tmp[2].x = c.l;
                tmp[3].x = gmax(tmp[3].x, c.l);
                p[0] = tmp[2]; p[1] = tmp[3];
            } else { vline(c.l, p[0].y, p[2].y, rev); return; }
        }
        if (p[2].x > c.r) {
            if (mono_quad_t(p[0].x, p[1].x, p[2].x, c.r, &t)) {
                split_quad(p, t, tmp);
                tmp[1].x = gmin(tmp[1].x, c.r);
                tmp[2].x = c.r;
                put_quad(tmp, rev);
                vline(c.r, tmp[2].y, tmp[4].y, rev);
            } else {
                p[1].x = gmin(p[1].x, c.r);
                p[2].x = gmin(p[2].x, c.r);
                put_quad(p, rev);
            }
        } else put_quad(p, rev);
    }

    GEO_HD void quad(const P s[3])
    {
        float miny = gmin(gmin(s[0].y, s[1].y), s[2].y), maxy = gmax(gmax(s[0].y, s[1].y), s[2].y);
        if (!(maxy > c.t && miny < c.b)) return;
        P my[5];
        int cy = quad_extrema<1>(s, my);
        for (int y = 0; y <= cy; y++) {
            P mx[5];
            int cx = quad_extrema<0>(&my[y * 2], mx);
            for (int x = 0; x <= cx; x++) mono_quad(&mx[x * 2]);
        }
    }

    // mono_cubic_closest_t over one coordinate (stride 2 floats)
    GEO_HD float closest_t(const float *s, float x)
    {
        float t = 0.5f, last_t, best = t, step = 0.25f;
        float d = s[0], a = s[6] + 3 * (s[2] - s[4]) - d, b = 3 * (s[4] - s[2] - s[2] + d), cc = 3 * (s[2] - d);
        x -= d;
        float closest = 3.402823466e+38f;
        do {
            float loc = ((a * t + b) * t + cc) * t;
            float dist = fabsf(loc - x);
            if (closest > dist) { closest = dist; best = t; }
            last_t = t;
            t += loc < x ? step : -step;
            step *= 0.5f;
        } while (closest > 0.25f && last_t != t);
        return best;
    }
    GEO_HD void chop_at(const P p[4], float v, int axis, P d[7]) { split_cubic(p, closest_t(axis ? &p[0].y : &p[0].x, v), d); }

    GEO_HD void mono_cubic(const P src[4])
    {
        P p[4];
        bool rev = src[0].y > src[3].y;
        if (rev) { p[0] = src[3]; p[1] = src[2]; p[2] = src[1]; p[3] = src[0]; } else memcpy(p, src, sizeof(p));
        if (p[3].y <= c.t || p[0].y >= c.b) return;
        P tmp[7];
        if (p[0].y < c.t) {
            chop_at(p, c.t, 1, tmp);
            if (tmp[3].y < c.t && tmp[4].y < c.t && tmp[5].y < c.t) {
                P t2[4] = {tmp[3], tmp[4], tmp[5], tmp[6]};
                chop_at(t2, c.t, 1, tmp);
            }
            tmp[3].y = c.t;
            tmp[4].y = gmax(tmp[4].y, c.t);
            p[0] = tmp[3]; p[1] = tmp[4]; p[2] = tmp[5];
        }
        if (p[3].y > c.b) {
            chop_at(p, c.b, 1, tmp);
            tmp[3].y = c.b;
            tmp[2].y = gmin(tmp[2].y, c.b);
            p[1] = tmp[1]; p[2] = tmp[2]; p[3] = tmp[3];
        }
        if (p[0].x > p[3].x) { gswap(p[0], p[3]); gswap(p[1], p[2]); rev = !rev; }
        if (p[3].x <= c.l) { vline(c.l, p[0].y, p[3].y, rev); return; }
        if (p[0].x >= c.r) { vline(c.r, p[0].y, p[3].y, rev); return; }
        if (p[0].x < c.l) {
            chop_at(p, c.l, 0, tmp);
            vline(c.l, tmp[0].y, tmp[3].y, rev);
            tmp[3].x = c.l;
            tmp[4].x = gmax(tmp[4].x, c.l);
            p[0] = tmp[3]; p[1] = tmp[4]; p[2] = tmp[5];
        }
        if (p[3].x > c.r) {
            chop_at(p, c.r, 0, tmp);
            tmp[3].x = c.r;
            tmp[2].x = gmin(tmp[2].x, c.r);
            put_cubic(tmp, rev);
            vline(c.r, tmp[3].y, tmp[6].y, rev);
        } else put_cubic(p, rev);
    }

    GEO_HD void cubic(const P s[4])
    {
        float minx = s[0].x, maxx = s[0].x, miny = s[0].y, maxy = s[0].y;
        for (int i = 1; i < 4; i++) {
            minx = gmin(minx, s[i].x); maxx = gmax(maxx, s[i].x);
            miny = gmin(miny, s[i].y); maxy = gmax(maxy, s[i].y);
        }
        if (!(maxy > c.t && miny < c.b)) return;
        const float limit = (float)(1 << 22);
        if (minx < -limit || miny < -limit || maxx > limit || maxy > limit) { line(s[0], s[3]); return; }
        P my[10];
        int cy = cubic_extrema<1>(s, my);
        for (int y = 0; y <= cy; y++) {
            P mx[10];
            int cx = cubic_extrema<0>(&my[y * 3], mx);
            for (int x = 0; x <= cx; x++) mono_cubic(&mx[x * 3]);
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// build_draw: scan::path_aa::fill_path / scan::path::fill_path up to (not including) walk_edges
// ---------------------------------------------------------------------------------------------------
GEO_HDI inline bool sect(IRect a, IRect b, IRect *o)
{
    int64_t l = gmax(a.x, b.x), t = gmax(a.y, b.y);
    int64_t r = gmin<int64_t>((int64_t)a.x + a.w, (int64_t)b.x + b.w);
    int64_t bt = gmin<int64_t>((int64_t)a.y + a.h, (int64_t)b.y + b.h);
    if (r <= l || bt <= t) return false;
    *o = IRect{(int32_t)l, (int32_t)t, (int32_t)(r - l), (int32_t)(bt - t)};
    return true;
}
GEO_HDI inline bool contains(IRect o, IRect in)
{
    return in.x >= o.x && in.y >= o.y && (int64_t)in.x + in.w <= (int64_t)o.x + o.w && (int64_t)in.y + in.h <= (int64_t)o.y + o.h;
}
GEO_HDI inline bool short_overflow(int32_t v, int s) { return ((int32_t)(int16_t)shl(v, s) >> s) != v; }

// Shared front end of both builders: bounds and clip decisions of tiny-skia's fill_path (painter.rs, scan/path.rs,
// scan/path_aa.rs), then PathEdgeIter + EdgeClipper feeding `sink`.  Returns false when nothing is to be drawn.
// Bounds of a path's points, with the probe for non-finite coordinates (tiny_skia_path::Path cannot hold a non-finite
// point — PathBuilder::finish fails — so such a path is never drawn; min / max would silently skip a NaN).
struct PathBounds {
    float l, t, r, b;
    bool finite;
};
template <class Pts> GEO_HD PathBounds path_bounds(const Pts &pts, int n_pts)
{
    PathBounds pb;
    pb.l = pb.r = pb.t = pb.b = 0.0f;
    pb.finite = false;
    if (n_pts == 0) return pb;
    const P p0 = pts[0];
    float l = p0.x, r = l, t = p0.y, b = t, probe = 0.0f;
    for (int i = 0; i < n_pts; i++) {
        const P p = pts[i];
        l = gmin(l, p.x); r = gmax(r, p.x);
        t = gmin(t, p.y); b = gmax(b, p.y);
        probe += p.x * 0.0f + p.y * 0.0f; // NaN as soon as one coordinate is NaN or infinite
    }
    pb.l = l; pb.t = t; pb.r = r; pb.b = b;
    pb.finite = probe == 0.0f;
    return pb;
}

// The decisions tiny-skia's fill_path takes from the path's bounds (painter.rs, scan/path.rs, scan/path_aa.rs) before any
// edge is built: the blitter rectangle, supersampling or not, whether the clipper is needed.  False: nothing is drawn.
struct FillPlan {
    IRect ir, sect;
    int shift;
    bool inside;
};
GEO_HD inline bool fill_plan(const PathBounds &pb, bool anti_alias, int32_t cw, int32_t ch, FillPlan *fp)
{
    if (!pb.finite) return false;
    const float l = pb.l, t = pb.t, r = pb.r, b = pb.b;
    if (!(gfinite(l) && gfinite(r) && gfinite(t) && gfinite(b))) return false;
    if (nearly_zero(r - l) || nearly_zero(b - t)) return false; // painter.rs: empty paths, h/v lines
    const IRect clip{0, 0, cw, ch};
    IRect ir;
    int shift = anti_alias ? 2 : 0;
    if (anti_alias) {
        int32_t il = f2i(floorf(l)), it = f2i(floorf(t)), irr = f2i(ceilf(r)), ib = f2i(ceilf(b));
        if ((int64_t)irr - il <= 0 || (int64_t)ib - it <= 0) return false;
        ir = IRect{il, it, (int32_t)((int64_t)irr - il), (int32_t)((int64_t)ib - it)};
        IRect s;
        if (!sect(ir, clip, &s)) return false;
        if (short_overflow(s.x, 2) || short_overflow(s.y, 2) || short_overflow(s.x + s.w, 2) || short_overflow(s.y + s.h, 2))
            shift = 0; // cannot supersample: non-AA fallback
        else if (cw > 32767 || ch > 32767) return false;
    }
    if (shift == 0) {
        const double bias = 0.5 + 1.5 / 64.0; // conservative_round_to_int
        int32_t il = d2i(ceil((double)l - bias)), it = d2i(ceil((double)t - bias));
        int32_t irr = d2i(floor((double)r + bias)), ib = d2i(floor((double)b + bias));
        if ((int64_t)irr - il <= 0 || (int64_t)ib - it <= 0) return false;
        ir = IRect{il, it, (int32_t)((int64_t)irr - il), (int32_t)((int64_t)ib - it)};
    }
    IRect s;
    if (!sect(ir, clip, &s)) return false;
    fp->ir = ir;
    fp->sect = s;
    fp->shift = shift;
    fp->inside = ir.x >= 0 && ir.y >= 0 && contains(clip, ir);
    return true;
}

// PathEdgeIter + EdgeClipper feeding `sink` (every contour is closed implicitly).  `verbs` / `pts` may be any run of whole
// contours of the path the plan was made for.
template <class Sink, class Pts>
GEO_HD void walk_verbs(const uint8_t *verbs, int n_verbs, const Pts &pts, bool inside, int32_t cw, int32_t ch, Sink &sink)
{
    Clipper<Sink> cl;
    cl.sink = &sink;
    cl.c = Clip{0.0f, 0.0f, (float)cw, (float)ch};
    int pi = 0;
    P move_to{0, 0}, last{0, 0};
    bool open = false;
    for (int vi = 0; vi <= n_verbs; vi++) {
        int verb = vi < n_verbs ? verbs[vi] : 4;
        if (verb == 0 || verb == 4) {
            if (open) {
                if (inside) sink.line(last, move_to); else cl.line(last, move_to);
                open = false;
            }
            if (verb == 0) { move_to = pts[pi++]; last = move_to; } else last = move_to;
            continue;
        }
        if (verb == 1) {
            P p1 = pts[pi++];
            if (inside) sink.line(last, p1); else cl.line(last, p1);
            last = p1;
        } else if (verb == 2) {
            P q[3] = {last, pts[pi], pts[pi + 1]};
            pi += 2;
            if (inside) {
                P m[5];
                int n = quad_extrema<1>(q, m);
                for (int i = 0; i <= n; i++) sink.quad(&m[i * 2]);
            } else cl.quad(q);
            last = q[2];
        } else if (verb == 3) {
            P q[4] = {last, pts[pi], pts[pi + 1], pts[pi + 2]};
            pi += 3;
            if (inside) {
                P m[10];
                int n = cubic_extrema<1>(q, m);
                for (int i = 0; i <= n; i++) sink.cubic(&m[i * 3]);
            } else cl.cubic(q);
            last = q[3];
        }
        open = true;
    }
}

// Shared front end of both builders: plan, then the edges.  Returns false when nothing is to be drawn.
template <class Sink, class Pts>
GEO_HD bool walk_path(const uint8_t *verbs, int n_verbs, const Pts &pts, int n_pts, bool anti_alias, int32_t cw, int32_t ch,
                      Sink &sink, DrawGeom *g, IRect *ir_out, bool *inside_out)
{
    if (n_pts == 0) return false;
    FillPlan fp;
    if (!fill_plan(path_bounds(pts, n_pts), anti_alias, cw, ch, &fp)) return false;
    sink.shift = fp.shift;
    walk_verbs(verbs, n_verbs, pts, fp.inside, cw, ch, sink);
    g->sect = fp.sect;
    g->shift = fp.shift;
    *ir_out = fp.ir;
    *inside_out = fp.inside;
    return true;
}

GEO_HD inline bool finish_geom(const IRect &ir, bool inside, int32_t ch, DrawGeom *g)
{
    int32_t start_y = shl(ir.y, g->shift), stop_y = shl(ir.y + ir.h, g->shift);
    if (!inside) {
        start_y = gmax(start_y, 0);
        stop_y = gmin(stop_y, shl(ch, g->shift));
    }
    if (start_y < 0 || stop_y <= start_y) return false;
    g->start_y = start_y;
    g->stop_y = stop_y;
    return true;
}


// Item form of the builder (curves recorded, see CurveRec): `lines` receives the final line edges with order = emission
// index, `curves` the recorded curves with item = emission index; `kinds` is scratch.  All three must be empty.
template <template <class> class Vec, class Pts>
GEO_HD bool build_items(const uint8_t *verbs, int n_verbs, const Pts &pts, int n_pts, bool anti_alias, int32_t cw, int32_t ch,
                        Sink<Vec> &sink, DrawGeom *g)
{
    IRect ir;
    bool inside;
    bool ok = walk_path(verbs, n_verbs, pts, n_pts, anti_alias, cw, ch, sink, g, &ir, &inside);
    // BasicEdgeBuilder::build: fewer than two edge objects (a curve is one) -> nothing to draw
    if (ok && (sink.out->size() - sink.base) + sink.curves->size() < 2) ok = false;
    if (ok) ok = finish_geom(ir, inside, ch, g);
    return ok;
}

// The draw's items laid out for the device: slots of the device-side edge array in emission order (one per line,
// 2^shift per curve; DevEdge::meta / CurveRec::item carry the slot), the capacity its tile-row edge lists need, and the
// y ranges of its chains (a line, or a whole curve) for the packed-winding bound.
struct Packed {
    uint32_t slots;   // slots of this draw in the device edge array
    size_t n_list;    // entries its tile-row edge lists can need
    bool too_large;   // slot indices would not fit DevEdge::meta
};
GEO_HDI inline int tile_row_of(int32_t suby, int shift, int oy, int r0, int nr) { return gmin(gmax((((suby >> shift) + oy) >> 3) - r0, 0), nr - 1); }

// `ends` receives (first_y, last_y) of every chain: Ends::operator()(int32_t first, int32_t last).
template <class Ends>
GEO_HD Packed pack_items(const Edge *src, size_t ne, const CurveRec *csrc, size_t ncv, DevEdge *dst, CurveRec *cdst, int shift, int oy,
                         int r0, int nr, Ends &ends)
{
    Packed po;
    po.n_list = 0;
    po.too_large = false;
    uint32_t slot = 0;
    size_t il = 0, ic = 0;
    while (il < ne || ic < ncv) {
        if (ic >= ncv || (il < ne && src[il].order < csrc[ic].item)) {
            const Edge e = src[il]; // by value: `dst` may be the same storage (the device packs in place)
            DevEdge de;
            de.x = e.x; de.dx = e.dx;
            de.ypack = ((uint32_t)e.first_y & 0xffffu) | ((uint32_t)e.last_y << 16);
            de.meta = (e.winding < 0 ? 1u : 0u) | (slot << 4);
            dst[il++] = de;
            po.n_list += (size_t)(tile_row_of(e.last_y, shift, oy, r0, nr) - tile_row_of(e.first_y, shift, oy, r0, nr) + 1);
            ends(e.first_y, e.last_y);
            slot += 1;
        } else {
            CurveRec c = csrc[ic];
            const int sh = (int)((c.info >> 4) & 0xfu);
            const int ylast = (c.info & 1u) ? c.p[7] : c.p[5];
            const int32_t top = (c.p[1] + 32) >> 6, bot = (ylast + 32) >> 6;
            c.item = slot;
            cdst[ic++] = c;
            // its segments partition [top, bot): at most one extra list entry per tile-row boundary
            po.n_list += ((size_t)1 << sh) + (size_t)(tile_row_of(bot - 1, shift, oy, r0, nr) - tile_row_of(top, shift, oy, r0, nr)) + 2;
            ends(top, bot - 1);
            slot += 1u << sh;
        }
        if (slot >= (1u << 28)) { po.too_large = true; break; } // meta keeps slot indices in 28 bits
    }
    po.slots = slot;
    return po;
}

} // namespace fl
} // namespace geo
