// raster_host.cpp — see raster_host.h.  Host-side geometry for the device rasteriser.
//
// Unlike the reference's scan converter, which keeps quadratic/cubic edges as stateful objects that the
// scanline walker advances (tiny-skia edge.rs QuadraticEdge::update / CubicEdge::update), every curve is
// forward-differenced here, once, into the line edges those updates would produce, so the device only
// ever sees independent `Edge {x, dx, first_y, last_y, winding}` records it can evaluate in closed form:
// x(y) = x + (y - first_y) * dx (wrapping i32).
#include "raster_host.h"

#include "edge_math.h"
#ifdef RB_HOST_PROFILE
#include <x86intrin.h>
#include <atomic>
#endif

#include <math.h>
#include <string.h>

#include <algorithm>

namespace rbh {

// ---------------------------------------------------------------------------------------------------
// numeric helpers: Rust `as` cast semantics and tiny-skia fixed point (fixed_point.rs)
// ---------------------------------------------------------------------------------------------------
static inline int32_t f2i(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static inline int32_t d2i(double v)
{
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}
using rbe::shl;
using rbe::fdot6_round;
static constexpr float kNearlyZero = 1.0f / 4096.0f;
static inline bool nearly_zero(float v, float tol = kNearlyZero) { return fabsf(v) <= tol; }

// ---------------------------------------------------------------------------------------------------
// transform (tiny-skia-path transform.rs)
// ---------------------------------------------------------------------------------------------------
bool Xform::is_finite() const
{
    return std::isfinite(sx) && std::isfinite(ky) && std::isfinite(kx) && std::isfinite(sy) && std::isfinite(tx)
           && std::isfinite(ty);
}
static inline float mam(float a, float b, float c, float d) { return (float)((double)a * b + (double)c * d); }

Xform concat(const Xform &a, const Xform &b)
{
    if (a.is_identity()) return b;
    if (b.is_identity()) return a;
    Xform r;
    if (!a.has_skew() && !b.has_skew()) {
        r.sx = a.sx * b.sx; r.ky = 0; r.kx = 0; r.sy = a.sy * b.sy;
        r.tx = a.sx * b.tx + a.tx;
        r.ty = a.sy * b.ty + a.ty;
    } else {
        r.sx = mam(a.sx, b.sx, a.kx, b.ky);
        r.ky = mam(a.ky, b.sx, a.sy, b.ky);
        r.kx = mam(a.sx, b.kx, a.kx, b.sy);
        r.sy = mam(a.ky, b.kx, a.sy, b.sy);
        r.tx = mam(a.sx, b.tx, a.kx, b.ty) + a.tx;
        r.ty = mam(a.ky, b.tx, a.sy, b.ty) + a.ty;
    }
    return r;
}

bool invert(const Xform &t, Xform *out)
{
    if (t.is_identity()) { *out = t; return true; }
    if (!t.has_skew()) {
        Xform r;
        if (t.has_scale()) {
            float ix = 1.0f / t.sx, iy = 1.0f / t.sy;
            r.sx = ix; r.sy = iy; r.tx = -t.tx * ix; r.ty = -t.ty * iy;
        } else {
            r.tx = -t.tx; r.ty = -t.ty;
        }
        *out = r;
        return true;
    }
    double det = (double)t.sx * t.sy - (double)t.kx * t.ky;
    if (nearly_zero((float)det, kNearlyZero * kNearlyZero * kNearlyZero)) return false;
    double inv = 1.0 / det;
    Xform r;
    r.sx = (float)((double)t.sy * inv);
    r.ky = (float)((double)(-t.ky) * inv);
    r.kx = (float)((double)(-t.kx) * inv);
    r.sy = (float)((double)t.sx * inv);
    r.tx = (float)(((double)t.kx * t.ty - (double)t.sy * t.tx) * inv);
    r.ty = (float)(((double)t.ky * t.tx - (double)t.sx * t.ty) * inv);
    if (!r.is_finite()) return false;
    *out = r;
    return true;
}

void map_points(const Xform &t, Pt *p, int n)
{
    if (t.is_identity()) return;
    if (t.is_translate()) {
        for (int i = 0; i < n; i++) { p[i].x += t.tx; p[i].y += t.ty; }
    } else if (!t.has_skew()) {
        for (int i = 0; i < n; i++) { p[i].x = p[i].x * t.sx + t.tx; p[i].y = p[i].y * t.sy + t.ty; }
    } else {
        for (int i = 0; i < n; i++) {
            float x = p[i].x * t.sx + p[i].y * t.kx + t.tx;
            float y = p[i].x * t.ky + p[i].y * t.sy + t.ty;
            p[i].x = x; p[i].y = y;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// curve chopping (tiny-skia-path path_geometry.rs)
// ---------------------------------------------------------------------------------------------------
static inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }
static inline Pt lerp(Pt a, Pt b, float t) { return Pt{lerpf(a.x, b.x, t), lerpf(a.y, b.y, t)}; }

static bool unit_divide(float numer, float denom, float *ratio)
{
    if (numer < 0) { numer = -numer; denom = -denom; }
    if (denom == 0 || numer == 0 || numer >= denom) return false;
    float r = numer / denom;
    if (!(r > 0.0f && r < 1.0f)) return false;
    *ratio = r;
    return true;
}

static int unit_quad_roots(float a, float b, float c, float roots[2])
{
    if (a == 0) return unit_divide(-c, b, roots) ? 1 : 0;
    double dr = (double)b * b - 4.0 * (double)a * c;
    if (dr < 0) return 0;
    float r = (float)sqrt(dr);
    if (!std::isfinite(r)) return 0;
    float q = (b < 0) ? -(b - r) / 2 : -(b + r) / 2;
    int n = 0;
    if (unit_divide(q, a, roots + n)) n++;
    if (unit_divide(c, q, roots + n)) n++;
    if (n == 2) {
        if (roots[0] > roots[1]) std::swap(roots[0], roots[1]);
        else if (roots[0] == roots[1]) n = 1;
    }
    return n;
}

static void split_quad(const Pt s[3], float t, Pt d[5])
{
    Pt p01 = lerp(s[0], s[1], t), p12 = lerp(s[1], s[2], t);
    d[0] = s[0]; d[1] = p01; d[2] = lerp(p01, p12, t); d[3] = p12; d[4] = s[2];
}
static void split_cubic(const Pt s[4], float t, Pt d[7])
{
    Pt ab = lerp(s[0], s[1], t), bc = lerp(s[1], s[2], t), cd = lerp(s[2], s[3], t);
    Pt abc = lerp(ab, bc, t), bcd = lerp(bc, cd, t);
    d[0] = s[0]; d[1] = ab; d[2] = abc; d[3] = lerp(abc, bcd, t); d[4] = bcd; d[5] = cd; d[6] = s[3];
}

template <int AXIS> static inline float &ax(Pt &p) { return AXIS ? p.y : p.x; }
template <int AXIS> static inline float axv(const Pt &p) { return AXIS ? p.y : p.x; }

template <int AXIS> static int quad_extrema(const Pt s[3], Pt d[5])
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

template <int AXIS> static int cubic_extrema(const Pt s[4], Pt d[10])
{
    float a = axv<AXIS>(s[0]), b = axv<AXIS>(s[1]), c = axv<AXIS>(s[2]), e = axv<AXIS>(s[3]);
    float tv[2];
    int roots = unit_quad_roots(e - a + 3 * (b - c), 2 * (a - b - b + c), b - a, tv);
    if (roots == 0) { memcpy(d, s, 4 * sizeof(Pt)); return 0; }
    Pt src[4];
    memcpy(src, s, sizeof(src));
    Pt *dst = d;
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
struct Sink {
    std::vector<Edge> *out;
    size_t base;  // first edge of this draw
    int shift;
    std::vector<uint8_t> kinds; // per emitted edge: 0 = from a line, 1 = from a curve (combine_vertical only
                                // ever looks at a preceding *line* edge)
    // item mode: curves are recorded (FDot6 control points + subdivision count) instead of being expanded; `order`
    // of a line / `item` of a curve is then the emission index shared by both kinds
    std::vector<CurveRec> *curves = nullptr;
    uint32_t n_items = 0;

    uint32_t next_order() { return curves ? n_items++ : (uint32_t)(out->size() - base); }

    // LineEdge::new / update tail: FDot6 end points, y0 <= y1
    bool emit(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int winding, Edge *e)
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

    void line(Pt p0, Pt p1)
    {
        float scale = (float)(1 << (shift + 6));
        int32_t x0 = f2i(p0.x * scale), y0 = f2i(p0.y * scale), x1 = f2i(p1.x * scale), y1 = f2i(p1.y * scale);
        int w = 1;
        if (y0 > y1) { std::swap(x0, x1); std::swap(y0, y1); w = -1; }
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
    static int combine_vertical(const Edge &edge, Edge &last)
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

    void push_segment(const rbe::RawEdge &r, int w, int32_t *link)
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
    void quad(const Pt p[3])
    {
        float scale = (float)(1 << (shift + 6));
        int32_t x0 = f2i(p[0].x * scale), y0 = f2i(p[0].y * scale), x1 = f2i(p[1].x * scale), y1 = f2i(p[1].y * scale);
        int32_t x2 = f2i(p[2].x * scale), y2 = f2i(p[2].y * scale);
        int w = 1;
        if (y0 > y2) { std::swap(x0, x2); std::swap(y0, y2); w = -1; }
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
    void cubic(const Pt p[4])
    {
        float scale = (float)(1 << (shift + 6));
        int32_t x0 = f2i(p[0].x * scale), y0 = f2i(p[0].y * scale), x1 = f2i(p[1].x * scale), y1 = f2i(p[1].y * scale);
        int32_t x2 = f2i(p[2].x * scale), y2 = f2i(p[2].y * scale), x3 = f2i(p[3].x * scale), y3 = f2i(p[3].y * scale);
        int w = 1;
        if (y0 > y3) { std::swap(x0, x3); std::swap(x1, x2); std::swap(y0, y3); std::swap(y1, y2); w = -1; }
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

static float pin(double v, double a, double b)
{
    if (a > b) std::swap(a, b);
    return (float)std::min(std::max(v, a), b);
}
static float cut_h(const Pt s[2], float y)
{
    float dy = s[1].y - s[0].y;
    if (nearly_zero(dy)) return (s[0].x + s[1].x) * 0.5f;
    double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
    return pin(x0 + ((double)y - y0) * (x1 - x0) / (y1 - y0), x0, x1);
}
static float cut_v(const Pt s[2], float x)
{
    float dx = s[1].x - s[0].x;
    float y;
    if (nearly_zero(dx)) y = (s[0].y + s[1].y) * 0.5f;
    else {
        double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
        y = (float)(y0 + ((double)x - x0) * (y1 - y0) / (x1 - x0));
    }
    float a = s[0].y, b = s[1].y;
    if (a > b) std::swap(a, b);
    return std::min(std::max(y, a), b);
}

struct Clipper {
    Sink *sink;
    Clip c;

    void line(Pt p0, Pt p1)
    {
        const Pt pts[2] = {p0, p1};
        int i0 = pts[0].y < pts[1].y ? 0 : 1, i1 = 1 - i0;
        if (pts[i1].y <= c.t || pts[i0].y >= c.b) return;
        Pt tmp[2] = {p0, p1};
        if (pts[i0].y < c.t) tmp[i0] = Pt{cut_h(pts, c.t), c.t};
        if (tmp[i1].y > c.b) tmp[i1] = Pt{cut_h(pts, c.b), c.b};
        Pt res[4];
        int n = 1;
        bool rev;
        if (pts[0].x < pts[1].x) { i0 = 0; i1 = 1; rev = false; } else { i0 = 1; i1 = 0; rev = true; }
        if (tmp[i1].x <= c.l) {
            res[0] = Pt{c.l, tmp[0].y}; res[1] = Pt{c.l, tmp[1].y}; rev = false;
        } else if (tmp[i0].x >= c.r) {
            res[0] = Pt{c.r, tmp[0].y}; res[1] = Pt{c.r, tmp[1].y}; rev = false;
        } else {
            Pt *r = res;
            if (tmp[i0].x < c.l) {
                *r++ = Pt{c.l, tmp[i0].y};
                *r = Pt{c.l, cut_v(tmp, c.l)};
            } else *r = tmp[i0];
            r++;
            if (tmp[i1].x > c.r) {
                *r++ = Pt{c.r, cut_v(tmp, c.r)};
                *r = Pt{c.r, tmp[i1].y};
            } else *r = tmp[i1];
            n = (int)(r - res);
        }
        if (rev) for (int i = n; i > 0; i--) sink->line(res[i], res[i - 1]);
        else for (int i = 0; i < n; i++) sink->line(res[i], res[i + 1]);
    }
    void vline(float x, float y0, float y1, bool rev)
    {
        if (rev) std::swap(y0, y1);
        sink->line(Pt{x, y0}, Pt{x, y1});
    }
    void put_quad(const Pt p[3], bool rev)
    {
        if (rev) { Pt r[3] = {p[2], p[1], p[0]}; sink->quad(r); } else sink->quad(p);
    }
    void put_cubic(const Pt p[4], bool rev)
    {
        if (rev) { Pt r[4] = {p[3], p[2], p[1], p[0]}; sink->cubic(r); } else sink->cubic(p);
    }

    static bool mono_quad_t(float c0, float c1, float c2, float target, float *t)
    {
        float roots[2];
        if (unit_quad_roots(c0 - c1 - c1 + c2, 2 * (c1 - c0), c0 - target, roots)) { *t = roots[0]; return true; }
        return false;
    }

    void mono_quad(const Pt src[3])
    {
        Pt p[3];
        bool rev = src[0].y > src[2].y;
        if (rev) { p[0] = src[2]; p[1] = src[1]; p[2] = src[0]; } else { p[0] = src[0]; p[1] = src[1]; p[2] = src[2]; }
        if (p[2].y <= c.t || p[0].y >= c.b) return;
        float t;
        Pt tmp[5];
        if (p[0].y < c.t) {
            if (mono_quad_t(p[0].y, p[1].y, p[2].y, c.t, &t)) {
                split_quad(p, t, tmp);
                tmp[2].y = c.t;
                tmp[3].y = std::max(tmp[3].y, c.t);
                p[0] = tmp[2]; p[1] = tmp[3];
            } else for (auto &q : p) if (q.y < c.t) q.y = c.t;
        }
        if (p[2].y > c.b) {
            if (mono_quad_t(p[0].y, p[1].y, p[2].y, c.b, &t)) {
                split_quad(p, t, tmp);
                tmp[1].y = std::min(tmp[1].y, c.b);
                tmp[2].y = c.b;
                p[1] = tmp[1]; p[2] = tmp[2];
            } else for (auto &q : p) if (q.y > c.b) q.y = c.b;
        }
        if (p[0].x > p[2].x) { std::swap(p[0], p[2]); rev = !rev; }
        if (p[2].x <= c.l) { vline(c.l, p[0].y, p[2].y, rev); return; }
        if (p[0].x >= c.r) { vline(c.r, p[0].y, p[2].y, rev); return; }
        if (p[0].x < c.l) {
            if (mono_quad_t(p[0].x, p[1].x, p[2].x, c.l, &t)) {
                split_quad(p, t, tmp);
                vline(c.l, tmp[0].y, tmp[2].y, rev);
                tmp[2].x = c.l;
                tmp[3].x = std::max(tmp[3].x, c.l);
                p[0] = tmp[2]; p[1] = tmp[3];
            } else { vline(c.l, p[0].y, p[2].y, rev); return; }
        }
        if (p[2].x > c.r) {
            if (mono_quad_t(p[0].x, p[1].x, p[2].x, c.r, &t)) {
                split_quad(p, t, tmp);
                tmp[1].x = std::min(tmp[1].x, c.r);
                tmp[2].x = c.r;
                put_quad(tmp, rev);
                vline(c.r, tmp[2].y, tmp[4].y, rev);
            } else {
                p[1].x = std::min(p[1].x, c.r);
                p[2].x = std::min(p[2].x, c.r);
                put_quad(p, rev);
            }
        } else put_quad(p, rev);
    }

    void quad(const Pt s[3])
    {
        float miny = std::min(std::min(s[0].y, s[1].y), s[2].y), maxy = std::max(std::max(s[0].y, s[1].y), s[2].y);
        if (!(maxy > c.t && miny < c.b)) return;
        Pt my[5];
        int cy = quad_extrema<1>(s, my);
        for (int y = 0; y <= cy; y++) {
            Pt mx[5];
            int cx = quad_extrema<0>(&my[y * 2], mx);
            for (int x = 0; x <= cx; x++) mono_quad(&mx[x * 2]);
        }
    }

    // mono_cubic_closest_t over one coordinate (stride 2 floats)
    static float closest_t(const float *s, float x)
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
    static void chop_at(const Pt p[4], float v, int axis, Pt d[7]) { split_cubic(p, closest_t(axis ? &p[0].y : &p[0].x, v), d); }

    void mono_cubic(const Pt src[4])
    {
        Pt p[4];
        bool rev = src[0].y > src[3].y;
        if (rev) { p[0] = src[3]; p[1] = src[2]; p[2] = src[1]; p[3] = src[0]; } else memcpy(p, src, sizeof(p));
        if (p[3].y <= c.t || p[0].y >= c.b) return;
        Pt tmp[7];
        if (p[0].y < c.t) {
            chop_at(p, c.t, 1, tmp);
            if (tmp[3].y < c.t && tmp[4].y < c.t && tmp[5].y < c.t) {
                Pt t2[4] = {tmp[3], tmp[4], tmp[5], tmp[6]};
                chop_at(t2, c.t, 1, tmp);
            }
            tmp[3].y = c.t;
            tmp[4].y = std::max(tmp[4].y, c.t);
            p[0] = tmp[3]; p[1] = tmp[4]; p[2] = tmp[5];
        }
        if (p[3].y > c.b) {
            chop_at(p, c.b, 1, tmp);
            tmp[3].y = c.b;
            tmp[2].y = std::min(tmp[2].y, c.b);
            p[1] = tmp[1]; p[2] = tmp[2]; p[3] = tmp[3];
        }
        if (p[0].x > p[3].x) { std::swap(p[0], p[3]); std::swap(p[1], p[2]); rev = !rev; }
        if (p[3].x <= c.l) { vline(c.l, p[0].y, p[3].y, rev); return; }
        if (p[0].x >= c.r) { vline(c.r, p[0].y, p[3].y, rev); return; }
        if (p[0].x < c.l) {
            chop_at(p, c.l, 0, tmp);
            vline(c.l, tmp[0].y, tmp[3].y, rev);
            tmp[3].x = c.l;
            tmp[4].x = std::max(tmp[4].x, c.l);
            p[0] = tmp[3]; p[1] = tmp[4]; p[2] = tmp[5];
        }
        if (p[3].x > c.r) {
            chop_at(p, c.r, 0, tmp);
            tmp[3].x = c.r;
            tmp[2].x = std::min(tmp[2].x, c.r);
            put_cubic(tmp, rev);
            vline(c.r, tmp[3].y, tmp[6].y, rev);
        } else put_cubic(p, rev);
    }

    void cubic(const Pt s[4])
    {
        float minx = s[0].x, maxx = s[0].x, miny = s[0].y, maxy = s[0].y;
        for (int i = 1; i < 4; i++) {
            minx = std::min(minx, s[i].x); maxx = std::max(maxx, s[i].x);
            miny = std::min(miny, s[i].y); maxy = std::max(maxy, s[i].y);
        }
        if (!(maxy > c.t && miny < c.b)) return;
        const float limit = (float)(1 << 22);
        if (minx < -limit || miny < -limit || maxx > limit || maxy > limit) { line(s[0], s[3]); return; }
        Pt my[10];
        int cy = cubic_extrema<1>(s, my);
        for (int y = 0; y <= cy; y++) {
            Pt mx[10];
            int cx = cubic_extrema<0>(&my[y * 3], mx);
            for (int x = 0; x <= cx; x++) mono_cubic(&mx[x * 3]);
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// build_draw: scan::path_aa::fill_path / scan::path::fill_path up to (not including) walk_edges
// ---------------------------------------------------------------------------------------------------
static bool sect(IRect a, IRect b, IRect *o)
{
    int64_t l = std::max(a.x, b.x), t = std::max(a.y, b.y);
    int64_t r = std::min<int64_t>((int64_t)a.x + a.w, (int64_t)b.x + b.w);
    int64_t bt = std::min<int64_t>((int64_t)a.y + a.h, (int64_t)b.y + b.h);
    if (r <= l || bt <= t) return false;
    *o = IRect{(int32_t)l, (int32_t)t, (int32_t)(r - l), (int32_t)(bt - t)};
    return true;
}
static bool contains(IRect o, IRect in)
{
    return in.x >= o.x && in.y >= o.y && (int64_t)in.x + in.w <= (int64_t)o.x + o.w && (int64_t)in.y + in.h <= (int64_t)o.y + o.h;
}
static inline bool short_overflow(int32_t v, int s) { return ((int32_t)(int16_t)shl(v, s) >> s) != v; }

#ifdef RB_HOST_PROFILE
std::atomic<uint64_t> g_bd_prof[4];
#define BD_T(i) do { uint64_t now__ = __rdtsc(); g_bd_prof[i] += now__ - bd_t__; bd_t__ = now__; } while (0)
#else
#define BD_T(i)
#endif
// Shared front end of both builders: bounds and clip decisions of tiny-skia's fill_path (painter.rs, scan/path.rs,
// scan/path_aa.rs), then PathEdgeIter + EdgeClipper feeding `sink`.  Returns false when nothing is to be drawn.
static bool walk_path(const uint8_t *verbs, int n_verbs, const Pt *pts, int n_pts, bool anti_alias, int32_t cw, int32_t ch,
                      Sink &sink, DrawGeom *g, IRect *ir_out, bool *inside_out)
{
    if (n_pts == 0) return false;
    // tiny_skia_path::Path cannot hold a non-finite point (PathBuilder::finish fails): such a path is never drawn.
    // min / max would silently skip a NaN, so every point is probed.
    float l = pts[0].x, r = l, t = pts[0].y, b = t, probe = 0.0f;
    for (int i = 0; i < n_pts; i++) {
        l = std::min(l, pts[i].x); r = std::max(r, pts[i].x);
        t = std::min(t, pts[i].y); b = std::max(b, pts[i].y);
        probe += pts[i].x * 0.0f + pts[i].y * 0.0f; // NaN as soon as one coordinate is NaN or infinite
    }
    if (!(probe == 0.0f)) return false;
    if (!(std::isfinite(l) && std::isfinite(r) && std::isfinite(t) && std::isfinite(b))) return false;
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
    const bool inside = ir.x >= 0 && ir.y >= 0 && contains(clip, ir);

    sink.shift = shift;
    Clipper cl{&sink, Clip{0.0f, 0.0f, (float)cw, (float)ch}};

    // PathEdgeIter: every contour is closed implicitly
    int pi = 0;
    Pt move_to{0, 0}, last{0, 0};
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
            Pt p1 = pts[pi++];
            if (inside) sink.line(last, p1); else cl.line(last, p1);
            last = p1;
        } else if (verb == 2) {
            Pt q[3] = {last, pts[pi], pts[pi + 1]};
            pi += 2;
            if (inside) {
                Pt m[5];
                int n = quad_extrema<1>(q, m);
                for (int i = 0; i <= n; i++) sink.quad(&m[i * 2]);
            } else cl.quad(q);
            last = q[2];
        } else if (verb == 3) {
            Pt q[4] = {last, pts[pi], pts[pi + 1], pts[pi + 2]};
            pi += 3;
            if (inside) {
                Pt m[10];
                int n = cubic_extrema<1>(q, m);
                for (int i = 0; i <= n; i++) sink.cubic(&m[i * 3]);
            } else cl.cubic(q);
            last = q[3];
        }
        open = true;
    }
    g->sect = s;
    g->shift = shift;
    *ir_out = ir;
    *inside_out = inside;
    return true;
}

static bool finish_geom(const IRect &ir, bool inside, int32_t ch, DrawGeom *g)
{
    int32_t start_y = shl(ir.y, g->shift), stop_y = shl(ir.y + ir.h, g->shift);
    if (!inside) {
        start_y = std::max(start_y, 0);
        stop_y = std::min(stop_y, shl(ch, g->shift));
    }
    if (start_y < 0 || stop_y <= start_y) return false;
    g->start_y = start_y;
    g->stop_y = stop_y;
    return true;
}

bool build_draw(const uint8_t *verbs, int n_verbs, const Pt *pts, int n_pts, bool anti_alias, int32_t cw, int32_t ch,
                std::vector<Edge> &out, DrawGeom *g)
{
#ifdef RB_HOST_PROFILE
    uint64_t bd_t__ = __rdtsc();
#endif
    Sink sink;
    sink.out = &out;
    sink.base = out.size();
    IRect ir;
    bool inside;
    if (!walk_path(verbs, n_verbs, pts, n_pts, anti_alias, cw, ch, sink, g, &ir, &inside)) { out.resize(sink.base); return false; }
    BD_T(0);
    size_t n = out.size() - sink.base;
    {
        // BasicEdgeBuilder::build: fewer than two edge objects (a curve is one) -> nothing to draw
        size_t objects = 0;
        for (size_t i = sink.base; i < out.size() && objects < 2; i++) objects += out[i].prev < 0 ? 1 : 0;
        if (objects < 2) { out.resize(sink.base); return false; }
    }
    // scan/path.rs: sort by (first_y, x); stable = builder order among ties
    std::stable_sort(out.begin() + (long)sink.base, out.end(), [](const Edge &a, const Edge &e) {
        if (a.first_y != e.first_y) return a.first_y < e.first_y;
        return a.x < e.x;
    });
    BD_T(1);
    {
        // `order` values may have gaps (combine_vertical pops), so map through a table sized by the maximum
        Edge *e = out.data() + sink.base;
        uint32_t max_order = 0;
        for (size_t i = 0; i < n; i++) max_order = std::max(max_order, e[i].order);
        std::vector<int32_t> pos((size_t)max_order + 1, -1);
        for (size_t i = 0; i < n; i++) pos[e[i].order] = (int32_t)i;
        for (size_t i = 0; i < n; i++) if (e[i].prev >= 0) e[i].prev = pos[(size_t)e[i].prev];
        // insert_new_edges batches: the non-continuation edges sharing one first_y, in sorted order
        size_t i = 0;
        while (i < n) {
            size_t j = i;
            bool have_first = false;
            int32_t first_x = 0;
            while (j < n && e[j].first_y == e[i].first_y) {
                if (e[j].prev < 0) {
                    if (!have_first) { have_first = true; first_x = e[j].x; }
                    else if (e[j].x > first_x) e[j].before = 1;
                }
                j++;
            }
            i = j;
        }
    }
    BD_T(2);
    if (!finish_geom(ir, inside, ch, g)) { out.resize(sink.base); return false; }
    return true;
}

// Item form of build_draw for the device-side expansion: line edges are final (appended to `lines`, order = emission
// index), curves are recorded with their FDot6 control points (appended to `curves`, item = emission index).
bool build_draw_items(const uint8_t *verbs, int n_verbs, const Pt *pts, int n_pts, bool anti_alias, int32_t cw, int32_t ch,
                      std::vector<Edge> &lines, std::vector<CurveRec> &curves, DrawGeom *g)
{
    Sink sink;
    sink.out = &lines;
    sink.base = lines.size();
    sink.curves = &curves;
    const size_t cbase = curves.size();
    IRect ir;
    bool inside;
    bool ok = walk_path(verbs, n_verbs, pts, n_pts, anti_alias, cw, ch, sink, g, &ir, &inside);
    // BasicEdgeBuilder::build: fewer than two edge objects (a curve is one) -> nothing to draw
    if (ok && (lines.size() - sink.base) + (curves.size() - cbase) < 2) ok = false;
    if (ok) ok = finish_geom(ir, inside, ch, g);
    if (!ok) { lines.resize(sink.base); curves.resize(cbase); return false; }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// paints (tiny-skia shaders/*.rs, pipeline/blitter.rs RasterPipelineBlitter::new)
// ---------------------------------------------------------------------------------------------------
static inline uint32_t f2u16(float v) { return !(v > 0.0f) ? 0u : (v >= 65535.0f ? 65535u : (uint32_t)v); }
static inline float clamp01(float v) { return std::min(std::max(v, 0.0f), 1.0f); }

static void set_solid(DevPaint *o, const float c[4])
{
    o->kind = 0;
    if (c[3] == 1.0f) { o->premul[0] = c[0]; o->premul[1] = c[1]; o->premul[2] = c[2]; }
    else { o->premul[0] = clamp01(c[0] * c[3]); o->premul[1] = clamp01(c[1] * c[3]); o->premul[2] = clamp01(c[2] * c[3]); }
    o->premul[3] = c[3];
    for (int k = 0; k < 4; k++) o->solid16[k] = f2u16(o->premul[k] * 255.0f + 0.5f);
}

static Xform translate(float x, float y) { Xform r; r.tx = x; r.ty = y; return r; }
static Xform scale(float x, float y) { Xform r; r.sx = x; r.sy = y; return r; }
static Xform from_poly2(Pt p0, Pt p1)
{
    Xform r;
    r.sx = p1.y - p0.y; r.ky = p0.x - p1.x; r.kx = p1.x - p0.x; r.sy = p1.y - p0.y; r.tx = p0.x; r.ty = p0.y;
    return r;
}
static bool poly_to_poly(Pt s0, Pt s1, Pt d0, Pt d1, Xform *out)
{
    Xform res;
    if (!invert(from_poly2(s0, s1), &res)) return false;
    *out = pre_concat(from_poly2(d0, d1), res);
    return true;
}

struct Stop { float pos; float c[4]; };

// Gradient::new + the colour part of Gradient::push_stages
static bool setup_stops(DevPaint *o, const float *in, int n_in, int spread, bool *opaque, std::vector<DevStop> &pool)
{
    float F[kMaxStops + 2][4], B[kMaxStops + 2][4], T[kMaxStops + 2];
    auto flush = [&](int len) {
        o->stop_off = (uint32_t)pool.size();
        for (int i = 0; i < len; i++) {
            DevStop d;
            memset(&d, 0, sizeof(d));
            memcpy(d.f, F[i], 16);
            memcpy(d.b, B[i], 16);
            d.t0 = T[i];
            pool.push_back(d);
        }
    };
    if (n_in > kMaxStops) return false;
    Stop st[kMaxStops + 2];
    int n = 0;
    float first = clamp01(in[0]), lastp = clamp01(in[(n_in - 1) * 5]);
    bool dfirst = first != 0.0f, dlast = lastp != 1.0f;
    if (dfirst) { st[n].pos = 0.0f; memcpy(st[n].c, in + 1, 16); n++; }
    for (int i = 0; i < n_in; i++) {
        float p = in[i * 5];
        st[n].pos = p != p ? 0.0f : clamp01(p);
        memcpy(st[n].c, in + i * 5 + 1, 16);
        n++;
    }
    if (dlast) { st[n].pos = 1.0f; memcpy(st[n].c, in + (n_in - 1) * 5 + 1, 16); n++; }
    *opaque = true;
    for (int i = 0; i < n; i++) if (st[i].c[3] != 1.0f) *opaque = false;
    int start = dfirst ? 0 : 1;
    float prev = 0.0f;
    bool uniform = true;
    float step = st[start].pos - prev;
    for (int i = start; i < n; i++) {
        float curr = (i + 1 == n) ? 1.0f : std::min(std::max(st[i].pos, prev), 1.0f);
        uniform = uniform && fabsf(step - (curr - prev)) <= kNearlyZero;
        st[i].pos = curr;
        prev = curr;
    }
    o->spread = spread;
    o->pad_x1 = (spread == 0 && uniform) ? 1 : 0;
    o->premul_after = *opaque ? 0 : 1;
    if (n == 2) {
        o->two_stop = 1;
        o->len = 1;
        for (int k = 0; k < 4; k++) { F[0][k] = st[1].c[k] - st[0].c[k]; B[0][k] = st[0].c[k]; }
        T[0] = 0.0f;
        flush(1);
        return true;
    }
    o->two_stop = 0;
    int first_stop = memcmp(st[0].c, st[1].c, 16) != 0 ? 0 : 1;
    int last_stop = (memcmp(st[n - 2].c, st[n - 1].c, 16) != 0 ? n : n - 1) - 1;
    float t_l = st[first_stop].pos;
    float c_l[4];
    memcpy(c_l, st[first_stop].c, 16);
    int len = 0;
    for (int k = 0; k < 4; k++) { F[len][k] = 0.0f; B[len][k] = c_l[k]; }
    T[len++] = 0.0f;
    for (int i = first_stop; i < last_stop; i++) {
        float t_r = st[i + 1].pos;
        const float *c_r = st[i + 1].c;
        if (t_l < t_r) {
            for (int k = 0; k < 4; k++) {
                float ff = (c_r[k] - c_l[k]) / (t_r - t_l);
                F[len][k] = ff;
                B[len][k] = c_l[k] - ff * t_l;
            }
            T[len++] = t_l;
        }
        t_l = t_r;
        memcpy(c_l, c_r, 16);
    }
    for (int k = 0; k < 4; k++) { F[len][k] = 0.0f; B[len][k] = c_l[k]; }
    T[len++] = t_l;
    o->len = len;
    for (int i = 0; i < 8; i++) o->t0s[i] = i < len ? T[i] : INFINITY;
    flush(len);
    return true;
}

static bool blend_is_lowp(int m) { return !(m == 18 || m == 19 || m == 21 || m >= 25); }

bool prepare_paint(const rb_paint *p, const Xform &ctm, DevPaint *o, std::vector<DevStop> &pool)
{
    memset(o, 0, sizeof(*o));
    bool lowp_ok = true, opaque = false;
    const float degenerate = 1.0f / (1 << 15);
    if (p->shader == 0) {
        set_solid(o, p->color);
        opaque = p->color[3] == 1.0f;
    } else {
        Xform local = post_concat(Xform::from(p->ts), ctm), inv;
        if (p->shader == 3) {
            if (!p->pattern || !invert(local, &inv)) return false;
            o->kind = 2;
            lowp_ok = false;
            memcpy(o->ts, &inv, sizeof(float) * 6);
            o->has_ts = inv.is_finite() && !inv.is_identity();
            o->pix = (const uint8_t *)rb_layer_device_ptr(const_cast<rb_layer *>(p->pattern));
            o->pw = rb_layer_width(p->pattern);
            o->ph = rb_layer_height(p->pattern);
            if (o->pw == 0 || o->ph == 0) return false;
            o->spread = p->spread;
            o->quality = (inv.is_identity() || inv.is_translate()) ? 0 : p->quality;
            o->opacity = p->opacity;
        } else {
            if (p->n_stops < 1 || !p->stops) return false;
            const float *lastc = p->stops + (p->n_stops - 1) * 5 + 1;
            bool solid = false;
            const float *solid_c = nullptr;
            Xform unit;
            Pt c0{p->x0, p->y0}, c1{p->x1, p->y1};
            if (p->n_stops == 1) { solid = true; solid_c = p->stops + 1; }
            else if (!invert(local, &inv)) return false;
            else if (p->shader == 1) {
                float dx = c1.x - c0.x, dy = c1.y - c0.y;
                float len = sqrtf(dx * dx + dy * dy);
                if (!std::isfinite(len)) return false;
                if (nearly_zero(len, degenerate)) { solid = true; solid_c = lastc; }
                else { // points_to_unit_ts
                    float im = 1.0f / len;
                    float vx = dx * im, vy = dy * im;
                    float sn = -vy, cs = vx, ci = 1.0f - cs;
                    Xform t;
                    t.sx = cs; t.ky = sn; t.kx = -sn; t.sy = cs;
                    t.tx = sn * c0.y + ci * c0.x;
                    t.ty = -sn * c0.x + ci * c0.y;
                    t = post_concat(t, translate(-c0.x, -c0.y));
                    unit = post_concat(t, scale(im, im));
                    o->geom = 0;
                }
            } else {
                float r0 = p->r0, r1 = p->r1;
                if (r0 < 0 || r1 < 0) return false;
                float dx = c0.x - c1.x, dy = c0.y - c1.y;
                float dlen = sqrtf(dx * dx + dy * dy);
                if (nearly_zero(dlen, degenerate)) {
                    if (nearly_zero(r0 - r1, degenerate)) { solid = true; solid_c = lastc; }
                    else if (nearly_zero(r0, degenerate)) {
                        float ir = 1.0f / r1;
                        unit = post_concat(translate(-c0.x, -c0.y), scale(ir, ir));
                        o->geom = 1;
                    } else {
                        float sc = 1.0f / std::max(r0, r1);
                        unit = post_concat(translate(-c1.x, -c1.y), scale(sc, sc));
                        float dr = r1 - r0;
                        o->conc_scale = std::max(r0, r1) / dr;
                        o->conc_bias = -r0 / dr;
                        o->geom = 4;
                        lowp_ok = false;
                    }
                } else {
                    if (!poly_to_poly(c0, c1, Pt{0, 0}, Pt{1, 0}, &unit)) return false;
                    lowp_ok = false;
                    if (nearly_zero(r1 - r0)) {
                        o->geom = 3;
                        float s0 = r0 / dlen;
                        o->p0 = s0 * s0;
                    } else {
                        o->geom = 2;
                        float fr0 = r0 / dlen, fr1 = r1 / dlen;
                        float fx = fr0 / (fr0 - fr1);
                        if (nearly_zero(fx - 1.0f)) {
                            unit = post_concat(unit, translate(-1.0f, 0.0f));
                            unit = post_concat(unit, scale(-1.0f, 1.0f));
                            std::swap(fr0, fr1);
                            fx = 0.0f;
                            o->swapped = 1;
                        }
                        Xform fm;
                        if (!poly_to_poly(Pt{fx, 0}, Pt{1, 0}, Pt{0, 0}, Pt{1, 0}, &fm)) return false;
                        unit = post_concat(unit, fm);
                        float fr = fr1 / fabsf(1.0f - fx);
                        o->focal_on_circle = nearly_zero(1.0f - fr);
                        o->well_behaved = !o->focal_on_circle && fr > 1.0f;
                        o->natively_focal = nearly_zero(fx);
                        if (o->focal_on_circle) unit = post_concat(unit, scale(0.5f, 0.5f));
                        else unit = post_concat(unit, scale(fr / (fr * fr - 1.0f), 1.0f / sqrtf(fabsf(fr * fr - 1.0f))));
                        float af = fabsf(1.0f - fx);
                        unit = post_concat(unit, scale(af, af));
                        o->p0 = 1.0f / fr;
                        o->p1 = fx;
                        o->negate_x = (1.0f - fx) < 0.0f;
                        o->smaller = o->swapped || o->negate_x;
                    }
                }
            }
            if (solid) {
                set_solid(o, solid_c);
                opaque = solid_c[3] == 1.0f;
            } else {
                o->kind = 1;
                Xform total = post_concat(inv, unit);
                memcpy(o->ts, &total, sizeof(float) * 6);
                o->has_ts = total.is_finite() && !total.is_identity();
                if (!setup_stops(o, p->stops, p->n_stops, p->spread, &opaque, pool)) return false;
            }
        }
    }
    int blend = p->blend_mode;
    if (blend == 2) return false;                // Destination
    if (blend == 6 && opaque) return false;      // DestinationIn with an opaque source
    if (opaque && blend == 3) blend = 1;         // SourceOver -> Source
    if (o->kind == 0 && blend == 1) {
        o->has_memset = 1;
        uint32_t c[4];
        for (int k = 0; k < 4; k++) c[k] = std::min(o->solid16[k], 255u);
        o->memset_color = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
    }
    if (blend == 0) { // Clear = transparent Source
        const float zero[4] = {0, 0, 0, 0};
        set_solid(o, zero);
        blend = 1;
        o->has_memset = 1;
        o->memset_color = 0;
        lowp_ok = true;
    }
    o->blend = blend;
    o->lowp = (lowp_ok && blend_is_lowp(blend) && !p->force_hq) ? 1 : 0;
    return true;
}

} // namespace rbh
