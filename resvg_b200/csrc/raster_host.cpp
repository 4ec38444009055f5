// raster_host.cpp — see raster_host.h.  Host-side geometry for the device rasteriser.
//
// Unlike the reference's scan converter, which keeps quadratic/cubic edges as stateful objects that the
// scanline walker advances (tiny-skia edge.rs QuadraticEdge::update / CubicEdge::update), every curve is
// forward-differenced here, once, into the line edges those updates would produce, so the device only
// ever sees independent `Edge {x, dx, first_y, last_y, winding}` records it can evaluate in closed form:
// x(y) = x + (y - first_y) * dx (wrapping i32).
#include "raster_host.h"

#include "edge_math.h"
#include "fill_core.h"
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
// curve chopping, edge emission, clipping, the fill_path front end: fill_core.h (shared with the device)
// ---------------------------------------------------------------------------------------------------
template <class T> using HVec = std::vector<T>;
typedef geo::fl::Sink<HVec> Sink;
using geo::fl::walk_path;
using geo::fl::finish_geom;

#ifdef RB_HOST_PROFILE
std::atomic<uint64_t> g_bd_prof[4];
#define BD_T(i) do { uint64_t now__ = __rdtsc(); g_bd_prof[i] += now__ - bd_t__; bd_t__ = now__; } while (0)
#else
#define BD_T(i)
#endif
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
    const geo::P *gp = reinterpret_cast<const geo::P *>(pts);
    if (!walk_path(verbs, n_verbs, gp, n_pts, anti_alias, cw, ch, sink, g, &ir, &inside)) { out.resize(sink.base); return false; }
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
    static thread_local Sink sink;
    sink.out = &lines;
    sink.base = lines.size();
    sink.curves = &curves;
    sink.kinds.clear();
    sink.n_items = 0;
    const size_t cbase = curves.size();
    if (cbase != 0) return false; // the builder's curve list is per draw
    const geo::P *gp = reinterpret_cast<const geo::P *>(pts);
    if (!geo::fl::build_items<HVec>(verbs, n_verbs, gp, n_pts, anti_alias, cw, ch, sink, g)) { lines.resize(sink.base); curves.resize(cbase); return false; }
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
    for (int i = 0; i < 12; i++) o->t0s[i] = i < len ? T[i] : INFINITY;
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
