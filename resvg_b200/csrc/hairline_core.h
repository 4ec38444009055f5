// hairline_core.h — anti-aliased hairline strokes: what PixmapMut::stroke_path draws when treat_as_hairline says the
// transformed stroke is at most one pixel wide (tiny-skia painter.rs stroke_path / stroke_hairline).
//
// Restates tiny-skia 0.12.0 scan/hairline.rs (stroke_path_impl, extend_pts, hair_quad, hair_cubic) and
// scan/hairline_aa.rs (anti_hair_line_rgn, do_anti_hairline and its four span blitters) — ports of Skia's
// SkScan_Hairline.cpp / SkScan_Antihair.cpp.  A hairline is not scan-converted: every path segment is walked in
// fixed point along its major axis and blitted two pixels at a time, each blit being a separate blend, so a pixel
// touched twice is blended twice.  The walker therefore produces the ordered list of (x, y, alpha) blits; the tile
// kernel applies them per pixel in that order.
//
// One source for both sides (geom_common.h): hairline.cpp instantiates it with std::vector, geo.cu with DVec.  `Pts` is
// anything indexable that yields the path's points in device space (a pointer, or a functor mapping them on the fly).
#pragma once

#include "edge_math.h"
#include "geom_common.h"
#include "stroker_core.h"

namespace geo {

// One blit of the hairline walker: blend the paint into pixel (x, y) with coverage alpha (1..255), in list order.
struct HairBlit { int32_t x, y; uint32_t alpha; };

namespace hl {

GEO_HDI inline bool is_zero(P p) { return p.x == 0.0f && p.y == 0.0f; }

typedef int32_t FDot6;
typedef int32_t FDot16;
constexpr FDot16 F16_HALF = 1 << 15, F16_ONE = 1 << 16;

GEO_HDI inline int32_t shl(int32_t v, int s) { return (int32_t)((uint32_t)v << s); }
GEO_HDI inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
GEO_HDI inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
GEO_HDI inline FDot6 fdot6_from_f32(float v) { return f2i(v * 64.0f); }
GEO_HDI inline int32_t fdot6_floor(FDot6 v) { return v >> 6; }
GEO_HDI inline int32_t fdot6_ceil(FDot6 v) { return wadd(v, 63) >> 6; }
GEO_HDI inline FDot16 fdot6_to_fdot16(FDot6 v) { return shl(v, 10); }
GEO_HDI inline FDot16 fast_div(FDot6 a, FDot6 b) { return shl(a, 16) / b; }
GEO_HDI inline uint32_t small_scale(uint32_t value, int32_t dot6) { return (uint32_t)(((int32_t)value * dot6) >> 6) & 0xffu; }
GEO_HDI inline uint32_t i32_to_alpha(int32_t a) { return (uint32_t)a & 0xffu; }

template <template <class> class Vec> struct HairSink {
    Vec<HairBlit> *out;
    int32_t w, h;
    // the sub-clip of the line being walked: (line bounds + 1) ∩ clip when the line's bounds leave the clip — what the
    // reference hands to do_anti_hairline and wraps the blitter in (RectClipBlitter); the whole clip otherwise
    int64_t sl = 0, st = 0, sr = 0, sb = 0;
    GEO_HDI void px(int64_t x, int64_t y, uint32_t a)
    {
        if (a == 0 || x < sl || y < st || x >= sr || y >= sb) return;
        out->push_back(HairBlit{(int32_t)x, (int32_t)y, a});
    }
    GEO_HDI void anti_h2(int64_t x, int64_t y, uint32_t a0, uint32_t a1) { px(x, y, a0); px(x + 1, y, a1); }
    GEO_HDI void anti_v2(int64_t x, int64_t y, uint32_t a0, uint32_t a1) { px(x, y, a0); px(x, y + 1, a1); }
    GEO_HD void v(int64_t x, int64_t y, int32_t height, uint32_t a) { for (int32_t i = 0; i < height; i++) px(x, y + i, a); }
    GEO_HD void hline(int64_t x, int64_t y, int32_t count, uint32_t a) { if (y < 0) return; for (int32_t i = 0; i < count; i++) px(x + i, y, a); }
};

// ---- the four span blitters of hairline_aa.rs ----------------------------------------------------------------------------
enum Kind { HLine, Horish, VLine, Vertish };

template <class Sink> GEO_HD FDot16 draw_cap(Sink &s, Kind k, int32_t at, FDot16 f, FDot16 slope, int32_t mod64)
{
    f = gmax(wadd(f, F16_HALF), 0);
    const int32_t lower = f >> 16;
    const uint32_t a = i32_to_alpha(f >> 8);
    switch (k) {
    case HLine: {
        uint32_t ma = small_scale(a, mod64);
        if (ma) s.hline(at, lower, 1, ma);
        ma = small_scale(255 - a, mod64);
        if (ma) s.hline(at, (int64_t)lower - 1, 1, ma); // y.checked_sub(1): nothing above row 0
        return f - F16_HALF;
    }
    case Horish:
        s.anti_v2(at, gmax(lower, 1) - 1, small_scale(255 - a, mod64), small_scale(a, mod64));
        return wadd(f, slope) - F16_HALF;
    case VLine: {
        uint32_t ma = small_scale(a, mod64);
        if (ma) s.v(lower, at, 1, ma);
        ma = small_scale(255 - a, mod64);
        if (ma) s.v(gmax(lower, 1) - 1, at, 1, ma);
        return f - F16_HALF;
    }
    default:
        s.anti_h2(gmax(lower, 1) - 1, at, small_scale(255 - a, mod64), small_scale(a, mod64));
        return wadd(f, slope) - F16_HALF;
    }
}

template <class Sink> GEO_HD FDot16 draw_line(Sink &s, Kind k, int32_t at, int32_t stop, FDot16 f, FDot16 slope)
{
    switch (k) {
    case HLine: {
        const int32_t count = stop - at;
        if (count == 0) return f;
        f = gmax(wadd(f, F16_HALF), 0);
        const int32_t y = f >> 16;
        uint32_t a = i32_to_alpha(f >> 8);
        if (a) s.hline(at, y, count, a);
        a = 255 - a;
        if (a) s.hline(at, (int64_t)y - 1, count, a);
        return f - F16_HALF;
    }
    case Horish: {
        f = wadd(f, F16_HALF);
        do {
            f = gmax(f, 0);
            const int32_t lower = f >> 16;
            const uint32_t a = i32_to_alpha(f >> 8);
            s.anti_v2(at, gmax(lower, 1) - 1, 255 - a, a);
            f = wadd(f, slope);
            at++;
        } while (at < stop);
        return f - F16_HALF;
    }
    case VLine: {
        const int32_t height = stop - at;
        if (height == 0) return f;
        f = gmax(wadd(f, F16_HALF), 0);
        const int32_t x = f >> 16;
        uint32_t a = i32_to_alpha(f >> 8);
        if (a) s.v(x, at, height, a);
        a = 255 - a;
        if (a) s.v(gmax(x, 1) - 1, at, height, a);
        return f - F16_HALF;
    }
    default: {
        f = wadd(f, F16_HALF);
        do {
            f = gmax(f, 0);
            const int32_t x = f >> 16;
            const uint32_t a = i32_to_alpha(f >> 8);
            s.anti_h2(gmax(x, 1) - 1, at, 255 - a, a);
            f = wadd(f, slope);
            at++;
        } while (at < stop);
        return f - F16_HALF;
    }
    }
}

// do_anti_hairline.  `clipped`: the line's bounds leave the clip and Sink::{sl,st,sr,sb} hold the sub-clip.  The reference
// then starts the walk at the clip edge (fstart jumps by slope * skipped columns — NOT the same as stepping there, because
// every step clamps the ordinate at 0), ends it at the far edge, and drops the remaining outside pixels in a RectClipBlitter
// (Sink::px).
template <class Sink> GEO_HD void do_anti_hairline(Sink &s, FDot6 x0, FDot6 y0, FDot6 x1, FDot6 y1, bool clipped)
{
    if (x0 == INT32_MIN || y0 == INT32_MIN || x1 == INT32_MIN || y1 == INT32_MIN) return; // any_bad_ints
    if (abs(x1 - x0) > (511 << 6) || abs(y1 - y0) > (511 << 6)) {
        // long lines are halved so that the FDot16 slope arithmetic cannot overflow
        const int32_t hx = (x0 >> 1) + (x1 >> 1), hy = (y0 >> 1) + (y1 >> 1);
        do_anti_hairline(s, x0, y0, hx, hy, clipped);
        do_anti_hairline(s, hx, hy, x1, y1, clipped);
        return;
    }
    int32_t scale_start, scale_stop, istart, istop;
    FDot16 fstart, slope;
    Kind kind;
    if (abs(x1 - x0) > abs(y1 - y0)) { // mostly horizontal
        if (x0 > x1) { gswap(x0, x1); gswap(y0, y1); }
        istart = fdot6_floor(x0);
        istop = fdot6_ceil(x1);
        fstart = fdot6_to_fdot16(y0);
        if (y0 == y1) { slope = 0; kind = HLine; }
        else {
            slope = fast_div(y1 - y0, x1 - x0);
            fstart = wadd(fstart, (wmul(slope, 32 - (x0 & 63)) + 32) >> 6);
            kind = Horish;
        }
        if (istop - istart == 1) { scale_start = x1 - x0; scale_stop = 0; } // within a single pixel
        else { scale_start = 64 - (x0 & 63); scale_stop = x1 & 63; }
        if (clipped) {
            if (istart >= s.sr || istop <= s.sl) return;
            if (istart < s.sl) {
                fstart = wadd(fstart, wmul(slope, (int32_t)(s.sl - istart)));
                istart = (int32_t)s.sl;
                scale_start = 64;
                if (istop - istart == 1) { scale_start = ((x1 - 1) & 63) + 1; scale_stop = 0; } // contribution_64
            }
            if (istop > s.sr) { istop = (int32_t)s.sr; scale_stop = 0; } // the last column is not drawn
            if (istart == istop) return;
            // rows the walk can touch, outset by one; wholly outside the clip -> nothing to draw
            int32_t top, bottom;
            const FDot16 fend = wadd(fstart, wmul(istop - istart - 1, slope));
            if (slope >= 0) { top = wadd(fstart, -F16_HALF) >> 16; bottom = (int32_t)(((int64_t)wadd(fend, F16_HALF) + 65535) >> 16); }
            else { bottom = (int32_t)(((int64_t)wadd(fstart, F16_HALF) + 65535) >> 16); top = wadd(fend, -F16_HALF) >> 16; }
            top -= 1; bottom += 1;
            if (top >= s.sb || bottom <= s.st) return;
        }
    } else { // mostly vertical
        if (y0 > y1) { gswap(x0, x1); gswap(y0, y1); }
        istart = fdot6_floor(y0);
        istop = fdot6_ceil(y1);
        fstart = fdot6_to_fdot16(x0);
        if (x0 == x1) {
            if (y0 == y1) return; // nothing to do
            slope = 0;
            kind = VLine;
        } else {
            slope = fast_div(x1 - x0, y1 - y0);
            fstart = wadd(fstart, (wmul(slope, 32 - (y0 & 63)) + 32) >> 6);
            kind = Vertish;
        }
        if (istop - istart == 1) { scale_start = y1 - y0; scale_stop = 0; }
        else { scale_start = 64 - (y0 & 63); scale_stop = y1 & 63; }
        if (clipped) {
            if (istart >= s.sb || istop <= s.st) return;
            if (istart < s.st) {
                fstart = wadd(fstart, wmul(slope, (int32_t)(s.st - istart)));
                istart = (int32_t)s.st;
                scale_start = 64;
                if (istop - istart == 1) { scale_start = ((y1 - 1) & 63) + 1; scale_stop = 0; }
            }
            if (istop > s.sb) { istop = (int32_t)s.sb; scale_stop = 0; }
            if (istart == istop) return;
            int32_t left, right;
            const FDot16 fend = wadd(fstart, wmul(istop - istart - 1, slope));
            if (slope >= 0) { left = wadd(fstart, -F16_HALF) >> 16; right = (int32_t)(((int64_t)wadd(fend, F16_HALF) + 65535) >> 16); }
            else { right = (int32_t)(((int64_t)wadd(fstart, F16_HALF) + 65535) >> 16); left = wadd(fend, -F16_HALF) >> 16; }
            left -= 1; right += 1;
            if (left >= s.sr || right <= s.sl) return;
        }
    }
    // the first pixel(s) are scaled by scale_start, the last by scale_stop, the full spans in between are not
    fstart = draw_cap(s, kind, istart, fstart, slope, scale_start);
    istart += 1;
    const int32_t full_spans = istop - istart - (scale_stop > 0 ? 1 : 0);
    if (full_spans > 0) fstart = draw_line(s, kind, istart, istart + full_spans, fstart, slope);
    if (scale_stop > 0) draw_cap(s, kind, istop - 1, fstart, slope, scale_stop);
}

// ---- line_clipper::intersect ------------------------------------------------------------------------------------------------
struct R { float l, t, r, b; };
GEO_HDI inline bool nested_lt(float a, float b, float dim) { return a <= b && (a < b || dim > 0.0f); }
GEO_HD inline float sect_with_horizontal(const P s[2], float y)
{
    const float dx = s[1].x - s[0].x;
    if (dx == 0.0f) return s[0].x;
    const double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
    double r = x0 + ((double)y - y0) * (x1 - x0) / (y1 - y0);
    const double lo = gmin(x0, x1), hi = gmax(x0, x1); // pin_unsorted
    r = r < lo ? lo : (r > hi ? hi : r);
    return (float)r;
}
GEO_HD inline float sect_with_vertical(const P s[2], float x)
{
    const float dy = s[1].y - s[0].y;
    if (dy == 0.0f) return s[0].y;
    const double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
    return (float)(y0 + ((double)x - x0) * (y1 - y0) / (x1 - x0));
}
GEO_HD inline bool line_intersect(const P src[2], const R &clip, P dst[2])
{
    const float bl = gmin(src[0].x, src[1].x), bt = gmin(src[0].y, src[1].y);
    const float br = gmax(src[0].x, src[1].x), bb = gmax(src[0].y, src[1].y);
    if (gfinite(bl) && gfinite(bt) && gfinite(br) && gfinite(bb)) {
        if (clip.l <= bl && clip.t <= bt && clip.r >= br && clip.b >= bb) { dst[0] = src[0]; dst[1] = src[1]; return true; }
        // no overlap; coincident edges are only allowed when the line runs along the clip edge
        if (nested_lt(br, clip.l, br - bl) || nested_lt(clip.r, bl, br - bl) || nested_lt(bb, clip.t, bb - bt) || nested_lt(clip.b, bt, bb - bt))
            return false;
    }
    int i0 = src[0].y < src[1].y ? 0 : 1, i1 = 1 - i0;
    P tmp[2] = {src[0], src[1]};
    if (tmp[i0].y < clip.t) tmp[i0] = P{sect_with_horizontal(src, clip.t), clip.t};
    if (tmp[i1].y > clip.b) tmp[i1] = P{sect_with_horizontal(src, clip.b), clip.b};
    i0 = tmp[0].x < tmp[1].x ? 0 : 1;
    i1 = 1 - i0;
    // quick reject in x again, now that the line may have been chopped
    if (tmp[i1].x <= clip.l || tmp[i0].x >= clip.r) {
        // a vertical line coincident with the clip edge survives
        if (tmp[0].x != tmp[1].x || tmp[0].x < clip.l || tmp[0].x > clip.r) return false;
    }
    if (tmp[i0].x < clip.l) tmp[i0] = P{clip.l, sect_with_vertical(src, clip.l)};
    if (tmp[i1].x > clip.r) tmp[i1] = P{clip.r, sect_with_vertical(src, clip.r)};
    dst[0] = tmp[0];
    dst[1] = tmp[1];
    return true;
}

// anti_hair_line_rgn for one line of a polyline; false = the reference leaves the polyline here (IntRect::from_ltrb -> None)
template <class Sink> GEO_HD bool anti_hair_line(Sink &s, P p0, P p1)
{
    const R fixed_bounds{-32767.0f, -32767.0f, 32767.0f, 32767.0f};
    // antialiased hairlines can draw up to half a pixel outside their bounds: the scalar pre-clip is outset by one
    const R clip_bounds{-1.0f, -1.0f, (float)s.w + 1.0f, (float)s.h + 1.0f};
    P a[2] = {p0, p1}, b[2], c[2];
    if (!line_intersect(a, fixed_bounds, b)) return true;
    if (!line_intersect(b, clip_bounds, c)) return true;
    const FDot6 x0 = fdot6_from_f32(c[0].x), y0 = fdot6_from_f32(c[0].y), x1 = fdot6_from_f32(c[1].x), y1 = fdot6_from_f32(c[1].y);
    // integral reject against the clip (the reference then narrows to a sub-clip, see do_anti_hairline)
    const int64_t il = (int64_t)fdot6_floor(gmin(x0, x1)) - 1, it = (int64_t)fdot6_floor(gmin(y0, y1)) - 1;
    const int64_t ir = (int64_t)fdot6_ceil(gmax(x0, x1)) + 1, ib = (int64_t)fdot6_ceil(gmax(y0, y1)) + 1;
    if (ir - il <= 0 || ib - it <= 0 || ir - il > INT32_MAX || ib - it > INT32_MAX) return false;
    if (il >= s.w || it >= s.h || ir <= 0 || ib <= 0) return true;
    // the walk below visits every column / row of the line; pixels outside the sub-clip are dropped by Sink::px.  The
    // border pairs that tiny-skia's unsigned coordinates shift inwards (max(1) - 1) are cut by it like any other pixel.
    s.sl = gmax<int64_t>(il, 0); s.st = gmax<int64_t>(it, 0);
    s.sr = gmin<int64_t>(ir, s.w); s.sb = gmin<int64_t>(ib, s.h);
    const bool contained = il >= 0 && it >= 0 && ir <= s.w && ib <= s.h;
    do_anti_hairline(s, x0, y0, x1, y1, !contained);
    return true;
}
template <class Sink> GEO_HD void anti_hair_lines(Sink &s, const P *pts, int n)
{
    for (int i = 0; i + 1 < n; i++) if (!anti_hair_line(s, pts[i], pts[i + 1])) return;
}

// ---- curves (hairline.rs) ----------------------------------------------------------------------------------------------------
constexpr int MAX_QUAD_LEVEL = 5, MAX_CUBIC_LEVEL = 9;

GEO_HDI inline int sat_ceil_i32(float v) { return f2i(ceilf(v)); }
GEO_HD inline uint32_t compute_int_quad_dist(const P p[3])
{
    const float dx = fabsf((p[0].x + p[2].x) * 0.5f - p[1].x), dy = fabsf((p[0].y + p[2].y) * 0.5f - p[1].y);
    const uint32_t idx = (uint32_t)sat_ceil_i32(dx), idy = (uint32_t)sat_ceil_i32(dy);
    return idx > idy ? idx + (idy >> 1) : idy + (idx >> 1);
}
GEO_HD inline int compute_quad_level(const P p[3])
{
    const uint32_t d = compute_int_quad_dist(p);
    int level = (33 - rbe::clz32(d)) / 2;
    return gmin(level, MAX_QUAD_LEVEL);
}
struct Cull { bool on; R inset, outset; };
GEO_HDI inline bool overlaps(const R &a, const R &b) { return a.l < b.r && b.l < a.r && a.t < b.b && b.t < a.b; }     // geometric_overlap
GEO_HDI inline bool contains(const R &o, const R &i) { return o.l <= i.l && o.t <= i.t && o.r >= i.r && o.b >= i.b; } // geometric_contains

template <class Sink> GEO_HD void hair_quad(Sink &s, const P p[3], const Cull &cull, int level)
{
    if (cull.on) {
        const R b{gmin(p[0].x, gmin(p[1].x, p[2].x)), gmin(p[0].y, gmin(p[1].y, p[2].y)),
                  gmax(p[0].x, gmax(p[1].x, p[2].x)), gmax(p[0].y, gmax(p[1].y, p[2].y))};
        if (!(b.l <= b.r && b.t <= b.b)) return;
        if (!overlaps(cull.outset, b)) return;
    }
    // QuadCoeff: (A t + B) t + C
    const float ax = p[2].x - 2.0f * p[1].x + p[0].x, ay = p[2].y - 2.0f * p[1].y + p[0].y;
    const float bx = 2.0f * (p[1].x - p[0].x), by = 2.0f * (p[1].y - p[0].y);
    const int lines = 1 << level;
    P tmp[(1 << MAX_QUAD_LEVEL) + 1];
    tmp[0] = p[0];
    float t = 0.0f;
    const float dt = 1.0f / (float)lines;
    for (int i = 1; i < lines; i++) {
        t = t + dt;
        tmp[i] = P{(ax * t + bx) * t + p[0].x, (ay * t + by) * t + p[0].y};
    }
    tmp[lines] = p[2];
    anti_hair_lines(s, tmp, lines + 1);
}

GEO_HD inline int compute_cubic_segments(const P p[4])
{
    const float third = 1.0f / 3.0f, two_third = 2.0f / 3.0f;
    const float p13x = third * p[3].x + two_third * p[0].x, p13y = third * p[3].y + two_third * p[0].y;
    const float p23x = third * p[0].x + two_third * p[3].x, p23y = third * p[0].y + two_third * p[3].y;
    const float dx = gmax(fabsf(p[1].x - p13x), fabsf(p[2].x - p23x)), dy = gmax(fabsf(p[1].y - p13y), fabsf(p[2].y - p23y));
    const float diff = gmax(dx, dy);
    float tol = 1.0f / 8.0f;
    for (int i = 0; i < MAX_CUBIC_LEVEL; i++) {
        if (diff < tol) return 1 << i;
        tol *= 4.0f;
    }
    return 1 << MAX_CUBIC_LEVEL;
}
template <class Sink> GEO_HD void hair_cubic2(Sink &s, const P p[4])
{
    const int lines = compute_cubic_segments(p);
    if (lines == 1) { const P l[2] = {p[0], p[3]}; anti_hair_lines(s, l, 2); return; }
    // CubicCoeff: ((A t + B) t + C) t + D
    const float ax = p[3].x + 3.0f * (p[1].x - p[2].x) - p[0].x, ay = p[3].y + 3.0f * (p[1].y - p[2].y) - p[0].y;
    const float bx = 3.0f * (p[2].x - 2.0f * p[1].x + p[0].x), by = 3.0f * (p[2].y - 2.0f * p[1].y + p[0].y);
    const float cx = 3.0f * (p[1].x - p[0].x), cy = 3.0f * (p[1].y - p[0].y);
    bool ok = finite(p[0]);
    {
        float t = 0.0f;
        const float dt = 1.0f / (float)lines;
        for (int i = 1; i < lines; i++) {
            t = t + dt;
            ok = ok && finite(P{((ax * t + bx) * t + cx) * t + p[0].x, ((ay * t + by) * t + cy) * t + p[0].y});
        }
    }
    if (!ok) return; // some point is not finite: nothing is drawn
    float t = 0.0f;
    const float dt = 1.0f / (float)lines;
    P prev = p[0];
    for (int i = 1; i < lines; i++) {
        t = t + dt;
        const P cur{((ax * t + bx) * t + cx) * t + p[0].x, ((ay * t + by) * t + cy) * t + p[0].y};
        if (!anti_hair_line(s, prev, cur)) return;
        prev = cur;
    }
    anti_hair_line(s, prev, p[3]);
}
GEO_HDI inline float dot(P a, P b) { return a.x * b.x + a.y * b.y; }
GEO_HDI inline bool lt_90(P p0, P pivot, P p2) { return dot(p0 - pivot, p2 - pivot) >= 0.0f; }

template <class Sink> GEO_HD void hair_cubic(Sink &s, const P p[4], const Cull &cull)
{
    if (cull.on) {
        R b{p[0].x, p[0].y, p[0].x, p[0].y};
        for (int i = 1; i < 4; i++) { b.l = gmin(b.l, p[i].x); b.t = gmin(b.t, p[i].y); b.r = gmax(b.r, p[i].x); b.b = gmax(b.b, p[i].y); }
        if (!(b.l <= b.r && b.t <= b.b)) return;
        if (!overlaps(cull.outset, b)) return;
    }
    // quick_cubic_niceness_check: the off-curve points lie "inside" the limits of the on-curve points
    if (lt_90(p[1], p[0], p[3]) && lt_90(p[2], p[0], p[3]) && lt_90(p[1], p[3], p[0]) && lt_90(p[2], p[3], p[0])) {
        hair_cubic2(s, p);
        return;
    }
    float tv[3];
    const int n = sk::cubic_max_curvature(p, tv);
    // chop_cubic_at(points, t_values): successive chops with renormalised t
    P cur[4] = {p[0], p[1], p[2], p[3]};
    float last_t = 0.0f;
    for (int i = 0; i < n; i++) {
        float t = tv[i];
        if (i > 0) {
            const float denom = 1.0f - last_t;
            t = denom == 0.0f ? 1.0f : (tv[i] - last_t) / denom;
            if (!(t >= 0.0f && t <= 1.0f)) t = 1.0f;
        }
        P d[7];
        sk::chop_cubic(cur, t, d);
        hair_cubic2(s, d);
        cur[0] = d[3]; cur[1] = d[4]; cur[2] = d[5]; cur[3] = d[6];
        last_t = tv[i];
    }
    hair_cubic2(s, cur);
}

// extend_pts: square / round caps of a hairline lengthen the end segments by (half a pixel / pi/8 of a pixel)
GEO_HD inline void extend_pts(int cap, int prev_verb, int next_verb /* -1: none */, P *pts, int n)
{
    const float cap_outset = cap == 2 ? 0.5f : 3.14159265358979323846f / 8.0f;
    if (prev_verb == V_MOVE) {
        const P first = pts[0];
        int offset = 0, controls = n - 1;
        P tangent{0, 0};
        do {
            offset++;
            tangent = first - pts[offset];
        } while (is_zero(tangent) && --controls > 0);
        if (is_zero(tangent)) { tangent = P{1.0f, 0.0f}; controls = n - 1; } // all points equal: move all but one
        else { const float len = sqrtf(tangent.x * tangent.x + tangent.y * tangent.y); tangent = P{tangent.x / len, tangent.y / len}; }
        offset = 0;
        do { // an end point equal to its control points moves in tandem with them
            pts[offset].x += tangent.x * cap_outset;
            pts[offset].y += tangent.y * cap_outset;
            offset++;
            controls++;
        } while (controls < n);
    }
    if (next_verb == V_MOVE || next_verb == V_CLOSE || next_verb < 0) {
        const P last = pts[n - 1];
        int offset = n - 1, controls = n - 1;
        P tangent{0, 0};
        do {
            offset--;
            tangent = last - pts[offset];
        } while (is_zero(tangent) && --controls > 0);
        if (is_zero(tangent)) { tangent = P{-1.0f, 0.0f}; controls = n - 1; }
        else { const float len = sqrtf(tangent.x * tangent.x + tangent.y * tangent.y); tangent = P{tangent.x / len, tangent.y / len}; }
        offset = n - 1;
        do {
            pts[offset].x += tangent.x * cap_outset;
            pts[offset].y += tangent.y * cap_outset;
            offset--;
            controls++;
        } while (controls < n);
    }
}


// hairline::stroke_path_impl with line_proc = anti_hair_line_rgn.  `pts` are in device space relative to the clip
// (0, 0, clip_w, clip_h); cap: 0 butt, 1 round, 2 square.
// only_verb >= 0: just the blits of that verb (the device walks the verbs of a path in parallel, one draw each: the blits of
// a path are those of its verbs one after the other, and a verb's blits depend on its neighbours only through the points
// carried along here).
// Bounds of the whole path (with the probe for non-finite points): what stroke_path_impl decides culling from.
struct HairBounds { float l, t, r, b; bool finite; };
template <class Pts> GEO_HD HairBounds hair_bounds(const Pts &pts, int n_pts)
{
    HairBounds hb;
    hb.l = hb.t = hb.r = hb.b = 0.0f;
    hb.finite = false;
    if (n_pts <= 0) return hb;
    const P p0 = pts[0];
    float l = p0.x, t = p0.y, r = l, b = t, probe = 0.0f;
    for (int i = 0; i < n_pts; i++) {
        const P p = pts[i];
        l = gmin(l, p.x); t = gmin(t, p.y); r = gmax(r, p.x); b = gmax(b, p.y);
        probe += p.x * 0.0f + p.y * 0.0f; // a Path never holds a non-finite point
    }
    hb.l = l; hb.t = t; hb.r = r; hb.b = b;
    hb.finite = probe == 0.0f;
    return hb;
}
// false: nothing of the path is drawn
GEO_HD inline bool hair_plan(const HairBounds &hb, int cap, int32_t clip_w, int32_t clip_h, Cull *cull)
{
    cull->on = false;
    cull->inset = R{0, 0, 0, 0};
    cull->outset = R{0, 0, 0, 0};
    if (!hb.finite) return false;
    const float l = hb.l, t = hb.t, r = hb.r, b = hb.b;
    if (!(gfinite(l) && gfinite(t) && gfinite(r) && gfinite(b))) return false;
    const float o = cap == 0 ? 1.0f : 2.0f;
    const double fl = floor((double)l - o), ft = floor((double)t - o), cr = ceil((double)r + o), cb = ceil((double)b + o); // round_out
    if (fl >= clip_w || ft >= clip_h || cr <= 0 || cb <= 0) return false;
    if (!(fl >= 0 && ft >= 0 && cr <= clip_w && cb <= clip_h)) {
        // per-segment culling rectangles: quick-accept inside the inset clip, quick-reject outside the outset clip
        if (clip_w <= 2 || clip_h <= 2) return false; // inset(1, 1) -> None
        cull->on = true;
        cull->outset = R{-1.0f, -1.0f, (float)clip_w + 1.0f, (float)clip_h + 1.0f};
        cull->inset = R{1.0f, 1.0f, (float)clip_w - 1.0f, (float)clip_h - 1.0f};
    }
    return true;
}

// The walk proper, for a path (or a run of whole contours of it) whose culling has been decided from the whole path.
template <template <class> class Vec, class Pts>
GEO_HD void hairline_walk(const uint8_t *verbs, int n_verbs, const Pts &pts, int cap, int32_t clip_w, int32_t clip_h, const Cull &cull,
                          Vec<HairBlit> &out, int only_verb = -1)
{
    HairSink<Vec> s;
    s.out = &out; s.w = clip_w; s.h = clip_h;
    s.sl = s.st = s.sr = s.sb = 0;
    int prev_verb = V_MOVE, pi = 0;
    P first_pt{0, 0}, last_pt{0, 0};
    for (int vi = 0; vi < n_verbs; vi++) {
        const int verb = verbs[vi];
        const bool walk = only_verb < 0 || only_verb == vi;
        const int next_verb = vi + 1 < n_verbs ? verbs[vi + 1] : -1;
        P last_pt2 = last_pt;
        switch (verb) {
        case V_MOVE:
            first_pt = last_pt = last_pt2 = pts[pi++];
            break;
        case V_LINE: {
            P l[2] = {last_pt, pts[pi++]};
            if (cap != 0) extend_pts(cap, prev_verb, next_verb, l, 2);
            if (walk) anti_hair_lines(s, l, 2);
            last_pt = l[0];
            last_pt2 = l[1];
            break;
        }
        case V_QUAD: {
            P q[3] = {last_pt, pts[pi], pts[pi + 1]};
            pi += 2;
            if (cap != 0) extend_pts(cap, prev_verb, next_verb, q, 3);
            if (walk) hair_quad(s, q, cull, compute_quad_level(q));
            last_pt = q[0];
            last_pt2 = q[2];
            break;
        }
        case V_CUBIC: {
            P c[4] = {last_pt, pts[pi], pts[pi + 1], pts[pi + 2]};
            pi += 3;
            if (cap != 0) extend_pts(cap, prev_verb, next_verb, c, 4);
            if (walk) hair_cubic(s, c, cull);
            last_pt = c[0];
            last_pt2 = c[3];
            break;
        }
        default: { // close
            P l[2] = {last_pt, first_pt};
            if (cap != 0 && prev_verb == V_MOVE) extend_pts(cap, prev_verb, next_verb, l, 2); // degenerate moveTo + close
            if (walk) anti_hair_lines(s, l, 2);
            last_pt2 = l[1];
            break;
        }
        }
        if (cap != 0) {
            if (prev_verb == V_MOVE && (verb == V_LINE || verb == V_QUAD || verb == V_CUBIC))
                first_pt = last_pt; // the cap moved the initial point: close to the moved one
            last_pt = last_pt2;
        } else {
            last_pt = last_pt2;
        }
        prev_verb = verb;
    }
}


template <template <class> class Vec, class Pts>
GEO_HD void hairline_blits(const uint8_t *verbs, int n_verbs, const Pts &pts, int n_pts, int cap, int32_t clip_w, int32_t clip_h, Vec<HairBlit> &out,
                           int only_verb = -1)
{
    if (n_pts <= 0 || n_verbs <= 0) return;
    Cull cull;
    if (!hair_plan(hair_bounds(pts, n_pts), cap, clip_w, clip_h, &cull)) return;
    hairline_walk<Vec>(verbs, n_verbs, pts, cap, clip_w, clip_h, cull, out, only_verb);
}

} // namespace hl
} // namespace geo
