/*
 * hairline.c — CPU oracle for tiny-skia's anti-aliased hairline stroking (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * PixmapMut::stroke_path (resvg path.rs:113) draws a stroke whose transformed width is at most one pixel with
 * scan::hairline_aa::stroke_path instead of filling an outline (tiny-skia painter.rs treat_as_hairline).  tiny-skia 0.12.0
 * is not under /root/reference: this restates the published algorithm of its scan/hairline.rs (path walk, cap extension,
 * quad / cubic subdivision) and scan/hairline_aa.rs (the fixed-point line walker and its four span blitters — the Rust
 * port of Skia's SkScan_Hairline.cpp / SkScan_Antihair.cpp) and returns the coverage blits in the walker's order.
 * Written independently of resvg_b200/csrc/hairline.cpp.  Pinned by the goldens with thin strokes
 * (painting/stroke-width, structure/style, shapes/line ...).
 */
#include "pathgeom.h"

typedef struct { int32_t *v; int n, cap; int cw, ch; } blit_list;

static void emit(blit_list *b, int x, int y, int alpha)
{
    if (alpha <= 0) return;                               /* transparent coverage blends nothing */
    if (x < 0 || y < 0 || x >= b->cw || y >= b->ch) return; /* the clip blitter / pixmap bounds */
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 256; b->v = (int32_t *)realloc(b->v, sizeof(int32_t) * 3 * (size_t)b->cap); }
    b->v[3 * b->n] = x; b->v[3 * b->n + 1] = y; b->v[3 * b->n + 2] = alpha > 255 ? 255 : alpha;
    b->n++;
}

typedef struct { int l, t, r, b; } irect;

/* ---- fixed point (fixed_point.rs) ---- */
static int32_t fdot6_from_f32(float v)
{
    float s = v * 64.0f;
    if (s != s) return 0;
    if (s >= 2147483648.0f) return INT32_MAX;
    if (s <= -2147483648.0f) return INT32_MIN;
    return (int32_t)s;
}
static int fdot6_floor(int32_t v) { return v >> 6; }
static int fdot6_ceil(int32_t v) { return (int)(((int64_t)v + 63) >> 6); }
static int32_t fdot6_to_fdot16(int32_t v) { return (int32_t)((uint32_t)v << 10); }
static int32_t fast_fix_div(int32_t a, int32_t b) { return (int32_t)((uint32_t)a << 16) / b; } /* |a| < 2^15 here */
static int fdot16_floor(int32_t v) { return v >> 16; }
static int fdot16_ceil(int32_t v) { return (int)(((int64_t)v + 65535) >> 16); }
#define FIX_HALF 32768
static int small_dot6_scale(int value, int dot6) { return (value * dot6) >> 6; }
static int contribution_64(int32_t ordinate) { return ((ordinate - 1) & 63) + 1; }

/* ---- the four span blitters: (cap, run) pairs for horizontal, mostly-horizontal, vertical, mostly-vertical lines ---- */
enum { HLINE, HORISH, VLINE, VERTISH };

/* tiny-skia keeps pixel coordinates unsigned, which shows at the top / left border (hairline_aa.rs): the stepped
 * ordinate is clamped at 0 after the half-pixel bias (`fy = fy.max(0)`), the upper / left pixel of a pair is addressed as
 * `lower.max(1) - 1` (so a pair whose first pixel would be off the canvas lands on pixels 0 and 1), and the horizontal
 * blitter skips its upper row when there is none (`y.checked_sub(1)`). */
static void emit_clipped(blit_list *b, const irect *clip, int x, int y, int a)
{
    if (clip && !(x >= clip->l && x < clip->r && y >= clip->t && y < clip->b)) return;
    emit(b, x, y, a);
}
static int32_t fix_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int max1m1(int v) { return (v > 1 ? v : 1) - 1; }

static int32_t draw_cap(blit_list *b, const irect *clip, int kind, int pos, int32_t f, int32_t slope, int mod64)
{
    f = fix_add(f, FIX_HALF);
    if (f < 0) f = 0;
    const int i = f >> 16;
    const int a = (f >> 8) & 0xFF;
    switch (kind) {
    case HLINE: /* pos = x: the lower row, then the row above */
        emit_clipped(b, clip, pos, i, small_dot6_scale(a, mod64));
        if (i - 1 >= 0) emit_clipped(b, clip, pos, i - 1, small_dot6_scale(255 - a, mod64));
        return f - FIX_HALF;
    case HORISH:
        emit_clipped(b, clip, pos, max1m1(i), small_dot6_scale(255 - a, mod64));
        emit_clipped(b, clip, pos, max1m1(i) + 1, small_dot6_scale(a, mod64));
        return fix_add(f, slope) - FIX_HALF;
    case VLINE: /* pos = y */
        emit_clipped(b, clip, i, pos, small_dot6_scale(a, mod64));
        emit_clipped(b, clip, max1m1(i), pos, small_dot6_scale(255 - a, mod64));
        return f - FIX_HALF;
    default:
        emit_clipped(b, clip, max1m1(i), pos, small_dot6_scale(255 - a, mod64));
        emit_clipped(b, clip, max1m1(i) + 1, pos, small_dot6_scale(a, mod64));
        return fix_add(f, slope) - FIX_HALF;
    }
}

static int32_t draw_line(blit_list *b, const irect *clip, int kind, int pos, int stop, int32_t f, int32_t slope)
{
    if (kind == HLINE) {
        f = fix_add(f, FIX_HALF);
        if (f < 0) f = 0;
        const int y = f >> 16;
        int a = (f >> 8) & 0xFF;
        if (a) for (int x = pos; x < stop; x++) emit_clipped(b, clip, x, y, a);
        a = 255 - a;
        if (a && y - 1 >= 0) for (int x = pos; x < stop; x++) emit_clipped(b, clip, x, y - 1, a);
        return f - FIX_HALF;
    }
    if (kind == VLINE) {
        f = fix_add(f, FIX_HALF);
        if (f < 0) f = 0;
        const int x = f >> 16;
        int a = (f >> 8) & 0xFF;
        if (a) for (int y = pos; y < stop; y++) emit_clipped(b, clip, x, y, a);
        a = 255 - a;
        if (a) for (int y = pos; y < stop; y++) emit_clipped(b, clip, max1m1(x), y, a);
        return f - FIX_HALF;
    }
    f = fix_add(f, FIX_HALF);
    if (kind == HORISH) {
        int x = pos;
        do {
            if (f < 0) f = 0;
            const int lower_y = f >> 16;
            const int a = (f >> 8) & 0xFF;
            emit_clipped(b, clip, x, max1m1(lower_y), 255 - a);
            emit_clipped(b, clip, x, max1m1(lower_y) + 1, a);
            f = fix_add(f, slope);
        } while (++x < stop);
        return f - FIX_HALF;
    }
    int y = pos;
    do {
        if (f < 0) f = 0;
        const int x = f >> 16;
        const int a = (f >> 8) & 0xFF;
        emit_clipped(b, clip, max1m1(x), y, 255 - a);
        emit_clipped(b, clip, max1m1(x) + 1, y, a);
        f = fix_add(f, slope);
    } while (++y < stop);
    return f - FIX_HALF;
}

static int iabs(int v) { return v < 0 ? -v : v; }

/* hairline_aa.rs do_anti_hairline: one line in FDot6, optionally against a clip rectangle */
static void do_anti_hairline(int32_t x0, int32_t y0, int32_t x1, int32_t y1, const irect *clip_in, blit_list *b)
{
    /* i32::MIN comes from converting an infinite or NaN float and cannot be negated: do not draw */
    if (x0 == INT32_MIN || y0 == INT32_MIN || x1 == INT32_MIN || y1 == INT32_MIN) return;
    if (iabs(x1 - x0) > (511 << 6) || iabs(y1 - y0) > (511 << 6)) {
        /* long lines are halved until the slope fits 16.16; each end shifted separately to avoid overflow */
        const int32_t hx = (x0 >> 1) + (x1 >> 1), hy = (y0 >> 1) + (y1 >> 1);
        do_anti_hairline(x0, y0, hx, hy, clip_in, b);
        do_anti_hairline(hx, hy, x1, y1, clip_in, b);
        return;
    }
    const irect *clip = clip_in;
    int scale_start, scale_stop, istart, istop, kind;
    int32_t fstart, slope;
    if (iabs(x1 - x0) > iabs(y1 - y0)) { /* mostly horizontal: walk left to right */
        if (x0 > x1) { int32_t t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
        istart = fdot6_floor(x0);
        istop = fdot6_ceil(x1);
        fstart = fdot6_to_fdot16(y0);
        if (y0 == y1) { slope = 0; kind = HLINE; }
        else {
            slope = fast_fix_div(y1 - y0, x1 - x0);
            fstart += (slope * (32 - (x0 & 63)) + 32) >> 6;
            kind = HORISH;
        }
        if (istop - istart == 1) { scale_start = x1 - x0; scale_stop = 0; } /* within a single pixel */
        else { scale_start = 64 - (x0 & 63); scale_stop = x1 & 63; }
        if (clip) {
            if (istart >= clip->r || istop <= clip->l) return;
            if (istart < clip->l) {
                fstart += slope * (clip->l - istart);
                istart = clip->l;
                scale_start = 64;
                if (istop - istart == 1) { scale_start = contribution_64(x1); scale_stop = 0; }
            }
            if (istop > clip->r) { istop = clip->r; scale_stop = 0; } /* the last column is not drawn */
            if (istart == istop) return;
            int top, bottom;
            if (slope >= 0) {
                top = fdot16_floor(fstart - FIX_HALF);
                bottom = fdot16_ceil(fstart + (istop - istart - 1) * slope + FIX_HALF);
            } else {
                bottom = fdot16_ceil(fstart + FIX_HALF);
                top = fdot16_floor(fstart + (istop - istart - 1) * slope - FIX_HALF);
            }
            top -= 1; bottom += 1; /* OUTSET_BEFORE_CLIP_TEST */
            if (top >= clip->b || bottom <= clip->t) return;
            if (clip->t <= top && clip->b >= bottom) clip = NULL;
        }
    } else { /* mostly vertical: walk top to bottom */
        if (y0 > y1) { int32_t t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
        istart = fdot6_floor(y0);
        istop = fdot6_ceil(y1);
        fstart = fdot6_to_fdot16(x0);
        if (x0 == x1) {
            if (y0 == y1) return; /* zero length */
            slope = 0;
            kind = VLINE;
        } else {
            slope = fast_fix_div(x1 - x0, y1 - y0);
            fstart += (slope * (32 - (y0 & 63)) + 32) >> 6;
            kind = VERTISH;
        }
        if (istop - istart == 1) { scale_start = y1 - y0; scale_stop = 0; }
        else { scale_start = 64 - (y0 & 63); scale_stop = y1 & 63; }
        if (clip) {
            if (istart >= clip->b || istop <= clip->t) return;
            if (istart < clip->t) {
                fstart += slope * (clip->t - istart);
                istart = clip->t;
                scale_start = 64;
                if (istop - istart == 1) { scale_start = contribution_64(y1); scale_stop = 0; }
            }
            if (istop > clip->b) { istop = clip->b; scale_stop = 0; }
            if (istart == istop) return;
            int left, right;
            if (slope >= 0) {
                left = fdot16_floor(fstart - FIX_HALF);
                right = fdot16_ceil(fstart + (istop - istart - 1) * slope + FIX_HALF);
            } else {
                right = fdot16_ceil(fstart + FIX_HALF);
                left = fdot16_floor(fstart + (istop - istart - 1) * slope - FIX_HALF);
            }
            left -= 1; right += 1;
            if (left >= clip->r || right <= clip->l) return;
            if (clip->l <= left && clip->r >= right) clip = NULL;
        }
    }
    fstart = draw_cap(b, clip, kind, istart, fstart, slope, scale_start);
    istart += 1;
    const int full_spans = istop - istart - (scale_stop > 0 ? 1 : 0);
    if (full_spans > 0) fstart = draw_line(b, clip, kind, istart, istart + full_spans, fstart, slope);
    if (scale_stop > 0) draw_cap(b, clip, kind, istop - 1, fstart, slope, scale_stop);
}

/* ---- line_clipper.rs intersect: the part of a segment inside a rectangle ---- */
typedef struct { float l, t, r, b; } frect;

static int nestedLT(float a, float b, float dim) { return a <= b && (a < b || dim > 0.0f); }
static int contains_no_empty_check(const frect *outer, const frect *inner)
{
    return outer->l <= inner->l && outer->t <= inner->t && outer->r >= inner->r && outer->b >= inner->b;
}
/* x where the segment crosses y (f64, pinned to the segment's x range) */
static float sect_with_horizontal(const pg_pt src[2], float y)
{
    const float dy = src[1].y - src[0].y;
    if (pg_nearly_zero(dy)) return (src[0].x + src[1].x) * 0.5f; /* average */
    const double x0 = src[0].x, y0 = src[0].y, x1 = src[1].x, y1 = src[1].y;
    double result = x0 + ((double)y - y0) * (x1 - x0) / (y1 - y0);
    /* the computed x can land outside the segment by rounding: pin it (in f64, unsorted limits) */
    const double lo = x0 < x1 ? x0 : x1, hi = x0 < x1 ? x1 : x0;
    if (result < lo) result = lo;
    if (result > hi) result = hi;
    return (float)result;
}
static float sect_with_vertical(const pg_pt src[2], float x)
{
    const float dx = src[1].x - src[0].x;
    if (pg_nearly_zero(dx)) return (src[0].y + src[1].y) * 0.5f;
    const double x0 = src[0].x, y0 = src[0].y, x1 = src[1].x, y1 = src[1].y;
    const double result = y0 + ((double)x - x0) * (y1 - y0) / (x1 - x0);
    return (float)result;
}

static int line_clip_intersect(const pg_pt src[2], const frect *clip, pg_pt dst[2])
{
    frect bounds;
    bounds.l = fminf(src[0].x, src[1].x); bounds.r = fmaxf(src[0].x, src[1].x);
    bounds.t = fminf(src[0].y, src[1].y); bounds.b = fmaxf(src[0].y, src[1].y);
    if (contains_no_empty_check(clip, &bounds)) { dst[0] = src[0]; dst[1] = src[1]; return 1; }
    /* a zero-width / zero-height line touching the clip edge is not rejected */
    if (nestedLT(bounds.r, clip->l, bounds.r - bounds.l) || nestedLT(clip->r, bounds.l, bounds.r - bounds.l)
        || nestedLT(bounds.b, clip->t, bounds.b - bounds.t) || nestedLT(clip->b, bounds.t, bounds.b - bounds.t))
        return 0;
    int index0, index1;
    if (src[0].y < src[1].y) { index0 = 0; index1 = 1; } else { index0 = 1; index1 = 0; }
    pg_pt tmp[2] = {src[0], src[1]};
    /* chop in y */
    if (tmp[index0].y < clip->t) tmp[index0] = pg_p(sect_with_horizontal(src, clip->t), clip->t);
    if (tmp[index1].y > clip->b) tmp[index1] = pg_p(sect_with_horizontal(src, clip->b), clip->b);
    if (tmp[0].x < tmp[1].x) { index0 = 0; index1 = 1; } else { index0 = 1; index1 = 0; }
    /* reject, then chop in x */
    if (tmp[index1].x <= clip->l || tmp[index0].x >= clip->r) {
        /* only reject a non-vertical line, or a vertical one strictly outside */
        if (tmp[0].x != tmp[1].x || tmp[0].x < clip->l || tmp[0].x > clip->r) return 0;
    }
    if (tmp[index0].x < clip->l) tmp[index0] = pg_p(clip->l, sect_with_vertical(src, clip->l));
    if (tmp[index1].x > clip->r) tmp[index1] = pg_p(clip->r, sect_with_vertical(src, clip->r));
    dst[0] = tmp[0];
    dst[1] = tmp[1];
    return 1;
}

/* hairline_aa.rs anti_hair_line_rgn: a polyline; clip = the pixmap when the path may reach beyond it */
static void anti_hair_line_rgn(const pg_pt *pts, int n, const irect *clip, blit_list *b)
{
    const float max = 32767.0f;
    const frect fixed_bounds = {-max, -max, max, max};
    frect clip_bounds = {0, 0, 0, 0};
    if (clip) {
        /* hairlines draw up to half a pixel outside their bounds: outset the scalar clip by one */
        clip_bounds.l = (float)clip->l - 1.0f; clip_bounds.t = (float)clip->t - 1.0f;
        clip_bounds.r = (float)clip->r + 1.0f; clip_bounds.b = (float)clip->b + 1.0f;
    }
    for (int i = 0; i + 1 < n; i++) {
        pg_pt seg[2] = {pts[i], pts[i + 1]}, p[2];
        if (!line_clip_intersect(seg, &fixed_bounds, p)) continue; /* must fit 16.16 */
        if (clip) {
            pg_pt q[2] = {p[0], p[1]};
            if (!line_clip_intersect(q, &clip_bounds, p)) continue;
        }
        const int32_t x0 = fdot6_from_f32(p[0].x), y0 = fdot6_from_f32(p[0].y), x1 = fdot6_from_f32(p[1].x), y1 = fdot6_from_f32(p[1].y);
        if (clip) {
            const int32_t left = x0 < x1 ? x0 : x1, top = y0 < y1 ? y0 : y1, right = x0 > x1 ? x0 : x1, bottom = y0 > y1 ? y0 : y1;
            irect ir = {fdot6_floor(left) - 1, fdot6_floor(top) - 1, fdot6_ceil(right) + 1, fdot6_ceil(bottom) + 1};
            if (ir.r <= ir.l || ir.b <= ir.t) continue;
            if (ir.l >= clip->r || ir.r <= clip->l || ir.t >= clip->b || ir.b <= clip->t) continue; /* quick reject */
            if (!(clip->l <= ir.l && clip->t <= ir.t && clip->r >= ir.r && clip->b >= ir.b)) {
                irect sub = {ir.l > clip->l ? ir.l : clip->l, ir.t > clip->t ? ir.t : clip->t, ir.r < clip->r ? ir.r : clip->r, ir.b < clip->b ? ir.b : clip->b};
                do_anti_hairline(x0, y0, x1, y1, &sub, b);
                continue;
            }
        }
        do_anti_hairline(x0, y0, x1, y1, NULL, b);
    }
}

/* ---- hairline.rs: quads and cubics become polylines ---- */
static int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static int32_t ceil_to_i32(float v)
{
    float c = ceilf(v);
    if (c != c) return 0;
    if (c >= 2147483648.0f) return INT32_MAX;
    if (c <= -2147483648.0f) return INT32_MIN;
    return (int32_t)c;
}

static void hair_quad(const pg_pt pts[3], const irect *clip, blit_list *b)
{
    /* distance of the control point from the chord's midpoint, in whole pixels (cheap norm) */
    const float dx = fabsf((pts[0].x + pts[2].x) * 0.5f - pts[1].x), dy = fabsf((pts[0].y + pts[2].y) * 0.5f - pts[1].y);
    const uint32_t idx = (uint32_t)ceil_to_i32(dx), idy = (uint32_t)ceil_to_i32(dy);
    const uint32_t d = idx > idy ? idx + (idy >> 1) : idy + (idx >> 1);
    /* each subdivision brings a quad 4x closer to its chord */
    int level = (33 - clz32(d)) >> 1;
    if (level > 5) level = 5; /* MAX_QUAD_SUBDIVIDE_LEVEL */
    const int lines = 1 << level;
    pg_pt tmp[(1 << 5) + 1];
    /* QuadCoeff, evaluated by forward stepping of t */
    const pg_pt bb = pg_sub(pts[1], pts[0]);
    const pg_pt A = pg_add(pg_sub(pts[2], pg_add(pts[1], pts[1])), pts[0]), B = pg_add(bb, bb), C = pts[0];
    const float dt = 1.0f / (float)lines;
    float t = 0.0f;
    tmp[0] = pts[0];
    for (int i = 1; i < lines; i++) {
        t += dt;
        tmp[i] = pg_p((A.x * t + B.x) * t + C.x, (A.y * t + B.y) * t + C.y);
    }
    tmp[lines] = pts[2];
    anti_hair_line_rgn(tmp, lines + 1, clip, b);
}

static int compute_cubic_segs(const pg_pt p[4])
{
    const float third = 1.0f / 3.0f, two_third = 2.0f / 3.0f;
    const pg_pt p13 = pg_p(third * p[3].x + two_third * p[0].x, third * p[3].y + two_third * p[0].y);
    const pg_pt p23 = pg_p(third * p[0].x + two_third * p[3].x, third * p[0].y + two_third * p[3].y);
    const float diff = fmaxf(fmaxf(fabsf(p[1].x - p13.x), fabsf(p[1].y - p13.y)), fmaxf(fabsf(p[2].x - p23.x), fabsf(p[2].y - p23.y)));
    float tol = 1.0f / 8.0f;
    for (int i = 0; i < 9; i++) { /* MAX_CUBIC_SUBDIVIDE_LEVEL */
        if (diff < tol) return 1 << i;
        tol *= 4.0f;
    }
    return 1 << 9;
}

static void hair_cubic_simple(const pg_pt pts[4], const irect *clip, blit_list *b)
{
    const int lines = compute_cubic_segs(pts);
    if (lines == 1) {
        const pg_pt tmp[2] = {pts[0], pts[3]};
        anti_hair_line_rgn(tmp, 2, clip, b);
        return;
    }
    const pg_pt A = pg_sub(pg_add(pts[3], pg_scale(pg_sub(pts[1], pts[2]), 3.0f)), pts[0]);
    const pg_pt B = pg_scale(pg_add(pg_sub(pts[2], pg_add(pts[1], pts[1])), pts[0]), 3.0f);
    const pg_pt C = pg_scale(pg_sub(pts[1], pts[0]), 3.0f), D = pts[0];
    const float dt = 1.0f / (float)lines;
    float t = 0.0f;
    pg_pt tmp[(1 << 9) + 1];
    tmp[0] = pts[0];
    int finite = 1;
    for (int i = 1; i < lines; i++) {
        t += dt;
        tmp[i] = pg_p(((A.x * t + B.x) * t + C.x) * t + D.x, ((A.y * t + B.y) * t + C.y) * t + D.y);
        finite = finite && pg_finite(tmp[i]);
    }
    if (finite) {
        tmp[lines] = pts[3];
        anti_hair_line_rgn(tmp, lines + 1, clip, b);
    }
}

/* max-curvature chop for cubics that turn sharply (path_geometry.rs chop_cubic_at_max_curvature) */
static void formulate(float s0, float s1, float s2, float s3, float c[4])
{
    const float a = s1 - s0, bq = s2 - 2.0f * s1 + s0, cq = s3 + 3.0f * (s1 - s2) - s0;
    c[0] = cq * cq; c[1] = 3.0f * bq * cq; c[2] = 2.0f * bq * bq + cq * a; c[3] = a * bq;
}
static float pin01f(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
static int solve_cubic(const float co[4], float t[3])
{
    if (pg_nearly_zero(co[0])) return pg_find_unit_quad_roots(co[1], co[2], co[3], t);
    const float inva = 1.0f / co[0];
    const float a = co[1] * inva, bq = co[2] * inva, c = co[3] * inva;
    const float q = (a * a - bq * 3.0f) / 9.0f, r = (2.0f * a * a * a - 9.0f * a * bq + 27.0f * c) / 54.0f;
    const float q3 = q * q * q, r2mq3 = r * r - q3, adiv3 = a / 3.0f;
    if (r2mq3 < 0.0f) {
        float cs = r / sqrtf(q3);
        cs = cs < -1.0f ? -1.0f : (cs > 1.0f ? 1.0f : cs);
        const float theta = acosf(cs), n2rq = -2.0f * sqrtf(q), pi = 3.14159265f;
        t[0] = pin01f(n2rq * cosf(theta / 3.0f) - adiv3);
        t[1] = pin01f(n2rq * cosf((theta + 2.0f * pi) / 3.0f) - adiv3);
        t[2] = pin01f(n2rq * cosf((theta - 2.0f * pi) / 3.0f) - adiv3);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2 - i; j++) if (t[j] > t[j + 1]) { float x = t[j]; t[j] = t[j + 1]; t[j + 1] = x; }
        int n = 3;
        if (t[1] == t[2]) n = 2;
        if (t[0] == t[1]) { t[1] = t[2]; n--; }
        return n;
    }
    float aa = cbrtf(fabsf(r) + sqrtf(r2mq3));
    if (r > 0.0f) aa = -aa;
    if (aa != 0.0f) aa += q / aa;
    t[0] = pin01f(aa - adiv3);
    return 1;
}

static int lt_90(pg_pt p0, pg_pt pivot, pg_pt p2) { return pg_dot(pg_sub(p0, pivot), pg_sub(p2, pivot)) >= 0.0f; }

static void hair_cubic(const pg_pt pts[4], const irect *clip_in, const frect *inset, const frect *outset, blit_list *b)
{
    const irect *clip = clip_in;
    if (inset) {
        frect bd;
        bd.l = fminf(fminf(pts[0].x, pts[1].x), fminf(pts[2].x, pts[3].x)); bd.r = fmaxf(fmaxf(pts[0].x, pts[1].x), fmaxf(pts[2].x, pts[3].x));
        bd.t = fminf(fminf(pts[0].y, pts[1].y), fminf(pts[2].y, pts[3].y)); bd.b = fmaxf(fmaxf(pts[0].y, pts[1].y), fmaxf(pts[2].y, pts[3].y));
        if (!(outset->l < bd.r && bd.l < outset->r && outset->t < bd.b && bd.t < outset->b)) return; /* geometric_overlap */
        if (bd.l >= inset->l && bd.r <= inset->r && bd.t >= inset->t && bd.b <= inset->b) clip = NULL; /* geometric_contains */
    }
    /* cubics whose control polygon turns by less than 90 degrees at both ends are subdivided uniformly */
    if (lt_90(pts[1], pts[0], pts[3]) && lt_90(pts[2], pts[0], pts[3]) && lt_90(pts[1], pts[3], pts[0]) && lt_90(pts[2], pts[3], pts[0])) {
        hair_cubic_simple(pts, clip, b);
        return;
    }
    float cx[4], cy[4], roots[3], tv[3];
    formulate(pts[0].x, pts[1].x, pts[2].x, pts[3].x, cx);
    formulate(pts[0].y, pts[1].y, pts[2].y, pts[3].y, cy);
    for (int i = 0; i < 4; i++) cx[i] += cy[i];
    const int rc = solve_cubic(cx, roots);
    int count = 0;
    for (int i = 0; i < rc; i++) if (0.0f < roots[i] && roots[i] < 1.0f) tv[count++] = roots[i];
    pg_pt dst[13];
    if (count == 0) memcpy(dst, pts, sizeof(pg_pt) * 4);
    else {
        /* chop_cubic_at with several t values: chop, renormalise the next t into the remainder, repeat */
        const pg_pt *src = pts;
        pg_pt tmp[4], *d = dst;
        float t = tv[0];
        for (int i = 0; i < count; i++) {
            pg_chop_cubic_at(src, t, d);
            if (i == count - 1) break;
            d += 3;
            memcpy(tmp, d, sizeof(pg_pt) * 4);
            src = tmp;
            if (!pg_valid_unit_divide(tv[i + 1] - tv[i], 1.0f - tv[i], &t)) {
                d[4] = d[5] = d[6] = src[3]; /* a degenerate remainder */
                break;
            }
        }
    }
    for (int i = 0; i <= count; i++) hair_cubic_simple(dst + i * 3, clip, b);
}

/* hairline.rs extend_pts: square and round caps lengthen the ends of a contour */
static void extend_pts(int cap, int prev_verb, int next_verb /* -1 = done */, pg_pt *pts, int n)
{
    const float cap_outset = cap == 2 ? 0.5f : 3.14159265f / 8.0f; /* round: half the area of a unit circle */
    if (prev_verb == PG_MOVE) {
        int first = 0, ctrl = 0, controls = n - 1;
        pg_pt tangent;
        do {
            ctrl++;
            tangent = pg_sub(pts[first], pts[ctrl]);
        } while (tangent.x == 0.0f && tangent.y == 0.0f && --controls > 0);
        if (tangent.x == 0.0f && tangent.y == 0.0f) { tangent = pg_p(1.0f, 0.0f); controls = n - 1; } /* all equal: move all but one */
        else pg_normalize(&tangent);
        do { /* an end point and the control points equal to it move together */
            pts[first].x += tangent.x * cap_outset;
            pts[first].y += tangent.y * cap_outset;
            first++;
        } while (++controls < n);
    }
    if (next_verb == PG_MOVE || next_verb == -1 || next_verb == PG_CLOSE) {
        int last = n - 1, ctrl = n - 1, controls = n - 1;
        pg_pt tangent;
        do {
            ctrl--;
            tangent = pg_sub(pts[last], pts[ctrl]);
        } while (tangent.x == 0.0f && tangent.y == 0.0f && --controls > 0);
        if (tangent.x == 0.0f && tangent.y == 0.0f) { tangent = pg_p(-1.0f, 0.0f); controls = n - 1; }
        else pg_normalize(&tangent);
        do {
            pts[last].x += tangent.x * cap_outset;
            pts[last].y += tangent.y * cap_outset;
            last--;
        } while (++controls < n);
    }
}

/* scan::hairline_aa::stroke_path over a path already in device space.  Returns the number of {x, y, alpha} blits
 * (malloc'ed int32 triples in *out_blits, free with orc_geom_free) in the walker's order. */
int32_t orc_path_hairline(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, int32_t cap, int32_t clip_w,
                          int32_t clip_h, int32_t **out_blits)
{
    *out_blits = NULL;
    if (n_verbs <= 0 || n_points <= 0 || clip_w <= 0 || clip_h <= 0) return 0;
    const pg_pt *P = (const pg_pt *)points;
    blit_list bl = {NULL, 0, 0, clip_w, clip_h};
    /* path bounds outset by the cap reach, rounded out, against the clip */
    float l = P[0].x, t = P[0].y, r = P[0].x, bt = P[0].y;
    for (int i = 1; i < n_points; i++) { l = fminf(l, P[i].x); r = fmaxf(r, P[i].x); t = fminf(t, P[i].y); bt = fmaxf(bt, P[i].y); }
    if (!(isfinite(l) && isfinite(t) && isfinite(r) && isfinite(bt))) return 0;
    const float cap_out = cap == 0 ? 1.0f : 2.0f;
    const double ibl = floor((double)(l - cap_out)), ibt = floor((double)(t - cap_out)), ibr = ceil((double)(r + cap_out)), ibb = ceil((double)(bt + cap_out));
    if (!(ibl < (double)clip_w && ibr > 0.0 && ibt < (double)clip_h && ibb > 0.0)) return 0; /* no intersection */
    const irect clip_rect = {0, 0, clip_w, clip_h};
    const irect *clip = NULL;
    frect inset_s, outset_s;
    const frect *inset = NULL, *outset = NULL;
    if (!(ibl >= 0.0 && ibt >= 0.0 && ibr <= (double)clip_w && ibb <= (double)clip_h)) {
        clip = &clip_rect;
        /* two scalar rects for culling cubics: hairlines may draw one pixel beyond their control points */
        outset_s.l = -1.0f; outset_s.t = -1.0f; outset_s.r = (float)clip_w + 1.0f; outset_s.b = (float)clip_h + 1.0f;
        /* tiny-skia: `clip.inset(1, 1)?` — an IntRect cannot be empty, so a clip 2 px wide or high ends the whole stroke */
        if (clip_w <= 2 || clip_h <= 2) return 0;
        inset_s.l = 1.0f; inset_s.t = 1.0f; inset_s.r = (float)clip_w - 1.0f; inset_s.b = (float)clip_h - 1.0f;
        inset = &inset_s;
        outset = &outset_s;
    }
    int prev_verb = -1, pi = 0;
    pg_pt first_pt = pg_p(0, 0), last_pt = pg_p(0, 0);
    for (int vi = 0; vi < n_verbs; vi++) {
        const int verb = verbs[vi];
        const int next_verb = vi + 1 < n_verbs ? verbs[vi + 1] : -1;
        pg_pt pts[4];
        switch (verb) {
        case PG_MOVE:
            first_pt = last_pt = P[pi++];
            pts[0] = first_pt;
            break;
        case PG_LINE:
            pts[0] = P[pi - 1]; pts[1] = P[pi]; pi += 1;
            if (cap != 0) extend_pts(cap, prev_verb, next_verb, pts, 2);
            anti_hair_line_rgn(pts, 2, clip, &bl);
            last_pt = pts[1];
            break;
        case PG_QUAD:
            pts[0] = P[pi - 1]; pts[1] = P[pi]; pts[2] = P[pi + 1]; pi += 2;
            if (cap != 0) extend_pts(cap, prev_verb, next_verb, pts, 3);
            hair_quad(pts, clip, &bl);
            last_pt = pts[2];
            break;
        case PG_CUBIC:
            pts[0] = P[pi - 1]; pts[1] = P[pi]; pts[2] = P[pi + 1]; pts[3] = P[pi + 2]; pi += 3;
            if (cap != 0) extend_pts(cap, prev_verb, next_verb, pts, 4);
            hair_cubic(pts, clip, inset, outset, &bl);
            last_pt = pts[3];
            break;
        default: /* close */
            pts[0] = last_pt; pts[1] = first_pt;
            if (cap != 0 && prev_verb == PG_MOVE) extend_pts(cap, prev_verb, next_verb, pts, 2); /* move + close: a capped dot */
            anti_hair_line_rgn(pts, 2, clip, &bl);
        }
        if (cap != 0) {
            if (prev_verb == PG_MOVE && verb >= PG_LINE && verb <= PG_CUBIC) first_pt = pts[0]; /* the cap moved the start: close to it */
            prev_verb = verb;
        }
    }
    *out_blits = bl.v;
    return bl.n;
}
