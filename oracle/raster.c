/*
 * raster.c — CPU oracle for the tiny-skia 0.12.0 fill path (TEST INFRASTRUCTURE ONLY; see raster.h).
 *
 * Sections: fixed point · transform · path geometry · edge clipper · edges · edge builder ·
 * scan walkers (non-AA, 4x4 supersampled AA) · raster pipeline (lowp u16 / highp f32) · shaders ·
 * painter entry points · masks.  Section headers name the tiny-skia module / Skia file restated.
 * Compile with -ffp-contract=off.
 */
#include "raster.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y; } pt;

/* ---- Rust cast semantics (saturating, NaN -> 0) ---- */
static inline int32_t f2i(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int32_t)v;
}
static inline int32_t d2i_sat(double v)
{
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int32_t)v;
}
static inline uint8_t f2u8(float v) { return !(v > 0.0f) ? 0 : (v >= 255.0f ? 255 : (uint8_t)v); }
static inline uint16_t f2u16(float v) { return !(v > 0.0f) ? 0 : (v >= 65535.0f ? 65535 : (uint16_t)v); }

#define SCALAR_NEARLY_ZERO (1.0f / 4096.0f)
static inline int nearly_zero(float v) { return fabsf(v) <= SCALAR_NEARLY_ZERO; }
static inline int nearly_zero_tol(float v, float tol) { return fabsf(v) <= tol; }

/* ==========================================================================================
 * fixed point — tiny-skia fixed_point.rs (Skia SkFDot6.h / SkFixed.h)
 * ======================================================================================== */
typedef int32_t fdot6;
typedef int32_t fdot16;

static inline int32_t lsh(int32_t v, int s) { return (int32_t)((uint32_t)v << s); }
static inline fdot6 fdot6_round(fdot6 n) { return (n + 32) >> 6; }
static inline fdot16 fdot6_to_fdot16(fdot6 n) { return lsh(n, 10); }
static inline fdot16 fdot16_mul(fdot16 a, fdot16 b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 16); }
static inline fdot16 fdot16_div(fdot6 a, fdot6 b)
{
    int64_t v = ((int64_t)a * 65536) / (int64_t)b;
    if (v > 2147483647LL) v = 2147483647LL;
    if (v < -2147483648LL) v = -2147483648LL;
    return (int32_t)v;
}
static inline fdot16 fdot6_div(fdot6 a, fdot6 b)
{
    if (a >= -32768 && a <= 32767) return lsh(a, 16) / b;
    return fdot16_div(a, b);
}
static inline int32_t fdot16_round_to_i32(fdot16 x) { return (int32_t)((uint32_t)x + 0x8000u) >> 16; }

/* ==========================================================================================
 * transform — tiny-skia-path transform.rs
 * ======================================================================================== */
typedef struct { float sx, ky, kx, sy, tx, ty; } xform;

static xform ts_from(const float t[6]) { xform r = {t[0], t[1], t[2], t[3], t[4], t[5]}; return r; }
static xform ts_identity(void) { xform r = {1, 0, 0, 1, 0, 0}; return r; }
static int ts_is_identity(xform t) { return t.sx == 1 && t.ky == 0 && t.kx == 0 && t.sy == 1 && t.tx == 0 && t.ty == 0; }
static int ts_has_skew(xform t) { return t.kx != 0 || t.ky != 0; }
static int ts_has_scale(xform t) { return t.sx != 1 || t.sy != 1; }
static int ts_is_translate(xform t) { return !ts_has_scale(t) && !ts_has_skew(t) && (t.tx != 0 || t.ty != 0); }
static int ts_is_finite(xform t)
{
    return isfinite(t.sx) && isfinite(t.ky) && isfinite(t.kx) && isfinite(t.sy) && isfinite(t.tx) && isfinite(t.ty);
}
static float mul_add_mul(float a, float b, float c, float d) { return (float)((double)a * (double)b + (double)c * (double)d); }
/* concat(a, b): b applied first */
static xform ts_concat(xform a, xform b)
{
    if (ts_is_identity(a)) return b;
    if (ts_is_identity(b)) return a;
    xform r;
    if (!ts_has_skew(a) && !ts_has_skew(b)) {
        r.sx = a.sx * b.sx; r.ky = 0; r.kx = 0; r.sy = a.sy * b.sy;
        r.tx = a.sx * b.tx + a.tx;
        r.ty = a.sy * b.ty + a.ty;
    } else {
        r.sx = mul_add_mul(a.sx, b.sx, a.kx, b.ky);
        r.ky = mul_add_mul(a.ky, b.sx, a.sy, b.ky);
        r.kx = mul_add_mul(a.sx, b.kx, a.kx, b.sy);
        r.sy = mul_add_mul(a.ky, b.kx, a.sy, b.sy);
        r.tx = mul_add_mul(a.sx, b.tx, a.kx, b.ty) + a.tx;
        r.ty = mul_add_mul(a.ky, b.tx, a.sy, b.ty) + a.ty;
    }
    return r;
}
static xform ts_pre_concat(xform self, xform other) { return ts_concat(self, other); }
static xform ts_post_concat(xform self, xform other) { return ts_concat(other, self); }
static xform ts_translate(float x, float y) { xform r = {1, 0, 0, 1, x, y}; return r; }
static xform ts_scale(float x, float y) { xform r = {x, 0, 0, y, 0, 0}; return r; }
static int ts_invert(xform t, xform *out)
{
    if (ts_is_identity(t)) { *out = t; return 1; }
    if (!ts_has_skew(t)) {
        if (ts_has_scale(t)) {
            float ix = 1.0f / t.sx, iy = 1.0f / t.sy;
            xform r = {ix, 0, 0, iy, -t.tx * ix, -t.ty * iy};
            *out = r;
        } else {
            *out = ts_translate(-t.tx, -t.ty);
        }
        return 1;
    }
    double det = (double)t.sx * (double)t.sy - (double)t.kx * (double)t.ky;
    float tol = SCALAR_NEARLY_ZERO * SCALAR_NEARLY_ZERO * SCALAR_NEARLY_ZERO;
    if (nearly_zero_tol((float)det, tol)) return 0;
    double inv = 1.0 / det;
    xform r;
    r.sx = (float)((double)t.sy * inv);
    r.ky = (float)((double)(-t.ky) * inv);
    r.kx = (float)((double)(-t.kx) * inv);
    r.sy = (float)((double)t.sx * inv);
    r.tx = (float)(((double)t.kx * (double)t.ty - (double)t.sy * (double)t.tx) * inv);
    r.ty = (float)(((double)t.ky * (double)t.tx - (double)t.sx * (double)t.ty) * inv);
    if (!ts_is_finite(r)) return 0;
    *out = r;
    return 1;
}
static void ts_map_points(xform t, pt *p, int n)
{
    if (ts_is_identity(t)) return;
    if (ts_is_translate(t)) {
        for (int i = 0; i < n; i++) { p[i].x += t.tx; p[i].y += t.ty; }
    } else if (!ts_has_skew(t)) {
        for (int i = 0; i < n; i++) { p[i].x = p[i].x * t.sx + t.tx; p[i].y = p[i].y * t.sy + t.ty; }
    } else {
        for (int i = 0; i < n; i++) {
            float x = p[i].x * t.sx + p[i].y * t.kx + t.tx;
            float y = p[i].x * t.ky + p[i].y * t.sy + t.ty;
            p[i].x = x; p[i].y = y;
        }
    }
}

/* ==========================================================================================
 * path geometry — tiny-skia-path path_geometry.rs (Skia SkGeometry.cpp)
 * ======================================================================================== */
static inline float interp(float a, float b, float t) { return a + (b - a) * t; }
static inline pt pinterp(pt a, pt b, float t) { pt r = {interp(a.x, b.x, t), interp(a.y, b.y, t)}; return r; }

/* returns 1 and *ratio in (0,1) or 0 */
static int valid_unit_divide(float numer, float denom, float *ratio)
{
    if (numer < 0) { numer = -numer; denom = -denom; }
    if (denom == 0 || numer == 0 || numer >= denom) return 0;
    float r = numer / denom;
    if (r != r) return 0;
    if (r == 0) return 0;
    if (!(r > 0.0f && r < 1.0f)) return 0;
    *ratio = r;
    return 1;
}

static int find_unit_quad_roots(float a, float b, float c, float roots[2])
{
    if (a == 0) return valid_unit_divide(-c, b, roots);
    double dr = (double)b * (double)b - 4.0 * (double)a * (double)c;
    if (dr < 0) return 0;
    dr = sqrt(dr);
    float r = (float)dr;
    if (!isfinite(r)) return 0;
    float q = (b < 0) ? -(b - r) / 2 : -(b + r) / 2;
    int n = 0;
    n += valid_unit_divide(q, a, roots + n);
    n += valid_unit_divide(c, q, roots + n);
    if (n == 2) {
        if (roots[0] > roots[1]) { float t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
        else if (roots[0] == roots[1]) n = 1;
    }
    return n;
}

static void chop_quad_at(const pt src[3], float t, pt dst[5])
{
    pt p01 = pinterp(src[0], src[1], t), p12 = pinterp(src[1], src[2], t);
    dst[0] = src[0]; dst[1] = p01; dst[2] = pinterp(p01, p12, t); dst[3] = p12; dst[4] = src[2];
}

static void chop_cubic_at2(const pt src[4], float t, pt dst[7])
{
    pt ab = pinterp(src[0], src[1], t), bc = pinterp(src[1], src[2], t), cd = pinterp(src[2], src[3], t);
    pt abc = pinterp(ab, bc, t), bcd = pinterp(bc, cd, t), abcd = pinterp(abc, bcd, t);
    dst[0] = src[0]; dst[1] = ab; dst[2] = abc; dst[3] = abcd; dst[4] = bcd; dst[5] = cd; dst[6] = src[3];
}

static int is_not_monotonic(float a, float b, float c)
{
    float ab = a - b, bc = b - c;
    if (ab < 0) bc = -bc;
    return ab == 0 || bc < 0;
}

/* axis: 0 = x, 1 = y */
#define AX(p, axis) ((axis) ? (p).y : (p).x)
static inline void set_ax(pt *p, int axis, float v) { if (axis) p->y = v; else p->x = v; }

static int chop_quad_at_extrema(const pt src[3], pt dst[5], int axis)
{
    float a = AX(src[0], axis), b = AX(src[1], axis), c = AX(src[2], axis);
    if (is_not_monotonic(a, b, c)) {
        float t;
        if (valid_unit_divide(a - b, a - b - b + c, &t)) {
            chop_quad_at(src, t, dst);
            /* flatten_double_quad_extrema */
            set_ax(&dst[1], axis, AX(dst[2], axis));
            set_ax(&dst[3], axis, AX(dst[2], axis));
            return 1;
        }
        b = fabsf(a - b) < fabsf(b - c) ? a : c;
    }
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    set_ax(&dst[0], axis, a); set_ax(&dst[1], axis, b); set_ax(&dst[2], axis, c);
    return 0;
}

static int find_cubic_extrema(float a, float b, float c, float d, float t[2])
{
    float A = d - a + 3 * (b - c);
    float B = 2 * (a - b - b + c);
    float C = b - a;
    return find_unit_quad_roots(A, B, C, t);
}

static void chop_cubic_at(const pt src_in[4], const float *tv, int roots, pt *dst)
{
    if (roots == 0) { memcpy(dst, src_in, 4 * sizeof(pt)); return; }
    pt src[4];
    memcpy(src, src_in, sizeof(src));
    float t = tv[0];
    for (int i = 0; i < roots; i++) {
        chop_cubic_at2(src, t, dst);
        if (i == roots - 1) break;
        dst += 3;
        memcpy(src, dst, 4 * sizeof(pt));
        if (!valid_unit_divide(tv[i + 1] - tv[i], 1.0f - tv[i], &t)) {
            dst[4] = dst[5] = dst[6] = src[3];
            break;
        }
    }
}

static int chop_cubic_at_extrema(const pt src[4], pt dst[10], int axis)
{
    float tv[2];
    int roots = find_cubic_extrema(AX(src[0], axis), AX(src[1], axis), AX(src[2], axis), AX(src[3], axis), tv);
    chop_cubic_at(src, tv, roots, dst);
    if (roots > 0) {
        set_ax(&dst[2], axis, AX(dst[3], axis));
        set_ax(&dst[4], axis, AX(dst[3], axis));
        if (roots == 2) {
            set_ax(&dst[5], axis, AX(dst[6], axis));
            set_ax(&dst[7], axis, AX(dst[6], axis));
        }
    }
    return roots;
}

/* ==========================================================================================
 * edges — tiny-skia edge.rs (Skia SkEdge.cpp)
 * ======================================================================================== */
typedef struct {
    int32_t prev, next;
    fdot16 x, dx;
    int32_t first_y, last_y;
    int8_t winding;
    int8_t kind; /* 0 line, 1 quad, 2 cubic */
    int8_t curve_count;
    uint8_t curve_shift, cubic_dshift;
    /* quad: qx,qy,qdx,qdy,qddx,qddy,qlastx,qlasty; cubic: cx,cy,cdx,cdy,cddx,cddy,cdddx,cdddy,clastx,clasty */
    fdot16 c[10];
} edge_t;

static int line_set(edge_t *e, fdot6 x0, fdot6 y0, fdot6 x1, fdot6 y1)
{
    fdot6 top = fdot6_round(y0), bottom = fdot6_round(y1);
    if (top == bottom) return 0;
    fdot16 slope = fdot6_div(x1 - x0, y1 - y0);
    fdot6 dy = lsh(top, 6) + 32 - y0;
    e->x = fdot6_to_fdot16(x0 + fdot16_mul(slope, dy));
    e->dx = slope;
    e->first_y = top;
    e->last_y = bottom - 1;
    return 1;
}

static int line_edge_new(edge_t *e, pt p0, pt p1, int shift)
{
    float scale = (float)(1 << (shift + 6));
    fdot6 x0 = f2i(p0.x * scale), y0 = f2i(p0.y * scale), x1 = f2i(p1.x * scale), y1 = f2i(p1.y * scale);
    int8_t winding = 1;
    if (y0 > y1) {
        fdot6 t = x0; x0 = x1; x1 = t;
        t = y0; y0 = y1; y1 = t;
        winding = -1;
    }
    memset(e, 0, sizeof(*e));
    e->kind = 0;
    e->winding = winding;
    return line_set(e, x0, y0, x1, y1);
}

/* LineEdge::update: inputs are FDot16 */
static int line_update(edge_t *e, fdot16 x0, fdot16 y0, fdot16 x1, fdot16 y1)
{
    return line_set(e, x0 >> 10, y0 >> 10, x1 >> 10, y1 >> 10);
}

static inline fdot6 cheap_distance(fdot6 dx, fdot6 dy)
{
    dx = dx < 0 ? -dx : dx;
    dy = dy < 0 ? -dy : dy;
    return dx > dy ? dx + (dy >> 1) : dy + (dx >> 1);
}
static inline int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int diff_to_shift(fdot6 dx, fdot6 dy, int shift_aa)
{
    fdot6 dist = cheap_distance(dx, dy);
    dist = (dist + (1 << 4)) >> (3 + shift_aa);
    return (32 - clz32((uint32_t)dist)) / 2;
}

static int quad_update(edge_t *e)
{
    int success;
    int count = e->curve_count;
    fdot16 oldx = e->c[0], oldy = e->c[1], dx = e->c[2], dy = e->c[3], newx, newy;
    int shift = e->curve_shift;
    do {
        if (--count > 0) {
            newx = oldx + (dx >> shift);
            dx += e->c[4];
            newy = oldy + (dy >> shift);
            dy += e->c[5];
        } else {
            newx = e->c[6];
            newy = e->c[7];
        }
        success = line_update(e, oldx, oldy, newx, newy);
        oldx = newx;
        oldy = newy;
    } while (count > 0 && !success);
    e->c[0] = newx; e->c[1] = newy; e->c[2] = dx; e->c[3] = dy;
    e->curve_count = (int8_t)count;
    return success;
}

#define MAX_COEFF_SHIFT 6

static int quad_edge_new(edge_t *e, const pt p[3], int shift)
{
    float scale = (float)(1 << (shift + 6));
    fdot6 x0 = f2i(p[0].x * scale), y0 = f2i(p[0].y * scale);
    fdot6 x1 = f2i(p[1].x * scale), y1 = f2i(p[1].y * scale);
    fdot6 x2 = f2i(p[2].x * scale), y2 = f2i(p[2].y * scale);
    int8_t winding = 1;
    if (y0 > y2) {
        fdot6 t = x0; x0 = x2; x2 = t;
        t = y0; y0 = y2; y2 = t;
        winding = -1;
    }
    fdot6 top = fdot6_round(y0), bottom = fdot6_round(y2);
    if (top == bottom) return 0;
    {
        fdot6 dx = (lsh(x1, 1) - x0 - x2) >> 2;
        fdot6 dy = (lsh(y1, 1) - y0 - y2) >> 2;
        shift = diff_to_shift(dx, dy, shift);
    }
    if (shift == 0) shift = 1;
    else if (shift > MAX_COEFF_SHIFT) shift = MAX_COEFF_SHIFT;
    memset(e, 0, sizeof(*e));
    e->kind = 1;
    e->winding = winding;
    e->curve_count = (int8_t)(1 << shift);
    e->curve_shift = (uint8_t)(shift - 1);
    fdot16 a = lsh(x0 - x1 - x1 + x2, 9);
    fdot16 b = fdot6_to_fdot16(x1 - x0);
    e->c[0] = fdot6_to_fdot16(x0);
    e->c[2] = b + (a >> shift);
    e->c[4] = a >> (shift - 1);
    a = lsh(y0 - y1 - y1 + y2, 9);
    b = fdot6_to_fdot16(y1 - y0);
    e->c[1] = fdot6_to_fdot16(y0);
    e->c[3] = b + (a >> shift);
    e->c[5] = a >> (shift - 1);
    e->c[6] = fdot6_to_fdot16(x2);
    e->c[7] = fdot6_to_fdot16(y2);
    return quad_update(e);
}

static int cubic_update(edge_t *e)
{
    int success;
    int count = e->curve_count;
    fdot16 oldx = e->c[0], oldy = e->c[1], newx, newy;
    int dd_shift = e->curve_shift, d_shift = e->cubic_dshift;
    do {
        if (++count < 0) {
            newx = oldx + (e->c[2] >> d_shift);
            e->c[2] += e->c[4] >> dd_shift;
            e->c[4] += e->c[6];
            newy = oldy + (e->c[3] >> d_shift);
            e->c[3] += e->c[5] >> dd_shift;
            e->c[5] += e->c[7];
        } else {
            newx = e->c[8];
            newy = e->c[9];
        }
        if (newy < oldy) newy = oldy;
        success = line_update(e, oldx, oldy, newx, newy);
        oldx = newx;
        oldy = newy;
    } while (count < 0 && !success);
    e->c[0] = newx; e->c[1] = newy;
    e->curve_count = (int8_t)count;
    return success;
}

static inline fdot6 cubic_delta_from_line(fdot6 a, fdot6 b, fdot6 c, fdot6 d)
{
    fdot6 one_third = ((a * 8 - b * 15 + 6 * c + d) * 19) >> 9;
    fdot6 two_third = ((a + 6 * b - c * 15 + d * 8) * 19) >> 9;
    one_third = one_third < 0 ? -one_third : one_third;
    two_third = two_third < 0 ? -two_third : two_third;
    return one_third > two_third ? one_third : two_third;
}

static int cubic_edge_new(edge_t *e, const pt p[4], int shift_aa)
{
    float scale = (float)(1 << (shift_aa + 6));
    fdot6 x0 = f2i(p[0].x * scale), y0 = f2i(p[0].y * scale), x1 = f2i(p[1].x * scale), y1 = f2i(p[1].y * scale);
    fdot6 x2 = f2i(p[2].x * scale), y2 = f2i(p[2].y * scale), x3 = f2i(p[3].x * scale), y3 = f2i(p[3].y * scale);
    int8_t winding = 1;
    if (y0 > y3) {
        fdot6 t;
        t = x0; x0 = x3; x3 = t;  t = x1; x1 = x2; x2 = t;
        t = y0; y0 = y3; y3 = t;  t = y1; y1 = y2; y2 = t;
        winding = -1;
    }
    fdot6 top = fdot6_round(y0), bot = fdot6_round(y3);
    if (top == bot) return 0;
    fdot6 dx = cubic_delta_from_line(x0, x1, x2, x3);
    fdot6 dy = cubic_delta_from_line(y0, y1, y2, y3);
    int shift = diff_to_shift(dx, dy, 2) + 1;
    if (shift > MAX_COEFF_SHIFT) shift = MAX_COEFF_SHIFT;
    int up_shift = 6;
    int down_shift = shift + up_shift - 10;
    if (down_shift < 0) { down_shift = 0; up_shift = 10 - shift; }
    memset(e, 0, sizeof(*e));
    e->kind = 2;
    e->winding = winding;
    e->curve_count = (int8_t)lsh(-1, shift);
    e->curve_shift = (uint8_t)shift;
    e->cubic_dshift = (uint8_t)down_shift;
    fdot16 b = lsh(3 * (x1 - x0), up_shift);
    fdot16 c = lsh(3 * (x0 - x1 - x1 + x2), up_shift);
    fdot16 d = lsh(x3 + 3 * (x1 - x2) - x0, up_shift);
    e->c[0] = fdot6_to_fdot16(x0);
    e->c[2] = b + (c >> shift) + (d >> (2 * shift));
    e->c[4] = 2 * c + ((3 * d) >> (shift - 1));
    e->c[6] = (3 * d) >> (shift - 1);
    b = lsh(3 * (y1 - y0), up_shift);
    c = lsh(3 * (y0 - y1 - y1 + y2), up_shift);
    d = lsh(y3 + 3 * (y1 - y2) - y0, up_shift);
    e->c[1] = fdot6_to_fdot16(y0);
    e->c[3] = b + (c >> shift) + (d >> (2 * shift));
    e->c[5] = 2 * c + ((3 * d) >> (shift - 1));
    e->c[7] = (3 * d) >> (shift - 1);
    e->c[8] = fdot6_to_fdot16(x3);
    e->c[9] = fdot6_to_fdot16(y3);
    return cubic_update(e);
}

/* ==========================================================================================
 * edge builder — tiny-skia edge_builder.rs (Skia SkEdgeBuilder.cpp)
 * ======================================================================================== */
typedef struct {
    edge_t *e;
    int n, cap;
    int shift;
} builder_t;

static edge_t *builder_push(builder_t *b)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 64;
        b->e = (edge_t *)realloc(b->e, sizeof(edge_t) * (size_t)b->cap);
    }
    return &b->e[b->n++];
}

/* 0 = No, 1 = Partial, 2 = Total */
static int combine_vertical(const edge_t *edge, edge_t *last)
{
    if (last->dx != 0 || edge->x != last->x) return 0;
    if (edge->winding == last->winding) {
        if (edge->last_y + 1 == last->first_y) { last->first_y = edge->first_y; return 1; }
        if (edge->first_y == last->last_y + 1) { last->last_y = edge->last_y; return 1; }
        return 0;
    }
    if (edge->first_y == last->first_y) {
        if (edge->last_y == last->last_y) return 2;
        if (edge->last_y < last->last_y) { last->first_y = edge->last_y + 1; return 1; }
        last->first_y = last->last_y + 1;
        last->last_y = edge->last_y;
        last->winding = edge->winding;
        return 1;
    }
    if (edge->last_y == last->last_y) {
        if (edge->first_y > last->first_y) { last->last_y = edge->first_y - 1; }
        else {
            last->last_y = last->first_y - 1;
            last->first_y = edge->first_y;
            last->winding = edge->winding;
        }
        return 1;
    }
    return 0;
}

static void push_line(builder_t *b, pt p0, pt p1)
{
    edge_t e;
    if (!line_edge_new(&e, p0, p1, b->shift)) return;
    int combine = 0;
    if (e.dx == 0 && b->n > 0 && b->e[b->n - 1].kind == 0) combine = combine_vertical(&e, &b->e[b->n - 1]);
    if (combine == 2) b->n--;
    else if (combine == 0) *builder_push(b) = e;
}
static void push_quad(builder_t *b, const pt p[3])
{
    edge_t e;
    if (quad_edge_new(&e, p, b->shift)) *builder_push(b) = e;
}
static void push_cubic(builder_t *b, const pt p[4])
{
    edge_t e;
    if (cubic_edge_new(&e, p, b->shift)) *builder_push(b) = e;
}

/* ==========================================================================================
 * edge clipper — tiny-skia edge_clipper.rs / line_clipper.rs (Skia SkEdgeClipper.cpp, SkLineClipper.cpp)
 * Emits clipped lines / quads / cubics straight into the builder.
 * ======================================================================================== */
typedef struct { float l, t, r, b; } rectf;

static float pin_unsorted_d(double v, double a, double b)
{
    if (a > b) { double t = a; a = b; b = t; }
    if (v < a) v = a; else if (v > b) v = b;
    return (float)v;
}
static float pin_unsorted_f(float v, float a, float b)
{
    if (a > b) { float t = a; a = b; b = t; }
    if (v < a) v = a; else if (v > b) v = b;
    return v;
}
static float sect_with_horizontal(const pt s[2], float y)
{
    float dy = s[1].y - s[0].y;
    if (nearly_zero(dy)) return (s[0].x + s[1].x) * 0.5f;
    double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
    double r = x0 + ((double)y - y0) * (x1 - x0) / (y1 - y0);
    return pin_unsorted_d(r, x0, x1);
}
static float sect_clamp_with_vertical(const pt s[2], float x)
{
    float dx = s[1].x - s[0].x;
    float y;
    if (nearly_zero(dx)) y = (s[0].y + s[1].y) * 0.5f;
    else {
        double x0 = s[0].x, y0 = s[0].y, x1 = s[1].x, y1 = s[1].y;
        y = (float)(y0 + ((double)x - x0) * (y1 - y0) / (x1 - x0));
    }
    return pin_unsorted_f(y, s[0].y, s[1].y);
}

/* SkLineClipper::ClipLine; returns number of line segments (points = n+1) */
static int clip_line_pts(const pt pts[2], rectf clip, pt lines[4], int cull_right)
{
    int i0, i1;
    if (pts[0].y < pts[1].y) { i0 = 0; i1 = 1; } else { i0 = 1; i1 = 0; }
    if (pts[i1].y <= clip.t) return 0;
    if (pts[i0].y >= clip.b) return 0;
    pt tmp[2] = {pts[0], pts[1]};
    if (pts[i0].y < clip.t) { tmp[i0].x = sect_with_horizontal(pts, clip.t); tmp[i0].y = clip.t; }
    if (tmp[i1].y > clip.b) { tmp[i1].x = sect_with_horizontal(pts, clip.b); tmp[i1].y = clip.b; }
    pt storage[4];
    pt *result;
    int line_count = 1;
    int reverse;
    if (pts[0].x < pts[1].x) { i0 = 0; i1 = 1; reverse = 0; } else { i0 = 1; i1 = 0; reverse = 1; }
    if (tmp[i1].x <= clip.l) {
        tmp[0].x = tmp[1].x = clip.l;
        result = tmp; reverse = 0;
    } else if (tmp[i0].x >= clip.r) {
        if (cull_right) return 0;
        tmp[0].x = tmp[1].x = clip.r;
        result = tmp; reverse = 0;
    } else {
        result = storage;
        pt *r = result;
        if (tmp[i0].x < clip.l) {
            r->x = clip.l; r->y = tmp[i0].y; r++;
            r->x = clip.l; r->y = sect_clamp_with_vertical(tmp, clip.l);
        } else *r = tmp[i0];
        r++;
        if (tmp[i1].x > clip.r) {
            r->x = clip.r; r->y = sect_clamp_with_vertical(tmp, clip.r); r++;
            r->x = clip.r; r->y = tmp[i1].y;
        } else *r = tmp[i1];
        line_count = (int)(r - result);
    }
    if (reverse) for (int i = 0; i <= line_count; i++) lines[line_count - i] = result[i];
    else memcpy(lines, result, sizeof(pt) * (size_t)(line_count + 1));
    return line_count;
}

typedef struct { builder_t *b; rectf clip; int cull_right; } clipper_t;

static void clip_line(clipper_t *c, pt p0, pt p1)
{
    pt pts[2] = {p0, p1}, lines[4];
    int n = clip_line_pts(pts, c->clip, lines, c->cull_right);
    for (int i = 0; i < n; i++) push_line(c->b, lines[i], lines[i + 1]);
}
static void append_vline(clipper_t *c, float x, float y0, float y1, int reverse)
{
    if (reverse) { float t = y0; y0 = y1; y1 = t; }
    pt a = {x, y0}, b = {x, y1};
    push_line(c->b, a, b);
}
static void append_quad(clipper_t *c, const pt p[3], int reverse)
{
    if (reverse) { pt r[3] = {p[2], p[1], p[0]}; push_quad(c->b, r); }
    else push_quad(c->b, p);
}
static void append_cubic(clipper_t *c, const pt p[4], int reverse)
{
    if (reverse) { pt r[4] = {p[3], p[2], p[1], p[0]}; push_cubic(c->b, r); }
    else push_cubic(c->b, p);
}

static int chop_mono_quad_at(float c0, float c1, float c2, float target, float *t)
{
    float A = c0 - c1 - c1 + c2, B = 2 * (c1 - c0), C = c0 - target;
    float roots[2];
    if (find_unit_quad_roots(A, B, C, roots)) { *t = roots[0]; return 1; }
    return 0;
}

static void chop_quad_in_y(rectf clip, pt p[3])
{
    float t;
    pt tmp[5];
    if (p[0].y < clip.t) {
        if (chop_mono_quad_at(p[0].y, p[1].y, p[2].y, clip.t, &t)) {
            chop_quad_at(p, t, tmp);
            tmp[2].y = clip.t;
            if (tmp[3].y < clip.t) tmp[3].y = clip.t;
            p[0] = tmp[2]; p[1] = tmp[3];
        } else {
            for (int i = 0; i < 3; i++) if (p[i].y < clip.t) p[i].y = clip.t;
        }
    }
    if (p[2].y > clip.b) {
        if (chop_mono_quad_at(p[0].y, p[1].y, p[2].y, clip.b, &t)) {
            chop_quad_at(p, t, tmp);
            if (tmp[1].y > clip.b) tmp[1].y = clip.b;
            tmp[2].y = clip.b;
            p[1] = tmp[1]; p[2] = tmp[2];
        } else {
            for (int i = 0; i < 3; i++) if (p[i].y > clip.b) p[i].y = clip.b;
        }
    }
}

static void clip_mono_quad(clipper_t *c, const pt src[3])
{
    pt p[3];
    int reverse;
    if (src[0].y > src[2].y) { p[0] = src[2]; p[1] = src[1]; p[2] = src[0]; reverse = 1; }
    else { p[0] = src[0]; p[1] = src[1]; p[2] = src[2]; reverse = 0; }
    rectf clip = c->clip;
    if (p[2].y <= clip.t || p[0].y >= clip.b) return;
    chop_quad_in_y(clip, p);
    if (p[0].x > p[2].x) { pt t = p[0]; p[0] = p[2]; p[2] = t; reverse = !reverse; }
    if (p[2].x <= clip.l) { append_vline(c, clip.l, p[0].y, p[2].y, reverse); return; }
    if (p[0].x >= clip.r) {
        if (!c->cull_right) append_vline(c, clip.r, p[0].y, p[2].y, reverse);
        return;
    }
    float t;
    pt tmp[5];
    if (p[0].x < clip.l) {
        if (chop_mono_quad_at(p[0].x, p[1].x, p[2].x, clip.l, &t)) {
            chop_quad_at(p, t, tmp);
            append_vline(c, clip.l, tmp[0].y, tmp[2].y, reverse);
            tmp[2].x = clip.l;
            if (tmp[3].x < clip.l) tmp[3].x = clip.l;
            p[0] = tmp[2]; p[1] = tmp[3];
        } else {
            append_vline(c, clip.l, p[0].y, p[2].y, reverse);
            return;
        }
    }
    if (p[2].x > clip.r) {
        if (chop_mono_quad_at(p[0].x, p[1].x, p[2].x, clip.r, &t)) {
            chop_quad_at(p, t, tmp);
            if (tmp[1].x > clip.r) tmp[1].x = clip.r;
            tmp[2].x = clip.r;
            append_quad(c, tmp, reverse);
            append_vline(c, clip.r, tmp[2].y, tmp[4].y, reverse);
        } else {
            if (p[1].x > clip.r) p[1].x = clip.r;
            if (p[2].x > clip.r) p[2].x = clip.r;
            append_quad(c, p, reverse);
        }
    } else {
        append_quad(c, p, reverse);
    }
}

static void clip_quad(clipper_t *c, const pt src[3])
{
    float minx = fminf(fminf(src[0].x, src[1].x), src[2].x), maxx = fmaxf(fmaxf(src[0].x, src[1].x), src[2].x);
    float miny = fminf(fminf(src[0].y, src[1].y), src[2].y), maxy = fmaxf(fmaxf(src[0].y, src[1].y), src[2].y);
    (void)minx; (void)maxx;
    if (!(maxy > c->clip.t && miny < c->clip.b)) return; /* quick_reject */
    pt mono_y[5];
    int cy = chop_quad_at_extrema(src, mono_y, 1);
    for (int y = 0; y <= cy; y++) {
        pt mono_x[5];
        int cx = chop_quad_at_extrema(&mono_y[y * 2], mono_x, 0);
        for (int x = 0; x <= cx; x++) clip_mono_quad(c, &mono_x[x * 2]);
    }
}

/* mono_cubic_closest_t over one coordinate (stride = 2 floats) */
static float mono_cubic_closest_t(const float *src, float x)
{
    float t = 0.5f, last_t, best_t = t, step = 0.25f;
    float d = src[0];
    float a = src[6] + 3 * (src[2] - src[4]) - d;
    float b = 3 * (src[4] - src[2] - src[2] + d);
    float cc = 3 * (src[2] - d);
    x -= d;
    float closest = 3.402823466e+38f;
    do {
        float loc = ((a * t + b) * t + cc) * t;
        float dist = fabsf(loc - x);
        if (closest > dist) { closest = dist; best_t = t; }
        last_t = t;
        t += loc < x ? step : -step;
        step *= 0.5f;
    } while (closest > 0.25f && last_t != t);
    return best_t;
}

static void chop_mono_cubic_at(const pt p[4], float v, int axis, pt dst[7])
{
    const float *src = axis ? &p[0].y : &p[0].x;
    chop_cubic_at2(p, mono_cubic_closest_t(src, v), dst);
}

static void chop_cubic_in_y(rectf clip, pt p[4])
{
    pt tmp[7];
    if (p[0].y < clip.t) {
        chop_mono_cubic_at(p, clip.t, 1, tmp);
        if (tmp[3].y < clip.t && tmp[4].y < clip.t && tmp[5].y < clip.t) {
            pt tmp2[4] = {tmp[3], tmp[4], tmp[5], tmp[6]};
            chop_mono_cubic_at(tmp2, clip.t, 1, tmp);
        }
        tmp[3].y = clip.t;
        if (tmp[4].y < clip.t) tmp[4].y = clip.t;
        p[0] = tmp[3]; p[1] = tmp[4]; p[2] = tmp[5];
    }
    if (p[3].y > clip.b) {
        chop_mono_cubic_at(p, clip.b, 1, tmp);
        tmp[3].y = clip.b;
        if (tmp[2].y > clip.b) tmp[2].y = clip.b;
        p[1] = tmp[1]; p[2] = tmp[2]; p[3] = tmp[3];
    }
}

static void clip_mono_cubic(clipper_t *c, const pt src[4])
{
    pt p[4];
    int reverse;
    if (src[0].y > src[3].y) { p[0] = src[3]; p[1] = src[2]; p[2] = src[1]; p[3] = src[0]; reverse = 1; }
    else { memcpy(p, src, sizeof(p)); reverse = 0; }
    rectf clip = c->clip;
    if (p[3].y <= clip.t || p[0].y >= clip.b) return;
    chop_cubic_in_y(clip, p);
    if (p[0].x > p[3].x) {
        pt t = p[0]; p[0] = p[3]; p[3] = t;
        t = p[1]; p[1] = p[2]; p[2] = t;
        reverse = !reverse;
    }
    if (p[3].x <= clip.l) { append_vline(c, clip.l, p[0].y, p[3].y, reverse); return; }
    if (p[0].x >= clip.r) {
        if (!c->cull_right) append_vline(c, clip.r, p[0].y, p[3].y, reverse);
        return;
    }
    pt tmp[7];
    if (p[0].x < clip.l) {
        chop_mono_cubic_at(p, clip.l, 0, tmp);
        append_vline(c, clip.l, tmp[0].y, tmp[3].y, reverse);
        tmp[3].x = clip.l;
        if (tmp[4].x < clip.l) tmp[4].x = clip.l;
        p[0] = tmp[3]; p[1] = tmp[4]; p[2] = tmp[5];
    }
    if (p[3].x > clip.r) {
        chop_mono_cubic_at(p, clip.r, 0, tmp);
        tmp[3].x = clip.r;
        if (tmp[2].x > clip.r) tmp[2].x = clip.r;
        append_cubic(c, tmp, reverse);
        append_vline(c, clip.r, tmp[3].y, tmp[6].y, reverse);
    } else {
        append_cubic(c, p, reverse);
    }
}

static void clip_cubic(clipper_t *c, const pt src[4])
{
    float minx = src[0].x, maxx = src[0].x, miny = src[0].y, maxy = src[0].y;
    for (int i = 1; i < 4; i++) {
        minx = fminf(minx, src[i].x); maxx = fmaxf(maxx, src[i].x);
        miny = fminf(miny, src[i].y); maxy = fmaxf(maxy, src[i].y);
    }
    if (!(maxy > c->clip.t && miny < c->clip.b)) return;
    const float limit = (float)(1 << 22);
    if (minx < -limit || miny < -limit || maxx > limit || maxy > limit) {
        clip_line(c, src[0], src[3]);
        return;
    }
    pt mono_y[10];
    int cy = chop_cubic_at_extrema(src, mono_y, 1);
    for (int y = 0; y <= cy; y++) {
        pt mono_x[10];
        int cx = chop_cubic_at_extrema(&mono_y[y * 3], mono_x, 0);
        for (int x = 0; x <= cx; x++) clip_mono_cubic(c, &mono_x[x * 3]);
    }
}

/* ---- BasicEdgeBuilder::build: iterate path edges (implicit close per contour, PathEdgeIter) ---- */
typedef struct {
    const uint8_t *verbs; int n_verbs;
    const pt *pts; int n_pts;
} path_t;

static int build_edges(builder_t *b, const path_t *path, const rectf *clip)
{
    clipper_t cl;
    if (clip) { cl.b = b; cl.clip = *clip; cl.cull_right = 0; }
    int pi = 0;
    pt move_to = {0, 0}, last = {0, 0};
    int needs_close = 0;
    for (int vi = 0; vi <= path->n_verbs; vi++) {
        int verb = vi < path->n_verbs ? path->verbs[vi] : ORC_CLOSE;
        if (verb == ORC_MOVE || verb == ORC_CLOSE) {
            if (needs_close) {
                if (clip) clip_line(&cl, last, move_to); else push_line(b, last, move_to);
                needs_close = 0;
            }
            if (verb == ORC_MOVE) { move_to = path->pts[pi++]; last = move_to; }
            else last = move_to;
            continue;
        }
        if (verb == ORC_LINE) {
            pt p1 = path->pts[pi++];
            if (clip) clip_line(&cl, last, p1); else push_line(b, last, p1);
            last = p1;
        } else if (verb == ORC_QUAD) {
            pt q[3] = {last, path->pts[pi], path->pts[pi + 1]};
            pi += 2;
            if (clip) clip_quad(&cl, q);
            else {
                pt mono[5];
                int n = chop_quad_at_extrema(q, mono, 1);
                for (int i = 0; i <= n; i++) push_quad(b, &mono[i * 2]);
            }
            last = q[2];
        } else if (verb == ORC_CUBIC) {
            pt q[4] = {last, path->pts[pi], path->pts[pi + 1], path->pts[pi + 2]};
            pi += 3;
            if (clip) clip_cubic(&cl, q);
            else {
                pt mono[10];
                int n = chop_cubic_at_extrema(q, mono, 1);
                for (int i = 0; i <= n; i++) push_cubic(b, &mono[i * 3]);
            }
            last = q[3];
        }
        needs_close = 1;
    }
    return b->n >= 2; /* build_edges: fewer than 2 edges -> nothing to draw */
}

/* ==========================================================================================
 * scan converter — tiny-skia scan/path.rs (Skia SkScan_Path.cpp)
 * ======================================================================================== */
typedef struct blitter blitter_t;
struct blitter {
    void (*blit_h)(blitter_t *self, int32_t x, int32_t y, int32_t width);
};

static void remove_edge(edge_t *e, int i)
{
    int p = e[i].prev, n = e[i].next;
    e[p].next = n;
    e[n].prev = p;
}
static void insert_edge_after(edge_t *e, int i, int after)
{
    e[i].prev = after;
    e[i].next = e[after].next;
    e[e[after].next].prev = i;
    e[after].next = i;
}
static void backward_insert_edge_based_on_x(edge_t *e, int i)
{
    fdot16 x = e[i].x;
    int prev = e[i].prev;
    while (e[prev].prev >= 0 && e[prev].x > x) prev = e[prev].prev;
    if (e[prev].next != i) {
        remove_edge(e, i);
        insert_edge_after(e, i, prev);
    }
}
static int backward_insert_start(edge_t *e, int prev, fdot16 x)
{
    while (e[prev].prev >= 0 && e[prev].x > x) prev = e[prev].prev;
    return prev;
}
static void insert_new_edges(edge_t *e, int new_edge, int32_t curr_y)
{
    if (e[new_edge].first_y != curr_y) return;
    int prev = e[new_edge].prev;
    if (e[prev].x <= e[new_edge].x) return;
    int start = backward_insert_start(e, prev, e[new_edge].x);
    do {
        int next = e[new_edge].next;
        int keep = 0;
        for (;;) {
            int after = e[start].next;
            if (after == new_edge) { keep = 1; break; }
            if (e[after].x >= e[new_edge].x) break;
            start = after;
        }
        if (!keep) {
            remove_edge(e, new_edge);
            insert_edge_after(e, new_edge, start);
        }
        start = new_edge;
        new_edge = next;
    } while (e[new_edge].first_y == curr_y);
}

static void walk_edges(int fill_rule, int32_t start_y, int32_t stop_y, int32_t right_clip, edge_t *e, blitter_t *bl)
{
    int32_t curr_y = start_y;
    int32_t mask = fill_rule == ORC_FILL_EVENODD ? 1 : -1;
    for (;;) {
        int32_t w = 0, left = 0;
        fdot16 prev_x = e[0].x;
        int curr = e[0].next;
        while (e[curr].first_y <= curr_y) {
            int32_t x = fdot16_round_to_i32(e[curr].x);
            if ((w & mask) == 0) left = x;
            w += e[curr].winding;
            if ((w & mask) == 0) {
                int32_t width = x - left;
                if (width > 0) bl->blit_h(bl, left, curr_y, width);
            }
            int next = e[curr].next;
            fdot16 new_x;
            int updated = 0;
            if (e[curr].last_y == curr_y) {
                if (e[curr].kind == 1 && e[curr].curve_count > 0 && quad_update(&e[curr])) updated = 1;
                else if (e[curr].kind == 2 && e[curr].curve_count < 0 && cubic_update(&e[curr])) updated = 1;
                else remove_edge(e, curr);
                if (updated) new_x = e[curr].x;
            } else {
                new_x = e[curr].x + e[curr].dx;
                e[curr].x = new_x;
                updated = 1;
            }
            if (updated) {
                if (new_x < prev_x) backward_insert_edge_based_on_x(e, curr);
                else prev_x = new_x;
            }
            curr = next;
        }
        if ((w & mask) != 0) {
            int32_t width = right_clip - left;
            if (width > 0) bl->blit_h(bl, left, curr_y, width);
        }
        curr_y += 1;
        if (curr_y >= stop_y) break;
        insert_new_edges(e, curr, curr_y);
    }
}

static int edge_cmp(const void *a, const void *b)
{
    const edge_t *ea = (const edge_t *)a, *eb = (const edge_t *)b;
    int32_t va = ea->first_y, vb = eb->first_y;
    if (va == vb) { va = ea->x; vb = eb->x; }
    if (va != vb) return va < vb ? -1 : 1;
    /* stable: `prev` temporarily holds the builder order */
    return ea->prev < eb->prev ? -1 : (ea->prev > eb->prev ? 1 : 0);
}

typedef struct { int32_t x, y, w, h; } irect;

/* scan::path::fill_path_impl */
static void fill_path_impl(const path_t *path, int fill_rule, irect clip, int32_t start_y, int32_t stop_y, int shift,
                           int contained, blitter_t *bl)
{
    builder_t b = {0, 0, 0, shift};
    rectf clipf = {(float)clip.x, (float)clip.y, (float)(clip.x + clip.w), (float)(clip.y + clip.h)};
    if (!build_edges(&b, path, contained ? NULL : &clipf)) { free(b.e); return; }
    int n = b.n;
    /* make room for the head (slot 0) and tail (slot n+1) sentinels */
    builder_push(&b);
    builder_push(&b);
    edge_t *e = b.e;
    memmove(e + 1, e, sizeof(edge_t) * (size_t)n);
    for (int i = 1; i <= n; i++) e[i].prev = i;
    qsort(e + 1, (size_t)n, sizeof(edge_t), edge_cmp);
    edge_t *tail = &e[n + 1];
    for (int i = 1; i <= n; i++) { e[i].prev = i - 1; e[i].next = i + 1; }
    memset(&e[0], 0, sizeof(edge_t));
    e[0].prev = -1; e[0].next = 1; e[0].x = (-2147483647 - 1); e[0].first_y = (-2147483647 - 1);
    memset(tail, 0, sizeof(edge_t));
    tail->prev = n; tail->next = -1; tail->first_y = 2147483647; tail->x = 2147483647;

    int32_t sclip_t = clip.y << shift, sclip_b = (clip.y + clip.h) << shift, sclip_r = (clip.x + clip.w) << shift;
    start_y = lsh(start_y, shift);
    stop_y = lsh(stop_y, shift);
    if (!contained && start_y < sclip_t) start_y = sclip_t;
    if (!contained && stop_y > sclip_b) stop_y = sclip_b;
    if (start_y >= 0 && stop_y > start_y) walk_edges(fill_rule, start_y, stop_y, sclip_r, e, bl);
    free(b.e);
}

/* path bounds over all points */
static int path_bounds(const path_t *p, rectf *out)
{
    if (p->n_pts == 0) return 0;
    /* tiny_skia_path::Path cannot hold a non-finite point (Rect::from_points fails in PathBuilder::finish): such a
     * path never reaches fill_path.  fmin/fmax would silently skip a NaN, so every point is tested. */
    float l = p->pts[0].x, r = l, t = p->pts[0].y, b = t, probe = 0.0f;
    for (int i = 0; i < p->n_pts; i++) {
        l = fminf(l, p->pts[i].x); r = fmaxf(r, p->pts[i].x);
        t = fminf(t, p->pts[i].y); b = fmaxf(b, p->pts[i].y);
        probe += p->pts[i].x * 0.0f + p->pts[i].y * 0.0f; /* NaN as soon as one coordinate is NaN or infinite */
    }
    if (!(probe == 0.0f)) return 0;
    if (!(isfinite(l) && isfinite(r) && isfinite(t) && isfinite(b))) return 0;
    out->l = l; out->t = t; out->r = r; out->b = b;
    return 1;
}

static int irect_intersect(irect a, irect b, irect *out)
{
    int64_t l = a.x > b.x ? a.x : b.x, t = a.y > b.y ? a.y : b.y;
    int64_t r = ((int64_t)a.x + a.w < (int64_t)b.x + b.w) ? (int64_t)a.x + a.w : (int64_t)b.x + b.w;
    int64_t bt = ((int64_t)a.y + a.h < (int64_t)b.y + b.h) ? (int64_t)a.y + a.h : (int64_t)b.y + b.h;
    if (r <= l || bt <= t) return 0;
    out->x = (int32_t)l; out->y = (int32_t)t; out->w = (int32_t)(r - l); out->h = (int32_t)(bt - t);
    return 1;
}
static int irect_contains(irect outer, irect in)
{
    return in.x >= outer.x && in.y >= outer.y && (int64_t)in.x + in.w <= (int64_t)outer.x + outer.w
           && (int64_t)in.y + in.h <= (int64_t)outer.y + outer.h;
}

/* scan::path::fill_path (non-AA) */
static void scan_fill_path(const path_t *path, int fill_rule, irect clip, blitter_t *bl)
{
    rectf bd;
    if (!path_bounds(path, &bd)) return;
    const double bias = 0.5 + 1.5 / 64.0;
    int32_t l = d2i_sat(ceil((double)bd.l - bias)), t = d2i_sat(ceil((double)bd.t - bias));
    int32_t r = d2i_sat(floor((double)bd.r + bias)), b = d2i_sat(floor((double)bd.b + bias));
    if ((int64_t)r - l <= 0 || (int64_t)b - t <= 0) return; /* IntRect::from_ltrb fails on empty */
    irect ir = {l, t, r - l, b - t};
    int contained = ir.x >= 0 && ir.y >= 0 && irect_contains(clip, ir);
    fill_path_impl(path, fill_rule, clip, ir.y, ir.y + ir.h, 0, contained, bl);
}

/* ==========================================================================================
 * anti-aliased scan converter — tiny-skia scan/path_aa.rs + alpha_runs.rs
 * (Skia SkScan_AntiPath.cpp SuperBlitter, SkAlphaRuns).  The run-length encoding of AlphaRuns is an
 * implementation detail; per-pixel accumulation below performs the same arithmetic per pixel.
 * ======================================================================================== */
#define SS_SHIFT 2
#define SS_SCALE 4
#define SS_MASK 3

typedef struct row_sink row_sink_t;
struct row_sink {
    /* coverage for pixels [x, x+n) of row y; cov[i] in 0..255 */
    void (*blit_row)(row_sink_t *self, int32_t x, int32_t y, const uint8_t *cov, int32_t n);
};

typedef struct {
    blitter_t base;
    row_sink_t *sink;
    int32_t left, super_left, width, top;
    int32_t curr_iy;
    uint16_t *alpha; /* width+1 */
    uint8_t *out;
    int dirty;
} super_blitter_t;

static void super_flush(super_blitter_t *s)
{
    if (s->curr_iy >= s->top && s->dirty) {
        for (int i = 0; i < s->width; i++) s->out[i] = (uint8_t)s->alpha[i];
        s->sink->blit_row(s->sink, s->left, s->curr_iy, s->out, s->width);
        memset(s->alpha, 0, sizeof(uint16_t) * (size_t)(s->width + 1));
        s->dirty = 0;
    }
}

static inline uint16_t catch_overflow(uint16_t a) { return (uint16_t)(a - (a >> 8)); }

static void super_blit_h(blitter_t *self, int32_t x, int32_t y, int32_t width)
{
    super_blitter_t *s = (super_blitter_t *)self;
    int32_t iy = y >> SS_SHIFT;
    x -= s->super_left;
    /* hack, until I figure out why my cubics (I think) go beyond the bounds */
    if (x < 0) { width += x; x = 0; }
    if (x + width > (s->width << SS_SHIFT)) width = (s->width << SS_SHIFT) - x;
    if (width <= 0) return;
    if (iy != s->curr_iy) {
        super_flush(s);
        s->curr_iy = iy;
    }
    int32_t start = x, stop = x + width;
    int32_t fb = start & SS_MASK, fe = stop & SS_MASK;
    int32_t n = (stop >> SS_SHIFT) - (start >> SS_SHIFT) - 1;
    if (n < 0) { fb = fe - fb; n = 0; fe = 0; }
    else {
        if (fb == 0) n += 1;
        else fb = SS_SCALE - fb;
    }
    uint16_t max_value = (uint16_t)((1 << (8 - SS_SHIFT)) - (((y & SS_MASK) + 1) >> SS_SHIFT));
    /* AlphaRuns::add(x >> SHIFT, fb << 4, n, fe << 4, max_value) */
    int32_t px = start >> SS_SHIFT;
    uint16_t *a = s->alpha + px;
    if (fb) {
        uint16_t tmp = (uint16_t)(a[0] + (fb << 4));
        a[0] = (uint16_t)(tmp - (tmp >> 8));
        a++;
    }
    for (int32_t i = 0; i < n; i++) a[i] = catch_overflow((uint16_t)(a[i] + max_value));
    a += n;
    if (fe) a[0] = (uint16_t)((a[0] + (fe << 4)) & 0xff);
    s->dirty = 1;
}

static int overflows_short_shift(int32_t v, int shift) { return ((int32_t)((int16_t)lsh(v, shift)) >> shift) != v; }

typedef struct { blitter_t base; row_sink_t *sink; irect clip; uint8_t *full; } direct_blitter_t;
static void direct_blit_h(blitter_t *self, int32_t x, int32_t y, int32_t width)
{
    direct_blitter_t *d = (direct_blitter_t *)self;
    /* spans are already inside the clip; clamp defensively */
    if (y < d->clip.y || y >= d->clip.y + d->clip.h) return;
    if (x < d->clip.x) { width -= d->clip.x - x; x = d->clip.x; }
    if (x + width > d->clip.x + d->clip.w) width = d->clip.x + d->clip.w - x;
    if (width <= 0) return;
    d->sink->blit_row(d->sink, x, y, d->full, width);
}

static void scan_fill_path_noaa_sink(const path_t *path, int fill_rule, irect clip, row_sink_t *sink)
{
    direct_blitter_t d;
    d.base.blit_h = direct_blit_h;
    d.sink = sink;
    d.clip = clip;
    d.full = (uint8_t *)malloc((size_t)clip.w + 1);
    memset(d.full, 255, (size_t)clip.w + 1);
    scan_fill_path(path, fill_rule, clip, &d.base);
    free(d.full);
}

/* scan::path_aa::fill_path */
static void scan_fill_path_aa(const path_t *path, int fill_rule, irect clip, row_sink_t *sink)
{
    rectf bd;
    if (!path_bounds(path, &bd)) return;
    float fl = floorf(bd.l), ft = floorf(bd.t), fr = ceilf(bd.r), fb = ceilf(bd.b);
    int32_t l = f2i(fl), t = f2i(ft), r = f2i(fr), b = f2i(fb);
    if ((int64_t)r - l <= 0 || (int64_t)b - t <= 0) return;
    irect ir = {l, t, (int32_t)((int64_t)r - l), (int32_t)((int64_t)b - t)};
    irect sect;
    if (!irect_intersect(ir, clip, &sect)) return;
    if (overflows_short_shift(sect.x, SS_SHIFT) || overflows_short_shift(sect.y, SS_SHIFT)
        || overflows_short_shift(sect.x + sect.w, SS_SHIFT) || overflows_short_shift(sect.y + sect.h, SS_SHIFT)) {
        scan_fill_path_noaa_sink(path, fill_rule, clip, sink);
        return;
    }
    if (clip.x + clip.w > 32767 || clip.y + clip.h > 32767) return;
    super_blitter_t s;
    s.base.blit_h = super_blit_h;
    s.sink = sink;
    s.left = sect.x;
    s.super_left = sect.x << SS_SHIFT;
    s.width = sect.w;
    s.top = sect.y;
    s.curr_iy = sect.y - 1;
    s.alpha = (uint16_t *)calloc((size_t)sect.w + 2, sizeof(uint16_t));
    s.out = (uint8_t *)malloc((size_t)sect.w + 1);
    s.dirty = 0;
    int contained = ir.x >= 0 && ir.y >= 0 && irect_contains(clip, ir);
    fill_path_impl(path, fill_rule, clip, ir.y, ir.y + ir.h, SS_SHIFT, contained, &s.base);
    super_flush(&s);
    free(s.alpha);
    free(s.out);
}

/* ==========================================================================================
 * raster pipeline — tiny-skia pipeline/{blitter,lowp,highp}.rs (Skia SkRasterPipeline_opts.h)
 * ======================================================================================== */
static inline uint16_t div255(uint32_t v) { return (uint16_t)((v + 255) >> 8); }
static inline uint16_t inv16(uint16_t v) { return (uint16_t)(255 - v); }

typedef struct { uint16_t r, g, b, a; } px16;
typedef struct { float r, g, b, a; } pxf;

static int blend_is_lowp(int m)
{
    switch (m) {
    case ORC_BLEND_COLOR_DODGE: case ORC_BLEND_COLOR_BURN: case ORC_BLEND_SOFT_LIGHT: case ORC_BLEND_HUE:
    case ORC_BLEND_SATURATION: case ORC_BLEND_COLOR: case ORC_BLEND_LUMINOSITY:
        return 0;
    default: return 1;
    }
}
static int blend_pre_scales(int m)
{
    switch (m) {
    case ORC_BLEND_DESTINATION: case ORC_BLEND_DESTINATION_OVER: case ORC_BLEND_PLUS: case ORC_BLEND_DESTINATION_OUT:
    case ORC_BLEND_SOURCE_ATOP: case ORC_BLEND_SOURCE_OVER: case ORC_BLEND_XOR:
        return 1;
    default: return 0;
    }
}

/* one colour channel of a lowp blend; s,d = channel, sa,da = alphas */
static inline uint16_t blend_ch_lowp(int m, uint32_t s, uint32_t d, uint32_t sa, uint32_t da)
{
    switch (m) {
    case ORC_BLEND_CLEAR: return 0;
    case ORC_BLEND_SOURCE: return (uint16_t)s;
    case ORC_BLEND_DESTINATION: return (uint16_t)d;
    case ORC_BLEND_SOURCE_OVER: return (uint16_t)(s + div255(d * (255 - sa)));
    case ORC_BLEND_DESTINATION_OVER: return (uint16_t)(d + div255(s * (255 - da)));
    case ORC_BLEND_SOURCE_IN: return div255(s * da);
    case ORC_BLEND_DESTINATION_IN: return div255(d * sa);
    case ORC_BLEND_SOURCE_OUT: return div255(s * (255 - da));
    case ORC_BLEND_DESTINATION_OUT: return div255(d * (255 - sa));
    case ORC_BLEND_SOURCE_ATOP: return div255(s * da + d * (255 - sa));
    case ORC_BLEND_DESTINATION_ATOP: return div255(d * sa + s * (255 - da));
    case ORC_BLEND_XOR: return div255(s * (255 - da) + d * (255 - sa));
    case ORC_BLEND_PLUS: return (uint16_t)((s + d) < 255 ? (s + d) : 255);
    case ORC_BLEND_MODULATE: return div255(s * d);
    case ORC_BLEND_SCREEN: return (uint16_t)(s + d - div255(s * d));
    case ORC_BLEND_MULTIPLY: return div255(s * (255 - da) + d * (255 - sa) + s * d);
    /* the following apply to colour channels only; alpha uses source-over */
    case ORC_BLEND_DARKEN: { uint32_t x = s * da, y = d * sa; return (uint16_t)(s + d - div255(x > y ? x : y)); }
    case ORC_BLEND_LIGHTEN: { uint32_t x = s * da, y = d * sa; return (uint16_t)(s + d - div255(x < y ? x : y)); }
    case ORC_BLEND_DIFFERENCE: { uint32_t x = s * da, y = d * sa; return (uint16_t)(s + d - 2 * div255(x < y ? x : y)); }
    case ORC_BLEND_EXCLUSION: return (uint16_t)(s + d - 2 * div255(s * d));
    case ORC_BLEND_HARD_LIGHT: {
        uint32_t t = (2 * s <= sa) ? 2 * s * d : sa * da - 2 * (sa - s) * (da - d);
        return div255(s * (255 - da) + d * (255 - sa) + t);
    }
    case ORC_BLEND_OVERLAY: {
        uint32_t t = (2 * d <= da) ? 2 * s * d : sa * da - 2 * (sa - s) * (da - d);
        return div255(s * (255 - da) + d * (255 - sa) + t);
    }
    default: return (uint16_t)s;
    }
}
static int blend_alpha_is_srcover(int m)
{
    switch (m) {
    case ORC_BLEND_DARKEN: case ORC_BLEND_LIGHTEN: case ORC_BLEND_DIFFERENCE: case ORC_BLEND_EXCLUSION:
    case ORC_BLEND_HARD_LIGHT: case ORC_BLEND_OVERLAY: case ORC_BLEND_COLOR_DODGE: case ORC_BLEND_COLOR_BURN:
    case ORC_BLEND_SOFT_LIGHT: case ORC_BLEND_HUE: case ORC_BLEND_SATURATION: case ORC_BLEND_COLOR:
    case ORC_BLEND_LUMINOSITY:
        return 1;
    default: return 0;
    }
}
static px16 blend_lowp(int m, px16 s, px16 d)
{
    px16 o;
    o.r = blend_ch_lowp(m, s.r, d.r, s.a, d.a);
    o.g = blend_ch_lowp(m, s.g, d.g, s.a, d.a);
    o.b = blend_ch_lowp(m, s.b, d.b, s.a, d.a);
    if (blend_alpha_is_srcover(m)) o.a = (uint16_t)(s.a + div255((uint32_t)d.a * (255 - s.a)));
    else o.a = blend_ch_lowp(m, s.a, d.a, s.a, d.a);
    return o;
}

/* ---- highp ---- */
static inline float finv(float v) { return 1.0f - v; }
static inline float two(float v) { return v + v; }
static inline float mad(float f, float m, float a) { return f * m + a; }
static inline float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
static inline float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
static inline float lum(float r, float g, float b) { return r * 0.30f + g * 0.59f + b * 0.11f; }
static inline float sat(float r, float g, float b) { return fmax3(r, g, b) - fmin3(r, g, b); }
static void set_sat(float *r, float *g, float *b, float s)
{
    float mn = fmin3(*r, *g, *b), mx = fmax3(*r, *g, *b), st = mx - mn;
    *r = st == 0.0f ? 0.0f : (*r - mn) * s / st;
    *g = st == 0.0f ? 0.0f : (*g - mn) * s / st;
    *b = st == 0.0f ? 0.0f : (*b - mn) * s / st;
}
static void set_lum(float *r, float *g, float *b, float l)
{
    float diff = l - lum(*r, *g, *b);
    *r += diff; *g += diff; *b += diff;
}
static void clip_color(float *r, float *g, float *b, float a)
{
    float mn = fmin3(*r, *g, *b), mx = fmax3(*r, *g, *b), l = lum(*r, *g, *b);
    float *ch[3] = {r, g, b};
    for (int i = 0; i < 3; i++) {
        float c = *ch[i];
        /* tiny-skia tests `mx >= 0` here (Skia tests mn); pinned by painting/mix-blend-mode/color.png */
        if (!(mx >= 0.0f)) c = l + (c - l) * l / (l - mn);
        if (mx > a) c = l + (c - l) * (a - l) / (mx - l);
        *ch[i] = fmaxf(c, 0.0f);
    }
}

static inline float blend_ch_highp(int m, float s, float d, float sa, float da)
{
    switch (m) {
    case ORC_BLEND_CLEAR: return 0.0f;
    case ORC_BLEND_SOURCE: return s;
    case ORC_BLEND_DESTINATION: return d;
    case ORC_BLEND_SOURCE_OVER: return mad(d, finv(sa), s);
    case ORC_BLEND_DESTINATION_OVER: return mad(s, finv(da), d);
    case ORC_BLEND_SOURCE_IN: return s * da;
    case ORC_BLEND_DESTINATION_IN: return d * sa;
    case ORC_BLEND_SOURCE_OUT: return s * finv(da);
    case ORC_BLEND_DESTINATION_OUT: return d * finv(sa);
    case ORC_BLEND_SOURCE_ATOP: return s * da + d * finv(sa);
    case ORC_BLEND_DESTINATION_ATOP: return d * sa + s * finv(da);
    case ORC_BLEND_XOR: return s * finv(da) + d * finv(sa);
    case ORC_BLEND_PLUS: return fminf(s + d, 1.0f);
    case ORC_BLEND_MODULATE: return s * d;
    case ORC_BLEND_SCREEN: return s + d - s * d;
    case ORC_BLEND_MULTIPLY: return s * finv(da) + d * finv(sa) + s * d;
    case ORC_BLEND_DARKEN: return s + d - fmaxf(s * da, d * sa);
    case ORC_BLEND_LIGHTEN: return s + d - fminf(s * da, d * sa);
    case ORC_BLEND_DIFFERENCE: return s + d - two(fminf(s * da, d * sa));
    case ORC_BLEND_EXCLUSION: return s + d - two(s * d);
    case ORC_BLEND_COLOR_BURN:
        if (d == da) return d + s * finv(da);
        if (s == 0.0f) return d * finv(sa);
        return sa * (da - fminf(da, (da - d) * sa * (1.0f / s))) + s * finv(da) + d * finv(sa);
    case ORC_BLEND_COLOR_DODGE:
        if (d == 0.0f) return s * finv(da);
        if (s == sa) return s + d * finv(sa);
        return sa * fminf(da, (d * sa) * (1.0f / (sa - s))) + s * finv(da) + d * finv(sa);
    case ORC_BLEND_HARD_LIGHT:
        return s * finv(da) + d * finv(sa) + (two(s) <= sa ? two(s * d) : sa * da - two((da - d) * (sa - s)));
    case ORC_BLEND_OVERLAY:
        return s * finv(da) + d * finv(sa) + (two(d) <= da ? two(s * d) : sa * da - two((da - d) * (sa - s)));
    case ORC_BLEND_SOFT_LIGHT: {
        float mm = da > 0.0f ? d / da : 0.0f, s2 = two(s), m4 = two(two(mm));
        float dark_src = d * (sa + (s2 - sa) * (1.0f - mm));
        float dark_dst = (m4 * m4 + m4) * (mm - 1.0f) + 7.0f * mm;
        float lite_dst = sqrtf(mm) - mm;
        float lite_src = d * sa + da * (s2 - sa) * (two(two(d)) <= da ? dark_dst : lite_dst);
        return s * finv(da) + d * finv(sa) + (s2 <= sa ? dark_src : lite_src);
    }
    default: return s;
    }
}

static pxf blend_highp(int m, pxf s, pxf d)
{
    pxf o;
    if (m >= ORC_BLEND_HUE) {
        float R, G, B;
        if (m == ORC_BLEND_HUE) {
            R = s.r * s.a; G = s.g * s.a; B = s.b * s.a;
            set_sat(&R, &G, &B, sat(d.r, d.g, d.b) * s.a);
            set_lum(&R, &G, &B, lum(d.r, d.g, d.b) * s.a);
        } else if (m == ORC_BLEND_SATURATION) {
            R = d.r * s.a; G = d.g * s.a; B = d.b * s.a;
            set_sat(&R, &G, &B, sat(s.r, s.g, s.b) * d.a);
            set_lum(&R, &G, &B, lum(d.r, d.g, d.b) * s.a);
        } else if (m == ORC_BLEND_COLOR) {
            R = s.r * d.a; G = s.g * d.a; B = s.b * d.a;
            set_lum(&R, &G, &B, lum(d.r, d.g, d.b) * s.a);
        } else {
            R = d.r * s.a; G = d.g * s.a; B = d.b * s.a;
            set_lum(&R, &G, &B, lum(s.r, s.g, s.b) * d.a);
        }
        clip_color(&R, &G, &B, s.a * d.a);
        o.r = s.r * finv(d.a) + d.r * finv(s.a) + R;
        o.g = s.g * finv(d.a) + d.g * finv(s.a) + G;
        o.b = s.b * finv(d.a) + d.b * finv(s.a) + B;
        o.a = s.a + d.a - s.a * d.a;
        return o;
    }
    o.r = blend_ch_highp(m, s.r, d.r, s.a, d.a);
    o.g = blend_ch_highp(m, s.g, d.g, s.a, d.a);
    o.b = blend_ch_highp(m, s.b, d.b, s.a, d.a);
    if (blend_alpha_is_srcover(m)) o.a = mad(d.a, finv(s.a), s.a);
    else o.a = blend_ch_highp(m, s.a, d.a, s.a, d.a);
    return o;
}

/* highp store: round(clamp(c,0,1)*255), SIMD round-to-nearest-even (cvtps2dq / vcvtnq) */
static inline uint8_t unnorm(float v)
{
    v = fmaxf(v, 0.0f);
    v = fminf(v, 1.0f);
    return (uint8_t)lrintf(v * 255.0f);
}

/* ==========================================================================================
 * shaders — tiny-skia shaders/{mod,gradient,linear_gradient,radial_gradient,pattern}.rs
 * ======================================================================================== */
#define MAX_STOPS 64

typedef struct {
    int kind;          /* 0 solid, 1 gradient, 2 pattern */
    int is_opaque;
    int lowp_ok;
    /* solid */
    pxf premul;        /* premultiplied colour */
    px16 solid16;      /* push_uniform_color: (c * 255 + 0.5) as u16 */
    /* common */
    xform ts;          /* pixel centre -> shader space (already inverted) */
    int has_ts;
    /* gradient */
    int geom;          /* 0 linear, 1 xy_to_radius, 2 two-point conical focal, 3 strip, 4 concentric */
    int spread;
    int pad_x1;
    int two_stop;
    float f[MAX_STOPS + 2][4], b[MAX_STOPS + 2][4], t[MAX_STOPS + 2];
    int len;
    int premul_after;
    /* two point conical */
    float p0, p1;
    int focal_on_circle, well_behaved, swapped, natively_focal, negate_x, smaller;
    float conc_scale, conc_bias;
    /* pattern */
    const uint8_t *pix; uint32_t pw, ph;
    int quality;
    float opacity;
} shader_t;

static pxf premultiply_color(const float c[4])
{
    pxf p;
    if (c[3] == 1.0f) { p.r = c[0]; p.g = c[1]; p.b = c[2]; p.a = c[3]; }
    else {
        p.r = fminf(fmaxf(c[0] * c[3], 0.0f), 1.0f);
        p.g = fminf(fmaxf(c[1] * c[3], 0.0f), 1.0f);
        p.b = fminf(fmaxf(c[2] * c[3], 0.0f), 1.0f);
        p.a = c[3];
    }
    return p;
}

typedef struct { float pos; float c[4]; } gstop;

/* Gradient::new + Gradient::push_stages colour set-up */
static void gradient_setup(shader_t *sh, const float *stops_in, int n_in, int spread)
{
    gstop st[MAX_STOPS + 2];
    int n = 0;
    if (n_in > MAX_STOPS) n_in = MAX_STOPS;
    float first_pos = fminf(fmaxf(stops_in[0], 0.0f), 1.0f);
    float last_pos = fminf(fmaxf(stops_in[(n_in - 1) * 5], 0.0f), 1.0f);
    int dummy_first = first_pos != 0.0f, dummy_last = last_pos != 1.0f;
    if (dummy_first) { st[n].pos = 0.0f; memcpy(st[n].c, stops_in + 1, 16); n++; }
    for (int i = 0; i < n_in; i++) {
        float p = stops_in[i * 5];
        st[n].pos = (p != p) ? 0.0f : fminf(fmaxf(p, 0.0f), 1.0f);
        memcpy(st[n].c, stops_in + i * 5 + 1, 16);
        n++;
    }
    if (dummy_last) { st[n].pos = 1.0f; memcpy(st[n].c, stops_in + (n_in - 1) * 5 + 1, 16); n++; }
    int opaque = 1;
    for (int i = 0; i < n; i++) if (st[i].c[3] != 1.0f) opaque = 0;
    int start_index = dummy_first ? 0 : 1;
    float prev = 0.0f;
    int uniform = 1;
    float uniform_step = st[start_index].pos - prev;
    for (int i = start_index; i < n; i++) {
        float curr = (i + 1 == n) ? 1.0f : fminf(fmaxf(st[i].pos, prev), 1.0f);
        uniform &= fabsf(uniform_step - (curr - prev)) <= SCALAR_NEARLY_ZERO;
        st[i].pos = curr;
        prev = curr;
    }
    sh->is_opaque = opaque;
    sh->spread = spread;
    sh->pad_x1 = (spread == ORC_SPREAD_PAD) && uniform;
    sh->premul_after = !opaque;
    if (n == 2) {
        sh->two_stop = 1;
        for (int k = 0; k < 4; k++) { sh->f[0][k] = st[1].c[k] - st[0].c[k]; sh->b[0][k] = st[0].c[k]; }
        sh->len = 1;
        return;
    }
    sh->two_stop = 0;
    int first_stop, last_stop;
    if (n > 2) {
        int first = memcmp(st[0].c, st[1].c, 16) != 0 ? 0 : 1;
        int last = memcmp(st[n - 2].c, st[n - 1].c, 16) != 0 ? n : n - 1;
        first_stop = first;
        last_stop = last - 1;
    } else { first_stop = 0; last_stop = 1; }
    float t_l = st[first_stop].pos;
    float c_l[4];
    memcpy(c_l, st[first_stop].c, 16);
    int len = 0;
    for (int k = 0; k < 4; k++) { sh->f[len][k] = 0.0f; sh->b[len][k] = c_l[k]; }
    sh->t[len] = 0.0f;
    len++;
    for (int i = first_stop; i < last_stop; i++) {
        float t_r = st[i + 1].pos;
        const float *c_r = st[i + 1].c;
        if (t_l < t_r) {
            for (int k = 0; k < 4; k++) {
                float ff = (c_r[k] - c_l[k]) / (t_r - t_l);
                sh->f[len][k] = ff;
                sh->b[len][k] = c_l[k] - ff * t_l;
            }
            sh->t[len] = t_l;
            len++;
        }
        t_l = t_r;
        memcpy(c_l, c_r, 16);
    }
    for (int k = 0; k < 4; k++) { sh->f[len][k] = 0.0f; sh->b[len][k] = c_l[k]; }
    sh->t[len] = t_l;
    len++;
    sh->len = len;
}

static xform ts_from_poly2(pt p0, pt p1)
{
    xform r = {p1.y - p0.y, p0.x - p1.x, p1.x - p0.x, p1.y - p0.y, p0.x, p0.y};
    return r;
}
static int ts_poly_to_poly(pt s0, pt s1, pt d0, pt d1, xform *out)
{
    xform tmp = ts_from_poly2(s0, s1), res;
    if (!ts_invert(tmp, &res)) return 0;
    tmp = ts_from_poly2(d0, d1);
    *out = ts_pre_concat(tmp, res);
    return 1;
}

#define DEGENERATE_THRESHOLD (1.0f / (1 << 15))

/* returns 1 ok, 0 = nothing to draw (None).  `total` = paint.shader.transform(ts) already applied to
 * the local transform. */
static int shader_prepare(shader_t *sh, const orc_paint *p, xform ctm)
{
    orc_paint solid_tmp;
    memset(sh, 0, sizeof(*sh));
    sh->lowp_ok = 1;
    if (p->shader == ORC_SHADER_SOLID) {
    solid:
        sh->kind = 0;
        sh->premul = premultiply_color(p->color);
        sh->is_opaque = p->color[3] == 1.0f;
        sh->solid16.r = f2u16(sh->premul.r * 255.0f + 0.5f);
        sh->solid16.g = f2u16(sh->premul.g * 255.0f + 0.5f);
        sh->solid16.b = f2u16(sh->premul.b * 255.0f + 0.5f);
        sh->solid16.a = f2u16(sh->premul.a * 255.0f + 0.5f);
        return 1;
    }
    xform local = ts_post_concat(ts_from(p->ts), ctm);
    if (p->shader == ORC_SHADER_PATTERN) {
        if (p->pattern_w == 0 || p->pattern_h == 0) return 0;
        xform inv;
        if (!ts_invert(local, &inv)) return 0;
        sh->kind = 2;
        sh->lowp_ok = 0;
        sh->ts = inv;
        sh->has_ts = ts_is_finite(inv) && !ts_is_identity(inv);
        sh->pix = p->pattern; sh->pw = p->pattern_w; sh->ph = p->pattern_h;
        sh->spread = p->spread;
        sh->quality = p->quality;
        if (ts_is_identity(inv) || ts_is_translate(inv)) sh->quality = ORC_QUALITY_NEAREST;
        sh->opacity = p->opacity;
        sh->is_opaque = 0;
        return 1;
    }
    /* gradients */
    if (p->n_stops < 1) return 0;
    if (p->n_stops == 1) {
        solid_tmp = *p;
        memcpy(solid_tmp.color, p->stops + 1, 16);
        p = &solid_tmp;
        goto solid;
    }
    xform inv;
    if (!ts_invert(local, &inv)) return 0;
    sh->kind = 1;
    xform unit;
    pt c0 = {p->x0, p->y0}, c1 = {p->x1, p->y1};
    if (p->shader == ORC_SHADER_LINEAR) {
        float dx = c1.x - c0.x, dy = c1.y - c0.y;
        float length = sqrtf(dx * dx + dy * dy);
        if (!isfinite(length)) return 0;
        if (nearly_zero_tol(length, DEGENERATE_THRESHOLD)) {
            /* degenerate: pad -> last colour; repeat/reflect -> average colour (approximated by last here) */
            solid_tmp = *p;
            memcpy(solid_tmp.color, p->stops + (p->n_stops - 1) * 5 + 1, 16);
            p = &solid_tmp;
            goto solid;
        }
        /* points_to_unit_ts */
        float mag = length, invm = mag != 0.0f ? 1.0f / mag : 0.0f;
        float vx = dx * invm, vy = dy * invm;
        float sn = -vy, cs = vx, cos_inv = 1.0f - cs;
        xform t = {cs, sn, -sn, cs, sn * c0.y + cos_inv * c0.x, -sn * c0.x + cos_inv * c0.y};
        t = ts_post_concat(t, ts_translate(-c0.x, -c0.y));
        t = ts_post_concat(t, ts_scale(invm, invm));
        unit = t;
        sh->geom = 0;
    } else {
        float r0 = p->r0, r1 = p->r1;
        if (r0 < 0 || r1 < 0) return 0;
        float dx = c0.x - c1.x, dy = c0.y - c1.y;
        float dlen = sqrtf(dx * dx + dy * dy);
        if (nearly_zero_tol(dlen, DEGENERATE_THRESHOLD)) {
            if (nearly_zero_tol(r0 - r1, DEGENERATE_THRESHOLD)) {
                solid_tmp = *p;
                memcpy(solid_tmp.color, p->stops + (p->n_stops - 1) * 5 + 1, 16);
                p = &solid_tmp;
                goto solid;
            }
            if (nearly_zero_tol(r0, DEGENERATE_THRESHOLD)) {
                /* simple radial */
                float ir = 1.0f / r1;
                unit = ts_post_concat(ts_translate(-c0.x, -c0.y), ts_scale(ir, ir));
                sh->geom = 1;
            } else {
                /* concentric two point conical */
                float scale = 1.0f / fmaxf(r0, r1);
                unit = ts_post_concat(ts_translate(-c1.x, -c1.y), ts_scale(scale, scale));
                float dr = r1 - r0;
                sh->conc_scale = fmaxf(r0, r1) / dr;
                sh->conc_bias = -r0 / dr;
                sh->geom = 4;
                sh->lowp_ok = 0;
            }
        } else {
            pt u0 = {0, 0}, u1 = {1, 0};
            if (!ts_poly_to_poly(c0, c1, u0, u1, &unit)) return 0;
            sh->lowp_ok = 0;
            if (nearly_zero(r1 - r0)) {
                sh->geom = 3;
                float sr0 = r0 / dlen;
                sh->p0 = sr0 * sr0;
            } else {
                sh->geom = 2;
                float fr0 = r0 / dlen, fr1 = r1 / dlen;
                float focal_x = fr0 / (fr0 - fr1);
                sh->swapped = 0;
                if (nearly_zero(focal_x - 1.0f)) {
                    unit = ts_post_concat(unit, ts_translate(-1.0f, 0.0f));
                    unit = ts_post_concat(unit, ts_scale(-1.0f, 1.0f));
                    float t = fr0; fr0 = fr1; fr1 = t;
                    focal_x = 0.0f;
                    sh->swapped = 1;
                }
                pt f0 = {focal_x, 0}, f1 = {1, 0};
                xform fm;
                if (!ts_poly_to_poly(f0, f1, u0, u1, &fm)) return 0;
                unit = ts_post_concat(unit, fm);
                float fr = fr1 / fabsf(1.0f - focal_x);
                sh->focal_on_circle = nearly_zero(1.0f - fr);
                sh->well_behaved = !sh->focal_on_circle && fr > 1.0f;
                sh->natively_focal = nearly_zero(focal_x);
                if (sh->focal_on_circle) unit = ts_post_concat(unit, ts_scale(0.5f, 0.5f));
                else unit = ts_post_concat(unit, ts_scale(fr / (fr * fr - 1.0f), 1.0f / sqrtf(fabsf(fr * fr - 1.0f))));
                float af = fabsf(1.0f - focal_x);
                unit = ts_post_concat(unit, ts_scale(af, af));
                sh->p0 = 1.0f / fr;
                sh->p1 = focal_x;
                sh->negate_x = (1.0f - focal_x) < 0.0f;
                sh->smaller = sh->swapped || sh->negate_x;
            }
        }
    }
    sh->ts = ts_post_concat(inv, unit);
    sh->has_ts = ts_is_finite(sh->ts) && !ts_is_identity(sh->ts);
    gradient_setup(sh, p->stops, p->n_stops, p->spread);
    return 1;
}

/* gradient t at pixel (x,y); *masked = 1 when the two-point-conical vector mask zeroes the pixel */
static float gradient_t(const shader_t *sh, int32_t px, int32_t py, int *masked)
{
    float x = (float)px + 0.5f, y = (float)py + 0.5f;
    if (sh->has_ts) {
        float nx = mad(x, sh->ts.sx, mad(y, sh->ts.kx, sh->ts.tx));
        float ny = mad(x, sh->ts.ky, mad(y, sh->ts.sy, sh->ts.ty));
        x = nx; y = ny;
    }
    *masked = 0;
    float t = x;
    switch (sh->geom) {
    case 0: break;
    case 1: t = sqrtf(x * x + y * y); break;
    case 4: t = sqrtf(x * x + y * y); t = t * sh->conc_scale + sh->conc_bias; break;
    case 3: {
        t = x + sqrtf(sh->p0 - y * y);
        if (t != t) { *masked = 1; t = 0.0f; }
        break;
    }
    case 2: {
        if (sh->focal_on_circle) t = x + y * y / x;
        else if (sh->well_behaved) t = sqrtf(x * x + y * y) - x * sh->p0;
        else if (sh->smaller) t = -sqrtf(x * x - y * y) - x * sh->p0;
        else t = sqrtf(x * x - y * y) - x * sh->p0;
        if (!sh->well_behaved) {
            if (t <= 0.0f || t != t) { *masked = 1; t = 0.0f; }
        }
        if (sh->negate_x) t = -t;
        if (!sh->natively_focal) t = t + sh->p1;
        if (sh->swapped) t = 1.0f - t;
        break;
    }
    }
    if (sh->spread == ORC_SPREAD_REFLECT) {
        float v = (t - 1.0f) - two(floorf((t - 1.0f) * 0.5f)) - 1.0f;
        v = fabsf(v);
        t = fminf(fmaxf(v, 0.0f), 1.0f);
    } else if (sh->spread == ORC_SPREAD_REPEAT) {
        float v = t - floorf(t);
        t = fminf(fmaxf(v, 0.0f), 1.0f);
    } else if (sh->pad_x1) {
        t = fminf(fmaxf(t, 0.0f), 1.0f);
    }
    return t;
}

static pxf gradient_color(const shader_t *sh, float t)
{
    int idx = 0;
    if (!sh->two_stop) for (int i = 1; i < sh->len; i++) if (t >= sh->t[i]) idx++;
    pxf c;
    c.r = mad(t, sh->f[idx][0], sh->b[idx][0]);
    c.g = mad(t, sh->f[idx][1], sh->b[idx][1]);
    c.b = mad(t, sh->f[idx][2], sh->b[idx][2]);
    c.a = mad(t, sh->f[idx][3], sh->b[idx][3]);
    return c;
}

static inline float normalize01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

static px16 shade_lowp(const shader_t *sh, int32_t x, int32_t y)
{
    px16 o;
    if (sh->kind == 0) return sh->solid16;
    int masked;
    float t = gradient_t(sh, x, y, &masked);
    pxf c = gradient_color(sh, t);
    /* round_f32_to_u16 */
    o.r = f2u16(normalize01(c.r) * 255.0f + 0.5f);
    o.g = f2u16(normalize01(c.g) * 255.0f + 0.5f);
    o.b = f2u16(normalize01(c.b) * 255.0f + 0.5f);
    o.a = f2u16(normalize01(c.a) * 255.0f + 0.5f);
    if (sh->premul_after) {
        o.r = div255((uint32_t)o.r * o.a);
        o.g = div255((uint32_t)o.g * o.a);
        o.b = div255((uint32_t)o.b * o.a);
    }
    return o;
}

/* ---- pattern sampling (highp only) ---- */
static inline float ulp_sub(float v)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    u -= 1;
    memcpy(&v, &u, 4);
    return v;
}
static inline pxf load_px(const uint8_t *p)
{
    pxf c = {(float)p[0] * (1.0f / 255.0f), (float)p[1] * (1.0f / 255.0f), (float)p[2] * (1.0f / 255.0f),
             (float)p[3] * (1.0f / 255.0f)};
    return c;
}
static inline float tile_coord(float v, int mode, float limit, float inv_limit)
{
    if (mode == ORC_SPREAD_REPEAT) return v - floorf(v * inv_limit) * limit;
    if (mode == ORC_SPREAD_REFLECT)
        return fabsf((v - limit) - (limit + limit) * floorf((v - limit) * (inv_limit * 0.5f)) - limit);
    return v;
}
static pxf gather(const shader_t *sh, float x, float y)
{
    float w = ulp_sub((float)sh->pw), h = ulp_sub((float)sh->ph);
    x = fminf(fmaxf(x, 0.0f), w);
    y = fminf(fmaxf(y, 0.0f), h);
    int32_t ix = f2i(x), iy = f2i(y);
    return load_px(sh->pix + ((size_t)iy * sh->pw + (size_t)ix) * 4);
}
static inline float bicubic_near(float t) { return mad(t, mad(t, mad(-21.0f / 18.0f, t, 27.0f / 18.0f), 9.0f / 18.0f), 1.0f / 18.0f); }
static inline float bicubic_far(float t) { return (t * t) * mad(7.0f / 18.0f, t, -6.0f / 18.0f); }

static pxf shade_pattern(const shader_t *sh, int32_t px, int32_t py)
{
    float x = (float)px + 0.5f, y = (float)py + 0.5f;
    if (sh->has_ts) {
        float nx = mad(x, sh->ts.sx, mad(y, sh->ts.kx, sh->ts.tx));
        float ny = mad(x, sh->ts.ky, mad(y, sh->ts.sy, sh->ts.ty));
        x = nx; y = ny;
    }
    float fw = (float)sh->pw, fh = (float)sh->ph, iw = 1.0f / fw, ih = 1.0f / fh;
    pxf c;
    if (sh->quality == ORC_QUALITY_NEAREST) {
        x = tile_coord(x, sh->spread, fw, iw);
        y = tile_coord(y, sh->spread, fh, ih);
        c = gather(sh, x, y);
    } else {
        int n = sh->quality == ORC_QUALITY_BILINEAR ? 2 : 4;
        float fx = (x + 0.5f) - floorf(x + 0.5f), fy = (y + 0.5f) - floorf(y + 0.5f);
        float wx[4], wy[4];
        if (n == 2) {
            wx[0] = 1.0f - fx; wx[1] = fx; wy[0] = 1.0f - fy; wy[1] = fy;
        } else {
            wx[0] = bicubic_far(1.0f - fx); wx[1] = bicubic_near(1.0f - fx); wx[2] = bicubic_near(fx); wx[3] = bicubic_far(fx);
            wy[0] = bicubic_far(1.0f - fy); wy[1] = bicubic_near(1.0f - fy); wy[2] = bicubic_near(fy); wy[3] = bicubic_far(fy);
        }
        float start = -0.5f * (float)(n - 1);
        c.r = c.g = c.b = c.a = 0.0f;
        float yy = y + start;
        for (int j = 0; j < n; j++) {
            float xx = x + start;
            for (int i = 0; i < n; i++) {
                pxf s = gather(sh, tile_coord(xx, sh->spread, fw, iw), tile_coord(yy, sh->spread, fh, ih));
                float w = wx[i] * wy[j];
                c.r = mad(w, s.r, c.r); c.g = mad(w, s.g, c.g); c.b = mad(w, s.b, c.b); c.a = mad(w, s.a, c.a);
                xx = xx + 1.0f;
            }
            yy = yy + 1.0f;
        }
        if (n == 4) {
            c.r = fmaxf(c.r, 0.0f); c.g = fmaxf(c.g, 0.0f); c.b = fmaxf(c.b, 0.0f); c.a = fmaxf(c.a, 0.0f);
            c.a = fminf(c.a, 1.0f);
            c.r = fminf(c.r, c.a); c.g = fminf(c.g, c.a); c.b = fminf(c.b, c.a);
        }
    }
    if (sh->opacity != 1.0f) { c.r *= sh->opacity; c.g *= sh->opacity; c.b *= sh->opacity; c.a *= sh->opacity; }
    return c;
}

static pxf shade_highp(const shader_t *sh, int32_t x, int32_t y)
{
    if (sh->kind == 0) return sh->premul;
    if (sh->kind == 2) return shade_pattern(sh, x, y);
    int masked;
    float t = gradient_t(sh, x, y, &masked);
    pxf c = gradient_color(sh, t);
    if (sh->premul_after) { c.r *= c.a; c.g *= c.a; c.b *= c.a; }
    if (masked) c.r = c.g = c.b = c.a = 0.0f;
    return c;
}

/* ==========================================================================================
 * RasterPipelineBlitter — tiny-skia pipeline/blitter.rs
 * ======================================================================================== */
typedef struct {
    row_sink_t base;
    uint8_t *px; uint32_t w, h;
    shader_t sh;
    int blend;
    int lowp;
    int has_memset; uint8_t memset_c[4];
    int noop;
} pix_blitter_t;

static void pix_blit_row(row_sink_t *self, int32_t x, int32_t y, const uint8_t *cov, int32_t n)
{
    pix_blitter_t *b = (pix_blitter_t *)self;
    if (y < 0 || y >= (int32_t)b->h) return;
    for (int32_t i = 0; i < n; i++) {
        int32_t xx = x + i;
        uint8_t c = cov[i];
        if (c == 0 || xx < 0 || xx >= (int32_t)b->w) continue;
        uint8_t *p = b->px + ((size_t)y * b->w + (size_t)xx) * 4;
        if (b->lowp) {
            px16 d = {p[0], p[1], p[2], p[3]};
            px16 o;
            if (c == 255) {
                if (b->has_memset) { memcpy(p, b->memset_c, 4); continue; }
                px16 s = shade_lowp(&b->sh, xx, y);
                o = b->blend == ORC_BLEND_SOURCE ? s : blend_lowp(b->blend, s, d);
            } else {
                px16 s = shade_lowp(&b->sh, xx, y);
                if (blend_pre_scales(b->blend)) {
                    s.r = div255((uint32_t)s.r * c); s.g = div255((uint32_t)s.g * c);
                    s.b = div255((uint32_t)s.b * c); s.a = div255((uint32_t)s.a * c);
                    o = blend_lowp(b->blend, s, d);
                } else {
                    px16 t = blend_lowp(b->blend, s, d);
                    o.r = div255((uint32_t)d.r * (255 - c) + (uint32_t)t.r * c);
                    o.g = div255((uint32_t)d.g * (255 - c) + (uint32_t)t.g * c);
                    o.b = div255((uint32_t)d.b * (255 - c) + (uint32_t)t.b * c);
                    o.a = div255((uint32_t)d.a * (255 - c) + (uint32_t)t.a * c);
                }
            }
            p[0] = (uint8_t)o.r; p[1] = (uint8_t)o.g; p[2] = (uint8_t)o.b; p[3] = (uint8_t)o.a;
        } else {
            pxf d = load_px(p);
            pxf s = shade_highp(&b->sh, xx, y);
            pxf o;
            if (c == 255) {
                if (b->has_memset) { memcpy(p, b->memset_c, 4); continue; }
                o = b->blend == ORC_BLEND_SOURCE ? s : blend_highp(b->blend, s, d);
            } else {
                float cf = (float)c * (1.0f / 255.0f);
                if (blend_pre_scales(b->blend)) {
                    s.r *= cf; s.g *= cf; s.b *= cf; s.a *= cf;
                    o = blend_highp(b->blend, s, d);
                } else {
                    pxf t = blend_highp(b->blend, s, d);
                    /* lerp(from, to, t) = mad(to - from, t, from) */
                    o.r = mad(t.r - d.r, cf, d.r); o.g = mad(t.g - d.g, cf, d.g);
                    o.b = mad(t.b - d.b, cf, d.b); o.a = mad(t.a - d.a, cf, d.a);
                }
            }
            p[0] = unnorm(o.r); p[1] = unnorm(o.g); p[2] = unnorm(o.b); p[3] = unnorm(o.a);
        }
    }
}

/* RasterPipelineBlitter::new; returns 0 when there is nothing to draw */
static int pix_blitter_init(pix_blitter_t *b, uint8_t *px, uint32_t w, uint32_t h, const orc_paint *paint, xform ctm)
{
    memset(b, 0, sizeof(*b));
    b->base.blit_row = pix_blit_row;
    b->px = px; b->w = w; b->h = h;
    if (!shader_prepare(&b->sh, paint, ctm)) return 0;
    int blend = paint->blend_mode;
    if (blend == ORC_BLEND_DESTINATION) return 0;
    if (blend == ORC_BLEND_DESTINATION_IN && b->sh.is_opaque) return 0;
    if (b->sh.is_opaque && blend == ORC_BLEND_SOURCE_OVER) blend = ORC_BLEND_SOURCE;
    if (b->sh.kind == 0 && blend == ORC_BLEND_SOURCE) {
        b->has_memset = 1;
        b->memset_c[0] = f2u8(b->sh.premul.r * 255.0f + 0.5f);
        b->memset_c[1] = f2u8(b->sh.premul.g * 255.0f + 0.5f);
        b->memset_c[2] = f2u8(b->sh.premul.b * 255.0f + 0.5f);
        b->memset_c[3] = f2u8(b->sh.premul.a * 255.0f + 0.5f);
    }
    if (blend == ORC_BLEND_CLEAR) {
        /* Clear is just a transparent colour memset */
        blend = ORC_BLEND_SOURCE;
        b->has_memset = 1;
        memset(b->memset_c, 0, 4);
        memset(&b->sh, 0, sizeof(b->sh));
        b->sh.kind = 0;
        b->sh.lowp_ok = 1;
    }
    b->blend = blend;
    b->lowp = b->sh.lowp_ok && blend_is_lowp(blend) && !paint->force_hq;
    return 1;
}

/* ==========================================================================================
 * painter — tiny-skia painter.rs
 * ======================================================================================== */
#define MAX_DIM 8191

typedef struct {
    row_sink_t base;
    row_sink_t *inner;
    int32_t dx, dy;
} offset_sink_t;
static void offset_blit_row(row_sink_t *self, int32_t x, int32_t y, const uint8_t *cov, int32_t n)
{
    offset_sink_t *o = (offset_sink_t *)self;
    o->inner->blit_row(o->inner, x + o->dx, y + o->dy, cov, n);
}

/* Draws `path` (device space, identity transform) through `sink`, applying the DrawTiler split. */
static void draw_path_tiled(uint32_t w, uint32_t h, const path_t *path, int fill_rule, int aa, row_sink_t *sink,
                            void (*retarget)(void *ud, float dx, float dy), void *ud)
{
    rectf bd0;
    if (!path_bounds(path, &bd0)) return;
    if (nearly_zero(bd0.r - bd0.l) || nearly_zero(bd0.b - bd0.t)) return;
    if (w <= MAX_DIM && h <= MAX_DIM) {
        irect clip = {0, 0, (int32_t)w, (int32_t)h};
        if (aa) scan_fill_path_aa(path, fill_rule, clip, sink);
        else scan_fill_path_noaa_sink(path, fill_rule, clip, sink);
        return;
    }
    pt *tmp = (pt *)malloc(sizeof(pt) * (size_t)(path->n_pts ? path->n_pts : 1));
    memcpy(tmp, path->pts, sizeof(pt) * (size_t)path->n_pts);
    path_t tp = *path;
    tp.pts = tmp;
    for (uint32_t ty = 0; ty < h; ty += MAX_DIM) {
        for (uint32_t tx = 0; tx < w; tx += MAX_DIM) {
            uint32_t tw = w - tx < MAX_DIM ? w - tx : MAX_DIM, th = h - ty < MAX_DIM ? h - ty : MAX_DIM;
            ts_map_points(ts_translate(-(float)tx, -(float)ty), tmp, path->n_pts);
            if (retarget) retarget(ud, -(float)tx, -(float)ty);
            offset_sink_t os;
            os.base.blit_row = offset_blit_row;
            os.inner = sink;
            os.dx = (int32_t)tx; os.dy = (int32_t)ty;
            irect clip = {0, 0, (int32_t)tw, (int32_t)th};
            if (aa) scan_fill_path_aa(&tp, fill_rule, clip, &os.base);
            else scan_fill_path_noaa_sink(&tp, fill_rule, clip, &os.base);
            ts_map_points(ts_translate((float)tx, (float)ty), tmp, path->n_pts);
            if (retarget) retarget(ud, (float)tx, (float)ty);
        }
    }
    free(tmp);
}

static int fill_path_device(uint8_t *px, uint32_t w, uint32_t h, const path_t *path, const orc_paint *paint,
                            int fill_rule, xform ctm)
{
    rectf bd;
    if (!path_bounds(path, &bd)) return 0;
    if (nearly_zero(bd.r - bd.l) || nearly_zero(bd.b - bd.t)) return 0;
    if (w <= MAX_DIM && h <= MAX_DIM) {
        pix_blitter_t b;
        if (!pix_blitter_init(&b, px, w, h, paint, ctm)) return 0;
        irect clip = {0, 0, (int32_t)w, (int32_t)h};
        if (paint->anti_alias) scan_fill_path_aa(path, fill_rule, clip, &b.base);
        else scan_fill_path_noaa_sink(path, fill_rule, clip, &b.base);
        return 1;
    }
    /* DrawTiler */
    pt *tmp = (pt *)malloc(sizeof(pt) * (size_t)(path->n_pts ? path->n_pts : 1));
    memcpy(tmp, path->pts, sizeof(pt) * (size_t)path->n_pts);
    path_t tp = *path;
    tp.pts = tmp;
    xform sts = ctm;
    for (uint32_t ty = 0; ty < h; ty += MAX_DIM) {
        for (uint32_t tx = 0; tx < w; tx += MAX_DIM) {
            uint32_t tw = w - tx < MAX_DIM ? w - tx : MAX_DIM, th = h - ty < MAX_DIM ? h - ty : MAX_DIM;
            xform tr = ts_translate(-(float)tx, -(float)ty);
            ts_map_points(tr, tmp, path->n_pts);
            sts = ts_post_concat(sts, tr);
            /* sub-pixmap view: rows of the tile inside the big pixmap.  The pixel blitter addresses a w-wide
             * buffer, so hand it the tile origin pointer with the full stride and tile-local sizes. */
            pix_blitter_t b;
            if (pix_blitter_init(&b, px + ((size_t)ty * w + tx) * 4, w, th, paint, sts)) {
                b.w = w; /* stride */
                irect clip = {0, 0, (int32_t)tw, (int32_t)th};
                /* restrict x writes to the tile: clip guarantees spans stay within [0, tw) */
                if (paint->anti_alias) scan_fill_path_aa(&tp, fill_rule, clip, &b.base);
                else scan_fill_path_noaa_sink(&tp, fill_rule, clip, &b.base);
            }
            tr = ts_translate((float)tx, (float)ty);
            ts_map_points(tr, tmp, path->n_pts);
            sts = ts_post_concat(sts, tr);
        }
    }
    free(tmp);
    return 1;
}

int orc_fill_path(uint8_t *px, uint32_t w, uint32_t h, const uint8_t *verbs, int32_t n_verbs, const float *pts,
                  int32_t n_pts, const orc_paint *paint, int32_t fill_rule, const float ts[6])
{
    xform ctm = ts ? ts_from(ts) : ts_identity();
    pt *p = (pt *)malloc(sizeof(pt) * (size_t)(n_pts ? n_pts : 1));
    memcpy(p, pts, sizeof(pt) * (size_t)n_pts);
    if (!ts_is_identity(ctm)) ts_map_points(ctm, p, n_pts);
    path_t path = {verbs, n_verbs, p, n_pts};
    int r = fill_path_device(px, w, h, &path, paint, fill_rule, ctm);
    free(p);
    return r;
}

/* Blitter::blit_anti_h with one pixel per call: the hairline walkers (scan/hairline_aa.rs) emit their coverage this way,
 * two pixels per step, every one a separate blend through the RasterPipelineBlitter (alpha 255 = the full-coverage
 * program).  blits = n x {x, y, alpha}; ts is the paint's (= the draw's) transform. */
int orc_blit_coverage(uint8_t *px, uint32_t w, uint32_t h, int32_t n, const int32_t *blits, const orc_paint *paint, const float ts[6])
{
    pix_blitter_t b;
    if (!pix_blitter_init(&b, px, w, h, paint, ts ? ts_from(ts) : ts_identity())) return 0;
    for (int32_t i = 0; i < n; i++) {
        const uint8_t c = (uint8_t)blits[3 * i + 2];
        b.base.blit_row(&b.base, blits[3 * i], blits[3 * i + 1], &c, 1);
    }
    return 1;
}

/* scan::fill_rect (non-AA): Rect::round() then intersect */
static int fill_int_rect(uint8_t *px, uint32_t w, uint32_t h, int32_t x, int32_t y, int32_t rw, int32_t rh,
                         const orc_paint *paint, xform ctm)
{
    pix_blitter_t b;
    if (!pix_blitter_init(&b, px, w, h, paint, ctm)) return 0;
    irect r = {x, y, rw, rh}, clip = {0, 0, (int32_t)w, (int32_t)h}, s;
    if (!irect_intersect(r, clip, &s)) return 0;
    uint8_t *full = (uint8_t *)malloc((size_t)s.w);
    memset(full, 255, (size_t)s.w);
    for (int32_t yy = s.y; yy < s.y + s.h; yy++) b.base.blit_row(&b.base, s.x, yy, full, s.w);
    free(full);
    return 1;
}

static inline int32_t sat_round(float v) { return f2i(roundf(v)); }

int orc_fill_rect(uint8_t *px, uint32_t w, uint32_t h, float x, float y, float rw, float rh, const orc_paint *paint,
                  const float ts[6])
{
    xform ctm = ts ? ts_from(ts) : ts_identity();
    if (ts_is_identity(ctm) && w <= MAX_DIM && h <= MAX_DIM && !paint->anti_alias) {
        int32_t ix = sat_round(x), iy = sat_round(y);
        int32_t iw = sat_round(rw), ih = sat_round(rh);
        if (iw < 1) iw = 1;
        if (ih < 1) ih = 1;
        return fill_int_rect(px, w, h, ix, iy, iw, ih, paint, ctm);
    }
    /* PathBuilder::from_rect + fill_path(Winding).  (fill_rect_aa for identity+AA is equivalent to the AA path
     * fill of the same rectangle up to the SuperBlitter's 1/4-pixel quantisation; resvg never issues it.) */
    uint8_t verbs[5] = {ORC_MOVE, ORC_LINE, ORC_LINE, ORC_LINE, ORC_CLOSE};
    float pts[8] = {x, y, x + rw, y, x + rw, y + rh, x, y + rh};
    return orc_fill_path(px, w, h, verbs, 5, pts, 4, paint, ORC_FILL_WINDING, ts);
}

int orc_draw_pixmap(uint8_t *dst, uint32_t dw, uint32_t dh, int32_t x, int32_t y, const uint8_t *src, uint32_t sw,
                    uint32_t sh, float opacity, int32_t blend_mode, int32_t quality, const float ts[6])
{
    orc_paint p;
    memset(&p, 0, sizeof(p));
    p.shader = ORC_SHADER_PATTERN;
    p.pattern = src; p.pattern_w = sw; p.pattern_h = sh;
    p.spread = ORC_SPREAD_PAD;
    p.quality = quality;
    p.opacity = opacity;
    p.ts[0] = 1; p.ts[3] = 1; p.ts[4] = (float)x; p.ts[5] = (float)y;
    p.blend_mode = blend_mode;
    p.anti_alias = 0;
    return orc_fill_rect(dst, dw, dh, (float)x, (float)y, (float)sw, (float)sh, &p, ts);
}

void orc_pixmap_fill(uint8_t *px, uint32_t w, uint32_t h, float r, float g, float b, float a)
{
    float c[4] = {r, g, b, a};
    pxf p = premultiply_color(c);
    uint8_t v[4] = {f2u8(p.r * 255.0f + 0.5f), f2u8(p.g * 255.0f + 0.5f), f2u8(p.b * 255.0f + 0.5f), f2u8(p.a * 255.0f + 0.5f)};
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) memcpy(px + i * 4, v, 4);
}

/* ==========================================================================================
 * masks — tiny-skia mask.rs
 * ======================================================================================== */
void orc_mask_from_pixmap(const uint8_t *px, uint32_t w, uint32_t h, int32_t type, uint8_t *mask)
{
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *p = px + i * 4;
        if (type == 0) { mask[i] = p[3]; continue; }
        float r = (float)p[0] / 255.0f, g = (float)p[1] / 255.0f, b = (float)p[2] / 255.0f, a = (float)p[3] / 255.0f;
        if (p[3] != 0) { r /= a; g /= a; b /= a; }
        /* Rec. 709 coefficients; pinned by masking/mask/simple-case.png (0 of 240 probe pixels differ,
         * whereas the 0.2125/0.7154/0.0721 set used by filter/color_matrix.rs gives 39 mismatches) */
        float luma = r * 0.2126f + g * 0.7152f + b * 0.0722f;
        float v = (luma * a) * 255.0f;
        v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        mask[i] = f2u8(ceilf(v));
    }
}

void orc_mask_invert(uint8_t *mask, uint32_t w, uint32_t h)
{
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) mask[i] = (uint8_t)(255 - mask[i]);
}

/* LoadMaskU8, LoadDestination, DestinationIn, Store in lowp: c' = div255(c * m) */
void orc_apply_mask(uint8_t *px, uint32_t w, uint32_t h, const uint8_t *mask)
{
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) {
        uint32_t m = mask[i];
        for (int c = 0; c < 4; c++) px[i * 4 + c] = (uint8_t)div255((uint32_t)px[i * 4 + c] * m);
    }
}

typedef struct { row_sink_t base; uint8_t *m; uint32_t w, h; int lerp; } mask_sink_t;
static void mask_blit_row(row_sink_t *self, int32_t x, int32_t y, const uint8_t *cov, int32_t n)
{
    mask_sink_t *s = (mask_sink_t *)self;
    if (y < 0 || y >= (int32_t)s->h) return;
    for (int32_t i = 0; i < n; i++) {
        int32_t xx = x + i;
        uint32_t c = cov[i];
        if (c == 0 || xx < 0 || xx >= (int32_t)s->w) continue;
        uint8_t *p = s->m + (size_t)y * s->w + (size_t)xx;
        if (!s->lerp) { *p = (uint8_t)c; continue; }
        /* new_mask pipelines: full coverage stores 255; partial: lerp(dst, 255, c) */
        if (c == 255) *p = 255;
        else *p = (uint8_t)div255((uint32_t)*p * (255 - c) + 255u * c);
    }
}

static int coverage_common(uint8_t *m, uint32_t w, uint32_t h, const uint8_t *verbs, int32_t n_verbs, const float *pts,
                           int32_t n_pts, int32_t fill_rule, int32_t aa, const float ts[6], int lerp)
{
    xform ctm = ts ? ts_from(ts) : ts_identity();
    pt *p = (pt *)malloc(sizeof(pt) * (size_t)(n_pts ? n_pts : 1));
    memcpy(p, pts, sizeof(pt) * (size_t)n_pts);
    if (!ts_is_identity(ctm)) ts_map_points(ctm, p, n_pts);
    path_t path = {verbs, n_verbs, p, n_pts};
    mask_sink_t s;
    s.base.blit_row = mask_blit_row;
    s.m = m; s.w = w; s.h = h; s.lerp = lerp;
    draw_path_tiled(w, h, &path, fill_rule, aa, &s.base, NULL, NULL);
    free(p);
    return 1;
}

int orc_mask_fill_path(uint8_t *mask, uint32_t w, uint32_t h, const uint8_t *verbs, int32_t n_verbs, const float *pts,
                       int32_t n_pts, int32_t fill_rule, int32_t anti_alias, const float ts[6])
{
    return coverage_common(mask, w, h, verbs, n_verbs, pts, n_pts, fill_rule, anti_alias, ts, 1);
}

int orc_path_coverage(uint8_t *cov, uint32_t w, uint32_t h, const uint8_t *verbs, int32_t n_verbs, const float *pts,
                      int32_t n_pts, int32_t fill_rule, int32_t anti_alias, const float ts[6])
{
    memset(cov, 0, (size_t)w * h);
    return coverage_common(cov, w, h, verbs, n_verbs, pts, n_pts, fill_rule, anti_alias, ts, 0);
}

/* Painter's-order loop over packed paths: the CPU baseline / checker for a whole scene in one call. */
int orc_fill_paths(uint8_t *px, uint32_t w, uint32_t h, int32_t n_paths, const uint32_t *verb_off, const uint32_t *pt_off,
                   const uint8_t *verbs, const float *pts, const orc_paint *paints, const uint8_t *rules, const float ts[6])
{
    int drawn = 0;
    for (int32_t i = 0; i < n_paths; i++)
        drawn += orc_fill_path(px, w, h, verbs + verb_off[i], (int32_t)(verb_off[i + 1] - verb_off[i]),
                               pts + 2 * (size_t)pt_off[i], (int32_t)(pt_off[i + 1] - pt_off[i]), &paints[i], rules[i], ts);
    return drawn;
}
