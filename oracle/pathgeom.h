/*
 * pathgeom.h — path container and curve helpers shared by the oracle's stroker, dasher and hairline walker
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * tiny-skia-path 0.12.0 (Cargo.lock:669-670) is a crates.io dependency whose source is not under /root/reference;
 * this restates the published algorithm of its path_builder.rs and path_geometry.rs (a Rust port of Skia's SkPath /
 * SkGeometry) sequentially in f32.  Written independently of resvg_b200/csrc: nothing is shared with the product.
 */
#ifndef RESVG_B200_ORACLE_PATHGEOM_H
#define RESVG_B200_ORACLE_PATHGEOM_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { PG_MOVE = 0, PG_LINE = 1, PG_QUAD = 2, PG_CUBIC = 3, PG_CLOSE = 4 };

#define PG_NEARLY_ZERO (1.0f / 4096.0f)   /* SCALAR_NEARLY_ZERO */
#define PG_ROOT2_OVER_2 0.707106781f      /* SCALAR_ROOT_2_OVER_2 */

typedef struct { float x, y; } pg_pt;

static inline pg_pt pg_p(float x, float y) { pg_pt p = {x, y}; return p; }
static inline pg_pt pg_add(pg_pt a, pg_pt b) { return pg_p(a.x + b.x, a.y + b.y); }
static inline pg_pt pg_sub(pg_pt a, pg_pt b) { return pg_p(a.x - b.x, a.y - b.y); }
static inline pg_pt pg_neg(pg_pt a) { return pg_p(-a.x, -a.y); }
static inline pg_pt pg_scale(pg_pt a, float s) { return pg_p(a.x * s, a.y * s); }
static inline float pg_dot(pg_pt a, pg_pt b) { return a.x * b.x + a.y * b.y; }
static inline float pg_cross(pg_pt a, pg_pt b) { return a.x * b.y - a.y * b.x; }
static inline int pg_eq(pg_pt a, pg_pt b) { return a.x == b.x && a.y == b.y; }
static inline int pg_finite(pg_pt a) { return isfinite(a.x) && isfinite(a.y); }
static inline float pg_len_sqd(pg_pt a) { return a.x * a.x + a.y * a.y; }
static inline float pg_dist_sqd(pg_pt a, pg_pt b) { return pg_len_sqd(pg_sub(a, b)); }
static inline int pg_nearly_zero(float v) { return fabsf(v) <= PG_NEARLY_ZERO; }
static inline float pg_interp(float a, float b, float t) { return a + (b - a) * t; }
static inline pg_pt pg_lerp(pg_pt a, pg_pt b, float t) { return pg_p(pg_interp(a.x, b.x, t), pg_interp(a.y, b.y, t)); }
/* Point::length: the f32 form, falling back to f64 when the square overflows */
static inline float pg_length(pg_pt a)
{
    float m2 = a.x * a.x + a.y * a.y;
    if (isfinite(m2)) return sqrtf(m2);
    double xx = a.x, yy = a.y;
    return (float)sqrt(xx * xx + yy * yy);
}
static inline float pg_distance(pg_pt a, pg_pt b) { return pg_length(pg_sub(a, b)); }
/* Point::set_length_from (SkPoint set_point_length): the scale is formed in f64 */
static inline int pg_set_length_from(pg_pt *p, float x, float y, float length)
{
    double xx = x, yy = y;
    double dmag = sqrt(xx * xx + yy * yy);
    double dscale = (double)length / dmag;
#ifdef PG_SETLEN_F64
    x = (float)((double)x * dscale);
    y = (float)((double)y * dscale);
#else
    x *= (float)dscale;
    y *= (float)dscale;
#endif
    if (!isfinite(x) || !isfinite(y) || (x == 0.0f && y == 0.0f)) { *p = pg_p(0, 0); return 0; }
    *p = pg_p(x, y);
    return 1;
}
static inline int pg_set_length(pg_pt *p, float length) { return pg_set_length_from(p, p->x, p->y, length); }
static inline int pg_normalize(pg_pt *p) { return pg_set_length_from(p, p->x, p->y, 1.0f); }
static inline pg_pt pg_rot_cw(pg_pt a) { return pg_p(-a.y, a.x); }   /* Point::rotate_cw */
static inline pg_pt pg_rot_ccw(pg_pt a) { return pg_p(a.y, -a.x); }  /* Point::rotate_ccw */
static inline int pg_can_normalize(pg_pt v) { return isfinite(v.x) && isfinite(v.y) && (v.x != 0.0f || v.y != 0.0f); }
static inline int pg_eq_within(pg_pt a, pg_pt b, float tol) { return fabsf(a.x - b.x) <= tol && fabsf(a.y - b.y) <= tol; }

/* ---- PathBuilder (path_builder.rs) ---- */
typedef struct {
    uint8_t *verbs;
    pg_pt *pts;
    int nv, np, cv, cp;
    int last_move_to_index;
    int move_to_required;
} pg_path;

static inline void pg_path_init(pg_path *b) { memset(b, 0, sizeof(*b)); b->move_to_required = 1; }
static inline void pg_path_free(pg_path *b) { free(b->verbs); free(b->pts); memset(b, 0, sizeof(*b)); }
static inline void pg_path_clear(pg_path *b) { b->nv = b->np = 0; b->last_move_to_index = 0; b->move_to_required = 1; }
static inline int pg_path_empty(const pg_path *b) { return b->nv == 0; }
static inline void pg_push_verb(pg_path *b, uint8_t v)
{
    if (b->nv == b->cv) { b->cv = b->cv ? b->cv * 2 : 64; b->verbs = (uint8_t *)realloc(b->verbs, (size_t)b->cv); }
    b->verbs[b->nv++] = v;
}
static inline void pg_push_pt(pg_path *b, pg_pt p)
{
    if (b->np == b->cp) { b->cp = b->cp ? b->cp * 2 : 64; b->pts = (pg_pt *)realloc(b->pts, sizeof(pg_pt) * (size_t)b->cp); }
    b->pts[b->np++] = p;
}
static inline void pg_move_to(pg_path *b, float x, float y)
{
    if (b->nv && b->verbs[b->nv - 1] == PG_MOVE) { b->pts[b->np - 1] = pg_p(x, y); return; }
    b->last_move_to_index = b->np;
    b->move_to_required = 0;
    pg_push_verb(b, PG_MOVE);
    pg_push_pt(b, pg_p(x, y));
}
static inline void pg_inject_move(pg_path *b)
{
    if (!b->move_to_required) return;
    if (b->last_move_to_index < b->np) pg_move_to(b, b->pts[b->last_move_to_index].x, b->pts[b->last_move_to_index].y);
    else pg_move_to(b, 0.0f, 0.0f);
}
static inline void pg_line_to(pg_path *b, float x, float y) { pg_inject_move(b); pg_push_verb(b, PG_LINE); pg_push_pt(b, pg_p(x, y)); }
static inline void pg_quad_to(pg_path *b, float x1, float y1, float x, float y)
{
    pg_inject_move(b);
    pg_push_verb(b, PG_QUAD);
    pg_push_pt(b, pg_p(x1, y1));
    pg_push_pt(b, pg_p(x, y));
}
static inline void pg_cubic_to(pg_path *b, float x1, float y1, float x2, float y2, float x, float y)
{
    pg_inject_move(b);
    pg_push_verb(b, PG_CUBIC);
    pg_push_pt(b, pg_p(x1, y1));
    pg_push_pt(b, pg_p(x2, y2));
    pg_push_pt(b, pg_p(x, y));
}
static inline void pg_close(pg_path *b)
{
    if (b->nv && b->verbs[b->nv - 1] != PG_CLOSE) pg_push_verb(b, PG_CLOSE);
    b->move_to_required = 1;
}
static inline int pg_last_pt(const pg_path *b, pg_pt *out)
{
    if (!b->np) return 0;
    *out = b->pts[b->np - 1];
    return 1;
}
static inline void pg_set_last_pt(pg_path *b, pg_pt p)
{
    if (b->np) b->pts[b->np - 1] = p;
    else pg_move_to(b, p.x, p.y);
}

/* ---- quadratic / cubic evaluation (path_geometry.rs) ---- */
static inline pg_pt pg_eval_quad(const pg_pt q[3], float t)
{
    /* QuadCoeff: (A t + B) t + C with A = p2 - 2 p1 + p0, B = 2 (p1 - p0) */
    pg_pt b1 = pg_sub(q[1], q[0]);
    pg_pt a = pg_sub(pg_sub(q[2], pg_add(q[1], q[1])), pg_neg(q[0])); /* p2 - 2p1 + p0 */
    pg_pt b = pg_add(b1, b1);
    return pg_p((a.x * t + b.x) * t + q[0].x, (a.y * t + b.y) * t + q[0].y);
}
static inline pg_pt pg_eval_quad_tangent(const pg_pt q[3], float t)
{
    if ((t == 0.0f && pg_eq(q[0], q[1])) || (t == 1.0f && pg_eq(q[1], q[2]))) return pg_sub(q[2], q[0]);
    pg_pt b = pg_sub(q[1], q[0]);
    pg_pt a = pg_sub(pg_sub(q[2], q[1]), b);
    pg_pt tt = pg_p(a.x * t + b.x, a.y * t + b.y);
    return pg_add(tt, tt);
}
static inline pg_pt pg_eval_cubic(const pg_pt c[4], float t)
{
    /* CubicCoeff: ((A t + B) t + C) t + D, A = p3 + 3 (p1 - p2) - p0, B = 3 (p2 - 2 p1 + p0), C = 3 (p1 - p0) */
    pg_pt d12 = pg_sub(c[1], c[2]);
    pg_pt a = pg_sub(pg_add(c[3], pg_scale(d12, 3.0f)), c[0]);
    pg_pt b = pg_scale(pg_add(pg_sub(c[2], pg_add(c[1], c[1])), c[0]), 3.0f);
    pg_pt cc = pg_scale(pg_sub(c[1], c[0]), 3.0f);
    return pg_p(((a.x * t + b.x) * t + cc.x) * t + c[0].x, ((a.y * t + b.y) * t + cc.y) * t + c[0].y);
}
static inline pg_pt pg_eval_cubic_derivative(const pg_pt c[4], float t)
{
    pg_pt a = pg_sub(pg_add(c[3], pg_scale(pg_sub(c[1], c[2]), 3.0f)), c[0]);
    pg_pt b0 = pg_add(pg_sub(c[2], pg_add(c[1], c[1])), c[0]);
    pg_pt b = pg_add(b0, b0);
    pg_pt cc = pg_sub(c[1], c[0]);
    return pg_p((a.x * t + b.x) * t + cc.x, (a.y * t + b.y) * t + cc.y);
}
static inline pg_pt pg_eval_cubic_tangent(const pg_pt c[4], float t)
{
    if ((t == 0.0f && pg_eq(c[0], c[1])) || (t == 1.0f && pg_eq(c[2], c[3]))) {
        pg_pt tan = t == 0.0f ? pg_sub(c[2], c[0]) : pg_sub(c[3], c[1]);
        if (tan.x == 0.0f && tan.y == 0.0f) tan = pg_sub(c[3], c[0]);
        return tan;
    }
    return pg_eval_cubic_derivative(c, t);
}
static inline void pg_chop_quad_at(const pg_pt s[3], float t, pg_pt d[5])
{
    pg_pt p01 = pg_lerp(s[0], s[1], t), p12 = pg_lerp(s[1], s[2], t);
    d[0] = s[0]; d[1] = p01; d[2] = pg_lerp(p01, p12, t); d[3] = p12; d[4] = s[2];
}
static inline void pg_chop_cubic_at(const pg_pt s[4], float t, pg_pt d[7])
{
    pg_pt ab = pg_lerp(s[0], s[1], t), bc = pg_lerp(s[1], s[2], t), cd = pg_lerp(s[2], s[3], t);
    pg_pt abc = pg_lerp(ab, bc, t), bcd = pg_lerp(bc, cd, t);
    d[0] = s[0]; d[1] = ab; d[2] = abc; d[3] = pg_lerp(abc, bcd, t); d[4] = bcd; d[5] = cd; d[6] = s[3];
}

/* valid_unit_divide / find_unit_quad_roots: roots strictly inside (0, 1), ascending, duplicates merged */
static inline int pg_valid_unit_divide(float numer, float denom, float *ratio)
{
    if (numer < 0.0f) { numer = -numer; denom = -denom; }
    if (denom == 0.0f || numer == 0.0f || numer >= denom) return 0;
    float r = numer / denom;
    if (r != r || r == 0.0f) return 0;
    *ratio = r;
    return 1;
}
static inline int pg_find_unit_quad_roots(float a, float b, float c, float roots[2])
{
    if (a == 0.0f) return pg_valid_unit_divide(-c, b, roots);
    double dr = (double)b * (double)b - 4.0 * (double)a * (double)c;
    if (dr < 0.0) return 0;
    dr = sqrt(dr);
    float r = (float)dr;
    if (!isfinite(r)) return 0;
    float q = b < 0.0f ? -(b - r) / 2.0f : -(b + r) / 2.0f;
    int n = 0;
    n += pg_valid_unit_divide(q, a, roots + n);
    n += pg_valid_unit_divide(c, q, roots + n);
    if (n == 2) {
        if (roots[0] > roots[1]) { float t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
        else if (roots[0] == roots[1]) n = 1;
    }
    return n;
}

/* ---- conics: only produced by round joins / caps / the cusp circle, always lowered to quads at once ---- */
typedef struct { pg_pt p[3]; float w; } pg_conic;

static inline void pg_conic_chop(const pg_conic *c, pg_conic d[2])
{
    float scale = 1.0f / (1.0f + c->w);
    pg_pt wp1 = pg_scale(c->p[1], c->w);
    pg_pt m = pg_p((c->p[0].x + (wp1.x + wp1.x) + c->p[2].x) * scale * 0.5f, (c->p[0].y + (wp1.y + wp1.y) + c->p[2].y) * scale * 0.5f);
    if (!pg_finite(m)) {
        double wd = c->w, w2 = wd * 2.0, sh = 1.0 / (1.0 + wd) * 0.5;
        m.x = (float)(((double)c->p[0].x + w2 * (double)c->p[1].x + (double)c->p[2].x) * sh);
        m.y = (float)(((double)c->p[0].y + w2 * (double)c->p[1].y + (double)c->p[2].y) * sh);
    }
    d[0].p[0] = c->p[0];
    d[0].p[1] = pg_scale(pg_add(c->p[0], wp1), scale);
    d[0].p[2] = d[1].p[0] = m;
    d[1].p[1] = pg_scale(pg_add(wp1, c->p[2]), scale);
    d[1].p[2] = c->p[2];
    d[0].w = d[1].w = sqrtf(0.5f + c->w * 0.5f);
}
static inline int pg_between(float a, float b, float c) { return (a - b) * (c - b) <= 0.0f; }
static inline pg_pt *pg_conic_subdivide(const pg_conic *src, pg_pt *pts, int level)
{
    if (level == 0) { pts[0] = src->p[1]; pts[1] = src->p[2]; return pts + 2; }
    pg_conic d[2];
    pg_conic_chop(src, d);
    const float start_y = src->p[0].y, end_y = src->p[2].y;
    if (pg_between(start_y, src->p[1].y, end_y)) {
        /* keep the chopped halves monotone in y when the input was */
        float mid_y = d[0].p[2].y;
        if (!pg_between(start_y, mid_y, end_y)) {
            float closer = fabsf(mid_y - start_y) < fabsf(mid_y - end_y) ? start_y : end_y;
            d[0].p[2].y = d[1].p[0].y = closer;
        }
        if (!pg_between(start_y, d[0].p[1].y, d[0].p[2].y)) d[0].p[1].y = start_y;
        if (!pg_between(d[1].p[0].y, d[1].p[1].y, end_y)) d[1].p[1].y = end_y;
    }
    pts = pg_conic_subdivide(&d[0], pts, level - 1);
    return pg_conic_subdivide(&d[1], pts, level - 1);
}
#define PG_MAX_CONIC_POW2 4
/* PathBuilder::conic_points_to via AutoConicToQuads (tolerance 0.25) */
static inline void pg_conic_to(pg_path *b, pg_pt p1, pg_pt p2, float w)
{
    if (!(w > 0.0f)) { pg_line_to(b, p2.x, p2.y); return; }
    if (!isfinite(w)) { pg_line_to(b, p1.x, p1.y); pg_line_to(b, p2.x, p2.y); return; }
    if (w == 1.0f) { pg_quad_to(b, p1.x, p1.y, p2.x, p2.y); return; }
    pg_inject_move(b);
    pg_conic c;
    c.p[0] = b->pts[b->np - 1]; c.p[1] = p1; c.p[2] = p2; c.w = w;
    /* Conic::compute_quad_pow2(0.25) */
    if (!pg_finite(c.p[0]) || !pg_finite(c.p[1]) || !pg_finite(c.p[2])) return;
    float a = w - 1.0f, k = a / (4.0f * (2.0f + a));
    float x = k * (c.p[0].x - 2.0f * c.p[1].x + c.p[2].x), y = k * (c.p[0].y - 2.0f * c.p[1].y + c.p[2].y);
    float error = sqrtf(x * x + y * y);
    int pow2 = 0;
    for (; pow2 < PG_MAX_CONIC_POW2; pow2++) {
        if (error <= 0.25f) break;
        error *= 0.25f;
    }
    if (pow2 < 1) pow2 = 1; /* tiny-skia: "at least one subdivision" — compute_quad_pow2 returns max(pow2, 1) */
    pg_pt pts[1 + 2 * (1 << PG_MAX_CONIC_POW2) + 2];
    pts[0] = c.p[0];
    int done = 0;
    if (pow2 == PG_MAX_CONIC_POW2) {
        pg_conic d[2];
        pg_conic_chop(&c, d);
        if (pg_eq_within(d[0].p[1], d[0].p[2], PG_NEARLY_ZERO) && pg_eq_within(d[1].p[0], d[1].p[1], PG_NEARLY_ZERO)) {
            pts[1] = pts[2] = pts[3] = d[0].p[1];
            pts[4] = d[1].p[2];
            pow2 = 1;
            done = 1;
        }
    }
    if (!done) pg_conic_subdivide(&c, pts + 1, pow2);
    const int quads = 1 << pow2, npts = 2 * quads + 1;
    int finite = 1;
    for (int i = 0; i < npts; i++) finite = finite && pg_finite(pts[i]);
    if (!finite) for (int i = 1; i < npts - 1; i++) pts[i] = c.p[1];
    for (int i = 0; i < quads; i++) pg_quad_to(b, pts[1 + 2 * i].x, pts[1 + 2 * i].y, pts[2 + 2 * i].x, pts[2 + 2 * i].y);
}

#endif
