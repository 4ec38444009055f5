/*
 * stroke.c — CPU oracle for tiny_skia_path::Path::stroke (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * resvg reaches it through PixmapMut::stroke_path (crates/resvg/src/path.rs:113, Stroke::to_tiny_skia
 * crates/usvg/src/tree/mod.rs:638-664) and usvg through Path::calculate_stroke_bbox (tree/mod.rs:1443-1456).
 * tiny-skia-path 0.12.0 (Cargo.lock:669-670) is not under /root/reference: this restates the published algorithm of its
 * stroker.rs — the Rust port of Skia's SkStroke.cpp / SkStrokerPriv.cpp (offset curves approximated by quads, checked by
 * perpendicular rays; bevel / round / miter / miter-clip joins; butt / round / square caps) — as one sequential pass.
 * Written independently of resvg_b200/csrc/stroker.cpp; tests compare the two outlines and their renders.
 * Pinned by the reference's golden PNGs: every corpus file strokes a frame, painting/stroke-* exercise the rest.
 */
#include "pathgeom.h"

typedef enum { CAP_BUTT = 0, CAP_ROUND = 1, CAP_SQUARE = 2 } cap_t;
typedef enum { JOIN_MITER = 0, JOIN_MITER_CLIP = 1, JOIN_ROUND = 2, JOIN_BEVEL = 3 } join_t;
typedef enum { RED_POINT, RED_LINE, RED_QUAD, RED_DEGENERATE, RED_DEGENERATE2, RED_DEGENERATE3 } reduction_t;
typedef enum { RES_SPLIT, RES_DEGENERATE, RES_QUAD } result_t;
typedef enum { ANGLE_NEARLY_180, ANGLE_SHARP, ANGLE_SHALLOW, ANGLE_NEARLY_LINE } angle_t;

/* the quad under construction that should run parallel to a stretch [start_t, end_t] of the source curve */
typedef struct {
    pg_pt quad[3];
    pg_pt tangent_start, tangent_end;
    float start_t, mid_t, end_t;
    int start_set, end_set, opposite_tangents;
} quad_construct;

typedef struct {
    float radius, inv_miter_limit, res_scale, inv_res_scale, inv_res_scale_sq;
    pg_pt first_normal, prev_normal, first_unit_normal, prev_unit_normal;
    pg_pt first_pt, prev_pt, first_outer_pt;
    int first_outer_pt_index;
    int segment_count;
    int prev_is_line;
    cap_t cap;
    join_t join;
    pg_path inner, outer, cusper;
    int stroke_type; /* +1 outer, -1 inner */
    int recursion_depth, found_tangents, join_completed;
} stroker;

/* ---- joins (stroker.rs: bevel_joiner, round_joiner, miter_joiner*) ---- */
static int is_clockwise(pg_pt before, pg_pt after) { return before.x * after.y > before.y * after.x; }

static angle_t dot_to_angle_type(float dot)
{
    if (dot >= 0.0f) return pg_nearly_zero(1.0f - dot) ? ANGLE_NEARLY_LINE : ANGLE_SHALLOW;
    return pg_nearly_zero(1.0f + dot) ? ANGLE_NEARLY_180 : ANGLE_SHARP;
}

static void handle_inner_join(pg_pt pivot, pg_pt after, pg_path *inner)
{
    /* through the pivot, so a radius larger than the segments does not show as a diagonal */
    pg_line_to(inner, pivot.x, pivot.y);
    pg_line_to(inner, pivot.x - after.x, pivot.y - after.y);
}

static void join_bevel(pg_pt before_un, pg_pt pivot, pg_pt after_un, float radius, pg_path *inner, pg_path *outer)
{
    pg_pt after = pg_scale(after_un, radius);
    if (!is_clockwise(before_un, after_un)) { pg_path *t = inner; inner = outer; outer = t; after = pg_neg(after); }
    pg_line_to(outer, pivot.x + after.x, pivot.y + after.y);
    handle_inner_join(pivot, after, inner);
}

/* 2x3 affine in tiny-skia's field order, with its concat / map rules (transform.rs) */
typedef struct { float sx, ky, kx, sy, tx, ty; } xf_t;
static float mul_add_mul(float a, float b, float c, float d) { return (float)((double)a * (double)b + (double)c * (double)d); }
static int xf_is_identity(xf_t t) { return t.sx == 1 && t.ky == 0 && t.kx == 0 && t.sy == 1 && t.tx == 0 && t.ty == 0; }
static int xf_has_skew(xf_t t) { return t.kx != 0 || t.ky != 0; }
static xf_t xf_concat(xf_t a, xf_t b) /* b first */
{
    if (xf_is_identity(a)) return b;
    if (xf_is_identity(b)) return a;
    xf_t r;
    if (!xf_has_skew(a) && !xf_has_skew(b)) {
        r.sx = a.sx * b.sx; r.ky = 0; r.kx = 0; r.sy = a.sy * b.sy; r.tx = a.sx * b.tx + a.tx; r.ty = a.sy * b.ty + a.ty;
    } else {
        r.sx = mul_add_mul(a.sx, b.sx, a.kx, b.ky); r.ky = mul_add_mul(a.ky, b.sx, a.sy, b.ky);
        r.kx = mul_add_mul(a.sx, b.kx, a.kx, b.sy); r.sy = mul_add_mul(a.ky, b.kx, a.sy, b.sy);
        r.tx = mul_add_mul(a.sx, b.tx, a.kx, b.ty) + a.tx; r.ty = mul_add_mul(a.ky, b.tx, a.sy, b.ty) + a.ty;
    }
    return r;
}
static pg_pt xf_map(xf_t t, pg_pt p)
{
    if (xf_is_identity(t)) return p;
    if (!xf_has_skew(t)) {
        if (t.sx == 1 && t.sy == 1) return pg_p(p.x + t.tx, p.y + t.ty);
        return pg_p(p.x * t.sx + t.tx, p.y * t.sy + t.ty);
    }
    return pg_p(p.x * t.sx + p.y * t.kx + t.tx, p.x * t.ky + p.y * t.sy + t.ty);
}

/* Conic::build_unit_arc: the arc from unit vector u_start to u_stop as at most 5 conics, mapped by `user` */
static int build_unit_arc(pg_pt u_start, pg_pt u_stop, int ccw, xf_t user, pg_conic dst[5])
{
    float x = pg_dot(u_start, u_stop);
    float y = pg_cross(u_start, u_stop);
    const float abs_y = fabsf(y);
    /* coincident vectors: angle nearly 0 (x > 0) — nothing to draw */
    if (abs_y <= PG_NEARLY_ZERO && x > 0.0f && ((y >= 0.0f && !ccw) || (y <= 0.0f && ccw))) return 0;
    if (ccw) y = -y;
    int quadrant = 0;
    if (y == 0.0f) quadrant = 2; /* 180 degrees */
    else if (x == 0.0f) quadrant = y > 0.0f ? 1 : 3;
    else {
        if (y < 0.0f) quadrant += 2;
        if ((x < 0.0f) != (y < 0.0f)) quadrant += 1;
    }
    static const pg_pt quadrant_pts[8] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
    int n = quadrant;
    for (int i = 0; i < n; i++) {
        dst[i].p[0] = quadrant_pts[i * 2];
        dst[i].p[1] = quadrant_pts[i * 2 + 1];
        dst[i].p[2] = quadrant_pts[(i * 2 + 2) % 8];
        dst[i].w = PG_ROOT2_OVER_2;
    }
    /* the remaining sub-90-degree arc */
    const pg_pt final_p = pg_p(x, y);
    const pg_pt last_q = quadrant_pts[quadrant * 2];
    const float dot = pg_dot(last_q, final_p);
    if (dot < 1.0f) {
        pg_pt off = pg_p(last_q.x + x, last_q.y + y);
        /* bisector rescaled to the off-curve point: length = 1 / cos(theta / 2), which is also the weight */
        const float cos_half = sqrtf((1.0f + dot) / 2.0f);
        pg_set_length(&off, 1.0f / cos_half);
        if (!pg_eq_within(last_q, off, PG_NEARLY_ZERO)) {
            dst[n].p[0] = last_q; dst[n].p[1] = off; dst[n].p[2] = final_p; dst[n].w = cos_half;
            n++;
        }
    }
    /* rotate by u_start, mirror for counter-clockwise, then the caller's matrix */
    xf_t m = {u_start.x, u_start.y, -u_start.y, u_start.x, 0, 0}; /* Transform::from_sin_cos(sin = y, cos = x) */
    if (ccw) { xf_t s = {1, 0, 0, -1, 0, 0}; m = xf_concat(m, s); } /* pre_scale(1, -1) */
    m = xf_concat(user, m);                                          /* post_concat(user) */
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) dst[i].p[k] = xf_map(m, dst[i].p[k]);
    return n;
}

static void join_round(pg_pt before_un, pg_pt pivot, pg_pt after_un, float radius, pg_path *inner, pg_path *outer)
{
    const float dot = pg_dot(before_un, after_un);
    if (dot_to_angle_type(dot) == ANGLE_NEARLY_LINE) return;
    pg_pt before = before_un, after = after_un;
    int ccw = 0;
    if (!is_clockwise(before, after)) {
        pg_path *t = inner; inner = outer; outer = t;
        before = pg_neg(before); after = pg_neg(after);
        ccw = 1;
    }
    xf_t ts = {radius, 0, 0, radius, pivot.x, pivot.y};
    pg_conic conics[5];
    int n = build_unit_arc(before, after, ccw, ts, conics);
    if (n > 0) {
        for (int i = 0; i < n; i++) pg_conic_to(outer, conics[i].p[1], conics[i].p[2], conics[i].w);
        handle_inner_join(pivot, pg_scale(after, radius), inner);
    }
}

static void join_miter(pg_pt before_un, pg_pt pivot, pg_pt after_un, float radius, float inv_miter_limit, int miter_clip,
                       int prev_is_line, int curr_is_line, pg_path *inner, pg_path *outer)
{
    const float dot = pg_dot(before_un, after_un);
    const angle_t angle = dot_to_angle_type(dot);
    pg_pt before = before_un, after = after_un, mid;
    int blunt = 0, ccw = 0;
    if (angle == ANGLE_NEARLY_LINE) return;
    if (angle == ANGLE_NEARLY_180) {
        curr_is_line = 0;
        mid = pg_scale(pg_sub(after, before), radius / 2.0f);
        blunt = 1;
    } else {
        ccw = !is_clockwise(before, after);
        if (ccw) {
            pg_path *t = inner; inner = outer; outer = t;
            before = pg_neg(before); after = pg_neg(after);
        }
        /* an upright right angle (rectangles) needs no square roots */
        if (dot == 0.0f && inv_miter_limit <= PG_ROOT2_OVER_2) {
            mid = pg_scale(pg_add(before, after), radius);
        } else {
            /* the dot product is built from normals, hence 1 + dot: sin of half the angle between the tangents */
            const float sin_half = sqrtf((1.0f + dot) * 0.5f);
            if (angle == ANGLE_SHARP) {
                mid = pg_p(after.y - before.y, before.x - after.x);
                if (ccw) mid = pg_neg(mid);
            } else {
                mid = pg_p(before.x + after.x, before.y + after.y);
            }
            if (sin_half < inv_miter_limit) { curr_is_line = 0; blunt = 1; }
            else pg_set_length(&mid, radius / sin_half);
        }
    }
    const pg_pt after_r = pg_scale(after, radius);
    if (blunt) {
        if (miter_clip) {
            /* cut the miter by the line perpendicular to its axis at miter_limit * radius from the pivot */
            pg_normalize(&mid);
            const float cos_beta = pg_dot(before, mid);
            const float sin_beta = pg_cross(before, mid);
            const float x = fabsf(sin_beta) <= PG_NEARLY_ZERO ? 1.0f / inv_miter_limit : ((1.0f / inv_miter_limit) - cos_beta) / sin_beta;
            const pg_pt before_r = pg_scale(before, radius);
            const pg_pt before_tangent = pg_rot_cw(before_r), after_tangent = pg_rot_ccw(after_r);
            const pg_pt c1 = pg_add(pg_add(pivot, before_r), pg_scale(before_tangent, x));
            const pg_pt c2 = pg_add(pg_add(pivot, after_r), pg_scale(after_tangent, x));
            if (prev_is_line) pg_set_last_pt(outer, c1);
            else pg_line_to(outer, c1.x, c1.y);
            pg_line_to(outer, c2.x, c2.y);
        }
    } else {
        if (prev_is_line) pg_set_last_pt(outer, pg_p(pivot.x + mid.x, pivot.y + mid.y));
        else pg_line_to(outer, pivot.x + mid.x, pivot.y + mid.y);
    }
    if (!curr_is_line) pg_line_to(outer, pivot.x + after_r.x, pivot.y + after_r.y);
    handle_inner_join(pivot, after_r, inner);
}

static void do_join(stroker *s, join_t join, pg_pt before_un, pg_pt pivot, pg_pt after_un, int prev_is_line, int curr_is_line)
{
    switch (join) {
    case JOIN_BEVEL: join_bevel(before_un, pivot, after_un, s->radius, &s->inner, &s->outer); break;
    case JOIN_ROUND: join_round(before_un, pivot, after_un, s->radius, &s->inner, &s->outer); break;
    case JOIN_MITER: join_miter(before_un, pivot, after_un, s->radius, s->inv_miter_limit, 0, prev_is_line, curr_is_line, &s->inner, &s->outer); break;
    default: join_miter(before_un, pivot, after_un, s->radius, s->inv_miter_limit, 1, prev_is_line, curr_is_line, &s->inner, &s->outer);
    }
}

/* ---- caps ---- */
static void do_cap(cap_t cap, pg_pt pivot, pg_pt normal, pg_pt stop, int other_is_line, pg_path *path)
{
    const pg_pt parallel = pg_rot_cw(normal);
    switch (cap) {
    case CAP_BUTT: pg_line_to(path, stop.x, stop.y); break;
    case CAP_ROUND: {
        const pg_pt c = pg_add(pivot, parallel);
        pg_conic_to(path, pg_add(c, normal), c, PG_ROOT2_OVER_2);
        pg_conic_to(path, pg_sub(c, normal), stop, PG_ROOT2_OVER_2);
        break;
    }
    default:
        if (other_is_line) {
            pg_set_last_pt(path, pg_p(pivot.x + normal.x + parallel.x, pivot.y + normal.y + parallel.y));
            pg_line_to(path, pivot.x - normal.x + parallel.x, pivot.y - normal.y + parallel.y);
        } else {
            pg_line_to(path, pivot.x + normal.x + parallel.x, pivot.y + normal.y + parallel.y);
            pg_line_to(path, pivot.x - normal.x + parallel.x, pivot.y - normal.y + parallel.y);
            pg_line_to(path, stop.x, stop.y);
        }
    }
}

/* ---- normals ---- */
static int set_normal_unit_normal(pg_pt before, pg_pt after, float scale, float radius, pg_pt *normal, pg_pt *unit_normal)
{
    if (!pg_set_length_from(unit_normal, (after.x - before.x) * scale, (after.y - before.y) * scale, 1.0f)) return 0;
    *unit_normal = pg_rot_ccw(*unit_normal);
    *normal = pg_scale(*unit_normal, radius);
    return 1;
}
static int set_normal_unit_normal2(pg_pt vec, float radius, pg_pt *normal, pg_pt *unit_normal)
{
    if (!pg_set_length_from(unit_normal, vec.x, vec.y, 1.0f)) return 0;
    *unit_normal = pg_rot_ccw(*unit_normal);
    *normal = pg_scale(*unit_normal, radius);
    return 1;
}

/* ---- contour bookkeeping ---- */
static void reverse_path_to(pg_path *dst, const pg_path *other)
{
    if (pg_path_empty(other)) return;
    int pi = other->np - 1;
    for (int vi = other->nv - 1; vi >= 0; vi--) {
        const uint8_t v = other->verbs[vi];
        if (v == PG_MOVE) break; /* only the last contour */
        if (v == PG_LINE) { pg_pt p = other->pts[pi - 1]; pi -= 1; pg_line_to(dst, p.x, p.y); }
        else if (v == PG_QUAD) { pg_pt a = other->pts[pi - 1], b = other->pts[pi - 2]; pi -= 2; pg_quad_to(dst, a.x, a.y, b.x, b.y); }
        else if (v == PG_CUBIC) {
            pg_pt a = other->pts[pi - 1], b = other->pts[pi - 2], c = other->pts[pi - 3];
            pi -= 3;
            pg_cubic_to(dst, a.x, a.y, b.x, b.y, c.x, c.y);
        }
    }
}

static void push_path(pg_path *dst, const pg_path *other)
{
    if (pg_path_empty(other)) return;
    if (dst->last_move_to_index != 0) dst->last_move_to_index = dst->np + other->last_move_to_index;
    for (int i = 0; i < other->nv; i++) pg_push_verb(dst, other->verbs[i]);
    for (int i = 0; i < other->np; i++) pg_push_pt(dst, other->pts[i]);
}

static void finish_contour(stroker *s, int close, int curr_is_line)
{
    if (s->segment_count > 0) {
        pg_pt pt = pg_p(0, 0);
        if (close) {
            do_join(s, s->join, s->prev_unit_normal, s->prev_pt, s->first_unit_normal, s->prev_is_line, curr_is_line);
            pg_close(&s->outer);
            /* the inner side becomes a contour of its own, reversed */
            pg_last_pt(&s->inner, &pt);
            pg_move_to(&s->outer, pt.x, pt.y);
            reverse_path_to(&s->outer, &s->inner);
            pg_close(&s->outer);
        } else {
            pg_last_pt(&s->inner, &pt);
            do_cap(s->cap, s->prev_pt, s->prev_normal, pt, curr_is_line, &s->outer);          /* cap the end */
            reverse_path_to(&s->outer, &s->inner);
            do_cap(s->cap, s->first_pt, pg_neg(s->first_normal), s->first_outer_pt, s->prev_is_line, &s->outer); /* and the start */
            pg_close(&s->outer);
        }
        if (!pg_path_empty(&s->cusper)) {
            push_path(&s->outer, &s->cusper);
            pg_path_clear(&s->cusper);
        }
    }
    pg_path_clear(&s->inner);
    s->segment_count = -1;
    s->first_outer_pt_index = s->outer.np;
}

static int pre_join_to(stroker *s, pg_pt p, int curr_is_line, pg_pt *normal, pg_pt *unit_normal)
{
    const float prev_x = s->prev_pt.x, prev_y = s->prev_pt.y;
    if (!set_normal_unit_normal(s->prev_pt, p, s->res_scale, s->radius, normal, unit_normal)) {
        if (s->cap == CAP_BUTT) return 0;
        /* square and round caps draw on a zero-length segment, upright by convention */
        *normal = pg_p(s->radius, 0.0f);
        *unit_normal = pg_p(1.0f, 0.0f);
    }
    if (s->segment_count == 0) {
        s->first_normal = *normal;
        s->first_unit_normal = *unit_normal;
        s->first_outer_pt = pg_p(prev_x + normal->x, prev_y + normal->y);
        pg_move_to(&s->outer, s->first_outer_pt.x, s->first_outer_pt.y);
        pg_move_to(&s->inner, prev_x - normal->x, prev_y - normal->y);
    } else {
        do_join(s, s->join, s->prev_unit_normal, s->prev_pt, *unit_normal, s->prev_is_line, curr_is_line);
    }
    s->prev_is_line = curr_is_line;
    return 1;
}

static void post_join_to(stroker *s, pg_pt p, pg_pt normal, pg_pt unit_normal)
{
    s->join_completed = 1;
    s->prev_pt = p;
    s->prev_unit_normal = unit_normal;
    s->prev_normal = normal;
    s->segment_count += 1;
}

/* ---- segment iterator with auto-close (path.rs PathSegmentsIter) ---- */
typedef struct {
    const uint8_t *verbs; const pg_pt *pts; int nv, np;
    int vi, pi;
    pg_pt last_move, last_pt;
    int pending_close; /* the closing line has been emitted, Close comes next */
} seg_iter;
typedef struct { int kind; pg_pt p[3]; } segment;

static int seg_next(seg_iter *it, segment *out)
{
    if (it->pending_close) { it->pending_close = 0; out->kind = PG_CLOSE; it->last_pt = it->last_move; return 1; }
    if (it->vi >= it->nv) return 0;
    const uint8_t v = it->verbs[it->vi++];
    out->kind = v;
    switch (v) {
    case PG_MOVE: out->p[0] = it->pts[it->pi++]; it->last_move = it->last_pt = out->p[0]; break;
    case PG_LINE: out->p[0] = it->pts[it->pi++]; it->last_pt = out->p[0]; break;
    case PG_QUAD: out->p[0] = it->pts[it->pi++]; out->p[1] = it->pts[it->pi++]; it->last_pt = out->p[1]; break;
    case PG_CUBIC: out->p[0] = it->pts[it->pi++]; out->p[1] = it->pts[it->pi++]; out->p[2] = it->pts[it->pi++]; it->last_pt = out->p[2]; break;
    default:
        /* auto close: a contour that does not end where it began gets the closing line first */
        if (!pg_eq(it->last_pt, it->last_move)) {
            out->kind = PG_LINE;
            out->p[0] = it->last_move;
            it->last_pt = it->last_move;
            it->pending_close = 1;
        } else {
            it->last_pt = it->last_move;
        }
    }
    return 1;
}

/* PathSegmentsIter::has_valid_tangent: does anything after this point of the contour have a direction? */
static int has_valid_tangent(const seg_iter *src)
{
    seg_iter it = *src;
    segment sg;
    for (;;) {
        const pg_pt prev = it.last_pt;
        if (!seg_next(&it, &sg)) return 0;
        switch (sg.kind) {
        case PG_MOVE: return 0;
        case PG_LINE: if (pg_eq(prev, sg.p[0])) continue; return 1;
        case PG_QUAD: if (pg_eq(prev, sg.p[0]) && pg_eq(prev, sg.p[1])) continue; return 1;
        case PG_CUBIC: if (pg_eq(prev, sg.p[0]) && pg_eq(prev, sg.p[1]) && pg_eq(prev, sg.p[2])) continue; return 1;
        default: return 0;
        }
    }
}

/* ---- lines ---- */
static void stroke_line_to(stroker *s, pg_pt p, const seg_iter *it)
{
    const int teeny = pg_eq_within(s->prev_pt, p, PG_NEARLY_ZERO * s->inv_res_scale);
    if (s->cap == CAP_BUTT && teeny) return;
    if (teeny && (s->join_completed || (it && has_valid_tangent(it)))) return;
    pg_pt normal = pg_p(0, 0), unit_normal = pg_p(0, 0);
    if (!pre_join_to(s, p, 1, &normal, &unit_normal)) return;
    pg_line_to(&s->outer, p.x + normal.x, p.y + normal.y);
    pg_line_to(&s->inner, p.x - normal.x, p.y - normal.y);
    post_join_to(s, p, normal, unit_normal);
}

/* ---- the ray machinery shared by quads and cubics ---- */
static void qc_init(quad_construct *q, float start, float end)
{
    q->start_t = start;
    q->mid_t = (start + end) * 0.5f;
    q->end_t = end;
    q->start_set = q->end_set = 0;
}
static int qc_valid(const quad_construct *q) { return q->start_t < q->mid_t && q->mid_t < q->end_t; }
static int qc_init_with_start(quad_construct *q, const quad_construct *parent)
{
    qc_init(q, parent->start_t, parent->mid_t);
    if (!qc_valid(q)) return 0;
    q->quad[0] = parent->quad[0];
    q->tangent_start = parent->tangent_start;
    q->start_set = 1;
    return 1;
}
static int qc_init_with_end(quad_construct *q, const quad_construct *parent)
{
    qc_init(q, parent->mid_t, parent->end_t);
    if (!qc_valid(q)) return 0;
    q->quad[2] = parent->quad[2];
    q->tangent_end = parent->tangent_end;
    q->end_set = 1;
    return 1;
}

static void stroker_init_side(stroker *s, int type, quad_construct *q, float t0, float t1)
{
    s->stroke_type = type;
    s->found_tangents = 0;
    qc_init(q, t0, t1);
}

/* the point on the offset curve perpendicular to the source curve at tp, and a point along the tangent there */
static void set_ray_points(const stroker *s, pg_pt tp, pg_pt *dxy, pg_pt *on_p, pg_pt *tangent)
{
    if (!pg_set_length(dxy, s->radius)) *dxy = pg_p(s->radius, 0.0f);
    const float axis_flip = (float)s->stroke_type;
    on_p->x = tp.x + axis_flip * dxy->y;
    on_p->y = tp.y - axis_flip * dxy->x;
    if (tangent) { tangent->x = on_p->x + dxy->x; tangent->y = on_p->y + dxy->y; }
}

static float pt_to_line(pg_pt pt, pg_pt line_start, pg_pt line_end)
{
    const pg_pt dxy = pg_sub(line_end, line_start), ab0 = pg_sub(pt, line_start);
    const float numer = pg_dot(dxy, ab0), denom = pg_dot(dxy, dxy);
    const float t = numer / denom;
    if (t >= 0.0f && t <= 1.0f) {
        const pg_pt hit = pg_p(line_start.x * (1.0f - t) + line_end.x * t, line_start.y * (1.0f - t) + line_end.y * t);
        return pg_dist_sqd(hit, pt);
    }
    return pg_dist_sqd(pt, line_start);
}

static int points_within_dist(pg_pt a, pg_pt b, float limit) { return pg_dist_sqd(a, b) <= limit * limit; }

static int sharp_angle(const pg_pt quad[3])
{
    pg_pt smaller = pg_sub(quad[1], quad[0]), larger = pg_sub(quad[1], quad[2]);
    const float smaller_len = pg_len_sqd(smaller);
    float larger_len = pg_len_sqd(larger);
    if (smaller_len > larger_len) { pg_pt t = smaller; smaller = larger; larger = t; larger_len = smaller_len; }
    if (!pg_set_length(&smaller, larger_len)) return 0;
    return pg_dot(smaller, larger) > 0.0f;
}

static int pt_in_quad_bounds(const stroker *s, const pg_pt q[3], pg_pt pt)
{
    const float e = s->inv_res_scale;
    if (pt.x + e < fminf(fminf(q[0].x, q[1].x), q[2].x)) return 0;
    if (pt.x - e > fmaxf(fmaxf(q[0].x, q[1].x), q[2].x)) return 0;
    if (pt.y + e < fminf(fminf(q[0].y, q[1].y), q[2].y)) return 0;
    if (pt.y - e > fmaxf(fmaxf(q[0].y, q[1].y), q[2].y)) return 0;
    return 1;
}

static int intersect_quad_ray(const pg_pt line[2], const pg_pt quad[3], float roots[2])
{
    const pg_pt vec = pg_sub(line[1], line[0]);
    float r[3];
    for (int n = 0; n < 3; n++) r[n] = (quad[n].y - line[0].y) * vec.x - (quad[n].x - line[0].x) * vec.y;
    float a = r[2], b = r[1];
    const float c = r[0];
    a += c - 2.0f * b;
    b -= c;
    return pg_find_unit_quad_roots(a, 2.0f * b, c, roots);
}

/* where do the end tangents of the construct meet?  with_ctrl: also place the quad's control point there */
static result_t intersect_ray(const stroker *s, quad_construct *q, int with_ctrl)
{
    const pg_pt start = q->quad[0], end = q->quad[2];
    const pg_pt a_len = pg_sub(q->tangent_start, start), b_len = pg_sub(q->tangent_end, end);
    const float denom = pg_cross(a_len, b_len);
    if (denom == 0.0f || !isfinite(denom)) {
        q->opposite_tangents = pg_dot(a_len, b_len) < 0.0f;
        return RES_DEGENERATE;
    }
    q->opposite_tangents = 0;
    const pg_pt ab0 = pg_sub(start, end);
    float numer_a = pg_cross(b_len, ab0);
    const float numer_b = pg_cross(a_len, ab0);
    if ((numer_a >= 0.0f) == (numer_b >= 0.0f)) {
        /* the control point would lie outside the ends: flat enough for a line, or split */
        const float dist1 = pt_to_line(start, end, q->tangent_end);
        const float dist2 = pt_to_line(end, start, q->tangent_start);
        if (fmaxf(dist1, dist2) <= s->inv_res_scale_sq) return RES_DEGENERATE;
        return RES_SPLIT;
    }
    numer_a /= denom;
    if (numer_a > numer_a - 1.0f) { /* the divide kept its precision */
        if (with_ctrl) {
            q->quad[1].x = start.x * (1.0f - numer_a) + q->tangent_start.x * numer_a;
            q->quad[1].y = start.y * (1.0f - numer_a) + q->tangent_start.y * numer_a;
        }
        return RES_QUAD;
    }
    q->opposite_tangents = pg_dot(a_len, b_len) < 0.0f;
    return RES_DEGENERATE; /* parallel tangents: a line will do */
}

static result_t stroke_close_enough(const stroker *s, const pg_pt stroke[3], const pg_pt ray[2], quad_construct *q)
{
    const pg_pt stroke_mid = pg_eval_quad(stroke, 0.5f);
    if (points_within_dist(ray[0], stroke_mid, s->inv_res_scale)) return sharp_angle(q->quad) ? RES_SPLIT : RES_QUAD;
    if (!pt_in_quad_bounds(s, stroke, ray[0])) return RES_SPLIT;
    float roots[2];
    if (intersect_quad_ray(ray, stroke, roots) != 1) return RES_SPLIT;
    const pg_pt quad_pt = pg_eval_quad(stroke, roots[0]);
    const float error = s->inv_res_scale * (1.0f - fabsf(roots[0] - 0.5f) * 2.0f);
    if (points_within_dist(ray[0], quad_pt, error)) return sharp_angle(q->quad) ? RES_SPLIT : RES_QUAD;
    return RES_SPLIT;
}

static void add_degenerate_line(stroker *s, const quad_construct *q)
{
    pg_path *path = s->stroke_type == 1 ? &s->outer : &s->inner;
    pg_line_to(path, q->quad[2].x, q->quad[2].y);
}
static void emit_quad(stroker *s, const quad_construct *q)
{
    pg_path *path = s->stroke_type == 1 ? &s->outer : &s->inner;
    pg_quad_to(path, q->quad[1].x, q->quad[1].y, q->quad[2].x, q->quad[2].y);
}

/* ---- quads ---- */
static float find_quad_max_curvature(const pg_pt src[3])
{
    const float ax = src[1].x - src[0].x, ay = src[1].y - src[0].y;
    const float bx = src[0].x - src[1].x - src[1].x + src[2].x, by = src[0].y - src[1].y - src[1].y + src[2].y;
    float numer = -(ax * bx + ay * by), denom = bx * bx + by * by;
    if (denom < 0.0f) { numer = -numer; denom = -denom; }
    if (numer <= 0.0f) return 0.0f;
    if (numer >= denom) return 1.0f;
    return numer / denom;
}

static int quad_in_line(const pg_pt quad[3])
{
    float pt_max = -1.0f;
    int outer1 = 0, outer2 = 0;
    for (int index = 0; index < 2; index++)
        for (int inner = index + 1; inner < 3; inner++) {
            const pg_pt d = pg_sub(quad[inner], quad[index]);
            const float m = fmaxf(fabsf(d.x), fabsf(d.y));
            if (pt_max < m) { outer1 = index; outer2 = inner; pt_max = m; }
        }
    const int mid = outer1 ^ outer2 ^ 3;
    const float line_slop = pt_max * pt_max * 0.000005f; /* "pulled out of the air" */
    return pt_to_line(quad[mid], quad[outer1], quad[outer2]) <= line_slop;
}

static reduction_t check_quad_linear(const pg_pt quad[3], pg_pt *reduction)
{
    const int deg_ab = !pg_can_normalize(pg_sub(quad[1], quad[0])), deg_bc = !pg_can_normalize(pg_sub(quad[2], quad[1]));
    if (deg_ab & deg_bc) return RED_POINT;
    if (deg_ab | deg_bc) return RED_LINE;
    if (!quad_in_line(quad)) return RED_QUAD;
    const float t = find_quad_max_curvature(quad);
    if (t == 0.0f || t == 1.0f) return RED_LINE;
    *reduction = pg_eval_quad(quad, t);
    return RED_DEGENERATE;
}

static void quad_perp_ray(const stroker *s, const pg_pt quad[3], float t, pg_pt *tp, pg_pt *on_p, pg_pt *tangent)
{
    *tp = pg_eval_quad(quad, t);
    pg_pt dxy = pg_eval_quad_tangent(quad, t);
    if (dxy.x == 0.0f && dxy.y == 0.0f) dxy = pg_sub(quad[2], quad[0]);
    set_ray_points(s, *tp, &dxy, on_p, tangent);
}

static result_t compare_quad_quad(stroker *s, const pg_pt quad[3], quad_construct *q)
{
    pg_pt tmp;
    if (!q->start_set) { quad_perp_ray(s, quad, q->start_t, &tmp, &q->quad[0], &q->tangent_start); q->start_set = 1; }
    if (!q->end_set) { quad_perp_ray(s, quad, q->end_t, &tmp, &q->quad[2], &q->tangent_end); q->end_set = 1; }
    const result_t r = intersect_ray(s, q, 1);
    if (r != RES_QUAD) return r;
    pg_pt ray[2];
    quad_perp_ray(s, quad, q->mid_t, &ray[1], &ray[0], NULL);
    return stroke_close_enough(s, q->quad, ray, q);
}

static int quad_stroke(stroker *s, const pg_pt quad[3], quad_construct *q)
{
    const result_t r = compare_quad_quad(s, quad, q);
    if (r == RES_QUAD) { emit_quad(s, q); return 1; }
    if (r == RES_DEGENERATE) { add_degenerate_line(s, q); return 1; }
    if (++s->recursion_depth > 11 * 3) return 0; /* RECURSIVE_LIMITS[quad] */
    quad_construct half;
    qc_init_with_start(&half, q);
    if (!quad_stroke(s, quad, &half)) return 0;
    qc_init_with_end(&half, q);
    if (!quad_stroke(s, quad, &half)) return 0;
    s->recursion_depth--;
    return 1;
}

static void stroke_quad_to(stroker *s, pg_pt p1, pg_pt p2)
{
    const pg_pt quad[3] = {s->prev_pt, p1, p2};
    pg_pt reduction = pg_p(0, 0);
    const reduction_t rt = check_quad_linear(quad, &reduction);
    if (rt == RED_POINT || rt == RED_LINE) { stroke_line_to(s, p2, NULL); return; }
    if (rt == RED_DEGENERATE) {
        /* a quad folded onto a line: out to the turning point and back, with a round join there */
        stroke_line_to(s, reduction, NULL);
        const join_t saved = s->join;
        s->join = JOIN_ROUND;
        stroke_line_to(s, p2, NULL);
        s->join = saved;
        return;
    }
    pg_pt normal_ab, unit_ab, normal_bc, unit_bc;
    if (!pre_join_to(s, p1, 0, &normal_ab, &unit_ab)) { stroke_line_to(s, p2, NULL); return; }
    quad_construct q;
    stroker_init_side(s, 1, &q, 0.0f, 1.0f);
    quad_stroke(s, quad, &q);
    stroker_init_side(s, -1, &q, 0.0f, 1.0f);
    quad_stroke(s, quad, &q);
    if (!set_normal_unit_normal(quad[1], quad[2], s->res_scale, s->radius, &normal_bc, &unit_bc)) { normal_bc = normal_ab; unit_bc = unit_ab; }
    post_join_to(s, p2, normal_bc, unit_bc);
}

/* ---- cubics ---- */
static int cubic_in_line(const pg_pt cubic[4])
{
    float pt_max = -1.0f;
    int outer1 = 0, outer2 = 0;
    for (int index = 0; index < 3; index++)
        for (int inner = index + 1; inner < 4; inner++) {
            const pg_pt d = pg_sub(cubic[inner], cubic[index]);
            const float m = fmaxf(fabsf(d.x), fabsf(d.y));
            if (pt_max < m) { outer1 = index; outer2 = inner; pt_max = m; }
        }
    const int mid1 = (1 + (2 >> outer2)) >> outer1;
    const int mid2 = outer1 ^ outer2 ^ mid1;
    const float line_slop = pt_max * pt_max * 0.00001f;
    return pt_to_line(cubic[mid1], cubic[outer1], cubic[outer2]) <= line_slop && pt_to_line(cubic[mid2], cubic[outer1], cubic[outer2]) <= line_slop;
}

static float pin01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); } /* NaN stays out: callers test 0 < t < 1 */

static int solve_cubic_poly(const float coeff[4], float t[3])
{
    if (pg_nearly_zero(coeff[0])) return pg_find_unit_quad_roots(coeff[1], coeff[2], coeff[3], t);
    const float inva = 1.0f / coeff[0];
    const float a = coeff[1] * inva, b = coeff[2] * inva, c = coeff[3] * inva;
    const float q = (a * a - b * 3.0f) / 9.0f;
    const float r = (2.0f * a * a * a - 9.0f * a * b + 27.0f * c) / 54.0f;
    const float q3 = q * q * q;
    const float r2_minus_q3 = r * r - q3;
    const float adiv3 = a / 3.0f;
    if (r2_minus_q3 < 0.0f) { /* three real roots */
        float cosv = r / sqrtf(q3);
        cosv = cosv < -1.0f ? -1.0f : (cosv > 1.0f ? 1.0f : cosv);
        const float theta = acosf(cosv);
        const float neg2_root_q = -2.0f * sqrtf(q);
        const float pi = 3.14159265f;
        t[0] = pin01(neg2_root_q * cosf(theta / 3.0f) - adiv3);
        t[1] = pin01(neg2_root_q * cosf((theta + 2.0f * pi) / 3.0f) - adiv3);
        t[2] = pin01(neg2_root_q * cosf((theta - 2.0f * pi) / 3.0f) - adiv3);
        /* sort, then collapse duplicates */
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2 - i; j++) if (t[j] > t[j + 1]) { float x = t[j]; t[j] = t[j + 1]; t[j + 1] = x; }
        int n = 3;
        if (t[1] == t[2]) n = 2;
        if (t[0] == t[1]) { t[1] = t[2]; n--; }
        return n;
    }
    float aa = fabsf(r) + sqrtf(r2_minus_q3);
    aa = cbrtf(aa);
    if (r > 0.0f) aa = -aa;
    if (aa != 0.0f) aa += q / aa;
    t[0] = pin01(aa - adiv3);
    return 1;
}

static void formulate_f1_dot_f2(const float s0, const float s1, const float s2, const float s3, float coeff[4])
{
    const float a = s1 - s0, b = s2 - 2.0f * s1 + s0, c = s3 + 3.0f * (s1 - s2) - s0;
    coeff[0] = c * c;
    coeff[1] = 3.0f * b * c;
    coeff[2] = 2.0f * b * b + c * a;
    coeff[3] = a * b;
}

static int find_cubic_max_curvature(const pg_pt src[4], float t[3])
{
    float cx[4], cy[4];
    formulate_f1_dot_f2(src[0].x, src[1].x, src[2].x, src[3].x, cx);
    formulate_f1_dot_f2(src[0].y, src[1].y, src[2].y, src[3].y, cy);
    for (int i = 0; i < 4; i++) cx[i] += cy[i];
    return solve_cubic_poly(cx, t);
}

static int find_cubic_inflections(const pg_pt src[4], float t[2])
{
    const float ax = src[1].x - src[0].x, ay = src[1].y - src[0].y;
    const float bx = src[2].x - 2.0f * src[1].x + src[0].x, by = src[2].y - 2.0f * src[1].y + src[0].y;
    const float cx = src[3].x + 3.0f * (src[1].x - src[2].x) - src[0].x, cy = src[3].y + 3.0f * (src[1].y - src[2].y) - src[0].y;
    return pg_find_unit_quad_roots(bx * cy - by * cx, ax * cy - ay * cx, ax * by - ay * bx, t);
}

static int on_same_side(const pg_pt src[4], int test_index, int line_index)
{
    const pg_pt origin = src[line_index];
    const pg_pt line = pg_sub(src[line_index + 1], origin);
    float crosses[2];
    for (int i = 0; i < 2; i++) crosses[i] = pg_cross(line, pg_sub(src[test_index + i], origin));
    return crosses[0] * crosses[1] >= 0.0f;
}

static float find_cubic_cusp(const pg_pt src[4])
{
    if (pg_eq(src[0], src[1]) || pg_eq(src[2], src[3])) return -1.0f;
    /* a cusp needs the two control legs to cross */
    if (on_same_side(src, 0, 2) || on_same_side(src, 2, 0)) return -1.0f;
    float mc[3];
    const int roots = find_cubic_max_curvature(src, mc);
    for (int i = 0; i < roots; i++) {
        const float t = mc[i];
        if (0.0f >= t || t >= 1.0f) continue;
        const pg_pt d = pg_eval_cubic_derivative(src, t);
        const float precision = (pg_dist_sqd(src[1], src[0]) + pg_dist_sqd(src[2], src[1]) + pg_dist_sqd(src[3], src[2])) * 1e-8f;
        if (pg_len_sqd(d) < precision) return t;
    }
    return -1.0f;
}

static reduction_t check_cubic_linear(const pg_pt cubic[4], pg_pt reduction[3], const pg_pt **tangent_pt)
{
    const int deg_ab = !pg_can_normalize(pg_sub(cubic[1], cubic[0]));
    const int deg_bc = !pg_can_normalize(pg_sub(cubic[2], cubic[1]));
    const int deg_cd = !pg_can_normalize(pg_sub(cubic[3], cubic[2]));
    if (deg_ab & deg_bc & deg_cd) return RED_POINT;
    if (deg_ab + deg_bc + deg_cd == 2) return RED_LINE;
    if (!cubic_in_line(cubic)) {
        *tangent_pt = deg_ab ? &cubic[2] : &cubic[1];
        return RED_QUAD;
    }
    float tv[3];
    const int count = find_cubic_max_curvature(cubic, tv);
    int r_count = 0;
    for (int i = 0; i < count; i++) {
        const float t = tv[i];
        if (0.0f >= t || t >= 1.0f) continue;
        reduction[r_count] = pg_eval_cubic(cubic, t);
        if (!pg_eq(reduction[r_count], cubic[0]) && !pg_eq(reduction[r_count], cubic[3])) r_count++;
    }
    if (r_count == 0) return RED_LINE;
    return (reduction_t)(RED_QUAD + r_count);
}

static void cubic_perp_ray(const stroker *s, const pg_pt cubic[4], float t, pg_pt *tp, pg_pt *on_p, pg_pt *tangent)
{
    *tp = pg_eval_cubic(cubic, t);
    pg_pt dxy = pg_eval_cubic_tangent(cubic, t);
    pg_pt chopped[7];
    if (dxy.x == 0.0f && dxy.y == 0.0f) {
        const pg_pt *c = cubic;
        if (pg_nearly_zero(t)) dxy = pg_sub(cubic[2], cubic[0]);
        else if (pg_nearly_zero(1.0f - t)) dxy = pg_sub(cubic[3], cubic[1]);
        else {
            /* the inflection sits on a cusp: take the tangent from the subdivided halves */
            pg_chop_cubic_at(cubic, t, chopped);
            dxy = pg_sub(chopped[3], chopped[2]);
            if (dxy.x == 0.0f && dxy.y == 0.0f) { dxy = pg_sub(chopped[3], chopped[1]); c = chopped; }
        }
        if (dxy.x == 0.0f && dxy.y == 0.0f) dxy = pg_sub(c[3], c[0]);
    }
    set_ray_points(s, *tp, &dxy, on_p, tangent);
}

static void cubic_quad_ends(const stroker *s, const pg_pt cubic[4], quad_construct *q)
{
    pg_pt tmp;
    if (!q->start_set) { cubic_perp_ray(s, cubic, q->start_t, &tmp, &q->quad[0], &q->tangent_start); q->start_set = 1; }
    if (!q->end_set) { cubic_perp_ray(s, cubic, q->end_t, &tmp, &q->quad[2], &q->tangent_end); q->end_set = 1; }
}

static int cubic_mid_on_line(const stroker *s, const pg_pt cubic[4], const quad_construct *q)
{
    pg_pt mid_pt, stroke_mid;
    cubic_perp_ray(s, cubic, q->mid_t, &mid_pt, &stroke_mid, NULL);
    return pt_to_line(stroke_mid, q->quad[0], q->quad[2]) < s->inv_res_scale_sq;
}

static result_t compare_quad_cubic(stroker *s, const pg_pt cubic[4], quad_construct *q)
{
    cubic_quad_ends(s, cubic, q);
    const result_t r = intersect_ray(s, q, 1);
    if (r != RES_QUAD) return r;
    pg_pt ray[2];
    cubic_perp_ray(s, cubic, q->mid_t, &ray[1], &ray[0], NULL);
    return stroke_close_enough(s, q->quad, ray, q);
}

static int cubic_stroke(stroker *s, const pg_pt cubic[4], quad_construct *q)
{
    if (!s->found_tangents) {
        cubic_quad_ends(s, cubic, q);
        const result_t r = intersect_ray(s, q, 0); /* tangents_meet */
        if (r != RES_QUAD) {
            if ((r == RES_DEGENERATE || points_within_dist(q->quad[0], q->quad[2], s->inv_res_scale)) && cubic_mid_on_line(s, cubic, q)) {
                add_degenerate_line(s, q);
                return 1;
            }
        } else {
            s->found_tangents = 1;
        }
    }
    if (s->found_tangents) {
        const result_t r = compare_quad_cubic(s, cubic, q);
        if (r == RES_QUAD) { emit_quad(s, q); return 1; }
        if (r == RES_DEGENERATE && !q->opposite_tangents) { add_degenerate_line(s, q); return 1; }
    }
    if (!isfinite(q->quad[2].x) || !isfinite(q->quad[2].y)) return 0; /* not representable */
    if (++s->recursion_depth > (s->found_tangents ? 26 * 3 : 5 * 3)) return 0; /* RECURSIVE_LIMITS[cubic / tangent] */
    quad_construct half;
    if (!qc_init_with_start(&half, q)) { add_degenerate_line(s, q); s->recursion_depth--; return 1; }
    if (!cubic_stroke(s, cubic, &half)) return 0;
    if (!qc_init_with_end(&half, q)) { add_degenerate_line(s, q); s->recursion_depth--; return 1; }
    if (!cubic_stroke(s, cubic, &half)) return 0;
    s->recursion_depth--;
    return 1;
}

/* PathBuilder::push_circle -> push_oval: four quarter conics, closed */
static void push_circle(pg_path *b, float x, float y, float r)
{
    const float left = x - r, top = y - r, right = (x - r) + (r + r), bottom = (y - r) + (r + r); /* Rect::from_xywh(x - r, y - r, r + r, r + r) */
    if (!(isfinite(left) && isfinite(top) && isfinite(right) && isfinite(bottom)) || !(left <= right && top <= bottom)) return;
    const float cx = left * 0.5f + right * 0.5f, cy = top * 0.5f + bottom * 0.5f;
    const pg_pt oval[4] = {{cx, bottom}, {left, cy}, {cx, top}, {right, cy}};
    const pg_pt rect[4] = {{right, bottom}, {left, bottom}, {left, top}, {right, top}};
    pg_move_to(b, oval[3].x, oval[3].y);
    for (int i = 0; i < 4; i++) pg_conic_to(b, rect[i], oval[i], PG_ROOT2_OVER_2);
    pg_close(b);
}

static void stroke_cubic_to(stroker *s, pg_pt p1, pg_pt p2, pg_pt p3)
{
    const pg_pt cubic[4] = {s->prev_pt, p1, p2, p3};
    pg_pt reduction[3];
    const pg_pt *tangent_pt = &cubic[1];
    const reduction_t rt = check_cubic_linear(cubic, reduction, &tangent_pt);
    if (rt == RED_POINT || rt == RED_LINE) { stroke_line_to(s, p3, NULL); return; }
    if (rt >= RED_DEGENERATE) {
        stroke_line_to(s, reduction[0], NULL);
        const join_t saved = s->join;
        s->join = JOIN_ROUND;
        if (rt >= RED_DEGENERATE2) stroke_line_to(s, reduction[1], NULL);
        if (rt == RED_DEGENERATE3) stroke_line_to(s, reduction[2], NULL);
        stroke_line_to(s, p3, NULL);
        s->join = saved;
        return;
    }
    pg_pt normal_ab, unit_ab, normal_cd, unit_cd;
    if (!pre_join_to(s, *tangent_pt, 0, &normal_ab, &unit_ab)) { stroke_line_to(s, p3, NULL); return; }
    float tv[2];
    const int count = find_cubic_inflections(cubic, tv);
    float last_t = 0.0f;
    for (int i = 0; i <= count; i++) {
        const float next_t = i < count ? tv[i] : 1.0f;
        quad_construct q;
        stroker_init_side(s, 1, &q, last_t, next_t);
        cubic_stroke(s, cubic, &q);
        stroker_init_side(s, -1, &q, last_t, next_t);
        cubic_stroke(s, cubic, &q);
        last_t = next_t;
    }
    const float cusp = find_cubic_cusp(cubic);
    if (cusp > 0.0f) {
        const pg_pt loc = pg_eval_cubic(cubic, cusp);
        push_circle(&s->cusper, loc.x, loc.y, s->radius);
    }
    /* set_cubic_end_normal: the join is emitted even if one side gave up */
    {
        pg_pt ab = pg_sub(cubic[1], cubic[0]), cd = pg_sub(cubic[3], cubic[2]);
        int deg_ab = !pg_can_normalize(ab), deg_cd = !pg_can_normalize(cd);
        int degenerate = deg_ab && deg_cd;
        if (!degenerate) {
            if (deg_ab) { ab = pg_sub(cubic[2], cubic[0]); deg_ab = !pg_can_normalize(ab); }
            if (deg_cd) { cd = pg_sub(cubic[3], cubic[1]); deg_cd = !pg_can_normalize(cd); }
            degenerate = deg_ab || deg_cd;
        }
        if (degenerate || !set_normal_unit_normal2(cd, s->radius, &normal_cd, &unit_cd)) { normal_cd = normal_ab; unit_cd = unit_ab; }
    }
    post_join_to(s, p3, normal_cd, unit_cd);
}

/* ---- driver (PathStroker::stroke_inner) ---- */
static int is_zero_length_since(const pg_path *b, int start)
{
    const int count = b->np - start;
    if (count < 2) return 1;
    for (int i = 1; i < count; i++) if (!pg_eq(b->pts[start], b->pts[start + i])) return 0;
    return 1;
}

/* Returns 1 and the outline, or 0 for the reference's None.  Outputs are malloc'ed (orc_geom_free). */
int orc_path_stroke(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, float width, float miter_limit,
                    int32_t cap, int32_t join, float res_scale, uint8_t **out_verbs, int32_t *out_n_verbs, float **out_points,
                    int32_t *out_n_points)
{
    if (!(width > 0.0f) || !isfinite(width)) return 0; /* NonZeroPositiveF32::new(stroke.width)? */
    stroker st;
    memset(&st, 0, sizeof(st));
    stroker *s = &st;
    join_t line_join = (join_t)join;
    float inv_miter_limit = 0.0f;
    if (line_join == JOIN_MITER) {
        if (miter_limit <= 1.0f) line_join = JOIN_BEVEL;
        else inv_miter_limit = 1.0f / miter_limit;
    }
    if (line_join == JOIN_MITER_CLIP) inv_miter_limit = 1.0f / miter_limit;
    s->res_scale = res_scale;
    s->inv_res_scale = 1.0f / (res_scale * 4.0f); /* the 4 matches the fill scan converter's error term */
    s->inv_res_scale_sq = s->inv_res_scale * s->inv_res_scale;
    s->radius = width * 0.5f;
    s->inv_miter_limit = inv_miter_limit;
    s->segment_count = -1;
    s->cap = (cap_t)cap;
    s->join = line_join;
    s->stroke_type = 1;
    pg_path_init(&s->inner);
    pg_path_init(&s->outer);
    pg_path_init(&s->cusper);

    seg_iter it;
    memset(&it, 0, sizeof(it));
    it.verbs = verbs; it.pts = (const pg_pt *)points; it.nv = n_verbs; it.np = n_points;
    int last_is_line = 0;
    segment sg;
    while (seg_next(&it, &sg)) {
        switch (sg.kind) {
        case PG_MOVE:
            if (s->segment_count > 0) finish_contour(s, 0, 0);
            s->segment_count = 0;
            s->first_pt = s->prev_pt = sg.p[0];
            s->join_completed = 0;
            break;
        case PG_LINE: stroke_line_to(s, sg.p[0], &it); last_is_line = 1; break;
        case PG_QUAD: stroke_quad_to(s, sg.p[0], sg.p[1]); last_is_line = 0; break;
        case PG_CUBIC: stroke_cubic_to(s, sg.p[0], sg.p[1], sg.p[2]); last_is_line = 0; break;
        default:
            if (s->cap != CAP_BUTT) {
                /* move + close, or move + zero-length verbs + close: a dot that still gets its caps */
                if (s->segment_count == 0) { stroke_line_to(s, s->first_pt, NULL); last_is_line = 1; continue; }
                if (is_zero_length_since(&s->inner, 0) && is_zero_length_since(&s->outer, s->first_outer_pt_index)) { last_is_line = 1; continue; }
            }
            finish_contour(s, 1, last_is_line);
        }
    }
    finish_contour(s, 0, last_is_line);

    int ok = 0;
    pg_path *o = &s->outer;
    if (o->nv > 1) { /* PathBuilder::finish: empty or a lone move is None; so are non-finite bounds */
        ok = 1;
        for (int i = 0; i < o->np; i++) if (!pg_finite(o->pts[i])) ok = 0;
    }
    if (ok) {
        *out_verbs = (uint8_t *)malloc((size_t)o->nv);
        memcpy(*out_verbs, o->verbs, (size_t)o->nv);
        *out_n_verbs = o->nv;
        *out_points = (float *)malloc(sizeof(pg_pt) * (size_t)(o->np ? o->np : 1));
        memcpy(*out_points, o->pts, sizeof(pg_pt) * (size_t)o->np);
        *out_n_points = o->np;
    }
    pg_path_free(&s->inner);
    pg_path_free(&s->outer);
    pg_path_free(&s->cusper);
    return ok;
}

void orc_geom_free(void *p) { free(p); }
