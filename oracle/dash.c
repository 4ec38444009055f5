/*
 * dash.c — CPU oracle for tiny_skia_path::Path::dash (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * PixmapMut::stroke_path dashes the path first when the stroke carries a StrokeDash (tiny-skia painter.rs; resvg
 * path.rs:113 with Stroke::to_tiny_skia, usvg tree/mod.rs:657-661).  tiny-skia-path 0.12.0 is not under /root/reference:
 * this restates the published algorithm of its dash.rs (StrokeDash::new, dash_impl, ContourMeasure — the Rust port of
 * Skia's SkDashPath.cpp and SkContourMeasure.cpp) sequentially.  Written independently of resvg_b200/csrc/dasher.cpp.
 * Pinned by the painting/stroke-dasharray, stroke-dashoffset goldens.
 */
#include "pathgeom.h"

#define MAX_T_VALUE 0x3FFFFFFFu
enum { SEG_LINE = 0, SEG_QUAD = 1, SEG_CUBIC = 2 };

typedef struct { float distance; int pt_index; uint32_t t_value; int kind; } cm_seg;

typedef struct {
    cm_seg *segs; int ns, cs;
    pg_pt *pts; int np, cp;
    float length;
    int is_closed;
    float tolerance;
} contour;

static void cm_push_seg(contour *c, float d, int pi, uint32_t t, int kind)
{
    if (c->ns == c->cs) { c->cs = c->cs ? c->cs * 2 : 64; c->segs = (cm_seg *)realloc(c->segs, sizeof(cm_seg) * (size_t)c->cs); }
    cm_seg s = {d, pi, t, kind};
    c->segs[c->ns++] = s;
}
static void cm_push_pt(contour *c, pg_pt p)
{
    if (c->np == c->cp) { c->cp = c->cp ? c->cp * 2 : 64; c->pts = (pg_pt *)realloc(c->pts, sizeof(pg_pt) * (size_t)c->cp); }
    c->pts[c->np++] = p;
}

static int tspan_big_enough(uint32_t tspan) { return (tspan >> 10) != 0; }

static int quad_too_curvy(const contour *c, pg_pt p0, pg_pt p1, pg_pt p2)
{
    /* distance between the curve's midpoint and the chord's: b/2 - a/4 - c/4 */
    const float dx = p1.x * 0.5f - (p0.x + p2.x) * 0.5f * 0.5f;
    const float dy = p1.y * 0.5f - (p0.y + p2.y) * 0.5f * 0.5f;
    return fmaxf(fabsf(dx), fabsf(dy)) > c->tolerance;
}
static int cheap_dist_exceeds(const contour *c, pg_pt pt, float x, float y)
{
    return fmaxf(fabsf(x - pt.x), fabsf(y - pt.y)) > c->tolerance;
}
static int cubic_too_curvy(const contour *c, pg_pt p0, pg_pt p1, pg_pt p2, pg_pt p3)
{
    const float third = 1.0f / 3.0f, two_thirds = 2.0f / 3.0f;
    return cheap_dist_exceeds(c, p1, pg_interp(p0.x, p3.x, third), pg_interp(p0.y, p3.y, third))
           || cheap_dist_exceeds(c, p2, pg_interp(p0.x, p3.x, two_thirds), pg_interp(p0.y, p3.y, two_thirds));
}

static float compute_line_seg(contour *c, pg_pt p0, pg_pt p1, float distance, int pt_index)
{
    const float d = pg_distance(p0, p1);
    const float prev = distance;
    distance += d;
    if (distance > prev) cm_push_seg(c, distance, pt_index, MAX_T_VALUE, SEG_LINE);
    return distance;
}
static float compute_quad_segs(contour *c, pg_pt p0, pg_pt p1, pg_pt p2, float distance, uint32_t mint, uint32_t maxt, int pt_index)
{
    if (tspan_big_enough(maxt - mint) && quad_too_curvy(c, p0, p1, p2)) {
        const pg_pt src[3] = {p0, p1, p2};
        pg_pt tmp[5];
        const uint32_t halft = (mint + maxt) >> 1;
        pg_chop_quad_at(src, 0.5f, tmp);
        distance = compute_quad_segs(c, tmp[0], tmp[1], tmp[2], distance, mint, halft, pt_index);
        distance = compute_quad_segs(c, tmp[2], tmp[3], tmp[4], distance, halft, maxt, pt_index);
    } else {
        const float d = pg_distance(p0, p2);
        const float prev = distance;
        distance += d;
        if (distance > prev) cm_push_seg(c, distance, pt_index, maxt, SEG_QUAD);
    }
    return distance;
}
static float compute_cubic_segs(contour *c, pg_pt p0, pg_pt p1, pg_pt p2, pg_pt p3, float distance, uint32_t mint, uint32_t maxt, int pt_index)
{
    if (tspan_big_enough(maxt - mint) && cubic_too_curvy(c, p0, p1, p2, p3)) {
        const pg_pt src[4] = {p0, p1, p2, p3};
        pg_pt tmp[7];
        const uint32_t halft = (mint + maxt) >> 1;
        pg_chop_cubic_at(src, 0.5f, tmp);
        distance = compute_cubic_segs(c, tmp[0], tmp[1], tmp[2], tmp[3], distance, mint, halft, pt_index);
        distance = compute_cubic_segs(c, tmp[3], tmp[4], tmp[5], tmp[6], distance, halft, maxt, pt_index);
    } else {
        const float d = pg_distance(p0, p3);
        const float prev = distance;
        distance += d;
        if (distance > prev) cm_push_seg(c, distance, pt_index, maxt, SEG_CUBIC);
    }
    return distance;
}

/* ContourMeasureIter: measures the contour starting at verb *vi / point *pi; returns 0 when the path is exhausted.
 * A contour of zero length is reported with length 0 (the caller skips it). */
static int build_contour(const uint8_t *verbs, int nv, const pg_pt *pts, int *vi, int *pi, float res_scale, contour *c)
{
    c->ns = c->np = 0;
    c->length = 0.0f;
    c->is_closed = 0;
    c->tolerance = 0.5f * (1.0f / res_scale); /* CHEAP_DIST_LIMIT * res_scale.invert() */
    if (*vi >= nv) return 0;
    int pt_index = -1;
    float distance = 0.0f;
    int have_seen_close = 0, have_seen_move = 0;
    pg_pt prev = pg_p(0, 0);
    while (*vi < nv) {
        const uint8_t v = verbs[*vi];
        if (v == PG_MOVE) {
            if (have_seen_move) break; /* the next contour */
            have_seen_move = 1;
            (*vi)++;
            pt_index += 1;
            prev = pts[(*pi)++];
            cm_push_pt(c, prev);
            continue;
        }
        (*vi)++;
        if (v == PG_LINE) {
            const pg_pt p = pts[(*pi)++];
            const float pd = distance;
            distance = compute_line_seg(c, prev, p, distance, pt_index);
            if (distance > pd) { cm_push_pt(c, p); pt_index += 1; }
            prev = p;
        } else if (v == PG_QUAD) {
            const pg_pt p1 = pts[(*pi)++], p2 = pts[(*pi)++];
            const float pd = distance;
            distance = compute_quad_segs(c, prev, p1, p2, distance, 0, MAX_T_VALUE, pt_index);
            if (distance > pd) { cm_push_pt(c, p1); cm_push_pt(c, p2); pt_index += 2; }
            prev = p2;
        } else if (v == PG_CUBIC) {
            const pg_pt p1 = pts[(*pi)++], p2 = pts[(*pi)++], p3 = pts[(*pi)++];
            const float pd = distance;
            distance = compute_cubic_segs(c, prev, p1, p2, p3, distance, 0, MAX_T_VALUE, pt_index);
            if (distance > pd) { cm_push_pt(c, p1); cm_push_pt(c, p2); cm_push_pt(c, p3); pt_index += 3; }
            prev = p3;
        } else {
            have_seen_close = 1;
        }
    }
    if (!isfinite(distance)) { c->ns = 0; c->length = 0.0f; return 1; }
    if (have_seen_close && c->np > 0) {
        const float pd = distance;
        const pg_pt first = c->pts[0];
        distance = compute_line_seg(c, c->pts[pt_index], first, distance, pt_index);
        if (distance > pd) cm_push_pt(c, first);
    }
    c->length = distance;
    c->is_closed = have_seen_close;
    return 1;
}

static float seg_scalar_t(const cm_seg *s) { return (float)s->t_value * (1.0f / (float)MAX_T_VALUE); }

/* distance -> (segment index, t inside the segment's curve) */
static int distance_to_segment(const contour *c, float distance, float *t)
{
    /* binary search for the first segment whose end distance is >= distance */
    int lo = 0, hi = c->ns - 1;
    while (lo < hi) {
        const int mid = (hi + lo) >> 1;
        if (c->segs[mid].distance < distance) lo = mid + 1;
        else hi = mid;
    }
    int index = hi;
    if (c->segs[hi].distance < distance) index = hi + 1; /* past the end: only with rounding; clamp */
    if (index >= c->ns) index = c->ns - 1;
    const cm_seg *seg = &c->segs[index];
    float start_t = 0.0f, start_d = 0.0f;
    if (index > 0) {
        start_d = c->segs[index - 1].distance;
        if (c->segs[index - 1].pt_index == seg->pt_index) start_t = seg_scalar_t(&c->segs[index - 1]);
    }
    *t = start_t + (seg_scalar_t(seg) - start_t) * (distance - start_d) / (seg->distance - start_d);
    return index;
}

static pg_pt compute_pos(const pg_pt *p, int kind, float t)
{
    if (kind == SEG_LINE) return pg_p(pg_interp(p[0].x, p[1].x, t), pg_interp(p[0].y, p[1].y, t));
    if (kind == SEG_QUAD) return pg_eval_quad(p, t);
    return pg_eval_cubic(p, t);
}

static void segment_to(const pg_pt *p, int kind, float start_t, float stop_t, pg_path *pb)
{
    if (start_t == stop_t) {
        /* a zero-length "on" interval: a zero-length line, so the stroker can still put caps on it */
        pg_pt last;
        if (pg_last_pt(pb, &last)) pg_line_to(pb, last.x, last.y);
        return;
    }
    if (kind == SEG_LINE) {
        if (stop_t == 1.0f) pg_line_to(pb, p[1].x, p[1].y);
        else pg_line_to(pb, pg_interp(p[0].x, p[1].x, stop_t), pg_interp(p[0].y, p[1].y, stop_t));
    } else if (kind == SEG_QUAD) {
        pg_pt tmp0[5], tmp1[5];
        if (start_t == 0.0f) {
            if (stop_t == 1.0f) pg_quad_to(pb, p[1].x, p[1].y, p[2].x, p[2].y);
            else { pg_chop_quad_at(p, stop_t, tmp0); pg_quad_to(pb, tmp0[1].x, tmp0[1].y, tmp0[2].x, tmp0[2].y); }
        } else {
            pg_chop_quad_at(p, start_t, tmp0);
            if (stop_t == 1.0f) pg_quad_to(pb, tmp0[3].x, tmp0[3].y, tmp0[4].x, tmp0[4].y);
            else {
                pg_chop_quad_at(tmp0 + 2, (stop_t - start_t) / (1.0f - start_t), tmp1);
                pg_quad_to(pb, tmp1[1].x, tmp1[1].y, tmp1[2].x, tmp1[2].y);
            }
        }
    } else {
        pg_pt tmp0[7], tmp1[7];
        if (start_t == 0.0f) {
            if (stop_t == 1.0f) pg_cubic_to(pb, p[1].x, p[1].y, p[2].x, p[2].y, p[3].x, p[3].y);
            else { pg_chop_cubic_at(p, stop_t, tmp0); pg_cubic_to(pb, tmp0[1].x, tmp0[1].y, tmp0[2].x, tmp0[2].y, tmp0[3].x, tmp0[3].y); }
        } else {
            pg_chop_cubic_at(p, start_t, tmp0);
            if (stop_t == 1.0f) pg_cubic_to(pb, tmp0[4].x, tmp0[4].y, tmp0[5].x, tmp0[5].y, tmp0[6].x, tmp0[6].y);
            else {
                pg_chop_cubic_at(tmp0 + 3, (stop_t - start_t) / (1.0f - start_t), tmp1);
                pg_cubic_to(pb, tmp1[1].x, tmp1[1].y, tmp1[2].x, tmp1[2].y, tmp1[3].x, tmp1[3].y);
            }
        }
    }
}

/* ContourMeasure::push_segment: the stretch [start_d, stop_d] of the contour appended to pb */
static void push_segment(const contour *c, float start_d, float stop_d, int start_with_move_to, pg_path *pb)
{
    if (start_d < 0.0f) start_d = 0.0f;
    if (stop_d > c->length) stop_d = c->length;
    if (!(start_d <= stop_d)) return; /* also catches NaN */
    if (c->ns == 0) return;
    float start_t, stop_t;
    int seg_index = distance_to_segment(c, start_d, &start_t);
    const int stop_index = distance_to_segment(c, stop_d, &stop_t);
    cm_seg seg = c->segs[seg_index];
    const cm_seg stop_seg = c->segs[stop_index];
    if (start_with_move_to) {
        const pg_pt p = compute_pos(c->pts + seg.pt_index, seg.kind, start_t);
        pg_move_to(pb, p.x, p.y);
    }
    if (seg.pt_index == stop_seg.pt_index) {
        segment_to(c->pts + seg.pt_index, seg.kind, start_t, stop_t, pb);
    } else {
        for (;;) {
            segment_to(c->pts + seg.pt_index, seg.kind, start_t, 1.0f, pb);
            const int old = seg.pt_index;
            do { seg_index++; } while (c->segs[seg_index].pt_index == old);
            seg = c->segs[seg_index];
            start_t = 0.0f;
            if (seg.pt_index >= stop_seg.pt_index) break;
        }
        segment_to(c->pts + seg.pt_index, seg.kind, 0.0f, stop_t, pb);
    }
}

/* Returns 1 and the dashed path, or 0 for the reference's None (a dash specification StrokeDash::new rejects, too many
 * dashes, nothing left).  Outputs are malloc'ed (orc_geom_free). */
int orc_path_dash(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, const float *dash_array,
                  int32_t n_dash, float dash_offset, float res_scale, uint8_t **out_verbs, int32_t *out_n_verbs,
                  float **out_points, int32_t *out_n_points)
{
    (void)n_points;
    /* StrokeDash::new */
    if (!isfinite(dash_offset)) return 0;
    if (n_dash < 2 || (n_dash % 2) != 0) return 0;
    float interval_len = 0.0f;
    for (int i = 0; i < n_dash; i++) {
        if (dash_array[i] < 0.0f) return 0;
        interval_len += dash_array[i];
    }
    if (!isfinite(interval_len) || interval_len <= 0.0f) return 0;
    /* adjust_dash_offset */
    float offset = dash_offset;
    if (offset < 0.0f) {
        offset = -offset;
        if (offset > interval_len) offset = fmodf(offset, interval_len);
        offset = interval_len - offset;
        if (offset == interval_len) offset = 0.0f; /* finite precision: len - tiny == len */
    } else if (offset >= interval_len) {
        offset = fmodf(offset, interval_len);
    }
    /* find_first_interval */
    float first_len = dash_array[0];
    int first_index = 0;
    {
        float o = offset;
        int found = 0;
        for (int i = 0; i < n_dash; i++) {
            const float gap = dash_array[i];
            if (o > gap || (o == gap && gap != 0.0f)) o -= gap;
            else { first_len = gap - o; first_index = i; found = 1; break; }
        }
        if (!found) { first_len = dash_array[0]; first_index = 0; }
    }

    pg_path pb;
    pg_path_init(&pb);
    contour c;
    memset(&c, 0, sizeof(c));
    float dash_count = 0.0f;
    int vi = 0, pi = 0, ok = 1;
    while (build_contour(verbs, n_verbs, (const pg_pt *)points, &vi, &pi, res_scale, &c)) {
        if (!(c.length > 0.0f) || c.ns == 0) continue; /* ContourMeasureIter skips zero-length contours */
        int skip_first_segment = c.is_closed;
        int added_segment = 0;
        const float length = c.length;
        int index = first_index;
        /* give up beyond a million dashes (Skia's guard against unbounded memory) */
        dash_count += length * (float)(n_dash >> 1) / interval_len;
        if (dash_count > 1000000.0f) { ok = 0; break; }
        float distance = 0.0f;
        float d_len = first_len;
        while (distance < length) {
            added_segment = 0;
            if ((index % 2) == 0 && !skip_first_segment) {
                added_segment = 1;
                push_segment(&c, distance, distance + d_len, 1, &pb);
            }
            distance += d_len;
            skip_first_segment = 0; /* only the first time around */
            index++;
            if (index == n_dash) index = 0;
            d_len = dash_array[index];
        }
        /* a closed contour that began inside an "on" interval: join the end up with the skipped start */
        if (c.is_closed && (first_index % 2) == 0 && first_len >= 0.0f) push_segment(&c, 0.0f, first_len, !added_segment, &pb);
    }
    free(c.segs);
    free(c.pts);
    if (ok && pb.nv > 1) {
        for (int i = 0; i < pb.np; i++) if (!pg_finite(pb.pts[i])) ok = 0;
    } else ok = 0;
    if (ok) {
        *out_verbs = (uint8_t *)malloc((size_t)pb.nv);
        memcpy(*out_verbs, pb.verbs, (size_t)pb.nv);
        *out_n_verbs = pb.nv;
        *out_points = (float *)malloc(sizeof(pg_pt) * (size_t)(pb.np ? pb.np : 1));
        memcpy(*out_points, pb.pts, sizeof(pg_pt) * (size_t)pb.np);
        *out_n_points = pb.np;
    }
    pg_path_free(&pb);
    return ok;
}
