// dasher_core.h — tiny_skia_path::Path::dash(&StrokeDash, res_scale): the path PixmapMut::stroke_path strokes when the
// stroke has a dash array (tiny-skia painter.rs stroke_path; usvg Stroke::to_tiny_skia tree/mod.rs:638-664).
//
// Restates tiny-skia-path 0.12.0 dash.rs (a port of SkDashPath / SkContourMeasure): every contour is measured into
// segments (lines as they are, curves subdivided until flat within 0.5 / res_scale, each segment tagged with the t
// value it ends at), and the "on" intervals of the dash pattern are cut out of it.
//
// One source for both sides (geom_common.h): dasher.cpp instantiates it with std::vector, geo.cu with DVec.
#pragma once

#include "geom_common.h"

namespace geo {
namespace ds {

GEO_HDI inline float dist(P a, P b)
{
    const float dx = a.x - b.x, dy = a.y - b.y;
    return sqrtf(dx * dx + dy * dy);
}
GEO_HDI inline float interp(float a, float b, float t) { return a + (b - a) * t; }
GEO_HDI inline P lerp(P a, P b, float t) { return P{interp(a.x, b.x, t), interp(a.y, b.y, t)}; }

// NormalizedF32Exclusive::new_bounded
GEO_HDI inline float bounded_t(float t)
{
    const float eps = 1.1920929e-7f;
    if (!(t > eps)) return eps;
    if (t > 1.0f - eps) return 1.0f - eps;
    return t;
}

GEO_HD inline void chop_quad_at(const P s[3], float t, P d[5])
{
    const P p01 = lerp(s[0], s[1], t), p12 = lerp(s[1], s[2], t);
    d[0] = s[0]; d[1] = p01; d[2] = lerp(p01, p12, t); d[3] = p12; d[4] = s[2];
}
GEO_HD inline void chop_cubic_at(const P s[4], float t, P d[7])
{
    const P ab = lerp(s[0], s[1], t), bc = lerp(s[1], s[2], t), cd = lerp(s[2], s[3], t);
    const P abc = lerp(ab, bc, t), bcd = lerp(bc, cd, t);
    d[0] = s[0]; d[1] = ab; d[2] = abc; d[3] = lerp(abc, bcd, t); d[4] = bcd; d[5] = cd; d[6] = s[3];
}
GEO_HD inline P eval_quad_at(const P s[3], float t)
{
    // QuadCoeff: (A t + B) t + C with A = p0 - 2 p1 + p2, B = 2 (p1 - p0)
    const float ax = s[2].x - 2.0f * s[1].x + s[0].x, ay = s[2].y - 2.0f * s[1].y + s[0].y;
    const float bx = 2.0f * (s[1].x - s[0].x), by = 2.0f * (s[1].y - s[0].y);
    return P{(ax * t + bx) * t + s[0].x, (ay * t + by) * t + s[0].y};
}
GEO_HD inline P eval_cubic_pos_at(const P s[4], float t)
{
    // CubicCoeff: ((A t + B) t + C) t + D
    const float ax = s[3].x + 3.0f * (s[1].x - s[2].x) - s[0].x, ay = s[3].y + 3.0f * (s[1].y - s[2].y) - s[0].y;
    const float bx = 3.0f * (s[2].x - 2.0f * s[1].x + s[0].x), by = 3.0f * (s[2].y - 2.0f * s[1].y + s[0].y);
    const float cx = 3.0f * (s[1].x - s[0].x), cy = 3.0f * (s[1].y - s[0].y);
    return P{((ax * t + bx) * t + cx) * t + s[0].x, ((ay * t + by) * t + cy) * t + s[0].y};
}

template <template <class> class Vec> struct DashOut {
    Vec<uint8_t> verbs;
    Vec<P> pts;
    bool move_required = true;
    size_t last_move = 0;
    GEO_HD void move_to(P p)
    {
        if (!verbs.empty() && verbs.back() == V_MOVE) { pts.back() = p; }
        else { last_move = pts.size(); verbs.push_back(V_MOVE); pts.push_back(p); }
        move_required = false;
    }
    GEO_HD void inject()
    {
        if (move_required) {
            if (pts.empty()) move_to(P{0, 0});
            else { P p = pts[last_move]; move_to(p); }
        }
    }
    GEO_HD void line_to(P p) { inject(); verbs.push_back(V_LINE); pts.push_back(p); }
    GEO_HD void quad_to(P a, P b) { inject(); verbs.push_back(V_QUAD); pts.push_back(a); pts.push_back(b); }
    GEO_HD void cubic_to(P a, P b, P c) { inject(); verbs.push_back(V_CUBIC); pts.push_back(a); pts.push_back(b); pts.push_back(c); }
    GEO_HDI bool last_point(P *p) const { if (pts.empty()) return false; *p = pts.back(); return true; }
};

enum Kind { KLine = 0, KQuad = 1, KCubic = 2 };
constexpr uint32_t MAX_T = 0x3FFFFFFF;
struct Seg { float distance; uint32_t point_index; uint32_t t_value; int kind; };
GEO_HDI inline float scalar_t(const Seg &s) { return (float)s.t_value * (1.0f / (float)MAX_T); }

template <template <class> class Vec> struct Contour {
    typedef DashOut<Vec> Out;
    Vec<Seg> segs;
    Vec<P> pts;
    float length = 0;
    bool closed = false;
    float tolerance = 0.5f;

    GEO_HDI static uint32_t t_span_big_enough(uint32_t span) { return span >> 10; }
    GEO_HDI bool quad_too_curvy(P a, P b, P c) const
    {
        const float dx = b.x * 0.5f - ((a.x + c.x) * 0.5f) * 0.5f, dy = b.y * 0.5f - ((a.y + c.y) * 0.5f) * 0.5f;
        return fmaxf(fabsf(dx), fabsf(dy)) > tolerance;
    }
    GEO_HDI bool cheap_exceeds(P pt, float x, float y) const { return fmaxf(fabsf(x - pt.x), fabsf(y - pt.y)) > tolerance; }
    GEO_HDI bool cubic_too_curvy(const P c[4]) const
    {
        return cheap_exceeds(c[1], interp(c[0].x, c[3].x, 1.0f / 3.0f), interp(c[0].y, c[3].y, 1.0f / 3.0f))
               || cheap_exceeds(c[2], interp(c[0].x, c[3].x, 2.0f / 3.0f), interp(c[0].y, c[3].y, 2.0f / 3.0f));
    }
    GEO_HD float quad_segs(P a, P b, P c, float distance, uint32_t mint, uint32_t maxt, uint32_t pi)
    {
        if (t_span_big_enough(maxt - mint) != 0 && quad_too_curvy(a, b, c)) {
            P tmp[5];
            const P src[3] = {a, b, c};
            const uint32_t half = (mint + maxt) >> 1;
            chop_quad_at(src, 0.5f, tmp);
            distance = quad_segs(tmp[0], tmp[1], tmp[2], distance, mint, half, pi);
            distance = quad_segs(tmp[2], tmp[3], tmp[4], distance, half, maxt, pi);
        } else {
            const float d = dist(a, c), prev = distance;
            distance += d;
            if (distance > prev) segs.push_back(Seg{distance, pi, maxt, KQuad});
        }
        return distance;
    }
    GEO_HD float cubic_segs(const P c[4], float distance, uint32_t mint, uint32_t maxt, uint32_t pi)
    {
        if (t_span_big_enough(maxt - mint) != 0 && cubic_too_curvy(c)) {
            P tmp[7];
            const uint32_t half = (mint + maxt) >> 1;
            chop_cubic_at(c, 0.5f, tmp);
            distance = cubic_segs(&tmp[0], distance, mint, half, pi);
            distance = cubic_segs(&tmp[3], distance, half, maxt, pi);
        } else {
            const float d = dist(c[0], c[3]), prev = distance;
            distance += d;
            if (distance > prev) segs.push_back(Seg{distance, pi, maxt, KCubic});
        }
        return distance;
    }

    // SkTKSearch over segment distances: index of the match, or ~index of the insertion point
    GEO_HD int find_segment(float key) const
    {
        int lo = 0, hi = (int)segs.size() - 1;
        while (lo < hi) {
            const int mid = (hi + lo) >> 1;
            if (segs[(size_t)mid].distance < key) lo = mid + 1;
            else hi = mid;
        }
        if (segs[(size_t)hi].distance < key) { hi += 1; hi = ~hi; }
        else if (key < segs[(size_t)hi].distance) hi = ~hi;
        return hi;
    }
    GEO_HD bool distance_to_segment(float distance, size_t *index, float *t) const
    {
        int i = find_segment(distance);
        i ^= i >> 31;
        if (i < 0 || (size_t)i >= segs.size()) return false;
        const Seg &seg = segs[(size_t)i];
        float start_t = 0.0f, start_d = 0.0f;
        if (i > 0) {
            start_d = segs[(size_t)i - 1].distance;
            if (segs[(size_t)i - 1].point_index == seg.point_index) start_t = scalar_t(segs[(size_t)i - 1]);
        }
        const float tv = start_t + (scalar_t(seg) - start_t) * (distance - start_d) / (seg.distance - start_d);
        if (!(tv >= 0.0f && tv <= 1.0f)) return false; // NormalizedF32::new
        *index = (size_t)i;
        *t = tv;
        return true;
    }
    GEO_HDI static void compute_pos(const P *p, int kind, float t, P *pos)
    {
        if (kind == KLine) *pos = lerp(p[0], p[1], t);
        else if (kind == KQuad) *pos = eval_quad_at(p, t);
        else *pos = eval_cubic_pos_at(p, t);
    }
    GEO_HD static void segment_to(const P *p, int kind, float start_t, float stop_t, Out &pb)
    {
        if (start_t == stop_t) {
            // a zero-length "on" interval: a zero-length line, so the stroker can still add caps
            P last;
            if (pb.last_point(&last)) pb.line_to(last);
            return;
        }
        if (kind == KLine) {
            if (stop_t == 1.0f) pb.line_to(p[1]);
            else pb.line_to(lerp(p[0], p[1], stop_t));
        } else if (kind == KQuad) {
            P t0[5], t1[5];
            if (start_t == 0.0f) {
                if (stop_t == 1.0f) pb.quad_to(p[1], p[2]);
                else { chop_quad_at(p, bounded_t(stop_t), t0); pb.quad_to(t0[1], t0[2]); }
            } else {
                chop_quad_at(p, bounded_t(start_t), t0);
                if (stop_t == 1.0f) pb.quad_to(t0[3], t0[4]);
                else {
                    const float nt = (stop_t - start_t) / (1.0f - start_t);
                    chop_quad_at(&t0[2], bounded_t(nt), t1);
                    pb.quad_to(t1[1], t1[2]);
                }
            }
        } else {
            P t0[7], t1[7];
            if (start_t == 0.0f) {
                if (stop_t == 1.0f) pb.cubic_to(p[1], p[2], p[3]);
                else { chop_cubic_at(p, bounded_t(stop_t), t0); pb.cubic_to(t0[1], t0[2], t0[3]); }
            } else {
                chop_cubic_at(p, bounded_t(start_t), t0);
                if (stop_t == 1.0f) pb.cubic_to(t0[4], t0[5], t0[6]);
                else {
                    const float nt = (stop_t - start_t) / (1.0f - start_t);
                    chop_cubic_at(&t0[3], bounded_t(nt), t1);
                    pb.cubic_to(t1[1], t1[2], t1[3]);
                }
            }
        }
    }
    GEO_HD void push_segment(float start_d, float stop_d, bool start_with_move_to, Out &pb) const
    {
        if (start_d < 0.0f) start_d = 0.0f;
        if (stop_d > length) stop_d = length;
        if (!(start_d <= stop_d)) return; // also catches NaN
        if (segs.empty()) return;
        size_t si, ei;
        float start_t, stop_t;
        if (!distance_to_segment(start_d, &si, &start_t)) return;
        if (!distance_to_segment(stop_d, &ei, &stop_t)) return;
        Seg seg = segs[si];
        const Seg stop_seg = segs[ei];
        if (start_with_move_to) {
            P p;
            compute_pos(&pts[seg.point_index], seg.kind, start_t, &p);
            pb.move_to(p);
        }
        if (seg.point_index == stop_seg.point_index) {
            segment_to(&pts[seg.point_index], seg.kind, start_t, stop_t, pb);
        } else {
            size_t ni = si;
            for (;;) {
                segment_to(&pts[seg.point_index], seg.kind, start_t, 1.0f, pb);
                const uint32_t old = seg.point_index;
                do { ni++; } while (ni < segs.size() && segs[ni].point_index == old);
                if (ni >= segs.size()) return;
                seg = segs[ni];
                start_t = 0.0f;
                if (seg.point_index >= stop_seg.point_index) break;
            }
            segment_to(&pts[seg.point_index], seg.kind, 0.0f, stop_t, pb);
        }
    }
};

// ContourMeasureIter::next: measures the contour that starts at verb index *vi; returns false at the end of the path.
template <template <class> class Vec>
GEO_HD bool next_contour(const uint8_t *verbs, int n_verbs, const P *pts, int *vi, int *pi, float tolerance, Contour<Vec> *c)
{
    while (*vi < n_verbs) {
        c->segs.clear();
        c->pts.clear();
        c->length = 0;
        c->closed = false;
        c->tolerance = tolerance;
        float distance = 0.0f;
        bool seen_close = false, seen_move = false;
        uint32_t point_index = 0;
        while (*vi < n_verbs) {
            const int v = verbs[*vi];
            if (v == V_MOVE) {
                if (seen_move) break; // the next contour starts here
                seen_move = true;
                c->pts.push_back(pts[(*pi)++]);
                (*vi)++;
            } else if (v == V_LINE) {
                const P p = pts[(*pi)++];
                (*vi)++;
                if (c->pts.empty()) { c->pts.push_back(p); continue; }
                const float prev = distance;
                distance += dist(c->pts[point_index], p);
                if (distance > prev) {
                    c->segs.push_back(Seg{distance, point_index, MAX_T, KLine});
                    c->pts.push_back(p);
                    point_index += 1;
                }
            } else if (v == V_QUAD) {
                const P a = pts[*pi], b2 = pts[*pi + 1];
                *pi += 2;
                (*vi)++;
                if (c->pts.empty()) { c->pts.push_back(b2); continue; }
                const float prev = distance;
                distance = c->quad_segs(c->pts[point_index], a, b2, distance, 0, MAX_T, point_index);
                if (distance > prev) {
                    c->pts.push_back(a);
                    c->pts.push_back(b2);
                    point_index += 2;
                }
            } else if (v == V_CUBIC) {
                const P a = pts[*pi], b2 = pts[*pi + 1], d = pts[*pi + 2];
                *pi += 3;
                (*vi)++;
                if (c->pts.empty()) { c->pts.push_back(d); continue; }
                const float prev = distance;
                const P cub[4] = {c->pts[point_index], a, b2, d};
                distance = c->cubic_segs(cub, distance, 0, MAX_T, point_index);
                if (distance > prev) {
                    c->pts.push_back(a);
                    c->pts.push_back(b2);
                    c->pts.push_back(d);
                    point_index += 3;
                }
            } else { // close
                seen_close = true;
                (*vi)++;
                break;
            }
        }
        if (!gfinite(distance)) return false;
        if (seen_close && !c->pts.empty()) {
            const float prev = distance;
            const P first = c->pts[0];
            distance += dist(c->pts[point_index], first);
            if (distance > prev) {
                c->segs.push_back(Seg{distance, point_index, MAX_T, KLine});
                c->pts.push_back(first);
            }
        }
        c->length = distance;
        c->closed = seen_close;
        if (!c->segs.empty() && distance > 0.0f) return true; // empty contours are skipped
    }
    return false;
}


// StrokeDash::new + adjust_dash_offset + find_first_interval
struct DashSpec {
    bool valid;
    float interval_len, first_len;
    int first_index;
};
GEO_HD inline DashSpec dash_spec(const float *dash_array, int n_dash, float dash_offset)
{
    DashSpec sp;
    sp.valid = false; sp.interval_len = 0.0f; sp.first_len = 0.0f; sp.first_index = 0;
    if (!gfinite(dash_offset)) return sp;
    if (n_dash < 2 || (n_dash & 1)) return sp;
    float interval_len = 0.0f;
    for (int i = 0; i < n_dash; i++) {
        if (dash_array[i] < 0.0f) return sp;
        interval_len += dash_array[i];
    }
    if (!gfinite(interval_len) || interval_len <= 0.0f) return sp;
    sp.valid = true;
    sp.interval_len = interval_len;
    // adjust_dash_offset
    float off = dash_offset;
    if (off < 0.0f) {
        off = -off;
        if (off > interval_len) off = fmodf(off, interval_len);
        off = interval_len - off;
        if (off == interval_len) off = 0.0f;
    } else if (off >= interval_len) {
        off = fmodf(off, interval_len);
    }
    // find_first_interval
    float first_len = dash_array[0];
    int first_index = 0;
    {
        bool found = false;
        for (int i = 0; i < n_dash; i++) {
            const float gap = dash_array[i];
            if (off > gap || (off == gap && gap != 0.0f)) off -= gap;
            else { first_len = gap - off; first_index = i; found = true; break; }
        }
        if (!found) { first_len = dash_array[0]; first_index = 0; }
    }
    sp.first_len = first_len;
    sp.first_index = first_index;
    return sp;
}

// The "on" intervals dash_impl cuts out of one measured contour, in order: f(start_d, stop_d, start_with_move_to).  A call
// without move_to continues the output contour of the previous call (a closed contour whose last dash runs into its first).
template <class F>
GEO_HD void dash_contour_ranges(const DashSpec &sp, const float *dash_array, int n_dash, float length, bool closed, F &f)
{
    bool skip_first = closed, added = false;
    int index = sp.first_index;
    float distance = 0.0f, d_len = sp.first_len;
    while (distance < length) {
        added = false;
        if ((index & 1) == 0 && !skip_first) {
            added = true;
            f(distance, distance + d_len, true);
        }
        distance += d_len;
        skip_first = false;
        index += 1;
        if (index == n_dash) index = 0;
        d_len = dash_array[index];
    }
    // extend if we ended on a segment and need to join up with the (skipped) initial segment
    if (closed && (sp.first_index & 1) == 0 && sp.first_len >= 0.0f) f(0.0f, sp.first_len, !added);
}

// Path::dash into pb (which must be empty); `c` is scratch.  *spec_valid = StrokeDash::new accepted the specification
// (a rejected one leaves the stroke solid).  Returns false when the specification is rejected or nothing is left.
template <template <class> class Vec>
GEO_HD bool dash_path(DashOut<Vec> &pb, Contour<Vec> &c, const uint8_t *verbs, int n_verbs, const P *pts, const float *dash_array, int n_dash,
                      float dash_offset, float res_scale, bool *spec_valid)
{
    const DashSpec sp = dash_spec(dash_array, n_dash, dash_offset);
    *spec_valid = sp.valid;
    if (!sp.valid) return false;
    const float tolerance = 0.5f * (1.0f / res_scale);
    int vi = 0, pi = 0;
    float dash_count = 0.0f;
    struct Push {
        Contour<Vec> *c;
        DashOut<Vec> *pb;
        GEO_HD void operator()(float a, float b, bool mv) { c->push_segment(a, b, mv, *pb); }
    } push{&c, &pb};
    while (next_contour(verbs, n_verbs, pts, &vi, &pi, tolerance, &c)) {
        dash_count += c.length * (float)(n_dash >> 1) / sp.interval_len;
        if (dash_count > 1000000.0f) return false;
        dash_contour_ranges(sp, dash_array, n_dash, c.length, c.closed, push);
    }
    // PathBuilder::finish: nothing but move_to's is no path
    return pb.verbs.size() > 1;
}

} // namespace ds
} // namespace geo
