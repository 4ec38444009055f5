// stroker_core.h — the path stroker: `path.stroke(&stroke, res_scale)` of tiny-skia-path 0.12.0 (stroker.rs, a port of
// Skia's SkStroke.cpp / SkStrokerPriv.cpp), which PixmapMut::stroke_path runs before filling the outline
// (crates/resvg/src/path.rs:113; Stroke fields set in crates/usvg/src/tree/mod.rs:638-664).  The reference source is not
// in /root/reference, so this restates the published algorithm: offset curves approximated by quads found by
// intersecting perpendicular rays, recursive subdivision until the approximation is within 1/(4*res_scale),
// miter / miter-clip / round / bevel joins, butt / round / square caps.
//
// One source for both sides (geom_common.h): stroker.cpp instantiates it with std::vector for the host builder and the
// rb_path_stroke export, geo.cu with DVec for the device geometry kernels.
#pragma once

#include "geom_common.h"

namespace geo {
namespace sk {

GEO_HDI inline float dot(P a, P b) { return a.x * b.x + a.y * b.y; }
GEO_HDI inline float cross(P a, P b) { return a.x * b.y - a.y * b.x; }
GEO_HDI inline float len_sqd(P a) { return dot(a, a); }
GEO_HDI inline float dist_sqd(P a, P b) { return len_sqd(a - b); }
GEO_HDI inline P rot_cw(P a) { return {-a.y, a.x}; }
GEO_HDI inline P rot_ccw(P a) { return {a.y, -a.x}; }

constexpr float kNearlyZero = 1.0f / 4096.0f;
constexpr float kRoot2Over2 = 0.707106781f;
GEO_HDI inline bool nearly_zero(float v, float tol = kNearlyZero) { return fabsf(v) <= tol; }

// Point::set_length / normalize (double precision magnitude, as tiny-skia-path point.rs)
GEO_HDI inline bool set_length(P &p, float x, float y, float length)
{
    double xx = x, yy = y;
    double dmag = sqrt(xx * xx + yy * yy);
    double dscale = (double)length / dmag;
    x *= (float)dscale; // tiny-skia-path point.rs set_point_length: `x *= dscale as f32` (Skia's C++ multiplies in double)
    y *= (float)dscale;
    if (!gfinite(x) || !gfinite(y) || (x == 0 && y == 0)) {
        p = {0, 0};
        return false;
    }
    p = {x, y};
    return true;
}
GEO_HDI inline bool set_length(P &p, float length) { return set_length(p, p.x, p.y, length); }
GEO_HDI inline bool normalize(P &p) { return set_length(p, p.x, p.y, 1.0f); }
GEO_HDI inline bool can_normalize(float dx, float dy) { return (gfinite(dx) && gfinite(dy)) && (dx != 0 || dy != 0); }

// ---------------------------------------------------------------------------------------------------
// PathBuilder subset (tiny-skia-path path_builder.rs)
// ---------------------------------------------------------------------------------------------------
template <template <class> class Vec> struct Builder {
    Vec<uint8_t> verbs;
    Vec<P> pts;
    size_t last_move = 0;
    bool move_required = true;

    GEO_HDI void clear() { verbs.clear(); pts.clear(); last_move = 0; move_required = true; }
    GEO_HDI bool empty() const { return verbs.empty(); }
    GEO_HD void move_to(float x, float y)
    {
        if (!verbs.empty() && verbs.back() == V_MOVE) pts.back() = {x, y};
        else {
            last_move = pts.size();
            move_required = false;
            verbs.push_back(V_MOVE);
            pts.push_back({x, y});
        }
    }
    GEO_HD void inject()
    {
        if (move_required) {
            if (pts.empty()) move_to(0, 0);
            else { P p = pts[last_move]; move_to(p.x, p.y); }
        }
    }
    GEO_HD void line_to(float x, float y) { inject(); verbs.push_back(V_LINE); pts.push_back({x, y}); }
    GEO_HD void line_to(P p) { line_to(p.x, p.y); }
    GEO_HD void quad_to(P a, P b) { inject(); verbs.push_back(V_QUAD); pts.push_back(a); pts.push_back(b); }
    GEO_HD void close()
    {
        if (!verbs.empty() && verbs.back() != V_CLOSE) verbs.push_back(V_CLOSE);
        move_required = true;
    }
    GEO_HDI bool last_point(P *p) const { if (pts.empty()) return false; *p = pts.back(); return true; }
    GEO_HD void set_last_point(P p) { if (pts.empty()) move_to(p.x, p.y); else pts.back() = p; }

    // conic -> quads (AutoConicToQuads, tolerance 0.25)
    GEO_HD void conic_to(P p1, P p2, float w);
    GEO_HD void reverse_path_to(const Builder &o)
    {
        if (o.verbs.empty()) return;
        size_t off = o.pts.size() - 1;
        for (size_t i = o.verbs.size(); i-- > 0;) {
            uint8_t v = o.verbs[i];
            if (v == V_MOVE) break;
            if (v == V_LINE) { line_to(o.pts[off - 1]); off -= 1; }
            else if (v == V_QUAD) { quad_to(o.pts[off - 1], o.pts[off - 2]); off -= 2; }
        }
    }
    GEO_HD void push_path(const Builder &o)
    {
        last_move = pts.size();
        for (size_t i = 0; i < o.verbs.size(); i++) verbs.push_back(o.verbs[i]);
        for (size_t i = 0; i < o.pts.size(); i++) pts.push_back(o.pts[i]);
    }
    GEO_HD void push_circle(float cx, float cy, float r)
    {
        float l = cx - r, t = cy - r, rr = cx + r, b = cy + r;
        float mx = l * 0.5f + rr * 0.5f, my = t * 0.5f + b * 0.5f;
        P oval[4] = {{mx, b}, {l, my}, {mx, t}, {rr, my}};
        P rect[4] = {{rr, b}, {l, b}, {l, t}, {rr, t}};
        move_to(oval[3].x, oval[3].y);
        for (int i = 0; i < 4; i++) conic_to(rect[i], oval[i], kRoot2Over2);
        close();
    }
    // true if all points from `start` on coincide (is_zero_length_since_point)
    GEO_HD bool zero_length_since(size_t start) const
    {
        size_t n = pts.size() - start;
        if (n < 2) return true;
        for (size_t i = 1; i < n; i++) if (pts[start] != pts[start + i]) return false;
        return true;
    }
};

struct Conic {
    P p[3];
    float w;
    GEO_HD void chop(Conic &a, Conic &b) const
    {
        float scale = 1.0f / (1.0f + w);
        float nw = sqrtf(0.5f + w * 0.5f);
        P wp1 = p[1] * w;
        P m = (p[0] + wp1 * 2.0f + p[2]) * scale * 0.5f;
        if (!finite(m)) {
            double wd = w, w2 = wd * 2, sh = 0.5 / (1 + wd);
            m.x = (float)((p[0].x + w2 * p[1].x + p[2].x) * sh);
            m.y = (float)((p[0].y + w2 * p[1].y + p[2].y) * sh);
        }
        a.p[0] = p[0]; a.p[1] = (p[0] + wp1) * scale; a.p[2] = m; a.w = nw;
        b.p[0] = m; b.p[1] = (wp1 + p[2]) * scale; b.p[2] = p[2]; b.w = nw;
    }
};
GEO_HDI inline bool between(float a, float b, float c) { return (a - b) * (c - b) <= 0; }

GEO_HD inline P *subdivide(const Conic &src, P *out, int level)
{
    if (level == 0) {
        out[0] = src.p[1];
        out[1] = src.p[2];
        return out + 2;
    }
    Conic a, b;
    src.chop(a, b);
    float sy = src.p[0].y, ey = src.p[2].y;
    if (between(sy, src.p[1].y, ey)) {
        float my = a.p[2].y;
        if (!between(sy, my, ey)) {
            float cy = fabsf(my - sy) < fabsf(my - ey) ? sy : ey;
            a.p[2].y = cy;
            b.p[0].y = cy;
        }
        if (!between(sy, a.p[1].y, a.p[2].y)) a.p[1].y = sy;
        if (!between(b.p[0].y, b.p[1].y, ey)) b.p[1].y = ey;
    }
    out = subdivide(a, out, level - 1);
    return subdivide(b, out, level - 1);
}

template <template <class> class Vec> GEO_HD void Builder<Vec>::conic_to(P p1, P p2, float w)
{
    if (!(w > 0.0f)) { line_to(p2); return; }
    if (!gfinite(w)) { line_to(p1); line_to(p2); return; }
    if (w == 1.0f) { quad_to(p1, p2); return; }
    inject();
    P last = pts.back();
    Conic c{{last, p1, p2}, w};
    if (!finite(last) || !finite(p1) || !finite(p2)) return;
    // compute_quad_pow2(0.25)
    float a = w - 1.0f, k = a / (4.0f * (2.0f + a));
    float ex = k * (last.x - 2.0f * p1.x + p2.x), ey = k * (last.y - 2.0f * p1.y + p2.y);
    float err = sqrtf(ex * ex + ey * ey);
    int pow2 = 0;
    for (int i = 0; i < 4; i++) {
        if (err <= 0.25f) break;
        err *= 0.25f;
        pow2++;
    }
    pow2 = gmax(pow2, 1);
    P q[64];
    q[0] = last;
    subdivide(c, q + 1, pow2);
    int quads = 1 << pow2, npt = 2 * quads + 1;
    bool bad = false;
    for (int i = 0; i < npt; i++) bad = bad || !finite(q[i]);
    if (bad) for (int i = 1; i < npt - 1; i++) q[i] = p1;
    for (int i = 0; i < quads; i++) quad_to(q[1 + 2 * i], q[2 + 2 * i]);
}

// SkConic::BuildUnitArc.  ccw: rotation direction flag.  Returns count (<= 5); conics are mapped by
// (radius scale, translate pivot).
GEO_HD inline int build_unit_arc(P u_start, P u_stop, bool ccw, float radius, P pivot, Conic out[5])
{
    float x = dot(u_start, u_stop), y = cross(u_start, u_stop);
    float abs_y = fabsf(y);
    if (abs_y <= kNearlyZero && x > 0 && ((y >= 0 && !ccw) || (y <= 0 && ccw))) return 0;
    if (ccw) y = -y;
    int quadrant = 0;
    if (y == 0) quadrant = 2;
    else if (x == 0) quadrant = y > 0 ? 1 : 3;
    else {
        if (y < 0) quadrant += 2;
        if ((x < 0) != (y < 0)) quadrant += 1;
    }
    const P qp[8] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
    int count = quadrant;
    for (int i = 0; i < count; i++) {
        out[i].p[0] = qp[i * 2]; out[i].p[1] = qp[i * 2 + 1]; out[i].p[2] = qp[(i * 2 + 2) % 8];
        out[i].w = kRoot2Over2;
    }
    P final_p = {x, y};
    P last_q = qp[quadrant * 2];
    float d = dot(last_q, final_p);
    if (d < 1.0f) {
        P off = {last_q.x + x, last_q.y + y};
        float cos_half = sqrtf((1.0f + d) / 2.0f);
        set_length(off, 1.0f / cos_half);
        if (!(nearly_zero(last_q.x - off.x) && nearly_zero(last_q.y - off.y))) {
            out[count].p[0] = last_q; out[count].p[1] = off; out[count].p[2] = final_p;
            out[count].w = cos_half;
            count++;
        }
    }
    // Transform::from_sin_cos(u_start.y, u_start.x), pre_scale(1, -1) for ccw, post_concat(scale(radius) + translate(pivot)),
    // then map_points: composed with tiny-skia's concat (f64 mul-add for the skewed case) before any point is touched
    auto mam = [](float a, float b, float c, float d) { return (float)((double)a * (double)b + (double)c * (double)d); };
    float m_sx = u_start.x, m_ky = u_start.y, m_kx = -u_start.y, m_sy = u_start.x; // rotation (has skew unless u_start.y == 0)
    if (ccw) {
        if (m_kx != 0 || m_ky != 0) { // concat(m, scale(1, -1)) through the general branch
            float sx = mam(m_sx, 1.0f, m_kx, 0.0f), ky = mam(m_ky, 1.0f, m_sy, 0.0f), kx = mam(m_sx, 0.0f, m_kx, -1.0f), sy = mam(m_ky, 0.0f, m_sy, -1.0f);
            m_sx = sx; m_ky = ky; m_kx = kx; m_sy = sy;
        } else {
            m_sy = m_sy * -1.0f;
        }
    }
    float f_sx, f_ky, f_kx, f_sy, f_tx, f_ty;
    const bool m_identity = m_sx == 1 && m_ky == 0 && m_kx == 0 && m_sy == 1;
    const bool u_identity = radius == 1 && pivot.x == 0 && pivot.y == 0;
    if (u_identity) { f_sx = m_sx; f_ky = m_ky; f_kx = m_kx; f_sy = m_sy; f_tx = 0; f_ty = 0; }
    else if (m_identity) { f_sx = radius; f_ky = 0; f_kx = 0; f_sy = radius; f_tx = pivot.x; f_ty = pivot.y; }
    else if (m_kx == 0 && m_ky == 0) { f_sx = radius * m_sx; f_ky = 0; f_kx = 0; f_sy = radius * m_sy; f_tx = radius * 0.0f + pivot.x; f_ty = radius * 0.0f + pivot.y; }
    else {
        f_sx = mam(radius, m_sx, 0.0f, m_ky); f_ky = mam(0.0f, m_sx, radius, m_ky);
        f_kx = mam(radius, m_kx, 0.0f, m_sy); f_sy = mam(0.0f, m_kx, radius, m_sy);
        f_tx = mam(radius, 0.0f, 0.0f, 0.0f) + pivot.x; f_ty = mam(0.0f, 0.0f, radius, 0.0f) + pivot.y;
    }
    const bool f_identity = f_sx == 1 && f_ky == 0 && f_kx == 0 && f_sy == 1 && f_tx == 0 && f_ty == 0;
    for (int i = 0; i < count; i++)
        for (int j = 0; j < 3; j++) {
            P p = out[i].p[j];
            if (f_identity) continue;
            if (f_kx == 0 && f_ky == 0) {
                if (f_sx == 1 && f_sy == 1) out[i].p[j] = {p.x + f_tx, p.y + f_ty};
                else out[i].p[j] = {p.x * f_sx + f_tx, p.y * f_sy + f_ty};
            } else {
                out[i].p[j] = {p.x * f_sx + p.y * f_kx + f_tx, p.x * f_ky + p.y * f_sy + f_ty};
            }
        }
    return count;
}

// ---------------------------------------------------------------------------------------------------
// curve helpers (tiny-skia-path path_geometry.rs)
// ---------------------------------------------------------------------------------------------------
GEO_HDI inline bool unit_divide(float numer, float denom, float *r)
{
    if (numer < 0) { numer = -numer; denom = -denom; }
    if (denom == 0 || numer == 0 || numer >= denom) return false;
    float v = numer / denom;
    if (!(v > 0.0f && v < 1.0f)) return false;
    *r = v;
    return true;
}
GEO_HDI inline int unit_quad_roots(float a, float b, float c, float roots[2])
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
GEO_HDI inline P eval_quad(const P q[3], float t)
{
    P a = q[2] - q[1] * 2.0f + q[0], b = (q[1] - q[0]) * 2.0f;
    return (a * t + b) * t + q[0];
}
GEO_HDI inline P eval_quad_tangent(const P q[3], float t)
{
    if ((t == 0 && q[0] == q[1]) || (t == 1 && q[1] == q[2])) return q[2] - q[0];
    P b = q[1] - q[0], a = q[2] - q[1] - b;
    P tt = a * t + b;
    return tt + tt;
}
GEO_HDI inline P eval_cubic(const P c[4], float t)
{
    P a = c[3] + (c[1] - c[2]) * 3.0f - c[0], b = (c[2] - c[1] * 2.0f + c[0]) * 3.0f, cc = (c[1] - c[0]) * 3.0f;
    return ((a * t + b) * t + cc) * t + c[0];
}
GEO_HDI inline P eval_cubic_derivative(const P c[4], float t)
{
    P a = c[3] + (c[1] - c[2]) * 3.0f - c[0], b = (c[2] - c[1] * 2.0f + c[0]) * 2.0f, cc = c[1] - c[0];
    return (a * t + b) * t + cc;
}
GEO_HDI inline P eval_cubic_tangent(const P c[4], float t)
{
    if ((t == 0 && c[0] == c[1]) || (t == 1 && c[2] == c[3])) {
        P tg = t == 0 ? c[2] - c[0] : c[3] - c[1];
        if (tg.x == 0 && tg.y == 0) tg = c[3] - c[0];
        return tg;
    }
    return eval_cubic_derivative(c, t);
}
GEO_HDI inline void chop_cubic(const P s[4], float t, P d[7])
{
    auto L = [&](P a, P b) { return P{a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t}; };
    P ab = L(s[0], s[1]), bc = L(s[1], s[2]), cd = L(s[2], s[3]);
    P abc = L(ab, bc), bcd = L(bc, cd);
    d[0] = s[0]; d[1] = ab; d[2] = abc; d[3] = L(abc, bcd); d[4] = bcd; d[5] = cd; d[6] = s[3];
}
GEO_HD inline float quad_max_curvature(const P q[3])
{
    float ax = q[1].x - q[0].x, ay = q[1].y - q[0].y;
    float bx = q[0].x - q[1].x - q[1].x + q[2].x, by = q[0].y - q[1].y - q[1].y + q[2].y;
    float numer = -(ax * bx + ay * by), denom = bx * bx + by * by;
    if (denom < 0) { numer = -numer; denom = -denom; }
    if (numer <= 0) return 0;
    if (numer >= denom) return 1;
    return numer / denom;
}
GEO_HD inline int cubic_inflections(const P c[4], float t[2])
{
    float ax = c[1].x - c[0].x, ay = c[1].y - c[0].y;
    float bx = c[2].x - 2 * c[1].x + c[0].x, by = c[2].y - 2 * c[1].y + c[0].y;
    float cx = c[3].x + 3 * (c[1].x - c[2].x) - c[0].x, cy = c[3].y + 3 * (c[1].y - c[2].y) - c[0].y;
    return unit_quad_roots(bx * cy - by * cx, ax * cy - ay * cx, ax * by - ay * bx, t);
}
GEO_HDI inline float pin01(float v) { return gmin(gmax(v, 0.0f), 1.0f); }
GEO_HD inline int solve_cubic_poly(const float co[4], float t[3])
{
    if (nearly_zero(co[0])) return unit_quad_roots(co[1], co[2], co[3], t);
    float inva = 1.0f / co[0];
    float a = co[1] * inva, b = co[2] * inva, c = co[3] * inva;
    float Q = (a * a - b * 3) / 9, R = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
    float Q3 = Q * Q * Q, R2mQ3 = R * R - Q3, adiv3 = a / 3;
    if (R2mQ3 < 0) {
        float theta = g_acosf(gmin(gmax(R / sqrtf(Q3), -1.0f), 1.0f));
        float n2rq = -2 * sqrtf(Q);
        const float pi = 3.14159265f;
        t[0] = pin01(n2rq * g_cosf(theta / 3) - adiv3);
        t[1] = pin01(n2rq * g_cosf((theta + 2 * pi) / 3) - adiv3);
        t[2] = pin01(n2rq * g_cosf((theta - 2 * pi) / 3) - adiv3);
        if (t[0] > t[1]) gswap(t[0], t[1]);
        if (t[1] > t[2]) gswap(t[1], t[2]);
        if (t[0] > t[1]) gswap(t[0], t[1]);
        int n = 3;
        if (t[1] == t[2]) n--;
        if (t[0] == t[1]) { t[1] = t[2]; n--; }
        return n;
    }
    float A = fabsf(R) + sqrtf(R2mQ3);
    A = g_cbrtf(A);
    if (R > 0) A = -A;
    if (A != 0) A += Q / A;
    t[0] = pin01(A - adiv3);
    return 1;
}
GEO_HD inline int cubic_max_curvature(const P c[4], float t[3])
{
    float co[4] = {0, 0, 0, 0};
    for (int axis = 0; axis < 2; axis++) {
        float s0 = axis ? c[0].y : c[0].x, s1 = axis ? c[1].y : c[1].x, s2 = axis ? c[2].y : c[2].x, s3 = axis ? c[3].y : c[3].x;
        float a = s1 - s0, b = s2 - 2 * s1 + s0, cc = s3 + 3 * (s1 - s2) - s0;
        co[0] += cc * cc; co[1] += 3 * b * cc; co[2] += 2 * b * b + cc * a; co[3] += a * b;
    }
    return solve_cubic_poly(co, t);
}
GEO_HD inline bool on_same_side(const P c[4], int test, int line)
{
    P origin = c[line], ln = c[line + 1] - origin;
    float cr[2];
    for (int i = 0; i < 2; i++) cr[i] = cross(ln, c[test + i] - origin);
    return cr[0] * cr[1] >= 0;
}
GEO_HD inline float cubic_cusp(const P c[4])
{
    if (c[0] == c[1] || c[2] == c[3]) return -1;
    if (on_same_side(c, 0, 2) || on_same_side(c, 2, 0)) return -1;
    float mc[3];
    int roots = cubic_max_curvature(c, mc);
    for (int i = 0; i < roots; i++) {
        float t = mc[i];
        if (0 >= t || t >= 1) continue;
        P d = eval_cubic_derivative(c, t);
        float precision = (dist_sqd(c[1], c[0]) + dist_sqd(c[2], c[1]) + dist_sqd(c[3], c[2])) * 1e-8f;
        if (len_sqd(d) < precision) return t;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------------
// the stroker
// ---------------------------------------------------------------------------------------------------
enum Cap { CapButt = 0, CapRound = 1, CapSquare = 2 };
enum Join { JoinMiter = 0, JoinMiterClip = 1, JoinRound = 2, JoinBevel = 3 };
enum Result { Split, Degenerate, Quad };
enum Reduction { RPoint, RLine, RQuad, RDegenerate, RDegenerate2, RDegenerate3 };

struct QuadConstruct {
    P quad[3], tangent_start, tangent_end;
    float start_t, mid_t, end_t;
    bool start_set, end_set, opposite_tangents;
    GEO_HDI bool init(float s, float e)
    {
        start_t = s; mid_t = (s + e) * 0.5f; end_t = e;
        start_set = end_set = false;
        return start_t < mid_t && mid_t < end_t;
    }
    GEO_HDI bool init_with_start(const QuadConstruct &p)
    {
        if (!init(p.start_t, p.mid_t)) return false;
        quad[0] = p.quad[0]; tangent_start = p.tangent_start; start_set = true;
        return true;
    }
    GEO_HDI bool init_with_end(const QuadConstruct &p)
    {
        if (!init(p.mid_t, p.end_t)) return false;
        quad[2] = p.quad[2]; tangent_end = p.tangent_end; end_set = true;
        return true;
    }
};

GEO_HDI inline float pt_to_line(P pt, P a, P b)
{
    P dxy = b - a, ab0 = pt - a;
    float numer = dot(dxy, ab0), denom = dot(dxy, dxy);
    float t = numer / denom;
    if (t >= 0 && t <= 1) {
        P hit = {a.x * (1 - t) + b.x * t, a.y * (1 - t) + b.y * t};
        return dist_sqd(hit, pt);
    }
    return dist_sqd(pt, a);
}
GEO_HDI inline bool degenerate_vector(P v) { return !can_normalize(v.x, v.y); }
GEO_HDI inline bool is_clockwise(P before, P after) { return before.x * after.y > before.y * after.x; }
enum Angle { Nearly180, Sharp, Shallow, NearlyLine };
GEO_HDI inline Angle dot_to_angle(float d)
{
    if (d >= 0) return nearly_zero(1.0f - d) ? NearlyLine : Shallow;
    return nearly_zero(1.0f + d) ? Nearly180 : Sharp;
}

template <template <class> class Vec> struct Stroker {
    typedef sk::Builder<Vec> Builder;
    float radius, inv_miter_limit, res_scale, inv_res_scale, inv_res_scale_sq;
    P first_normal, prev_normal, first_unit_normal, prev_unit_normal, first_pt, prev_pt, first_outer_pt;
    size_t first_outer_idx = 0;
    int segment_count = -1;
    bool prev_is_line = false;
    Cap cap;
    Join join;
    Builder inner, outer, cusper;
    int stroke_type = 1; // 1 outer, -1 inner
    int recursion_depth = 0;
    bool found_tangents = false, join_completed = false;
    bool too_deep = false; // device only: a recursion went past the device stack budget (the batch then goes to the host builder)

    // Back to the freshly constructed state, keeping the builders' storage.
    GEO_HD void reset()
    {
        first_normal = prev_normal = first_unit_normal = prev_unit_normal = first_pt = prev_pt = first_outer_pt = P{0, 0};
        first_outer_idx = 0;
        segment_count = -1;
        prev_is_line = false;
        stroke_type = 1;
        recursion_depth = 0;
        found_tangents = false;
        join_completed = false;
        inner.clear();
        outer.clear();
        cusper.clear();
    }

    GEO_HDI bool set_normal_unitnormal(P before, P after, float scale, P &normal, P &unit)
    {
        if (!set_length(unit, (after.x - before.x) * scale, (after.y - before.y) * scale, 1.0f)) return false;
        unit = rot_ccw(unit);
        normal = unit * radius;
        return true;
    }
    GEO_HDI bool set_normal_unitnormal2(P vec, P &normal, P &unit)
    {
        if (!set_length(unit, vec.x, vec.y, 1.0f)) return false;
        unit = rot_ccw(unit);
        normal = unit * radius;
        return true;
    }

    // ---- joins ----
    GEO_HD static void handle_inner_join(P pivot, P after, Builder &in)
    {
        in.line_to(pivot);
        in.line_to(pivot - after);
    }
    GEO_HD void do_join(Join j, P before_un, P pivot, P after_un, bool prev_line, bool curr_line)
    {
        Builder *out = &outer, *in = &inner;
        if (j == JoinBevel) {
            P after = after_un * radius;
            if (!is_clockwise(before_un, after_un)) { gswap(out, in); after = -after; }
            out->line_to(pivot + after);
            handle_inner_join(pivot, after, *in);
            return;
        }
        if (j == JoinRound) {
            float dp = dot(before_un, after_un);
            if (dot_to_angle(dp) == NearlyLine) return;
            P before = before_un, after = after_un;
            bool ccw = false;
            if (!is_clockwise(before, after)) { gswap(out, in); before = -before; after = -after; ccw = true; }
            Conic cs[5];
            int n = build_unit_arc(before, after, ccw, radius, pivot, cs);
            if (n > 0) {
                for (int i = 0; i < n; i++) out->conic_to(cs[i].p[1], cs[i].p[2], cs[i].w);
                handle_inner_join(pivot, after * radius, *in);
            }
            return;
        }
        const bool miter_clip = j == JoinMiterClip;
        float dp = dot(before_un, after_un);
        Angle at = dot_to_angle(dp);
        P before = before_un, after = after_un, mid;
        if (at == NearlyLine) return;
        auto blunt_or_clipped = [&](bool curr_is_line) {
            P a = after * radius;
            if (miter_clip) {
                P m = mid;
                normalize(m);
                float cos_b = dot(before, m), sin_b = cross(before, m);
                float x = fabsf(sin_b) <= kNearlyZero ? 1.0f / inv_miter_limit : ((1.0f / inv_miter_limit) - cos_b) / sin_b;
                P b = before * radius;
                P bt = rot_cw(b), atg = rot_ccw(a);
                P c1 = pivot + b + bt * x, c2 = pivot + a + atg * x;
                if (prev_line) out->set_last_point(c1); else out->line_to(c1);
                out->line_to(c2);
            }
            if (!curr_is_line) out->line_to(pivot + a);
            handle_inner_join(pivot, a, *in);
        };
        auto do_miter = [&]() {
            P a = after * radius;
            if (prev_line) out->set_last_point(pivot + mid); else out->line_to(pivot + mid);
            if (!curr_line) out->line_to(pivot + a);
            handle_inner_join(pivot, a, *in);
        };
        if (at == Nearly180) {
            mid = (after - before) * (radius / 2.0f);
            blunt_or_clipped(false);
            return;
        }
        bool ccw = !is_clockwise(before, after);
        if (ccw) { gswap(out, in); before = -before; after = -after; }
        if (dp == 0.0f && inv_miter_limit <= kRoot2Over2) {
            mid = (before + after) * radius;
            do_miter();
            return;
        }
        if (at == Sharp) {
            mid = {after.y - before.y, before.x - after.x};
            if (ccw) mid = -mid;
        } else mid = {before.x + after.x, before.y + after.y};
        float sin_half = sqrtf((1.0f + dp) * 0.5f);
        if (sin_half < inv_miter_limit) { blunt_or_clipped(false); return; }
        set_length(mid, radius / sin_half);
        do_miter();
    }

    // ---- caps ----
    GEO_HD void do_cap(P pivot, P normal, P stop, Builder *other)
    {
        if (cap == CapButt) { outer.line_to(stop); return; }
        P parallel = rot_cw(normal);
        if (cap == CapRound) {
            P pc = pivot + parallel;
            outer.conic_to(pc + normal, pc, kRoot2Over2);
            outer.conic_to(pc - normal, stop, kRoot2Over2);
            return;
        }
        if (other) {
            outer.set_last_point(pivot + normal + parallel);
            outer.line_to(pivot - normal + parallel);
        } else {
            outer.line_to(pivot + normal + parallel);
            outer.line_to(pivot - normal + parallel);
            outer.line_to(stop);
        }
    }

    GEO_HD bool pre_join_to(P curr, P &normal, P &unit, bool curr_is_line)
    {
        float px = prev_pt.x, py = prev_pt.y;
        if (!set_normal_unitnormal(prev_pt, curr, res_scale, normal, unit)) {
            if (cap == CapButt) return false;
            normal = {radius, 0};
            unit = {1, 0};
        }
        if (segment_count == 0) {
            first_normal = normal;
            first_unit_normal = unit;
            first_outer_pt = {px + normal.x, py + normal.y};
            outer.move_to(first_outer_pt.x, first_outer_pt.y);
            inner.move_to(px - normal.x, py - normal.y);
        } else {
            do_join(join, prev_unit_normal, prev_pt, unit, prev_is_line, curr_is_line);
        }
        prev_is_line = curr_is_line;
        return true;
    }
    GEO_HD void post_join_to(P curr, P normal, P unit)
    {
        join_completed = true;
        prev_pt = curr;
        prev_unit_normal = unit;
        prev_normal = normal;
        segment_count += 1;
    }
    GEO_HD void finish_contour(bool close, bool curr_is_line)
    {
        if (segment_count > 0) {
            P pt;
            if (close) {
                do_join(join, prev_unit_normal, prev_pt, first_unit_normal, prev_is_line, curr_is_line);
                outer.close();
                inner.last_point(&pt);
                outer.move_to(pt.x, pt.y);
                outer.reverse_path_to(inner);
                outer.close();
            } else {
                inner.last_point(&pt);
                do_cap(prev_pt, prev_normal, pt, curr_is_line ? &inner : nullptr);
                outer.reverse_path_to(inner);
                do_cap(first_pt, -first_normal, first_outer_pt, prev_is_line ? &inner : nullptr);
                outer.close();
            }
            if (!cusper.empty()) {
                outer.push_path(cusper);
                cusper.clear();
            }
        }
        inner.clear();
        segment_count = -1;
        first_outer_idx = outer.pts.size();
    }
    GEO_HD void move_to(P p)
    {
        if (segment_count > 0) finish_contour(false, false);
        segment_count = 0;
        first_pt = prev_pt = p;
        join_completed = false;
    }
    GEO_HD void line_to(P p, bool has_valid_tangent_ahead)
    {
        float tol = kNearlyZero * inv_res_scale;
        bool teeny = nearly_zero(prev_pt.x - p.x, tol) && nearly_zero(prev_pt.y - p.y, tol);
        if (cap == CapButt && teeny) return;
        if (teeny && (join_completed || has_valid_tangent_ahead)) return;
        P normal, unit;
        if (!pre_join_to(p, normal, unit, true)) return;
        outer.line_to(p + normal);
        inner.line_to(p - normal);
        post_join_to(p, normal, unit);
    }

    // ---- quad / cubic offsetting ----
    GEO_HDI void set_ray_pts(P tp, P &dxy, P *on, P *tangent)
    {
        if (!set_length(dxy, radius)) dxy = {radius, 0};
        float flip = (float)stroke_type;
        on->x = tp.x + flip * dxy.y;
        on->y = tp.y - flip * dxy.x;
        if (tangent) { tangent->x = on->x + dxy.x; tangent->y = on->y + dxy.y; }
    }
    GEO_HDI void quad_perp_ray(const P q[3], float t, P *tp, P *on, P *tangent)
    {
        *tp = eval_quad(q, t);
        P dxy = eval_quad_tangent(q, t);
        if (dxy.x == 0 && dxy.y == 0) dxy = q[2] - q[0];
        set_ray_pts(*tp, dxy, on, tangent);
    }
    GEO_HDI void cubic_perp_ray(const P c[4], float t, P *tp, P *on, P *tangent)
    {
        *tp = eval_cubic(c, t);
        P dxy = eval_cubic_tangent(c, t);
        P chopped[7];
        if (dxy.x == 0 && dxy.y == 0) {
            const P *cp = c;
            if (nearly_zero(t)) dxy = c[2] - c[0];
            else if (nearly_zero(1 - t)) dxy = c[3] - c[1];
            else {
                chop_cubic(c, t, chopped);
                dxy = chopped[3] - chopped[2];
                if (dxy.x == 0 && dxy.y == 0) { dxy = chopped[3] - chopped[1]; cp = chopped; }
            }
            if (dxy.x == 0 && dxy.y == 0) dxy = cp[3] - cp[0];
        }
        set_ray_pts(*tp, dxy, on, tangent);
    }
    GEO_HDI Result intersect_ray(QuadConstruct &qp, bool want_ctrl)
    {
        P start = qp.quad[0], end = qp.quad[2];
        P a_len = qp.tangent_start - start, b_len = qp.tangent_end - end;
        float denom = cross(a_len, b_len);
        if (denom == 0 || !gfinite(denom)) {
            qp.opposite_tangents = dot(a_len, b_len) < 0;
            return Degenerate;
        }
        qp.opposite_tangents = false;
        P ab0 = start - end;
        float numer_a = cross(b_len, ab0), numer_b = cross(a_len, ab0);
        if ((numer_a >= 0) == (numer_b >= 0)) {
            float d1 = pt_to_line(start, end, qp.tangent_end), d2 = pt_to_line(end, start, qp.tangent_start);
            if (gmax(d1, d2) <= inv_res_scale_sq) return Degenerate;
            return Split;
        }
        numer_a /= denom;
        bool valid = numer_a > numer_a - 1;
        if (valid) {
            if (want_ctrl) {
                qp.quad[1].x = start.x * (1 - numer_a) + qp.tangent_start.x * numer_a;
                qp.quad[1].y = start.y * (1 - numer_a) + qp.tangent_start.y * numer_a;
            }
            return Quad;
        }
        qp.opposite_tangents = dot(a_len, b_len) < 0;
        return Degenerate;
    }
    GEO_HDI static bool points_within_dist(P a, P b, float limit) { return dist_sqd(a, b) <= limit * limit; }
    GEO_HDI static bool sharp_angle(const P q[3])
    {
        P smaller = q[1] - q[0], larger = q[1] - q[2];
        float sl = len_sqd(smaller), ll = len_sqd(larger);
        if (sl > ll) { gswap(smaller, larger); ll = sl; }
        if (!set_length(smaller, ll)) return false;
        return dot(smaller, larger) > 0;
    }
    GEO_HDI bool pt_in_quad_bounds(const P q[3], P pt)
    {
        float xmin = gmin(gmin(q[0].x, q[1].x), q[2].x);
        if (pt.x + inv_res_scale < xmin) return false;
        float xmax = gmax(gmax(q[0].x, q[1].x), q[2].x);
        if (pt.x - inv_res_scale > xmax) return false;
        float ymin = gmin(gmin(q[0].y, q[1].y), q[2].y);
        if (pt.y + inv_res_scale < ymin) return false;
        float ymax = gmax(gmax(q[0].y, q[1].y), q[2].y);
        if (pt.y - inv_res_scale > ymax) return false;
        return true;
    }
    GEO_HDI Result stroke_close_enough(const P stroke[3], const P ray[2], QuadConstruct &qp)
    {
        P mid = eval_quad(stroke, 0.5f);
        if (points_within_dist(ray[0], mid, inv_res_scale)) return sharp_angle(qp.quad) ? Split : Quad;
        if (!pt_in_quad_bounds(stroke, ray[0])) return Split;
        // intersect_quad_ray
        P vec = ray[1] - ray[0];
        float r[3];
        for (int n = 0; n < 3; n++) r[n] = (stroke[n].y - ray[0].y) * vec.x - (stroke[n].x - ray[0].x) * vec.y;
        float A = r[2], B = r[1], C = r[0];
        A += C - 2 * B;
        B -= C;
        float roots[2];
        if (unit_quad_roots(A, 2 * B, C, roots) != 1) return Split;
        P qpnt = eval_quad(stroke, roots[0]);
        float error = inv_res_scale * (1.0f - fabsf(roots[0] - 0.5f) * 2);
        if (points_within_dist(ray[0], qpnt, error)) return sharp_angle(qp.quad) ? Split : Quad;
        return Split;
    }
    GEO_HDI Builder &side() { return stroke_type == 1 ? outer : inner; }
    GEO_HDI Result compare_quad_quad(const P q[3], QuadConstruct &qp)
    {
        if (!qp.start_set) { P t; quad_perp_ray(q, qp.start_t, &t, &qp.quad[0], &qp.tangent_start); qp.start_set = true; }
        if (!qp.end_set) { P t; quad_perp_ray(q, qp.end_t, &t, &qp.quad[2], &qp.tangent_end); qp.end_set = true; }
        Result r = intersect_ray(qp, true);
        if (r != Quad) return r;
        P ray[2];
        quad_perp_ray(q, qp.mid_t, &ray[1], &ray[0], nullptr);
        return stroke_close_enough(qp.quad, ray, qp);
    }
    GEO_HD bool quad_stroke(const P q[3], QuadConstruct &qp)
    {
        Result r = compare_quad_quad(q, qp);
        if (r == Quad) { side().quad_to(qp.quad[1], qp.quad[2]); return true; }
        if (r == Degenerate) { side().line_to(qp.quad[2]); return true; }
        if (++recursion_depth > 11 * 3) return false;
#if defined(__CUDA_ARCH__)
        if (recursion_depth > 80) { too_deep = true; return false; } // never: the limits above are lower (GEO_STACK holds 80 levels)
#endif
        QuadConstruct half;
        half.init_with_start(qp);
        if (!quad_stroke(q, half)) return false;
        half.init_with_end(qp);
        if (!quad_stroke(q, half)) return false;
        --recursion_depth;
        return true;
    }
    GEO_HDI void cubic_quad_ends(const P c[4], QuadConstruct &qp)
    {
        if (!qp.start_set) { P t; cubic_perp_ray(c, qp.start_t, &t, &qp.quad[0], &qp.tangent_start); qp.start_set = true; }
        if (!qp.end_set) { P t; cubic_perp_ray(c, qp.end_t, &t, &qp.quad[2], &qp.tangent_end); qp.end_set = true; }
    }
    GEO_HDI bool cubic_mid_on_line(const P c[4], const QuadConstruct &qp)
    {
        P t, mid;
        cubic_perp_ray(c, qp.mid_t, &t, &mid, nullptr);
        return pt_to_line(mid, qp.quad[0], qp.quad[2]) < inv_res_scale_sq;
    }
    GEO_HDI Result compare_quad_cubic(const P c[4], QuadConstruct &qp)
    {
        cubic_quad_ends(c, qp);
        Result r = intersect_ray(qp, true);
        if (r != Quad) return r;
        P ray[2];
        cubic_perp_ray(c, qp.mid_t, &ray[1], &ray[0], nullptr);
        return stroke_close_enough(qp.quad, ray, qp);
    }
    GEO_HD bool cubic_stroke(const P c[4], QuadConstruct &qp)
    {
        if (!found_tangents) {
            cubic_quad_ends(c, qp);
            Result r = intersect_ray(qp, false);
            if (r != Quad) {
                if ((r == Degenerate || points_within_dist(qp.quad[0], qp.quad[2], inv_res_scale)) && cubic_mid_on_line(c, qp)) {
                    side().line_to(qp.quad[2]);
                    return true;
                }
            } else found_tangents = true;
        }
        if (found_tangents) {
            Result r = compare_quad_cubic(c, qp);
            if (r == Quad) { side().quad_to(qp.quad[1], qp.quad[2]); return true; }
            if (r == Degenerate && !qp.opposite_tangents) { side().line_to(qp.quad[2]); return true; }
        }
        if (!finite(qp.quad[2])) return false;
        const int limits[2] = {5 * 3, 26 * 3};
        if (++recursion_depth > limits[found_tangents ? 1 : 0]) return false;
#if defined(__CUDA_ARCH__)
        if (recursion_depth > 80) { too_deep = true; return false; } // never: the limits above are lower (GEO_STACK holds 80 levels)
#endif
        QuadConstruct half;
        if (!half.init_with_start(qp)) { side().line_to(qp.quad[2]); --recursion_depth; return true; }
        if (!cubic_stroke(c, half)) return false;
        if (!half.init_with_end(qp)) { side().line_to(qp.quad[2]); --recursion_depth; return true; }
        if (!cubic_stroke(c, half)) return false;
        --recursion_depth;
        return true;
    }

    GEO_HD static bool quad_in_line(const P q[3])
    {
        float pt_max = -1;
        int o1 = 0, o2 = 0;
        for (int i = 0; i < 2; i++)
            for (int j = i + 1; j < 3; j++) {
                P d = q[j] - q[i];
                float m = gmax(fabsf(d.x), fabsf(d.y));
                if (pt_max < m) { o1 = i; o2 = j; pt_max = m; }
            }
        int mid = o1 ^ o2 ^ 3;
        return pt_to_line(q[mid], q[o1], q[o2]) <= pt_max * pt_max * 0.000005f;
    }
    GEO_HD static bool cubic_in_line(const P c[4])
    {
        float pt_max = -1;
        int o1 = 0, o2 = 0;
        for (int i = 0; i < 3; i++)
            for (int j = i + 1; j < 4; j++) {
                P d = c[j] - c[i];
                float m = gmax(fabsf(d.x), fabsf(d.y));
                if (pt_max < m) { o1 = i; o2 = j; pt_max = m; }
            }
        int m1 = (1 + (2 >> o2)) >> o1, m2 = o1 ^ o2 ^ m1;
        float slop = pt_max * pt_max * 0.00001f;
        return pt_to_line(c[m1], c[o1], c[o2]) <= slop && pt_to_line(c[m2], c[o1], c[o2]) <= slop;
    }

    GEO_HD void quad_to(P p1, P p2)
    {
        P q[3] = {prev_pt, p1, p2};
        bool dab = degenerate_vector(q[1] - q[0]), dbc = degenerate_vector(q[2] - q[1]);
        Reduction red;
        P reduction{0, 0};
        if (dab && dbc) red = RPoint;
        else if (dab || dbc) red = RLine;
        else if (!quad_in_line(q)) red = RQuad;
        else {
            float t = quad_max_curvature(q);
            if (t == 0 || t == 1) red = RLine;
            else { reduction = eval_quad(q, t); red = RDegenerate; }
        }
        if (red == RPoint || red == RLine) { line_to(p2, false); return; }
        if (red == RDegenerate) {
            line_to(reduction, false);
            Join save = join;
            join = JoinRound;
            line_to(p2, false);
            join = save;
            return;
        }
        P nab, uab, nbc, ubc;
        if (!pre_join_to(p1, nab, uab, false)) { line_to(p2, false); return; }
        QuadConstruct qp;
        stroke_type = 1; found_tangents = false; qp.init(0, 1);
        quad_stroke(q, qp);
        stroke_type = -1; found_tangents = false; qp.init(0, 1);
        quad_stroke(q, qp);
        if (!set_normal_unitnormal(q[1], q[2], res_scale, nbc, ubc)) { nbc = nab; ubc = uab; }
        post_join_to(p2, nbc, ubc);
    }

    GEO_HD void cubic_to(P p1, P p2, P p3)
    {
        P c[4] = {prev_pt, p1, p2, p3};
        bool dab = degenerate_vector(c[1] - c[0]), dbc = degenerate_vector(c[2] - c[1]), dcd = degenerate_vector(c[3] - c[2]);
        P reduction[3];
        const P *tangent_pt = &c[1];
        int red;
        if (dab && dbc && dcd) red = RPoint;
        else if ((int)dab + (int)dbc + (int)dcd == 2) red = RLine;
        else if (!cubic_in_line(c)) { tangent_pt = dab ? &c[2] : &c[1]; red = RQuad; }
        else {
            float tv[3];
            int count = cubic_max_curvature(c, tv), rc = 0;
            for (int i = 0; i < count; i++) {
                float t = tv[i];
                if (0 >= t || t >= 1) continue;
                reduction[rc] = eval_cubic(c, t);
                if (reduction[rc] != c[0] && reduction[rc] != c[3]) rc++;
            }
            red = rc == 0 ? (int)RLine : (int)RQuad + rc;
        }
        if (red == RPoint || red == RLine) { line_to(p3, false); return; }
        if (red >= RDegenerate) {
            line_to(reduction[0], false);
            Join save = join;
            join = JoinRound;
            if (red >= RDegenerate2) line_to(reduction[1], false);
            if (red == RDegenerate3) line_to(reduction[2], false);
            line_to(p3, false);
            join = save;
            return;
        }
        P nab, uab, ncd, ucd;
        if (!pre_join_to(*tangent_pt, nab, uab, false)) { line_to(p3, false); return; }
        float tv[2];
        int count = cubic_inflections(c, tv);
        float last_t = 0;
        for (int i = 0; i <= count; i++) {
            float next_t = i < count ? tv[i] : 1.0f;
            QuadConstruct qp;
            stroke_type = 1; found_tangents = false; qp.init(last_t, next_t);
            cubic_stroke(c, qp);
            stroke_type = -1; found_tangents = false; qp.init(last_t, next_t);
            cubic_stroke(c, qp);
            last_t = next_t;
        }
        float cusp = cubic_cusp(c);
        if (cusp > 0) {
            P loc = eval_cubic(c, cusp);
            cusper.push_circle(loc.x, loc.y, radius);
        }
        // set_cubic_end_normal
        P ab = c[1] - c[0], cd = c[3] - c[2];
        bool d_ab = degenerate_vector(ab), d_cd = degenerate_vector(cd);
        bool fallback = d_ab && d_cd;
        if (!fallback) {
            if (d_ab) { ab = c[2] - c[0]; d_ab = degenerate_vector(ab); }
            if (d_cd) { cd = c[3] - c[1]; d_cd = degenerate_vector(cd); }
            fallback = d_ab || d_cd;
        }
        if (fallback || !set_normal_unitnormal2(cd, ncd, ucd)) { ncd = nab; ucd = uab; }
        post_join_to(p3, ncd, ucd);
    }
};

// has_valid_tangent: is there a non-degenerate segment before the contour ends?
GEO_HD inline bool has_valid_tangent(const uint8_t *verbs, int n_verbs, int vi, const P *pts, int pi, P last)
{
    for (; vi < n_verbs; vi++) {
        switch (verbs[vi]) {
        case V_MOVE: return false;
        case V_LINE:
            if (pts[pi] == last) { pi += 1; continue; }
            return true;
        case V_QUAD:
            if (pts[pi] == last && pts[pi + 1] == last) { pi += 2; continue; }
            return true;
        case V_CUBIC:
            if (pts[pi] == last && pts[pi + 1] == last && pts[pi + 2] == last) { pi += 3; continue; }
            return true;
        default: return false;
        }
    }
    return false;
}


// Path::stroke(&Stroke{width, miter_limit, line_cap, line_join}, res_scale) into s.outer.  cap: 0 butt, 1 round, 2 square;
// join: 0 miter, 1 miter-clip, 2 round, 3 bevel.  `s` must be in its reset() state (its three builders empty); returns
// false when the stroke is empty (Option::None in the reference).
template <template <class> class Vec>
GEO_HD bool stroke_path(Stroker<Vec> &s, const uint8_t *verbs, int n_verbs, const P *pts, float width, float miter_limit, int cap,
                        int join, float res_scale)
{
    if (!(width > 0.0f) || !gfinite(width) || n_verbs <= 0) return false;
    Join j = (Join)join;
    float inv_miter = 0.0f;
    if (j == JoinMiter) {
        if (miter_limit <= 1.0f) j = JoinBevel;
        else inv_miter = 1.0f / miter_limit;
    }
    if (j == JoinMiterClip) inv_miter = 1.0f / miter_limit;
    s.res_scale = res_scale;
    s.inv_res_scale = 1.0f / (res_scale * 4.0f);
    s.inv_res_scale_sq = s.inv_res_scale * s.inv_res_scale;
    s.radius = width * 0.5f;
    s.inv_miter_limit = inv_miter;
    s.cap = (Cap)cap;
    s.join = j;
    int pi = 0;
    bool last_is_line = false;
    P move_pt{0, 0}, last_pt{0, 0};
    for (int vi = 0; vi < n_verbs; vi++) {
        switch (verbs[vi]) {
        case V_MOVE:
            move_pt = last_pt = pts[pi++];
            s.move_to(move_pt);
            break;
        case V_LINE: {
            P p = pts[pi++];
            s.line_to(p, has_valid_tangent(verbs, n_verbs, vi + 1, pts, pi, p));
            last_pt = p;
            last_is_line = true;
            break;
        }
        case V_QUAD:
            s.quad_to(pts[pi], pts[pi + 1]);
            last_pt = pts[pi + 1];
            pi += 2;
            last_is_line = false;
            break;
        case V_CUBIC:
            s.cubic_to(pts[pi], pts[pi + 1], pts[pi + 2]);
            last_pt = pts[pi + 2];
            pi += 3;
            last_is_line = false;
            break;
        case V_CLOSE:
            // auto-close: a line back to the contour start when the pen is elsewhere
            if (last_pt != move_pt) {
                s.line_to(move_pt, has_valid_tangent(verbs, n_verbs, vi, pts, pi, move_pt));
                last_pt = move_pt;
                last_is_line = true;
            }
            if (s.cap != CapButt) {
                if (s.segment_count == 0) { // only a move_to so far: zero-length line
                    s.line_to(move_pt, false);
                    last_is_line = true;
                    break;
                }
                if (s.inner.zero_length_since(0) && s.outer.zero_length_since(s.first_outer_idx)) {
                    last_is_line = true;
                    break;
                }
            }
            s.finish_contour(true, last_is_line);
            break;
        default: return false;
        }
    }
    s.finish_contour(false, last_is_line);
    return s.outer.verbs.size() > 1;
}

} // namespace sk
} // namespace geo
