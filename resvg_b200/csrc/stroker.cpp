// stroker.cpp — host instantiation of the path stroker (stroker_core.h) and the rb_path_stroke export.
// The same source is compiled for the device by geo.cu.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/resvg_b200.h"
#include "stroker_core.h"

namespace {
template <class T> using HVec = std::vector<T>;
typedef geo::sk::Stroker<HVec> HostStroker;
}

// path_geometry helpers shared with the hairline walker
namespace rbs {
int cubic_max_curvature_ts(const float pts[8], float t[3]) { return geo::sk::cubic_max_curvature(reinterpret_cast<const geo::P *>(pts), t); }
void chop_cubic_at_t(const float src[8], float t, float dst[14])
{
    geo::sk::chop_cubic(reinterpret_cast<const geo::P *>(src), t, reinterpret_cast<geo::P *>(dst));
}
}

// Internal form: the outline stays in a thread-local stroker (valid until the next call on this thread).
// Returns RB_OK, or RB_ERR_INVALID when the stroke is empty (Option::None in the reference).
int rb_path_stroke_view(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, float width,
                        float miter_limit, int32_t cap, int32_t join, float res_scale, const uint8_t **out_verbs,
                        int32_t *out_n_verbs, const float **out_points, int32_t *out_n_points)
{
    (void)n_points;
    if (!verbs || !points || !out_verbs || !out_points || !out_n_verbs || !out_n_points) return RB_ERR_INVALID;
    *out_verbs = nullptr; *out_points = nullptr; *out_n_verbs = 0; *out_n_points = 0;
    static thread_local HostStroker tls;
    HostStroker &s = tls;
    s.reset();
    if (!geo::sk::stroke_path(s, verbs, n_verbs, reinterpret_cast<const geo::P *>(points), width, miter_limit, cap, join, res_scale))
        return RB_ERR_INVALID;
    *out_verbs = s.outer.verbs.data();
    *out_n_verbs = (int32_t)s.outer.verbs.size();
    *out_points = reinterpret_cast<const float *>(s.outer.pts.data());
    *out_n_points = (int32_t)s.outer.pts.size();
    return RB_OK;
}

// Path::stroke(&Stroke{width, miter_limit, line_cap, line_join}, res_scale).  Outputs are malloc'ed (free with
// rb_path_free).
extern "C" int rb_path_stroke(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, float width,
                              float miter_limit, int32_t cap, int32_t join, float res_scale, uint8_t **out_verbs,
                              int32_t *out_n_verbs, float **out_points, int32_t *out_n_points)
{
    if (!out_verbs || !out_points || !out_n_verbs || !out_n_points) return RB_ERR_INVALID;
    *out_verbs = nullptr; *out_points = nullptr; *out_n_verbs = 0; *out_n_points = 0;
    const uint8_t *vv;
    const float *pp;
    int32_t cv = 0, cp = 0;
    int st = rb_path_stroke_view(verbs, n_verbs, points, n_points, width, miter_limit, cap, join, res_scale, &vv, &cv, &pp, &cp);
    if (st != RB_OK) return st;
    size_t nv = (size_t)cv, np = (size_t)cp;
    uint8_t *ov = (uint8_t *)malloc(nv);
    float *op = (float *)malloc(np * sizeof(geo::P));
    if (!ov || !op) { free(ov); free(op); return RB_ERR_OOM; }
    memcpy(ov, vv, nv);
    memcpy(op, pp, np * sizeof(geo::P));
    *out_verbs = ov; *out_points = op; *out_n_verbs = (int32_t)nv; *out_n_points = (int32_t)np;
    return RB_OK;
}

extern "C" void rb_path_free(void *p) { free(p); }

// ---- the dash-by-dash decomposition of a dashed stroke, on the host -------------------------------------------------------
// What the device does for a dashed stroke (geo.cu k_geo_plan / k_geo_unit_path) restated with the host instantiation of
// the same cores, so that the CPU test-suite can compare it with dash-then-stroke of the whole path: measure the contours,
// list the dashes (dash_contour_ranges), cut every dash out on its own, stroke it on its own (a dash that is not the last one
// is followed by the next dash's move_to, as in the whole path), and put the outlines one after the other.
#include "dasher_core.h"

extern "C" int rb_debug_stroke_dashed_in_units(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                                               const float *dash_array, int32_t n_dash, float dash_offset, float width, float miter_limit,
                                               int32_t cap, int32_t join, float res_scale, uint8_t **out_verbs, int32_t *out_n_verbs,
                                               float **out_points, int32_t *out_n_points)
{
    (void)n_points;
    if (!verbs || !points || !dash_array || !out_verbs || !out_points || !out_n_verbs || !out_n_points || n_verbs <= 0) return RB_ERR_INVALID;
    *out_verbs = nullptr; *out_points = nullptr; *out_n_verbs = 0; *out_n_points = 0;
    using namespace geo;
    const ds::DashSpec sp = ds::dash_spec(dash_array, n_dash, dash_offset);
    if (!sp.valid) return RB_ERR_INVALID;
    const P *pts = reinterpret_cast<const P *>(points);
    const float tolerance = 0.5f * (1.0f / res_scale);
    std::vector<ds::Contour<HVec>> contours;
    {
        int vi = 0, pi = 0;
        float dash_count = 0.0f;
        for (;;) {
            ds::Contour<HVec> c;
            if (!ds::next_contour(verbs, n_verbs, pts, &vi, &pi, tolerance, &c)) break;
            dash_count += c.length * (float)(n_dash >> 1) / sp.interval_len;
            if (dash_count > 1000000.0f) return RB_ERR_INVALID;
            contours.push_back(c);
        }
    }
    struct Unit { size_t contour; float a0, a1, b0, b1; bool has_b; };
    std::vector<Unit> units;
    for (size_t k = 0; k < contours.size(); k++) {
        struct Collect {
            std::vector<Unit> *u; size_t contour; size_t first;
            void operator()(float a, float b, bool mv)
            {
                if (mv || u->size() == first) u->push_back(Unit{contour, a, b, 0.0f, 0.0f, false});
                else { u->back().b0 = a; u->back().b1 = b; u->back().has_b = true; }
            }
        } collect{&units, k, units.size()};
        ds::dash_contour_ranges(sp, dash_array, n_dash, contours[k].length, contours[k].closed, collect);
    }
    std::vector<uint8_t> ov;
    std::vector<P> op;
    for (size_t u = 0; u < units.size(); u++) {
        ds::DashOut<HVec> pb;
        const ds::Contour<HVec> &c = contours[units[u].contour];
        c.push_segment(units[u].a0, units[u].a1, true, pb);
        if (units[u].has_b) c.push_segment(units[u].b0, units[u].b1, false, pb);
        // every contour but the path's last is ended by the next one's move_to (finish_contour(false, false))
        if (u + 1 != units.size()) { pb.verbs.push_back(V_MOVE); pb.pts.push_back(P{0.0f, 0.0f}); }
        HostStroker s;
        s.reset();
        if (!sk::stroke_path(s, pb.verbs.data(), (int)pb.verbs.size(), pb.pts.data(), width, miter_limit, cap, join, res_scale)) continue;
        ov.insert(ov.end(), s.outer.verbs.begin(), s.outer.verbs.end());
        op.insert(op.end(), s.outer.pts.begin(), s.outer.pts.end());
    }
    if (ov.size() <= 1) return RB_ERR_INVALID;
    uint8_t *v = (uint8_t *)malloc(ov.size());
    float *p = (float *)malloc(op.size() * sizeof(P));
    if (!v || !p) { free(v); free(p); return RB_ERR_OOM; }
    memcpy(v, ov.data(), ov.size());
    memcpy(p, op.data(), op.size() * sizeof(P));
    *out_verbs = v; *out_n_verbs = (int32_t)ov.size();
    *out_points = p; *out_n_points = (int32_t)op.size();
    return RB_OK;
}
