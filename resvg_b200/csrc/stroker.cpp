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
