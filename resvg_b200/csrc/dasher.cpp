// dasher.cpp — host instantiation of the path dasher (dasher_core.h) and the rb_path_dash export.
// The same source is compiled for the device by geo.cu.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/resvg_b200.h"
#include "dasher_core.h"

namespace {
template <class T> using HVec = std::vector<T>;
}

// Internal form used by the batch builder: appends nothing, replaces `ov`/`op`.  Returns RB_OK, or RB_ERR_INVALID when
// the dash specification is rejected (StrokeDash::new -> None: the stroke is then not dashed) or nothing is left.
int rb_path_dash_into(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, const float *dash_array,
                      int32_t n_dash, float dash_offset, float res_scale, std::vector<uint8_t> &ov, std::vector<float> &op,
                      bool *spec_valid)
{
    (void)n_points;
    static thread_local geo::ds::DashOut<HVec> pb;
    static thread_local geo::ds::Contour<HVec> c;
    pb.verbs.clear();
    pb.pts.clear();
    pb.move_required = true;
    pb.last_move = 0;
    if (!geo::ds::dash_path(pb, c, verbs, n_verbs, reinterpret_cast<const geo::P *>(points), dash_array, n_dash, dash_offset, res_scale, spec_valid))
        return RB_ERR_INVALID;
    ov.assign(pb.verbs.begin(), pb.verbs.end());
    op.resize(pb.pts.size() * 2);
    memcpy(op.data(), pb.pts.data(), pb.pts.size() * sizeof(geo::P));
    return RB_OK;
}

extern "C" int rb_path_dash(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, const float *dash_array,
                            int32_t n_dash, float dash_offset, float res_scale, uint8_t **out_verbs, int32_t *out_n_verbs,
                            float **out_points, int32_t *out_n_points)
{
    if (!verbs || !points || !dash_array || !out_verbs || !out_points || !out_n_verbs || !out_n_points || n_verbs <= 0) return RB_ERR_INVALID;
    *out_verbs = nullptr; *out_points = nullptr; *out_n_verbs = 0; *out_n_points = 0;
    std::vector<uint8_t> ov;
    std::vector<float> op;
    bool valid = false;
    int st = rb_path_dash_into(verbs, n_verbs, points, n_points, dash_array, n_dash, dash_offset, res_scale, ov, op, &valid);
    if (st != RB_OK) return st;
    uint8_t *v = (uint8_t *)malloc(ov.size());
    float *p = (float *)malloc(op.size() * sizeof(float));
    if (!v || !p) { free(v); free(p); return RB_ERR_OOM; }
    memcpy(v, ov.data(), ov.size());
    memcpy(p, op.data(), op.size() * sizeof(float));
    *out_verbs = v; *out_n_verbs = (int32_t)ov.size();
    *out_points = p; *out_n_points = (int32_t)(op.size() / 2);
    return RB_OK;
}
