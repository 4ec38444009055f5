// hairline.cpp — host instantiation of the anti-aliased hairline walker (hairline_core.h) and the rb_path_hairline
// export.  The same source is compiled for the device by geo.cu.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "hairline.h"
#include "hairline_core.h"

namespace {
template <class T> using HVec = std::vector<T>;
}

namespace rbh {

void hairline_blits(const uint8_t *verbs, int n_verbs, const float *points, int n_pts, int cap, int32_t clip_w, int32_t clip_h,
                    std::vector<HairBlit> &out)
{
    const geo::P *pts = reinterpret_cast<const geo::P *>(points);
    geo::hl::hairline_blits<HVec>(verbs, n_verbs, pts, n_pts, cap, clip_w, clip_h, out);
}

} // namespace rbh

// Host-side export for the test front end (shared geometry, like rb_path_stroke): the ordered blits as
// {x, y, alpha} int32 triples, malloc'ed.
extern "C" int rb_path_hairline(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, int32_t cap,
                                int32_t clip_w, int32_t clip_h, int32_t **out_blits, int32_t *out_n)
{
    if (!verbs || !points || !out_blits || !out_n || clip_w <= 0 || clip_h <= 0) return RB_ERR_INVALID;
    std::vector<rbh::HairBlit> v;
    rbh::hairline_blits(verbs, n_verbs, points, n_points, cap, clip_w, clip_h, v);
    int32_t *o = (int32_t *)malloc(std::max<size_t>(v.size(), 1) * 3 * sizeof(int32_t));
    if (!o) return RB_ERR_OOM;
    for (size_t i = 0; i < v.size(); i++) { o[3 * i] = v[i].x; o[3 * i + 1] = v[i].y; o[3 * i + 2] = (int32_t)v[i].alpha; }
    *out_blits = o;
    *out_n = (int32_t)v.size();
    return RB_OK;
}
