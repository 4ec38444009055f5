// hairline.h — host-side generation of the blits of an anti-aliased hairline stroke (see hairline.cpp).
#pragma once

#include <stdint.h>

#include <vector>

#include "../../include/resvg_b200.h"
#include "hairline_core.h"

namespace rbh {

typedef geo::HairBlit HairBlit; // {x, y, alpha}: blend the paint into pixel (x, y) with coverage alpha (1..255), in list order

void hairline_blits(const uint8_t *verbs, int n_verbs, const float *points, int n_pts, int cap, int32_t clip_w, int32_t clip_h,
                    std::vector<HairBlit> &out);

} // namespace rbh
