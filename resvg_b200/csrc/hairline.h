// hairline.h — host-side generation of the blits of an anti-aliased hairline stroke (see hairline.cpp).
#pragma once

#include <stdint.h>

#include <vector>

#include "../../include/resvg_b200.h"

namespace rbh {

// One blit of the hairline walker: blend the paint into pixel (x, y) with coverage alpha (1..255), in list order.
struct HairBlit { int32_t x, y; uint32_t alpha; };

void hairline_blits(const uint8_t *verbs, int n_verbs, const float *points, int n_pts, int cap, int32_t clip_w, int32_t clip_h,
                    std::vector<HairBlit> &out);

} // namespace rbh
