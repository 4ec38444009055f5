// raster_host.h — host half of the B200 fill path: path transform, Y-monotone chopping, clipping,
// fixed-point edge construction (curves pre-expanded into their line edges), shader preparation.
// The north star keeps "curve flattening and edge building on the host"; everything per-pixel runs on
// the device (raster.cu).
//
// Semantics follow tiny-skia 0.12.0 (edge.rs, edge_builder.rs, edge_clipper.rs, scan/path*.rs,
// shaders/*.rs — a Rust port of Skia's SkEdge/SkEdgeBuilder/SkEdgeClipper/SkScan_*), reached from
// crates/resvg/src/path.rs:73 (`pixmap.fill_path`).
#pragma once

#include <stdint.h>

#include <vector>

#include "../../include/resvg_b200.h"

namespace rbh {

struct Pt { float x, y; };

struct Xform {
    float sx = 1, ky = 0, kx = 0, sy = 1, tx = 0, ty = 0;
    static Xform from(const float t[6]) { Xform r; r.sx = t[0]; r.ky = t[1]; r.kx = t[2]; r.sy = t[3]; r.tx = t[4]; r.ty = t[5]; return r; }
    bool is_identity() const { return sx == 1 && ky == 0 && kx == 0 && sy == 1 && tx == 0 && ty == 0; }
    bool has_skew() const { return kx != 0 || ky != 0; }
    bool has_scale() const { return sx != 1 || sy != 1; }
    bool is_translate() const { return !has_scale() && !has_skew() && (tx != 0 || ty != 0); }
    bool is_finite() const;
};
Xform concat(const Xform &a, const Xform &b); // b applied first
inline Xform pre_concat(const Xform &self, const Xform &other) { return concat(self, other); }
inline Xform post_concat(const Xform &self, const Xform &other) { return concat(other, self); }
bool invert(const Xform &t, Xform *out);
void map_points(const Xform &t, Pt *p, int n);

// One line edge in (super-sampled) fixed point, exactly the state tiny-skia's LineEdge carries.
struct Edge {
    int32_t x;       // FDot16 at first_y
    int32_t dx;      // FDot16 per scanline
    int32_t first_y; // inclusive
    int32_t last_y;  // inclusive
    int32_t winding; // +1 / -1
    // Bookkeeping for the one place where the scanline walker's list order is observable (two crossings
    // with the identical FDot16 x): `prev` = index (within the draw, after sorting) of the previous segment
    // of the same curve, -1 for an edge that enters the walker through insert_new_edges; `before` = 1 when
    // insert_new_edges would place it in front of already-active edges with the same x (scan/path.rs:
    // every new edge but the first of its scanline batch stops at the first active edge with x >= its x).
    int32_t prev;
    int32_t before;
    uint32_t order;  // emission order (internal)
};

struct IRect { int32_t x, y, w, h; };

// Result of building one draw for one DrawTiler tile.
struct DrawGeom {
    IRect sect;   // pixels the blitter may touch, in tile-local coordinates (bounds ∩ clip)
    int shift;    // 2 = 4x4 supersampled AA, 0 = non-AA (also the AA overflow fallback)
    int32_t start_y, stop_y; // walker range in (super-sampled) scanlines
};

// A recorded quadratic / cubic edge for the device-side expansion: FDot6 control points (y ascending),
// info = kind (0 quad, 1 cubic) | log2 subdivisions << 4 | upward << 8; item = emission index among the draw's items.
struct CurveRec { int32_t p[8]; uint32_t info; uint32_t item; };

// Builds the sorted line-edge list for `path` (device space, tile-local) against clip (0,0,cw,ch).
// Appends to `out`; returns false when nothing is to be drawn.
bool build_draw(const uint8_t *verbs, int n_verbs, const Pt *pts, int n_pts, bool anti_alias, int32_t cw, int32_t ch,
                std::vector<Edge> &out, DrawGeom *geom);

// Same front end, but curves are only recorded (see CurveRec); `lines` receives the final line edges with
// order = emission index.  Nothing is sorted: the device orders nothing either.
bool build_draw_items(const uint8_t *verbs, int n_verbs, const Pt *pts, int n_pts, bool anti_alias, int32_t cw, int32_t ch,
                      std::vector<Edge> &lines, std::vector<CurveRec> &curves, DrawGeom *geom);

// ---- shaders ------------------------------------------------------------------------------------
constexpr int kMaxStops = 32;

// One gradient interval: colour(t) = t * f + b for t >= t0 (tiny-skia GradientCtx factors/biases/t_values).
struct DevStop { float f[4]; float b[4]; float t0; float pad[3]; };

// Device-consumable paint (plain data, uploaded as is).
struct DevPaint {
    int32_t kind;        // 0 solid, 1 gradient, 2 pattern
    int32_t blend;       // tiny_skia::BlendMode after strength reduction
    int32_t lowp;        // 1: u16 integer pipeline, 0: f32 pipeline
    int32_t has_memset;  // full-coverage spans store memset_color
    uint32_t memset_color;
    float premul[4];     // solid: premultiplied colour
    uint32_t solid16[4]; // solid: (c*255+0.5) as u16
    // gradient
    float ts[6];
    int32_t has_ts;
    int32_t geom;        // 0 linear, 1 radial, 2 focal, 3 strip, 4 concentric
    int32_t spread, pad_x1, two_stop, len, premul_after;
    float p0, p1;
    int32_t focal_on_circle, well_behaved, swapped, natively_focal, negate_x, smaller;
    float conc_scale, conc_bias;
    uint32_t stop_off;   // first entry of this gradient in the pooled DevStop array
    alignas(16) float t0s[12]; // t0 of the first 12 intervals (+inf beyond len): the interval search reads these, not the pool
    // pattern
    const uint8_t *pix;
    uint32_t pw, ph;
    int32_t quality;
    float opacity;
};

// RasterPipelineBlitter::new + Shader::push_stages.  Returns false when the draw is a no-op.
bool prepare_paint(const rb_paint *paint, const Xform &ctm, DevPaint *out, std::vector<DevStop> &stop_pool);

} // namespace rbh
