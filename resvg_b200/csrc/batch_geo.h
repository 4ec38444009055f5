// batch_geo.h — device path geometry: what the host hands to the geometry kernels and what they return.
//
// With the host builder (batch_host.cpp) the host threads transform, dash, stroke, chop, clip and set up the edges of
// every draw and upload the result; for the 100 000-path scene that is about one core-second per frame and bounds the
// end-to-end rate.  Here the host only classifies and culls the draws, prepares their paints and uploads the RAW paths
// (`GeoTask` + verbs + points + dash arrays); the geometry kernels (geo.cu) run the same dasher / stroker / hairline
// walker / fill front end — the shared cores of geom_common.h — one thread per draw, and leave exactly what the host
// builder would have uploaded: DevDraw[], line edges, curve records (item mode of batch.h), ready for k_row_lists.
#pragma once

#include <stdint.h>

#include "batch.h"

// One (draw, DrawTiler tile): 1:1 with the DevDraw of the same index, painter's order.
struct GeoTask {
    uint32_t verb_off, n_verbs, pt_off, n_pts; // into the uploaded verbs / points (the draw's recorded form)
    float ctm[6];            // strokes: local -> device (applied to the outline); bulk fills: painter.rs path.transform(ts)
    float tile_tx, tile_ty;  // added to the device-space points: minus the tile's origin inside the viewport
    int32_t tw, th;          // the tile = the pixmap the draw is clipped against
    int32_t ox, oy;          // tile origin inside the target
    uint32_t paint;          // DevPaint index
    uint32_t flags;          // GT_*
    float width, miter, res_scale, dash_offset;
    uint32_t dash_off, n_dash;
    uint32_t hint;           // expected number of edge items (reserve hint, not a bound)
    int32_t sub;             // hairline strokes without dashes: the one verb this task walks (every verb is a draw of its own); -1: the whole path
    uint32_t draw, n_draws;  // its DevDraw(s): one, except a dashed hairline, whose every dash is a draw of its own (n_draws = the bound below)
    uint32_t max_units, pad; // dashed strokes: upper bound of the contours dashing can produce (GT_UNITS)
};
static_assert(sizeof(GeoTask) == 120, "GeoTask is uploaded as is");
enum {
    GT_STROKE = 1, GT_HAIR = 2, GT_AA = 4, GT_EVENODD = 8, GT_DASH = 16, GT_MAP = 32, GT_TILE = 64,
    GT_UNITS = 128, // a dashed stroke built dash by dash (one thread per output contour) rather than by one thread
    GT_CAP_SHIFT = 8, GT_JOIN_SHIFT = 10
};

// Running totals of a geometry launch (device memory, copied back once): sizes of the structures the raster pre-pass
// builds next, and what went wrong.
struct GeoTotals {
    unsigned long long heap_cursor;
    unsigned long long n_slots, n_list, n_row_off, n_row_ent, n_wpairs;
    unsigned int overflow;   // the heap ran out: repeat with a larger one
    unsigned int wide;       // some draw may exceed the packed winding range: the fallback builder has to take the batch
    unsigned int too_large;  // a draw the device structures cannot index
    unsigned int n_wide_q;   // draws queued for the exact winding bound
    unsigned int deep;       // a recursion went deeper than the device stack allows (or a bound did not hold): host fallback
    unsigned int n_units;    // output contours of the dashed strokes built in units
    unsigned int n_seq;      // dashed strokes whose edges were rebuilt by one thread (combine_vertical across units)
    unsigned int pad[1];
};

// Host half (batch_geo.cpp): tasks, paints, stops and the raw path data of draws [begin, end), laid out in one staging
// block.  Offsets are bytes from the block start.
struct GeoBlock {
    size_t o_tasks = 0, o_verbs = 0, o_pts = 0, o_dashes = 0, o_paints = 0, o_stops = 0, total = 0;
    size_t n_draws = 0;       // DevDraw entries (>= n_tasks: a dashed hairline reserves one per possible dash)
    size_t max_units = 0;     // sum of the tasks' max_units
    size_t n_units_l = 0;     // GT_UNITS tasks (their list follows the four others)
    size_t n_tasks = 0, n_verbs = 0, n_pts = 0, n_dashes = 0, n_paints = 0, n_stops = 0;
    // task indices per kernel, heaviest first: [dash | stroke | hair | fill (plain fills) | units | outline fills]
    size_t o_lists = 0, n_dash_l = 0, n_stroke_l = 0, n_hair_l = 0, n_fill_l = 0, n_outline_l = 0;
    size_t heap_hint = 0; // bytes the geometry is expected to take from the heap
    bool has_hair = false;
};
int rb_geo_host_build(rb_batch *b, int W, int H, int n_threads, rb_stage_alloc alloc, void *user, void **block, GeoBlock *gb,
                      size_t begin, size_t end);

// 0: device geometry for large batches on layers (default), 1: for every eligible batch, 2: never (host builder).
extern int g_geo_mode;
#include <atomic>
extern std::atomic<uint64_t> g_geo_counts[6]; // ranges built on the device, ranges handed back, heap retries; last range: device wait us, host build us, tasks
