// batch.h — a recorded batch of fill_path / stroke_path calls and its device-consumable form.
//
// The host half (batch_host.cpp) records draws, builds edges on host threads, bins draws into device tiles and
// lays everything out in ONE contiguous block (pinned staging) that raster.cu uploads with a single copy and the
// tile kernel consumes in place.
#pragma once

#include <stdint.h>

#include <vector>

#include "raster_host.h"

struct rb_ctx;
struct rb_layer;
struct rb_mask;
struct GeoPending;

constexpr int TW = 64; // device tile width  (pixels)
constexpr int TH = 16; // device tile height (pixels)

// ypack = first_y | last_y << 16.  meta: bit 0 = upward edge (winding -1), bit 1 = continuation segment of a
// curve, bit 2 = insert_new_edges places it before equal-x active edges, bits 4.. = index of the previous
// segment of the same curve (valid when bit 1 is set).
struct DevEdge { int32_t x, dx; uint32_t ypack; uint32_t meta; };
struct DevDraw {
    uint32_t edge_off, edge_cnt;
    int32_t ox, oy;         // DrawTiler tile origin inside the layer
    int32_t sx, sy, sw, sh; // pixels the blitter may touch (DrawTiler-tile local)
    int32_t shift, rule;    // 2 = AA / 0 = non-AA; 0 winding / 1 even-odd
    uint32_t paint;
    // per-draw tile-row edge lists (built on the device by k_row_lists): tile row r of the draw covers layer pixel
    // rows [8 (r0 + r), 8 (r0 + r) + 8); its edges are row_edges[list_off + row_off[row_base + r] ..
    // list_off + row_off[row_base + r + 1])
    uint32_t list_off, row_base, r0, n_rows;
    uint32_t list_cap;      // entries reserved for this draw's lists
    // item mode (curves expanded on the device): edge_off / edge_cnt address the draw's SLOTS in the device-side
    // edge array; the uploaded line edges (meta >> 4 = slot) and curve records (item = first slot) are here:
    uint32_t line_off, line_cnt, curve_off, curve_cnt;
};
static_assert(sizeof(DevDraw) == 80, "DevDraw is uploaded as is");

struct RecordedDraw {
    uint32_t verb_off, n_verbs; // into rb_batch::verbs
    uint32_t pt_off, n_pts;     // into rb_batch::pts — fills: device space (transform applied); strokes: local space
    uint32_t stop_off, n_stops; // into rb_batch::stops (5 floats per stop)
    rb_paint paint;
    rbh::Xform ctm;
    int rule;
    bool is_stroke;
    rb_stroke stroke;           // dash_array points into rb_batch::dashes (dash_off) for individually recorded draws
    uint32_t dash_off;
};

// rb_batch_draw_paths records by reference: the caller's packed arrays are used in place by the host build.
struct BulkSeg {
    int32_t n;
    const uint32_t *verb_off, *point_off;
    const uint8_t *verbs;
    const float *points;
    const rb_paint *paints;
    const uint8_t *fill_rules;
    const rb_stroke *strokes; // may be null
    rbh::Xform ctm;
};
// Draws in call order: a span is either a run of individually recorded draws (recs[first ..]) or one bulk segment.
// vp_*: the viewport the span's draws are rendered into (vp_w == 0: the whole target), see rb_batch_set_viewport.
struct DrawSpan { size_t start, count; int bulk; size_t first; int32_t vp_x, vp_y, vp_w, vp_h; };

// Byte offsets of the arrays inside the contiguous block (identical on host staging and device).
struct BatchLayout {
    size_t o_edges = 0, o_draws = 0, o_paints = 0, o_stops = 0, o_toff = 0, o_tdraws = 0, o_tids = 0, o_curves = 0, total = 0;
    size_t n_edges = 0, n_draws = 0, n_paints = 0, n_stops = 0, n_tiles = 0, n_pairs = 0, n_tile_ids = 0;
    // item mode: the block carries final line edges (n_edges of them) and curve records; the device expands the
    // curves into an edge array of n_slots entries
    bool items = false;
    size_t n_curves = 0, n_slots = 0;
    int tiles_x = 0;
    bool wide = false; // some draw may reach |winding| > 127: use k_raster_tiles_wide
    bool has_hair = false; // some draw is a hairline stroke (its "edges" are blits)
    // warp-tile path (k_raster_warp): device-built structures, sizes known on the host
    int wtiles_x = 0, wtiles_y = 0;
    size_t n_row_off = 0;   // entries of row_off: sum over draws of (n_rows + 1)
    size_t n_list = 0;      // entries of row_edges: sum over edges of the tile rows they touch
    size_t n_row_ent = 0;   // sum over draws of n_rows
    size_t n_wpairs = 0;    // (draw, warp tile) pairs
};

// phases of the last host build, microseconds: [0] edge build (threads), [1] layout + count, [2] pack (threads),
// [3] binning, [4] staging allocation / wait, [5] total
enum { RB_PHASES = 6 };

struct rb_batch {
    rb_ctx *ctx = nullptr; // retained for the batch's lifetime: destroying it must not look at the (maybe gone) target
    rb_layer *layer = nullptr;
    rb_mask *mask = nullptr;
    int host_w = 0, host_h = 0; // host-only batches (rb_debug_batch_begin_host): no target, no upload
    std::vector<uint8_t> verbs;
    std::vector<rbh::Pt> pts;
    std::vector<float> stops;
    std::vector<float> dashes;
    std::vector<RecordedDraw> recs;
    std::vector<BulkSeg> bulk;
    std::vector<DrawSpan> spans;
    int32_t vp_x = 0, vp_y = 0, vp_w = 0, vp_h = 0; // current viewport for the draws recorded next (vp_w == 0: whole target)
    size_t n_total = 0; // draws recorded so far (recs + bulk)
    size_t n_hair = 0;  // how many of them are hairline strokes (drawn by k_hair_blits, between the fill runs)
    uint64_t stats[6] = {0, 0, 0, 0, 0, 0};
    uint64_t phases[RB_PHASES] = {0, 0, 0, 0, 0, 0};
    // device-resident form produced by rb_batch_prepare
    BatchLayout lay;
    uint8_t *dev = nullptr;
    uint8_t *dev_scratch = nullptr; // row lists + warp-tile bins (device-built)
    bool scratch_owned = false;     // dev_scratch is an allocation of its own (device geometry) rather than the tail of `dev`
    void *host_block = nullptr; // host-only batches: malloc'ed copy of the block
    struct GeoPending *geo = nullptr; // a geometry launch in flight (geo.cu rb_geo_begin / rb_geo_finish)
    // rb_batch_submit_download: the host pixmap the layer goes to.  The LAST raster launch of the submit is cut into bands of
    // tile rows and every band is copied out (context's copy stream) while the next one is rendered.
    uint8_t *dl_host = nullptr;
    bool dl_arm = false;  // the next batch_run is the submit's last one
    bool dl_done = false; // the banded copies were enqueued
};

// rb_batch_host_build: the range holds hairline strokes the chosen builder cannot draw inline (internal status).
enum { RB_NEEDS_RUN_SPLIT = 104 };
// rb_geo_prepare: the range has to be built by the host builder (internal status).
enum { RB_GEO_FALLBACK = 105 };

// Staging allocator: returns `bytes` of host memory the block is assembled in (pinned, owned by the context; or
// malloc'ed for host-only batches).
typedef void *(*rb_stage_alloc)(void *user, size_t bytes);

// Edge build + binning + layout.  Returns RB_OK with lay.n_draws == 0 when nothing is to be drawn.
// Draws [begin, end) of the batch only (end = 0: all of them).
int rb_batch_host_build(rb_batch *b, int W, int H, bool mask_target, int n_threads, rb_stage_alloc alloc, void *user,
                        void **block, size_t begin = 0, size_t end = 0);

// ---- hairline strokes (hairline.cpp): blits grouped per pixel, applied in order by one thread per pixel ----------------
struct HairGroup { uint32_t x, y, first, count; };            // layer pixel and its range of blits
struct HairDevBlit { uint32_t alpha, paint; int32_t ox, oy; }; // coverage, DevPaint index, DrawTiler tile origin
struct HairBuilt {
    std::vector<HairGroup> groups;
    std::vector<HairDevBlit> blits;
    std::vector<rbh::DevPaint> paints;
    std::vector<rbh::DevStop> stops;
};
// painter.rs treat_as_hairline: the coverage the hairline is modulated with, or < 0 when the stroke is a real outline.
float rb_hairline_coverage(const rb_paint &paint, const rb_stroke &stroke, const rbh::Xform &ctm);
// Is draw i of the batch a hairline stroke?
bool rb_batch_draw_is_hairline(const rb_batch *b, size_t i);
// Builds the blits of the consecutive hairline draws [begin, end).
int rb_batch_hair_build(const rb_batch *b, size_t begin, size_t end, int W, int H, HairBuilt *out);

int rb_batch_record(rb_batch *b, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                    const rb_paint *paint, int32_t rule, const float ts[6]);
