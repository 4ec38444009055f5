// geo.cu — device path geometry (crates/resvg/src/path.rs:73,113; tiny-skia painter.rs fill_path /
// stroke_path): dashing, stroking, hairline walking and the fill front end (transform, y-monotone chop, clip, edge
// set-up) as CUDA kernels, one thread per draw, running the SAME source as the host builder (geom_common.h).
//
// Pipeline of one batch range (all on the context's stream):
//   k_geo_dash    thread per dashed stroke    Path::dash                                 -> the path to stroke / walk
//   k_geo_stroke  thread per stroke           PathStroker::stroke (local coordinates)    -> the outline to fill
//   k_geo_hair    thread per hairline stroke  hairline::stroke_path + hairline_aa        -> ordered blits, DevDraw
//   k_geo_fill    thread per fill / outline   fill_path up to the walker                 -> line edges, curve records, DevDraw
//   k_geo_wide    CTA per many-chain draw     exact bound of simultaneously active edges (packed winding range)
// Every kernel takes its task indices from a list the host sorted heaviest first, so that the long draws (dashed strokes:
// thousands of edges) start first and the short ones fill in behind them.  Dynamic storage comes from a bump heap
// (geom_common.h DVec); a launch that exhausts it is repeated with a larger heap.  The kernels leave exactly the block
// the host builder would have uploaded in item mode, so k_row_lists and everything after it run unchanged.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "batch.h"
#include "batch_geo.h"
#include "dasher_core.h"
#include "fill_core.h"
#include "hairline_core.h"
#include "rb_internal.h"
#include "stroker_core.h"

using geo::DVec;
using geo::GeoHeap;
using geo::P;

// The path a later stage consumes: status 0 = the recorded path, 1 = `verbs` / `pts` below, 2 = nothing is drawn.
struct ContourRec { // one measured contour of a dashed stroke (dasher_core.h Contour), kept for the threads that cut its dashes
    const geo::ds::Seg *segs;
    const P *pts;
    uint32_t n_segs, n_pts;
    float length;
    uint32_t closed;
};
struct GeoMid {
    const uint8_t *verbs;
    const P *pts;
    uint32_t n_verbs, n_pts;
    uint32_t status, pad;
    // dashed strokes built in units (GT_UNITS)
    const ContourRec *contours;
    uint32_t unit_first, unit_count;
    uint32_t plan_ok, n_list_row0;
    geo::fl::FillPlan fp;   // stroke outlines: the fill decisions taken from the bounds of the WHOLE outline
    rbh::IRect sect;        //   its blitter rectangle inside the target
    int32_t r0, nr;         //   its warp-tile rows
    geo::hl::Cull cull;     // hairlines: the culling decided from the bounds of the whole dashed path
    DevEdge *out_lines;     // the outline's merged edge items
    rbh::CurveRec *out_curves;
};
// One output contour of a dashed stroke: up to two "on" intervals of one measured contour (the second when the last dash of
// a closed contour runs into its first).  contour == ~0u: the whole recorded path (dash specification rejected).
struct GeoUnit {
    uint32_t task, contour;
    float a0, a1, b0, b1;
    uint32_t has_b, local;
};
struct GeoPiece { // what the pipeline knows about a unit
    const uint8_t *verbs; // its path: the dash (hairlines) or the outline of the dash (strokes)
    const P *pts;
    uint32_t n_verbs, n_pts;
    float l, t, r, b;     // bounds of its points in device space
    uint32_t finite, has_pts;
    rbh::Edge *lines;     // its edge items (packed in place into DevEdge / slots local to the unit by k_geo_unit_fill)
    rbh::CurveRec *curves;
    uint32_t ne, ncv, slots, n_list;
    uint32_t first_v, last_v; // its first / last item is a vertical line (what combine_vertical could merge across units)
    int32_t first_x, last_x;
    int32_t first_y0, first_y1, first_w, last_y0, last_y1, last_w;
    uint32_t line_base, curve_base, slot_base, pad;
};

struct GeoArgs {
    const GeoTask *tasks;
    const uint8_t *verbs;
    const P *pts;
    const float *dashes;
    GeoMid *mid;
    DevDraw *draws;
    GeoTotals *tot;
    GeoHeap heap;
    int W, H;
    GeoUnit *units;
    GeoPiece *pieces;
    uint32_t max_units;
    uint32_t lane_shift;     // log2 of the lanes a task occupies (one of them works)
    unsigned long long *dbg; // RB_GEO_TIMES: cycles per task of k_geo_stroke / k_geo_hair / k_geo_fill (3 arrays of n_tasks)
    uint32_t n_tasks;
};

// The geometry code is one long data-dependent control flow per draw (recursive subdivision, clipping cases): the 32 lanes
// of a warp would each take their own path and be executed one after the other, so a warp is no faster than a thread but
// ties up 32 draws.  Every draw therefore gets a WARP of its own with lane 0 working: the same issue slots, but up to 32
// times as many independent instruction streams per SM to hide latency with (measured on the 100 000-path scene: stroke
// kernel 17.7 -> see DESIGN.md).
constexpr int GEO_THREADS = 64;
constexpr int GEO_TASKS_PER_CTA = GEO_THREADS / 32;
// recursion guards (the device stack is GEO_STACK bytes per thread): deeper than this and the batch goes to the host builder
constexpr int GEO_STACK = 24576; // cubic_stroke recurses up to 78 levels of 232 bytes (the reference's own limit)

__device__ __forceinline__ void empty_draw_at(const GeoArgs &a, uint32_t draw, const GeoTask &t)
{
    DevDraw d;
    memset(&d, 0, sizeof(d));
    d.ox = t.ox; d.oy = t.oy;
    d.shift = 2;
    d.paint = t.paint;
    d.r0 = 1; // r0 > r0 + n_rows - 1: the binning kernels never see it
    d.n_rows = 0;
    d.row_base = (uint32_t)atomicAdd(&a.tot->n_row_off, 1ull);
    a.draws[draw] = d;
}
__device__ __forceinline__ void empty_draw(const GeoArgs &a, uint32_t, const GeoTask &t) { empty_draw_at(a, t.draw, t); }

// ---- Path::dash ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEO_THREADS) k_geo_dash(GeoArgs a, const uint32_t *__restrict__ list, uint32_t n)
{
    const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per task, one of them working: see GEO_THREADS
    if (li >= n) return;
    const uint32_t ti = list[li];
    const GeoTask t = a.tasks[ti];
    GeoHeap heap = a.heap;
    geo::ds::DashOut<DVec> pb;
    geo::ds::Contour<DVec> c;
    pb.verbs.init(&heap, t.hint);
    pb.pts.init(&heap, t.hint * 2);
    pb.move_required = true;
    pb.last_move = 0;
    c.segs.init(&heap, t.n_verbs * 8);
    c.pts.init(&heap, t.n_pts + 4);
    GeoMid m;
    m.verbs = nullptr; m.pts = nullptr; m.n_verbs = 0; m.n_pts = 0; m.status = 2; m.pad = 0;
    if (pb.verbs.ok() && pb.pts.ok() && c.segs.ok() && c.pts.ok()) {
        bool valid = false;
        const bool ok = geo::ds::dash_path(pb, c, a.verbs + t.verb_off, (int)t.n_verbs, a.pts + t.pt_off, a.dashes + t.dash_off, (int)t.n_dash,
                                           t.dash_offset, t.res_scale, &valid);
        if (!valid) m.status = 0; // StrokeDash::new -> None: the stroke stays solid
        else if (ok) {
            m.verbs = pb.verbs.data(); m.pts = pb.pts.data();
            m.n_verbs = (uint32_t)pb.verbs.size(); m.n_pts = (uint32_t)pb.pts.size();
            m.status = 1;
        }
    }
    a.mid[ti] = m;
}

// ---- PathStroker::stroke --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEO_THREADS, 10) k_geo_stroke(GeoArgs a, const uint32_t *__restrict__ list, uint32_t n)
{
    const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per task, one of them working: see GEO_THREADS
    if (li >= n) return;
    const uint32_t ti = list[li];
    const GeoTask t = a.tasks[ti];
    struct Tm { const GeoArgs &a; uint32_t ti; long long t0; __device__ ~Tm() { if (a.dbg) a.dbg[0 * (size_t)a.n_tasks + ti] = (unsigned long long)(clock64() - t0); } } tm__{a, ti, clock64()};
    const uint8_t *verbs = a.verbs + t.verb_off;
    const P *pts = a.pts + t.pt_off;
    int n_verbs = (int)t.n_verbs;
    GeoMid m;
    m.verbs = nullptr; m.pts = nullptr; m.n_verbs = 0; m.n_pts = 0; m.status = 2; m.pad = 0;
    if (t.flags & GT_DASH) {
        const GeoMid in = a.mid[ti];
        if (in.status == 2) { a.mid[ti] = m; return; } // "path dashing failed": nothing is drawn
        if (in.status == 1) { verbs = in.verbs; pts = in.pts; n_verbs = (int)in.n_verbs; }
    }
    GeoHeap heap = a.heap;
    geo::sk::Stroker<DVec> s;
    s.outer.verbs.init(&heap, t.hint);
    s.outer.pts.init(&heap, t.hint * 2);
    s.inner.verbs.init(&heap, t.hint / 2 + 8);
    s.inner.pts.init(&heap, t.hint + 8);
    s.cusper.verbs.init(&heap, 8);
    s.cusper.pts.init(&heap, 8);
    if (s.outer.verbs.ok() && s.outer.pts.ok() && s.inner.verbs.ok() && s.inner.pts.ok() && s.cusper.verbs.ok() && s.cusper.pts.ok()) {
        s.reset();
        if (geo::sk::stroke_path(s, verbs, n_verbs, pts, t.width, t.miter, (int)((t.flags >> GT_CAP_SHIFT) & 3u), (int)((t.flags >> GT_JOIN_SHIFT) & 3u),
                                 t.res_scale)) {
            m.verbs = s.outer.verbs.data(); m.pts = s.outer.pts.data();
            m.n_verbs = (uint32_t)s.outer.verbs.size(); m.n_pts = (uint32_t)s.outer.pts.size();
            m.status = 1;
        }
        if (s.too_deep) atomicOr(&a.tot->deep, 1u);
    }
    a.mid[ti] = m;
}

// The path's points in device space relative to the DrawTiler tile: tiny-skia's path.transform(ts) (map_points: identity /
// translate / scale + translate / affine, each its own expression) followed by the tile shift.
struct MapPts {
    const P *p;
    float sx, ky, kx, sy, tx, ty, ttx, tty;
    int mode; // 0 identity, 1 translate, 2 scale + translate, 3 affine
    bool tile;
    __device__ P operator[](int i) const
    {
        P q = p[i];
        if (mode == 1) { q.x += tx; q.y += ty; }
        else if (mode == 2) { q.x = q.x * sx + tx; q.y = q.y * sy + ty; }
        else if (mode == 3) {
            const float x = q.x * sx + q.y * kx + tx;
            const float y = q.x * ky + q.y * sy + ty;
            q.x = x; q.y = y;
        }
        if (tile) { q.x += ttx; q.y += tty; }
        return q;
    }
};
__device__ __forceinline__ MapPts map_for(const GeoTask &t, const P *p)
{
    MapPts m;
    m.p = p;
    m.sx = t.ctm[0]; m.ky = t.ctm[1]; m.kx = t.ctm[2]; m.sy = t.ctm[3]; m.tx = t.ctm[4]; m.ty = t.ctm[5];
    m.ttx = t.tile_tx; m.tty = t.tile_ty;
    m.tile = (t.flags & GT_TILE) != 0;
    m.mode = 0;
    if (t.flags & GT_MAP) {
        const bool skew = m.kx != 0 || m.ky != 0, scale = m.sx != 1 || m.sy != 1;
        m.mode = skew ? 3 : (scale ? 2 : ((m.tx != 0 || m.ty != 0) ? 1 : 0));
    }
    return m;
}

// The ordered blits of one hairline draw -> its DevDraw: blits outside the target dropped, bounding box, warp-tile cells, the
// rank of every blit inside its cell (k_row_lists copies blit k of a cell to the cell's start + rank).
__device__ void hair_emit(const GeoArgs &a, GeoHeap &heap, const GeoTask &t, uint32_t draw, DVec<geo::HairBlit> &blits)
{
    const int W = a.W, H = a.H, ox = t.ox, oy = t.oy;
    if (ox < 0 || oy < 0 || ox + t.tw > W || oy + t.th > H) { // keep the blits that land inside the target
        uint32_t keep = 0;
        for (uint32_t k = 0; k < blits.n; k++) {
            const geo::HairBlit hb = blits.p[k];
            if (hb.x + ox >= 0 && hb.y + oy >= 0 && hb.x + ox < W && hb.y + oy < H) blits.p[keep++] = hb;
        }
        blits.n = keep;
    }
    const uint32_t nb = blits.n;
    if (nb == 0) { empty_draw_at(a, draw, t); return; }
    if (nb >= (1u << 28)) { a.tot->too_large = 1u; empty_draw_at(a, draw, t); return; }
    int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
    for (uint32_t k = 0; k < nb; k++) {
        const geo::HairBlit hb = blits.p[k];
        x0 = min(x0, hb.x); x1 = max(x1, hb.x);
        y0 = min(y0, hb.y); y1 = max(y1, hb.y);
    }
    DevDraw d;
    memset(&d, 0, sizeof(d));
    d.ox = ox; d.oy = oy;
    d.sx = x0; d.sy = y0; d.sw = x1 - x0 + 1; d.sh = y1 - y0 + 1;
    d.shift = 2;
    d.rule = 2; // hairline
    d.paint = t.paint;
    const int r0 = (oy + y0) >> 3, r1 = (oy + y1) >> 3, nr = r1 - r0 + 1;
    const int c0 = (ox + x0) / 32, c1 = (ox + x1) / 32, ncols = c1 - c0 + 1;
    // the draw's "tile rows" are its warp-tile CELLS (row-major over its bounding box); rank = order of the blit inside its cell
    DVec<uint32_t> rank;
    rank.init(&heap, (uint32_t)nr * (uint32_t)ncols);
    DVec<DevEdge> out;
    out.init(&heap, nb);
    if (!rank.ok() || !out.ok()) { empty_draw_at(a, draw, t); return; }
    rank.resize((size_t)nr * (size_t)ncols);
    for (uint32_t k = 0; k < nb; k++) {
        const geo::HairBlit hb = blits.p[k];
        const uint32_t cell = (uint32_t)(((oy + hb.y) >> 3) - r0) * (uint32_t)ncols + (uint32_t)(((ox + hb.x) >> 5) - c0);
        DevEdge e; // a blit in an edge-sized record: layer pixel, coverage, rank inside its cell, cell
        e.x = (int32_t)((uint32_t)(hb.x + ox) | ((uint32_t)(hb.y + oy) << 16));
        e.dx = (int32_t)hb.alpha;
        e.ypack = rank.p[cell]++;
        e.meta = cell;
        out.p[k] = e;
    }
    d.curve_off = (uint32_t)c0; // hairline draws: first cell column / cells per row
    d.curve_cnt = (uint32_t)ncols;
    d.edge_off = 0;
    d.edge_cnt = 0;
    d.line_off = (uint32_t)(((const uint8_t *)out.p - heap.base) / sizeof(DevEdge));
    d.line_cnt = nb;
    d.r0 = (uint32_t)r0;
    d.n_rows = (uint32_t)nr;
    d.list_off = (uint32_t)atomicAdd(&a.tot->n_list, (unsigned long long)nb);
    d.row_base = (uint32_t)atomicAdd(&a.tot->n_row_off, (unsigned long long)nr * (unsigned long long)ncols + 1ull);
    d.list_cap = nb;
    atomicAdd(&a.tot->n_row_ent, (unsigned long long)nr);
    atomicAdd(&a.tot->n_wpairs, (unsigned long long)nr * (unsigned long long)ncols);
    a.draws[draw] = d;
}

// ---- hairline strokes -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEO_THREADS, 10) k_geo_hair(GeoArgs a, const uint32_t *__restrict__ list, uint32_t n)
{
    const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per task, one of them working: see GEO_THREADS
    if (li >= n) return;
    const uint32_t ti = list[li];
    const GeoTask t = a.tasks[ti];
    struct Tm { const GeoArgs &a; uint32_t ti; long long t0; __device__ ~Tm() { if (a.dbg) a.dbg[1 * (size_t)a.n_tasks + ti] = (unsigned long long)(clock64() - t0); } } tm__{a, ti, clock64()};
    const uint8_t *verbs = a.verbs + t.verb_off;
    const P *pts = a.pts + t.pt_off;
    int n_verbs = (int)t.n_verbs, n_pts = (int)t.n_pts;
    if (t.flags & GT_DASH) {
        const GeoMid in = a.mid[ti];
        if (in.status == 2) { empty_draw(a, ti, t); return; }
        if (in.status == 1) { verbs = in.verbs; pts = in.pts; n_verbs = (int)in.n_verbs; n_pts = (int)in.n_pts; }
    }
    GeoHeap heap = a.heap;
    DVec<geo::HairBlit> blits;
    blits.init(&heap, (t.flags & GT_DASH) ? t.hint * 16 : 256u);
    if (!blits.ok()) { empty_draw(a, ti, t); return; }
    const MapPts mp = map_for(t, pts);
    geo::hl::hairline_blits<DVec>(verbs, n_verbs, mp, n_pts, (int)((t.flags >> GT_CAP_SHIFT) & 3u), t.tw, t.th, blits, (t.flags & GT_DASH) ? -1 : t.sub);
    hair_emit(a, heap, t, t.draw, blits);
}

// ---- fill_path up to the scanline walker ------------------------------------------------------------------------------------------
struct NoEnds {
    uint32_t chains;
    __device__ void operator()(int32_t, int32_t) { chains++; }
};

__global__ void __launch_bounds__(GEO_THREADS, 10) k_geo_fill(GeoArgs a, const uint32_t *__restrict__ list, uint32_t n, uint32_t *__restrict__ wide_q)
{
    const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per task, one of them working: see GEO_THREADS
    if (li >= n) return;
    const uint32_t ti = list[li];
    const GeoTask t = a.tasks[ti];
    struct Tm { const GeoArgs &a; uint32_t ti; long long t0; __device__ ~Tm() { if (a.dbg) a.dbg[2 * (size_t)a.n_tasks + ti] = (unsigned long long)(clock64() - t0); } } tm__{a, ti, clock64()};
    const uint8_t *verbs = a.verbs + t.verb_off;
    const P *pts = a.pts + t.pt_off;
    int n_verbs = (int)t.n_verbs, n_pts = (int)t.n_pts;
    int rule = (t.flags & GT_EVENODD) ? 1 : 0;
    if (t.flags & GT_STROKE) {
        const GeoMid in = a.mid[ti];
        if (in.status != 1) { empty_draw(a, ti, t); return; }
        verbs = in.verbs; pts = in.pts; n_verbs = (int)in.n_verbs; n_pts = (int)in.n_pts;
        rule = 0;
    }
    GeoHeap heap = a.heap;
    DVec<rbh::Edge> lines;
    DVec<rbh::CurveRec> curves;
    geo::fl::Sink<DVec> sink;
    lines.init(&heap, t.hint);
    curves.init(&heap, t.hint);
    sink.kinds.init(&heap, t.hint * 2);
    if (!lines.ok() || !curves.ok() || !sink.kinds.ok()) { empty_draw(a, ti, t); return; }
    sink.out = &lines;
    sink.base = 0;
    sink.curves = &curves;
    sink.n_items = 0;
    const MapPts mp = map_for(t, pts);
    rbh::DrawGeom g;
    if (!geo::fl::build_items<DVec>(verbs, n_verbs, mp, n_pts, (t.flags & GT_AA) != 0, t.tw, t.th, sink, &g)) { empty_draw(a, ti, t); return; }
    const int W = a.W, H = a.H, ox = t.ox, oy = t.oy;
    if (ox < 0 || oy < 0 || ox + t.tw > W || oy + t.th > H) { // blitter rectangle ∩ target (tile-local coordinates)
        const int cx0 = max(g.sect.x, -ox), cy0 = max(g.sect.y, -oy);
        const int cx1 = min(g.sect.x + g.sect.w, W - ox), cy1 = min(g.sect.y + g.sect.h, H - oy);
        if (cx1 <= cx0 || cy1 <= cy0) { empty_draw(a, ti, t); return; }
        g.sect.x = cx0; g.sect.y = cy0; g.sect.w = cx1 - cx0; g.sect.h = cy1 - cy0;
    }
    const uint32_t ne = lines.n, ncv = curves.n;
    DevDraw d;
    memset(&d, 0, sizeof(d));
    d.ox = ox; d.oy = oy;
    d.sx = g.sect.x; d.sy = g.sect.y; d.sw = g.sect.w; d.sh = g.sect.h;
    d.shift = g.shift;
    d.rule = rule;
    d.paint = t.paint;
    const int r0 = (oy + g.sect.y) >> 3, r1 = (oy + g.sect.y + g.sect.h - 1) >> 3, nr = r1 - r0 + 1;
    const int c0 = (ox + g.sect.x) / 32, c1 = (ox + g.sect.x + g.sect.w - 1) / 32;
    // packed in place: a 16-byte DevEdge over the first half of each 32-byte Edge already consumed, curves where they are
    NoEnds ends{0};
    const geo::fl::Packed po = geo::fl::pack_items(lines.p, ne, curves.p, ncv, reinterpret_cast<DevEdge *>(lines.p), curves.p, g.shift, oy, r0, nr, ends);
    if (po.too_large) { a.tot->too_large = 1u; empty_draw(a, ti, t); return; }
    if (ends.chains >= 128u) wide_q[atomicAdd(&a.tot->n_wide_q, 1u)] = t.draw; // the exact bound is taken by k_geo_wide
    d.edge_cnt = po.slots;
    d.edge_off = (uint32_t)atomicAdd(&a.tot->n_slots, (unsigned long long)po.slots);
    d.line_off = (uint32_t)(((const uint8_t *)lines.p - heap.base) / sizeof(DevEdge));
    d.line_cnt = ne;
    d.curve_off = (uint32_t)(((const uint8_t *)curves.p - heap.base) / sizeof(rbh::CurveRec));
    d.curve_cnt = ncv;
    d.r0 = (uint32_t)r0;
    d.n_rows = (uint32_t)nr;
    d.list_off = (uint32_t)atomicAdd(&a.tot->n_list, (unsigned long long)po.n_list);
    d.row_base = (uint32_t)atomicAdd(&a.tot->n_row_off, (unsigned long long)nr + 1ull);
    d.list_cap = (uint32_t)po.n_list;
    atomicAdd(&a.tot->n_row_ent, (unsigned long long)nr);
    atomicAdd(&a.tot->n_wpairs, (unsigned long long)nr * (unsigned long long)(c1 - c0 + 1));
    a.draws[t.draw] = d;
}


// ---- dashed strokes, dash by dash ---------------------------------------------------------------------------------------------
// A dashed stroke is by far the longest draw (a hundred dashes, thousands of edges) and one thread per draw made the whole
// launch wait for it.  Dashing cuts the path into independent open contours: the stroker treats every contour on its own,
// and so does the fill front end up to one rule (combine_vertical looks at the last edge emitted, which may belong to the
// previous contour — checked below).  So: k_geo_plan measures the contours of a dashed stroke and lists the dashes
// ("units"); one thread per unit cuts the dash out, strokes it and takes its bounds (k_geo_unit_path); the fill decisions
// are taken per draw from the bounds of all its units (k_geo_unit_bounds); one thread per unit builds its edges
// (k_geo_unit_fill) or, for a hairline, walks it into a draw of its own (k_geo_unit_hair); k_geo_unit_merge lays the
// units' edge items out one after the other and k_geo_unit_pack moves them there.
struct CountRanges {
    uint32_t n;
    __device__ void operator()(float, float, bool mv) { if (mv || n == 0) n++; }
};
struct WriteRanges {
    GeoUnit *units;
    uint32_t task, contour, n;
    __device__ void operator()(float a, float b, bool mv)
    {
        if (mv || n == 0) {
            GeoUnit u;
            u.task = task; u.contour = contour; u.a0 = a; u.a1 = b; u.b0 = 0.0f; u.b1 = 0.0f; u.has_b = 0; u.local = n;
            units[n++] = u;
        } else {
            units[n - 1].b0 = a; units[n - 1].b1 = b; units[n - 1].has_b = 1;
        }
    }
};

__global__ void __launch_bounds__(GEO_THREADS) k_geo_plan(GeoArgs a, const uint32_t *__restrict__ list, uint32_t n)
{
    const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per task, one of them working: see GEO_THREADS
    if (li >= n) return;
    const uint32_t ti = list[li];
    const GeoTask t = a.tasks[ti];
    GeoMid m;
    memset(&m, 0, sizeof(m));
    m.status = 2;
    const float *dash = a.dashes + t.dash_off;
    const geo::ds::DashSpec sp = geo::ds::dash_spec(dash, (int)t.n_dash, t.dash_offset);
    GeoHeap heap = a.heap;
    uint32_t cnt = 0;
    if (!sp.valid) {
        // StrokeDash::new -> None: the stroke stays solid, one unit holding the recorded path
        cnt = 1;
        m.unit_first = atomicAdd(&a.tot->n_units, 1u);
        GeoUnit u;
        u.task = ti; u.contour = ~0u; u.a0 = u.a1 = u.b0 = u.b1 = 0.0f; u.has_b = 0; u.local = 0;
        a.units[m.unit_first] = u;
        m.unit_count = 1;
        m.status = 1;
    } else {
        DVec<ContourRec> recs;
        recs.init(&heap, 4);
        const float tolerance = 0.5f * (1.0f / t.res_scale);
        const uint8_t *verbs = a.verbs + t.verb_off;
        const P *pts = a.pts + t.pt_off;
        int vi = 0, pi = 0;
        float dash_count = 0.0f;
        bool ok = recs.ok();
        while (ok) {
            geo::ds::Contour<DVec> c;
            c.segs.init(&heap, t.n_verbs * 8);
            c.pts.init(&heap, t.n_pts + 4);
            if (!c.segs.ok() || !c.pts.ok()) { ok = false; break; }
            if (!geo::ds::next_contour(verbs, (int)t.n_verbs, pts, &vi, &pi, tolerance, &c)) break;
            dash_count += c.length * (float)(t.n_dash >> 1) / sp.interval_len;
            if (dash_count > 1000000.0f) { ok = false; break; } // dash_impl gives up: nothing is drawn
            ContourRec r;
            r.segs = c.segs.data(); r.pts = c.pts.data(); r.n_segs = (uint32_t)c.segs.size(); r.n_pts = (uint32_t)c.pts.size();
            r.length = c.length; r.closed = c.closed ? 1u : 0u;
            recs.push_back(r);
            CountRanges cr{0};
            geo::ds::dash_contour_ranges(sp, dash, (int)t.n_dash, c.length, c.closed, cr);
            cnt += cr.n;
        }
        if (ok && cnt > t.max_units) { atomicOr(&a.tot->deep, 2u); ok = false; } // the host's bound did not hold: the host builder takes the batch
        if (ok && cnt > 0) {
            m.unit_first = atomicAdd(&a.tot->n_units, cnt);
            uint32_t at = m.unit_first;
            for (uint32_t k = 0; k < recs.n; k++) {
                WriteRanges wr{a.units + at, ti, k, 0};
                geo::ds::dash_contour_ranges(sp, dash, (int)t.n_dash, recs.p[k].length, recs.p[k].closed != 0, wr);
                for (uint32_t j = 0; j < wr.n; j++) a.units[at + j].local = at + j - m.unit_first;
                at += wr.n;
            }
            m.unit_count = cnt;
            m.contours = recs.data();
            m.status = 1;
        } else cnt = 0;
    }
    a.mid[ti] = m;
    if (t.flags & GT_HAIR) { // every dash is a draw of its own: the reserved draws beyond the dashes there are stay empty
        for (uint32_t k = cnt; k < t.n_draws; k++) empty_draw_at(a, t.draw + k, t);
    } else if (cnt == 0) empty_draw(a, ti, t);
}

// The dash (hairlines) or its outline (strokes), and the bounds of its points in device space.
__global__ void __launch_bounds__(GEO_THREADS, 10) k_geo_unit_path(GeoArgs a)
{
    const uint32_t ui = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per unit, one of them working: see GEO_THREADS
    if (ui >= a.tot->n_units) return;
    const GeoUnit U = a.units[ui];
    const GeoTask t = a.tasks[U.task];
    GeoHeap heap = a.heap;
    GeoPiece pc;
    memset(&pc, 0, sizeof(pc));
    const uint8_t *verbs = a.verbs + t.verb_off;
    const P *pts = a.pts + t.pt_off;
    int n_verbs = (int)t.n_verbs, n_pts = (int)t.n_pts;
    bool ok = true;
    if (U.contour != ~0u) {
        const ContourRec r = a.mid[U.task].contours[U.contour];
        geo::ds::Contour<DVec> c;
        c.segs.p = const_cast<geo::ds::Seg *>(r.segs); c.segs.n = c.segs.cap = r.n_segs; c.segs.h = &heap;
        c.pts.p = const_cast<P *>(r.pts); c.pts.n = c.pts.cap = r.n_pts; c.pts.h = &heap;
        c.length = r.length; c.closed = r.closed != 0;
        geo::ds::DashOut<DVec> pb;
        pb.verbs.init(&heap, 16);
        pb.pts.init(&heap, 40);
        pb.move_required = true;
        pb.last_move = 0;
        ok = pb.verbs.ok() && pb.pts.ok();
        if (ok) {
            c.push_segment(U.a0, U.a1, true, pb);
            if (U.has_b) c.push_segment(U.b0, U.b1, false, pb);
            // The stroker ends every contour but the path's last through its move_to (finish_contour(false, false)) and the
            // last one with finish_contour(false, last_is_line), which shapes a square cap after a line differently: a dash
            // that is not the last one is followed by the next dash's move_to.
            if (!(t.flags & GT_HAIR) && U.local + 1 != a.mid[U.task].unit_count) { pb.verbs.push_back(geo::V_MOVE); pb.pts.push_back(P{0.0f, 0.0f}); }
            verbs = pb.verbs.data(); pts = pb.pts.data();
            n_verbs = (int)pb.verbs.size(); n_pts = (int)pb.pts.size();
        }
    }
    if (ok && !(t.flags & GT_HAIR)) {
        geo::sk::Stroker<DVec> s;
        s.outer.verbs.init(&heap, 48);
        s.outer.pts.init(&heap, 96);
        s.inner.verbs.init(&heap, 24);
        s.inner.pts.init(&heap, 48);
        s.cusper.verbs.init(&heap, 8);
        s.cusper.pts.init(&heap, 8);
        ok = s.outer.verbs.ok() && s.outer.pts.ok() && s.inner.verbs.ok() && s.inner.pts.ok() && s.cusper.verbs.ok() && s.cusper.pts.ok();
        if (ok) {
            s.reset();
            ok = geo::sk::stroke_path(s, verbs, n_verbs, pts, t.width, t.miter, (int)((t.flags >> GT_CAP_SHIFT) & 3u), (int)((t.flags >> GT_JOIN_SHIFT) & 3u),
                                      t.res_scale);
            if (s.too_deep) atomicOr(&a.tot->deep, 1u);
            verbs = s.outer.verbs.data(); pts = s.outer.pts.data();
            n_verbs = (int)s.outer.verbs.size(); n_pts = (int)s.outer.pts.size();
        }
    }
    if (ok && n_pts > 0 && n_verbs > 1) {
        pc.verbs = verbs; pc.pts = pts; pc.n_verbs = (uint32_t)n_verbs; pc.n_pts = (uint32_t)n_pts;
        const MapPts mp = map_for(t, pts);
        const geo::fl::PathBounds pb = geo::fl::path_bounds(mp, n_pts);
        pc.l = pb.l; pc.t = pb.t; pc.r = pb.r; pc.b = pb.b;
        pc.finite = pb.finite ? 1u : 0u;
        pc.has_pts = 1u;
    }
    a.pieces[ui] = pc;
}

// Per dashed stroke: the bounds of all its units -> the decisions fill_path / stroke_hairline take from the whole path.
__global__ void __launch_bounds__(GEO_THREADS) k_geo_unit_bounds(GeoArgs a, const uint32_t *__restrict__ list, uint32_t n)
{
    const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per task, one of them working: see GEO_THREADS
    if (li >= n) return;
    const uint32_t ti = list[li];
    const GeoTask t = a.tasks[ti];
    GeoMid *M = a.mid + ti;
    if (M->status != 1) return;
    bool any = false, finite = true;
    float l = 0, tt = 0, r = 0, b = 0;
    for (uint32_t k = 0; k < M->unit_count; k++) {
        const GeoPiece *pc = a.pieces + M->unit_first + k;
        if (!pc->has_pts) continue;
        finite = finite && pc->finite != 0;
        if (!any) { l = pc->l; tt = pc->t; r = pc->r; b = pc->b; any = true; }
        else { l = geo::gmin(l, pc->l); tt = geo::gmin(tt, pc->t); r = geo::gmax(r, pc->r); b = geo::gmax(b, pc->b); }
    }
    uint32_t ok = 0;
    if (any) {
        if (t.flags & GT_HAIR) {
            geo::hl::HairBounds hb;
            hb.l = l; hb.t = tt; hb.r = r; hb.b = b; hb.finite = finite;
            geo::hl::Cull cull;
            if (geo::hl::hair_plan(hb, (int)((t.flags >> GT_CAP_SHIFT) & 3u), t.tw, t.th, &cull)) { M->cull = cull; ok = 1; }
        } else {
            geo::fl::PathBounds pb;
            pb.l = l; pb.t = tt; pb.r = r; pb.b = b; pb.finite = finite;
            geo::fl::FillPlan fp;
            if (geo::fl::fill_plan(pb, (t.flags & GT_AA) != 0, t.tw, t.th, &fp)) {
                rbh::IRect sc = fp.sect;
                const int W = a.W, H = a.H, ox = t.ox, oy = t.oy;
                bool vis = true;
                if (ox < 0 || oy < 0 || ox + t.tw > W || oy + t.th > H) { // blitter rectangle ∩ target (tile-local coordinates)
                    const int cx0 = max(sc.x, -ox), cy0 = max(sc.y, -oy);
                    const int cx1 = min(sc.x + sc.w, W - ox), cy1 = min(sc.y + sc.h, H - oy);
                    if (cx1 <= cx0 || cy1 <= cy0) vis = false;
                    sc.x = cx0; sc.y = cy0; sc.w = cx1 - cx0; sc.h = cy1 - cy0;
                }
                rbh::DrawGeom g;
                g.shift = fp.shift;
                if (vis && geo::fl::finish_geom(fp.ir, fp.inside, t.th, &g)) {
                    M->fp = fp;
                    M->sect = sc;
                    M->r0 = (oy + sc.y) >> 3;
                    M->nr = ((oy + sc.y + sc.h - 1) >> 3) - M->r0 + 1;
                    ok = 1;
                }
            }
        }
    }
    M->plan_ok = ok;
    if (!ok) {
        if (t.flags & GT_HAIR) { for (uint32_t k = 0; k < M->unit_count; k++) empty_draw_at(a, t.draw + k, t); }
        else empty_draw(a, ti, t);
    }
}

// A dash of a hairline stroke: walked into a draw of its own.
__global__ void __launch_bounds__(GEO_THREADS) k_geo_unit_hair(GeoArgs a)
{
    const uint32_t ui = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per unit, one of them working: see GEO_THREADS
    if (ui >= a.tot->n_units) return;
    const GeoUnit U = a.units[ui];
    const GeoTask t = a.tasks[U.task];
    if (!(t.flags & GT_HAIR)) return;
    const GeoMid *M = a.mid + U.task;
    if (!M->plan_ok) return; // its draws were emptied by k_geo_unit_bounds
    const GeoPiece pc = a.pieces[ui];
    const uint32_t draw = t.draw + U.local;
    if (!pc.has_pts) { empty_draw_at(a, draw, t); return; }
    GeoHeap heap = a.heap;
    DVec<geo::HairBlit> blits;
    blits.init(&heap, 128);
    if (!blits.ok()) { empty_draw_at(a, draw, t); return; }
    const MapPts mp = map_for(t, pc.pts);
    geo::hl::hairline_walk<DVec>(pc.verbs, (int)pc.n_verbs, mp, (int)((t.flags >> GT_CAP_SHIFT) & 3u), t.tw, t.th, M->cull, blits);
    hair_emit(a, heap, t, draw, blits);
}

// The edge items of one dash outline.
__global__ void __launch_bounds__(GEO_THREADS) k_geo_unit_fill(GeoArgs a)
{
    const uint32_t ui = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per unit, one of them working: see GEO_THREADS
    if (ui >= a.tot->n_units) return;
    const GeoUnit U = a.units[ui];
    const GeoTask t = a.tasks[U.task];
    if (t.flags & GT_HAIR) return;
    const GeoMid *M = a.mid + U.task;
    GeoPiece *pc = a.pieces + ui;
    if (!M->plan_ok || !pc->has_pts) return; // ne = ncv = 0
    GeoHeap heap = a.heap;
    DVec<rbh::Edge> lines;
    DVec<rbh::CurveRec> curves;
    geo::fl::Sink<DVec> sink;
    lines.init(&heap, 32);
    curves.init(&heap, 32);
    sink.kinds.init(&heap, 64);
    if (!lines.ok() || !curves.ok() || !sink.kinds.ok()) return;
    sink.out = &lines;
    sink.base = 0;
    sink.curves = &curves;
    sink.n_items = 0;
    sink.shift = M->fp.shift;
    const MapPts mp = map_for(t, pc->pts);
    geo::fl::walk_verbs(pc->verbs, (int)pc->n_verbs, mp, M->fp.inside, t.tw, t.th, sink);
    const uint32_t ne = lines.n, ncv = curves.n;
    if (ne + ncv == 0) return;
    // what combine_vertical could have merged with the neighbouring units' edges
    const bool first_line = sink.kinds.p[0] == 0, last_line = sink.kinds.back() == 0;
    pc->first_v = (first_line && lines.p[0].dx == 0) ? 1u : 0u;
    pc->first_x = first_line ? lines.p[0].x : 0;
    pc->last_v = (last_line && lines.back().dx == 0) ? 1u : 0u;
    pc->last_x = last_line ? lines.back().x : 0;
    if (first_line) { pc->first_y0 = lines.p[0].first_y; pc->first_y1 = lines.p[0].last_y; pc->first_w = lines.p[0].winding; }
    if (last_line) { pc->last_y0 = lines.back().first_y; pc->last_y1 = lines.back().last_y; pc->last_w = lines.back().winding; }
    NoEnds ends{0};
    const geo::fl::Packed po = geo::fl::pack_items(lines.p, ne, curves.p, ncv, reinterpret_cast<DevEdge *>(lines.p), curves.p, M->fp.shift, t.oy, M->r0, M->nr, ends);
    if (po.too_large) { a.tot->too_large = 1u; return; }
    pc->lines = lines.p; pc->curves = curves.p;
    pc->ne = ne; pc->ncv = ncv; pc->slots = po.slots; pc->n_list = (uint32_t)po.n_list;
}

// Per dashed stroke: its units' items one after the other -> the draw.
__global__ void __launch_bounds__(GEO_THREADS) k_geo_unit_merge(GeoArgs a, const uint32_t *__restrict__ list, uint32_t n, uint32_t *__restrict__ wide_q)
{
    const uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per task, one of them working: see GEO_THREADS
    if (li >= n) return;
    const uint32_t ti = list[li];
    const GeoTask t = a.tasks[ti];
    if (t.flags & GT_HAIR) return;
    GeoMid *M = a.mid + ti;
    if (M->status != 1 || !M->plan_ok) return; // already an empty draw
    uint32_t ne = 0, ncv = 0, slots = 0;
    unsigned long long n_list = 0;
    bool have_prev = false, prev_v = false, sequential = false;
    int32_t prev_x = 0, prev_y0 = 0, prev_y1 = 0, prev_w = 0;
    for (uint32_t k = 0; k < M->unit_count; k++) {
        GeoPiece *pc = a.pieces + M->unit_first + k;
        if (pc->ne + pc->ncv == 0) continue;
        // edge_builder.rs combine_vertical merges a new vertical line with the last edge emitted when that is a vertical line
        // on the same x with a matching end.  Across units nobody looked: if it would have happened, the draw's edges are
        // built again below, by this thread, unit after unit into one list.
        if (have_prev && prev_v && pc->first_v && prev_x == pc->first_x) {
            rbh::Edge e, last;
            memset(&e, 0, sizeof(e)); memset(&last, 0, sizeof(last));
            e.x = pc->first_x; e.first_y = pc->first_y0; e.last_y = pc->first_y1; e.winding = pc->first_w;
            last.x = prev_x; last.first_y = prev_y0; last.last_y = prev_y1; last.winding = prev_w;
            if (geo::fl::Sink<DVec>::combine_vertical(e, last) != 0) sequential = true;
        }
        have_prev = true; prev_v = pc->last_v != 0; prev_x = pc->last_x; prev_y0 = pc->last_y0; prev_y1 = pc->last_y1; prev_w = pc->last_w;
        pc->line_base = ne; pc->curve_base = ncv; pc->slot_base = slots;
        ne += pc->ne; ncv += pc->ncv; slots += pc->slots; n_list += pc->n_list;
        if (slots >= (1u << 28)) { a.tot->too_large = 1u; empty_draw(a, ti, t); M->plan_ok = 0; return; }
    }
    if (sequential) {
        GeoHeap heap = a.heap;
        DVec<rbh::Edge> lines;
        DVec<rbh::CurveRec> curves;
        geo::fl::Sink<DVec> sink;
        lines.init(&heap, ne + 16);
        curves.init(&heap, ncv + 16);
        sink.kinds.init(&heap, ne + ncv + 16);
        M->plan_ok = 0; // the units' own lists are not used: k_geo_unit_pack skips this draw
        if (!lines.ok() || !curves.ok() || !sink.kinds.ok()) { empty_draw(a, ti, t); return; }
        sink.out = &lines;
        sink.base = 0;
        sink.curves = &curves;
        sink.n_items = 0;
        sink.shift = M->fp.shift;
        for (uint32_t k = 0; k < M->unit_count; k++) {
            const GeoPiece *pc = a.pieces + M->unit_first + k;
            if (!pc->has_pts) continue;
            const MapPts mp = map_for(t, pc->pts);
            geo::fl::walk_verbs(pc->verbs, (int)pc->n_verbs, mp, M->fp.inside, t.tw, t.th, sink);
        }
        const uint32_t sne = lines.n, sncv = curves.n;
        if (sne + sncv < 2) { empty_draw(a, ti, t); return; }
        NoEnds ends{0};
        const geo::fl::Packed po = geo::fl::pack_items(lines.p, sne, curves.p, sncv, reinterpret_cast<DevEdge *>(lines.p), curves.p, M->fp.shift, t.oy, M->r0, M->nr, ends);
        if (po.too_large) { a.tot->too_large = 1u; empty_draw(a, ti, t); return; }
        const rbh::IRect sc = M->sect;
        DevDraw d;
        memset(&d, 0, sizeof(d));
        d.ox = t.ox; d.oy = t.oy;
        d.sx = sc.x; d.sy = sc.y; d.sw = sc.w; d.sh = sc.h;
        d.shift = M->fp.shift;
        d.rule = 0;
        d.paint = t.paint;
        const int c0 = (t.ox + sc.x) / 32, c1 = (t.ox + sc.x + sc.w - 1) / 32;
        if (ends.chains >= 128u) wide_q[atomicAdd(&a.tot->n_wide_q, 1u)] = t.draw;
        d.edge_cnt = po.slots;
        d.edge_off = (uint32_t)atomicAdd(&a.tot->n_slots, (unsigned long long)po.slots);
        d.line_off = (uint32_t)(((const uint8_t *)lines.p - heap.base) / sizeof(DevEdge));
        d.line_cnt = sne;
        d.curve_off = (uint32_t)(((const uint8_t *)curves.p - heap.base) / sizeof(rbh::CurveRec));
        d.curve_cnt = sncv;
        d.r0 = (uint32_t)M->r0;
        d.n_rows = (uint32_t)M->nr;
        d.list_off = (uint32_t)atomicAdd(&a.tot->n_list, (unsigned long long)po.n_list);
        d.row_base = (uint32_t)atomicAdd(&a.tot->n_row_off, (unsigned long long)M->nr + 1ull);
        d.list_cap = (uint32_t)po.n_list;
        atomicAdd(&a.tot->n_row_ent, (unsigned long long)M->nr);
        atomicAdd(&a.tot->n_wpairs, (unsigned long long)M->nr * (unsigned long long)(c1 - c0 + 1));
        atomicAdd(&a.tot->n_seq, 1u);
        a.draws[t.draw] = d;
        return;
    }
    if (ne + ncv < 2) { empty_draw(a, ti, t); M->plan_ok = 0; return; } // BasicEdgeBuilder::build: fewer than two edge objects
    GeoHeap heap = a.heap;
    DVec<DevEdge> out_l;
    DVec<rbh::CurveRec> out_c;
    out_l.init(&heap, ne);
    out_c.init(&heap, ncv);
    if (!out_l.ok() || !out_c.ok()) { empty_draw(a, ti, t); M->plan_ok = 0; return; }
    M->out_lines = out_l.p;
    M->out_curves = out_c.p;
    const rbh::IRect sc = M->sect;
    DevDraw d;
    memset(&d, 0, sizeof(d));
    d.ox = t.ox; d.oy = t.oy;
    d.sx = sc.x; d.sy = sc.y; d.sw = sc.w; d.sh = sc.h;
    d.shift = M->fp.shift;
    d.rule = 0;
    d.paint = t.paint;
    const int c0 = (t.ox + sc.x) / 32, c1 = (t.ox + sc.x + sc.w - 1) / 32;
    if (ne + ncv >= 128u) wide_q[atomicAdd(&a.tot->n_wide_q, 1u)] = t.draw;
    d.edge_cnt = slots;
    d.edge_off = (uint32_t)atomicAdd(&a.tot->n_slots, (unsigned long long)slots);
    d.line_off = (uint32_t)(((const uint8_t *)out_l.p - heap.base) / sizeof(DevEdge));
    d.line_cnt = ne;
    d.curve_off = (uint32_t)(((const uint8_t *)out_c.p - heap.base) / sizeof(rbh::CurveRec));
    d.curve_cnt = ncv;
    d.r0 = (uint32_t)M->r0;
    d.n_rows = (uint32_t)M->nr;
    d.list_off = (uint32_t)atomicAdd(&a.tot->n_list, n_list);
    d.row_base = (uint32_t)atomicAdd(&a.tot->n_row_off, (unsigned long long)M->nr + 1ull);
    d.list_cap = (uint32_t)n_list;
    if (n_list > 0xfffffff0ull) a.tot->too_large = 1u;
    atomicAdd(&a.tot->n_row_ent, (unsigned long long)M->nr);
    atomicAdd(&a.tot->n_wpairs, (unsigned long long)M->nr * (unsigned long long)(c1 - c0 + 1));
    a.draws[t.draw] = d;
}

// Every unit moves its items to their place in the draw (slots become draw-wide).
__global__ void __launch_bounds__(GEO_THREADS) k_geo_unit_pack(GeoArgs a)
{
    const uint32_t ui = (blockIdx.x * blockDim.x + threadIdx.x) >> a.lane_shift;
    if (threadIdx.x & ((1u << a.lane_shift) - 1u)) return; // 2^lane_shift lanes per unit, one of them working: see GEO_THREADS
    if (ui >= a.tot->n_units) return;
    const GeoUnit U = a.units[ui];
    const GeoMid *M = a.mid + U.task;
    const GeoPiece pc = a.pieces[ui];
    if ((a.tasks[U.task].flags & GT_HAIR) || !M->plan_ok || pc.ne + pc.ncv == 0 || !M->out_lines) return;
    const DevEdge *src = reinterpret_cast<const DevEdge *>(pc.lines);
    DevEdge *dst = M->out_lines + pc.line_base;
    for (uint32_t i = 0; i < pc.ne; i++) {
        DevEdge e = src[i];
        e.meta += pc.slot_base << 4;
        dst[i] = e;
    }
    rbh::CurveRec *cd = M->out_curves + pc.curve_base;
    for (uint32_t i = 0; i < pc.ncv; i++) {
        rbh::CurveRec c = pc.curves[i];
        c.item += pc.slot_base;
        cd[i] = c;
    }
}

// ---- packed winding range ------------------------------------------------------------------------------------------------------------
// The tile kernel counts crossings in balanced base-256 digits: a draw is only safe there when fewer than 128 edges can
// be active on one sub-scanline.  Chains (a line, or a whole curve) never overlap themselves in y, so the largest number of
// chains covering one sub-scanline bounds it.  One CTA per queued draw: difference array over the draw's sub-scanlines in
// shared memory, chunk by chunk, then a block scan for the running maximum.
constexpr int GW_THREADS = 256, GW_CHUNK = 8192;
__global__ void __launch_bounds__(GW_THREADS) k_geo_wide(GeoArgs a, const uint32_t *__restrict__ wide_q)
{
    __shared__ int diff[GW_CHUNK + 1];
    __shared__ int warp_tot[GW_THREADS / 32], warp_max[GW_THREADS / 32];
    __shared__ int worst_s;
    const DevEdge *lines_base = reinterpret_cast<const DevEdge *>(a.heap.base);
    const rbh::CurveRec *curves_base = reinterpret_cast<const rbh::CurveRec *>(a.heap.base);
    const uint32_t nq = a.tot->n_wide_q;
    const int tid = threadIdx.x;
    for (uint32_t q = blockIdx.x; q < nq; q += gridDim.x) {
        const DevDraw D = a.draws[wide_q[q]];
        if (tid == 0) worst_s = 0;
        // the chains' own range
        int lo = INT32_MAX, hi = INT32_MIN;
        for (uint32_t i = tid; i < D.line_cnt; i += GW_THREADS) {
            const DevEdge E = lines_base[D.line_off + i];
            lo = min(lo, (int)(E.ypack & 0xffffu)); hi = max(hi, (int)(E.ypack >> 16));
        }
        for (uint32_t i = tid; i < D.curve_cnt; i += GW_THREADS) {
            const rbh::CurveRec C = curves_base[D.curve_off + i];
            const int ylast = (C.info & 1u) ? C.p[7] : C.p[5];
            lo = min(lo, (C.p[1] + 32) >> 6); hi = max(hi, ((ylast + 32) >> 6) - 1);
        }
        for (int d = 16; d >= 1; d >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d)); }
        if ((tid & 31) == 0) { warp_tot[tid >> 5] = lo; warp_max[tid >> 5] = hi; }
        __syncthreads();
        for (int w = 0; w < GW_THREADS / 32; w++) { lo = min(lo, warp_tot[w]); hi = max(hi, warp_max[w]); }
        __syncthreads();
        for (int base = lo; base <= hi; base += GW_CHUNK) {
            const int top = min(base + GW_CHUNK - 1, hi); // chunk = [base, top]
            for (int i = tid; i <= GW_CHUNK; i += GW_THREADS) diff[i] = 0;
            __syncthreads();
            auto add = [&](int f, int l) {
                if (l < f || l < base || f > top) return;
                atomicAdd(&diff[max(f, base) - base], 1);
                atomicAdd(&diff[min(l, top) + 1 - base], -1);
            };
            for (uint32_t i = tid; i < D.line_cnt; i += GW_THREADS) {
                const DevEdge E = lines_base[D.line_off + i];
                add((int)(E.ypack & 0xffffu), (int)(E.ypack >> 16));
            }
            for (uint32_t i = tid; i < D.curve_cnt; i += GW_THREADS) {
                const rbh::CurveRec C = curves_base[D.curve_off + i];
                const int ylast = (C.info & 1u) ? C.p[7] : C.p[5];
                add((C.p[1] + 32) >> 6, ((ylast + 32) >> 6) - 1);
            }
            __syncthreads();
            // running sum over the chunk: contiguous pieces per thread, warp scan of the piece sums
            const int per = GW_CHUNK / GW_THREADS;
            int s = 0, mx_local = INT32_MIN, run = 0;
            for (int k = 0; k < per; k++) s += diff[tid * per + k];
            int incl = s;
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= d) incl += v; }
            if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
            __syncthreads();
            int before = incl - s;
            for (int w = 0; w < (tid >> 5); w++) before += warp_tot[w];
            run = before; // sum of the chunk's entries before this thread's piece (entries cut at `base` already count from there)
            for (int k = 0; k < per; k++) { run += diff[tid * per + k]; mx_local = max(mx_local, run); }
            for (int d = 16; d >= 1; d >>= 1) mx_local = max(mx_local, __shfl_xor_sync(0xffffffffu, mx_local, d));
            if ((tid & 31) == 0) warp_max[tid >> 5] = mx_local;
            __syncthreads();
            if (tid == 0) {
                int m = worst_s;
                for (int w = 0; w < GW_THREADS / 32; w++) m = max(m, warp_max[w]);
                worst_s = m;
            }
            __syncthreads();
        }
        if (tid == 0 && worst_s >= 128) a.tot->wide = 1u;
        __syncthreads();
    }
}

// ---- host orchestration ----------------------------------------------------------------------------------------------------------
struct StageReq2 { rb_ctx *ctx; int status; };
static void *geo_stage_pinned(void *user, size_t bytes)
{
    StageReq2 *r = (StageReq2 *)user;
    void *p = nullptr;
    r->status = rb_staging(r->ctx, bytes, &p);
    return r->status == RB_OK ? p : nullptr;
}

// Builds draws [begin, end) of the batch on the device, in two steps so that the caller can do other work (build another
// range on the host threads, rasterise it) while the geometry kernels run: rb_geo_begin builds the tasks, uploads them and
// enqueues the kernels on the geometry stream; rb_geo_finish waits, repeats the launch with a larger heap if it ran out,
// and on RB_OK leaves the block in b->dev (b->lay describes it: item mode, lines and curves addressed from the heap base) —
// the caller allocates the raster scratch.  RB_GEO_FALLBACK: the range needs the host builder (a draw beyond the packed
// winding range, a recursion deeper than the device stack, memory); nothing is left allocated.
struct GeoPending {
    GeoBlock G;
    void *blk = nullptr;      // pinned staging (first attempt only)
    uint8_t *dev = nullptr;
    size_t heap_bytes = 0, o_draws = 0, o_tot = 0, o_heap = 0;
    int attempt = 0, W = 0, H = 0;
    unsigned long long *dbg = nullptr;
    double t_start = 0, t_built = 0, t_enq = 0;
    cudaStream_t gs = nullptr;
    // several launches of one batch may be in flight (rb_batch_submit enqueues the sub-ranges of its device share one after
    // the other): a FIFO hanging off rb_batch::geo, each with its own slot of the pinned totals and its own completion event
    GeoPending *next = nullptr;
    int slot = 0;
    cudaEvent_t done = nullptr;
};
static double geo_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int geo_enqueue(rb_ctx *ctx, GeoPending *gp)
{
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    cudaStream_t gs = gp->gs;
    const int W = gp->W, H = gp->H;
    {
        // [uploaded block | DevDraw[] | GeoMid[] | wide queue | units | pieces | totals | heap]
        const size_t n_tasks = gp->G.n_tasks;
        const size_t n_draws = gp->G.n_draws, max_units = gp->G.max_units;
        const size_t o_draws = al(gp->G.total), o_mid = o_draws + al(n_draws * sizeof(DevDraw)), o_wq = o_mid + al(n_tasks * sizeof(GeoMid));
        const size_t o_units = o_wq + al(n_draws * 4), o_pieces = o_units + al((max_units + 1) * sizeof(GeoUnit));
        const size_t o_tot = o_pieces + al((max_units + 1) * sizeof(GeoPiece)), o_heap = o_tot + 256, total = o_heap + gp->heap_bytes + 65536;
        uint8_t *dev = nullptr;
        if (cudaMallocAsync((void **)&dev, total, gs) != cudaSuccess) {
            cudaGetLastError();
            if (gp->dev) { cudaFreeAsync(gp->dev, gs); gp->dev = nullptr; }
            return RB_GEO_FALLBACK;
        }
        if (gp->attempt == 0) {
            RB_CUDA(ctx, cudaMemcpyAsync(dev, gp->blk, gp->G.total, cudaMemcpyHostToDevice, gs));
            ctx->h2d_bytes += gp->G.total;
            { int st__ = rb_staging_mark(ctx, gs); if (st__ != RB_OK) return st__; }
        } else {
            RB_CUDA(ctx, cudaMemcpyAsync(dev, gp->dev, gp->G.total, cudaMemcpyDeviceToDevice, gs));
            RB_CUDA(ctx, cudaFreeAsync(gp->dev, gs));
        }
        gp->dev = dev;
        RB_CUDA(ctx, cudaMemsetAsync(dev + o_tot, 0, 256, gs));
        GeoArgs a;
        a.tasks = (const GeoTask *)(dev + gp->G.o_tasks);
        a.verbs = dev + gp->G.o_verbs;
        a.pts = (const P *)(dev + gp->G.o_pts);
        a.dashes = (const float *)(dev + gp->G.o_dashes);
        a.mid = (GeoMid *)(dev + o_mid);
        a.draws = (DevDraw *)(dev + o_draws);
        a.tot = (GeoTotals *)(dev + o_tot);
        a.heap.base = dev + o_heap;
        a.heap.cursor = &a.tot->heap_cursor;
        a.heap.size = gp->heap_bytes;
        a.heap.overflow = &a.tot->overflow;
        a.W = W; a.H = H;
        a.units = (GeoUnit *)(dev + o_units);
        a.pieces = (GeoPiece *)(dev + o_pieces);
        a.max_units = (uint32_t)max_units;
        a.n_tasks = (uint32_t)n_tasks;
        static const int lane_shift = getenv("RB_GEO_LANE_SHIFT") ? atoi(getenv("RB_GEO_LANE_SHIFT")) : 3;
        a.lane_shift = (uint32_t)lane_shift;
        a.dbg = nullptr;
        unsigned long long *dbg = nullptr;
        if (getenv("RB_GEO_TIMES")) { cudaMalloc((void **)&dbg, 3 * n_tasks * 8); cudaMemset(dbg, 0, 3 * n_tasks * 8); a.dbg = dbg; }
        gp->dbg = dbg;
        const uint32_t *lists = (const uint32_t *)(dev + gp->G.o_lists);
        const uint32_t nd = (uint32_t)gp->G.n_dash_l, ns = (uint32_t)gp->G.n_stroke_l, nh = (uint32_t)gp->G.n_hair_l, nf = (uint32_t)gp->G.n_fill_l, nu = (uint32_t)gp->G.n_units_l;
        const uint32_t *l_dash = lists, *l_stroke = l_dash + nd, *l_hair = l_stroke + ns, *l_fill = l_hair + nh, *l_units = l_fill + nf;
        uint32_t *wide_q = (uint32_t *)(dev + o_wq);
        const size_t per_cta = (size_t)GEO_THREADS >> lane_shift;
        auto grid = [&](size_t n) { return (unsigned)((n + per_cta - 1) / per_cta); };
        const unsigned g_units = grid(max_units);
        // Four chains that do not depend on each other — the dashed strokes built in units; stroke -> fill of the outlines;
        // hairlines; plain fills — run side by side on streams of their own (every kernel here is bound by latency and
        // instruction fetch, not by issue slots), joined before the winding check.
        static const bool use_streams = getenv("RB_GEO_STREAMS") && atoi(getenv("RB_GEO_STREAMS")) != 0; // measured: no gain, every kernel fills the SMs' warp slots by itself
        cudaStream_t s0 = gs, s1 = s0, s2 = s0, s3 = s0;
        // dashed strokes beyond the unit bound (rare): dashed by one thread each, before anything that strokes or walks them
        if (nd) { k_geo_dash<<<grid(nd), GEO_THREADS, 0, s0>>>(a, l_dash, nd); RB_LAUNCHED(ctx, "geo_dash"); }
        if (use_streams) {
            s1 = ctx->geo_streams[1]; s2 = ctx->geo_streams[2]; s3 = ctx->geo_streams[3];
            RB_CUDA(ctx, cudaEventRecord(ctx->geo_events[0], s0));
            RB_CUDA(ctx, cudaStreamWaitEvent(s1, ctx->geo_events[0], 0));
            RB_CUDA(ctx, cudaStreamWaitEvent(s2, ctx->geo_events[0], 0));
            RB_CUDA(ctx, cudaStreamWaitEvent(s3, ctx->geo_events[0], 0));
        }
        const uint32_t no = (uint32_t)gp->G.n_outline_l;
        const uint32_t *l_outline = l_units + nu;
        // chain 1: the long strokes first
        if (ns) { k_geo_stroke<<<grid(ns), GEO_THREADS, 0, s1>>>(a, l_stroke, ns); RB_LAUNCHED(ctx, "geo_stroke"); }
        if (no) { k_geo_fill<<<grid(no), GEO_THREADS, 0, s1>>>(a, l_outline, no, wide_q); RB_LAUNCHED(ctx, "geo_fill_outlines"); }
        // chain 2: dashed strokes, dash by dash
        if (nu) {
            k_geo_plan<<<grid(nu), GEO_THREADS, 0, s2>>>(a, l_units, nu); RB_LAUNCHED(ctx, "geo_plan");
            k_geo_unit_path<<<g_units, GEO_THREADS, 0, s2>>>(a); RB_LAUNCHED(ctx, "geo_unit_path");
            k_geo_unit_bounds<<<grid(nu), GEO_THREADS, 0, s2>>>(a, l_units, nu); RB_LAUNCHED(ctx, "geo_unit_bounds");
            k_geo_unit_hair<<<g_units, GEO_THREADS, 0, s2>>>(a); RB_LAUNCHED(ctx, "geo_unit_hair");
            k_geo_unit_fill<<<g_units, GEO_THREADS, 0, s2>>>(a); RB_LAUNCHED(ctx, "geo_unit_fill");
            k_geo_unit_merge<<<grid(nu), GEO_THREADS, 0, s2>>>(a, l_units, nu, wide_q); RB_LAUNCHED(ctx, "geo_unit_merge");
            k_geo_unit_pack<<<g_units, GEO_THREADS, 0, s2>>>(a); RB_LAUNCHED(ctx, "geo_unit_pack");
        }
        // chain 3: hairlines
        if (nh) {
            k_geo_hair<<<grid(nh), GEO_THREADS, 0, s3>>>(a, l_hair, nh); RB_LAUNCHED(ctx, "geo_hair");
        }
        // chain 0: plain fills
        if (nf) { k_geo_fill<<<grid(nf), GEO_THREADS, 0, s0>>>(a, l_fill, nf, wide_q); RB_LAUNCHED(ctx, "geo_fill"); }
        if (use_streams) {
            RB_CUDA(ctx, cudaEventRecord(ctx->geo_events[1], s1)); RB_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->geo_events[1], 0));
            RB_CUDA(ctx, cudaEventRecord(ctx->geo_events[2], s2)); RB_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->geo_events[2], 0));
            RB_CUDA(ctx, cudaEventRecord(ctx->geo_events[3], s3)); RB_CUDA(ctx, cudaStreamWaitEvent(s0, ctx->geo_events[3], 0));
        }
        if (nf || no || nu) {
            k_geo_wide<<<std::min<uint32_t>((uint32_t)n_draws, (uint32_t)ctx->sm_count * 4u), GW_THREADS, 0, s0>>>(a, wide_q);
            RB_LAUNCHED(ctx, "geo_wide");
        }
        GeoTotals *ht = (GeoTotals *)((uint8_t *)ctx->geo_pinned + 128 * gp->slot);
        RB_CUDA(ctx, cudaMemcpyAsync(ht, dev + o_tot, sizeof(GeoTotals), cudaMemcpyDeviceToHost, gs));
        RB_CUDA(ctx, cudaEventRecord(gp->done, gs));
        gp->o_draws = o_draws; gp->o_tot = o_tot; gp->o_heap = o_heap;
        gp->t_enq = geo_now_ms();
    }
    return RB_OK;
}

int rb_geo_begin(rb_batch *b, int32_t n_threads, size_t begin, size_t end)
{
    rb_ctx *ctx = b->layer->ctx;
    cudaSetDevice(ctx->device);
    if (!(ctx->attr_bits & RB_ATTR_GEO)) {
        RB_CUDA(ctx, cudaDeviceSetLimit(cudaLimitStackSize, GEO_STACK));
        ctx->attr_bits |= RB_ATTR_GEO;
    }
    if (!ctx->geo_pinned) RB_CUDA(ctx, cudaHostAlloc(&ctx->geo_pinned, 4096, cudaHostAllocDefault));
    // The geometry runs on a stream of its own: the host waits for that stream only, and the context's stream is free to
    // rasterise another range of the batch meanwhile (rb_batch_submit).  What rb_geo_finish leaves behind is used on the
    // context's stream after that wait.
    for (auto &st : ctx->geo_streams) if (!st) RB_CUDA(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &ev : ctx->geo_events) if (!ev) RB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    static const bool own_stream = !(getenv("RB_GEO_OWN_STREAM") && atoi(getenv("RB_GEO_OWN_STREAM")) == 0);
    GeoPending *gp = new GeoPending();
    gp->gs = own_stream ? ctx->geo_streams[0] : ctx->stream;
    {
        int n_pending = 0;
        for (GeoPending *q = b->geo; q; q = q->next) n_pending++;
        if (n_pending >= 16) { delete gp; return RB_GEO_FALLBACK; }
        static std::atomic<int> slot_counter{0};
        gp->slot = slot_counter.fetch_add(1) & 15;
        for (GeoPending *q = b->geo; q; q = q->next) if (q->slot == gp->slot) { gp->slot = (gp->slot + 1) & 15; q = b->geo; }
    }
    if (cudaEventCreateWithFlags(&gp->done, cudaEventDisableTiming) != cudaSuccess) { delete gp; return RB_ERR_CUDA; }
    gp->W = (int)b->layer->w; gp->H = (int)b->layer->h;
    StageReq2 req{ctx, RB_OK};
    int st;
    gp->t_start = geo_now_ms();
    { rb_prof_scope prof__(RB_T_BUILD); st = rb_geo_host_build(b, gp->W, gp->H, n_threads, geo_stage_pinned, &req, &gp->blk, &gp->G, begin, end); }
    if (req.status != RB_OK) { cudaEventDestroy(gp->done); delete gp; return req.status; }
    if (st != RB_OK) { cudaEventDestroy(gp->done); delete gp; return rb_fail(ctx, st, "geometry task build failed"); }
    gp->t_built = geo_now_ms();
    auto unlink = [&]() { // on failure: take gp out of the FIFO again
        if (b->geo == gp) b->geo = gp->next;
        else for (GeoPending *q = b->geo; q; q = q->next) if (q->next == gp) { q->next = gp->next; break; }
        cudaEventDestroy(gp->done);
        delete gp;
    };
    if (!b->geo) b->geo = gp;
    else { GeoPending *q = b->geo; while (q->next) q = q->next; q->next = gp; }
    if (!gp->blk || gp->G.n_tasks == 0) return RB_OK; // nothing to draw: rb_geo_finish reports an empty layout
    rb_prof_scope prof_up__(RB_T_UPLOAD);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    gp->heap_bytes = al(gp->G.heap_hint + (64u << 20));
    if (const char *e = getenv("RB_GEO_HEAP_BYTES")) gp->heap_bytes = al((size_t)std::max(1024ll, atoll(e))); // tests: force heap retries
    st = geo_enqueue(ctx, gp);
    if (st != RB_OK) { unlink(); if (st == RB_GEO_FALLBACK) g_geo_counts[1]++; }
    return st;
}

int rb_geo_finish(rb_batch *b)
{
    rb_ctx *ctx = b->layer->ctx;
    GeoPending *gp = b->geo;
    if (!gp) return RB_ERR_INVALID;
    struct Drop { rb_batch *b; GeoPending *gp; ~Drop() { b->geo = gp->next; if (gp->done) cudaEventDestroy(gp->done); delete gp; } } drop{b, gp}; // pops the head
    cudaSetDevice(ctx->device);
    b->lay = BatchLayout();
    if (!gp->blk || gp->G.n_tasks == 0) return RB_OK;
    const bool diag = getenv("RB_GEO_DIAG") != nullptr;
    cudaStream_t gs = gp->gs;
    const GeoBlock &G = gp->G;
    const size_t n_tasks = G.n_tasks, n_draws = G.n_draws;
    const int W = gp->W, H = gp->H;
    for (;;) {
        const GeoTotals *ht = (const GeoTotals *)((const uint8_t *)ctx->geo_pinned + 128 * gp->slot);
        RB_CUDA(ctx, cudaEventSynchronize(gp->done));
        const double t_done = geo_now_ms();
        const GeoTotals T = *ht;
        if (gp->dbg) {
            unsigned long long *dbg = gp->dbg;
            std::vector<unsigned long long> h(3 * n_tasks);
            cudaMemcpy(h.data(), dbg, 3 * n_tasks * 8, cudaMemcpyDeviceToHost);
            cudaFree(dbg);
            gp->dbg = nullptr;
            const char *names[3] = {"stroke", "hair", "fill"};
            for (int k = 0; k < 3; k++) {
                std::vector<unsigned long long> v;
                for (size_t i = 0; i < n_tasks; i++) if (h[k * n_tasks + i]) v.push_back(h[k * n_tasks + i]);
                if (v.empty()) continue;
                std::sort(v.begin(), v.end());
                double sum = 0; for (auto x : v) sum += (double)x;
                auto q = [&](double f) { return (double)v[std::min(v.size() - 1, (size_t)(f * v.size()))] / 1.9e3; };
                fprintf(stderr, "[geo times] %s: n %zu, us: p50 %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f, sum %.1f ms\n", names[k], v.size(), q(0.5), q(0.9), q(0.99), q(0.999),
                        (double)v.back() / 1.9e3, sum / 1.9e6);
            }
        }
        if (diag)
            fprintf(stderr, "[geo] tasks %zu draws %zu (dash %zu stroke %zu hair %zu fill %zu + %zu units-tasks %zu units %u / %zu) upload %zu B heap %llu / %zu B slots %llu list %llu wide_q %u overflow %u wide %u deep %u too_large %u seq %u | host %.2f ms, enqueue %.2f ms, kernels done %.2f ms after the enqueue\n",
                    n_tasks, n_draws, G.n_dash_l, G.n_stroke_l, G.n_hair_l, G.n_fill_l, G.n_outline_l, G.n_units_l, T.n_units, G.max_units, G.total, T.heap_cursor, gp->heap_bytes, T.n_slots,
                    T.n_list, T.n_wide_q, T.overflow, T.wide, T.deep, T.too_large, T.n_seq, gp->t_built - gp->t_start, gp->t_enq - gp->t_built, t_done - gp->t_enq);
        if (T.overflow && gp->attempt < 5) { // heap exhausted: again with eight times the heap
            g_geo_counts[2]++;
            gp->attempt++;
            gp->heap_bytes *= 8;
            int st = geo_enqueue(ctx, gp);
            if (st == RB_GEO_FALLBACK) break;
            if (st != RB_OK) return st;
            continue;
        }
        if (T.overflow || T.wide || T.too_large || T.deep) break;
        if (T.n_slots > 0xfffffff0ull || T.n_list > 0xfffffff0ull || T.n_wpairs > 0xfffffff0ull || T.n_row_off > 0xfffffff0ull) break;
        BatchLayout L;
        L.items = true;
        L.n_draws = n_draws;
        L.n_paints = G.n_paints; L.n_stops = G.n_stops;
        L.o_draws = gp->o_draws; L.o_paints = G.o_paints; L.o_stops = G.o_stops;
        L.o_edges = gp->o_heap; L.o_curves = gp->o_heap;
        L.total = G.total;
        L.n_slots = (size_t)T.n_slots; L.n_list = (size_t)T.n_list; L.n_row_off = (size_t)T.n_row_off; L.n_row_ent = (size_t)T.n_row_ent;
        L.n_wpairs = (size_t)T.n_wpairs;
        L.has_hair = G.has_hair;
        L.wtiles_x = (W + 31) / 32;
        L.wtiles_y = (H + 7) / 8;
        L.tiles_x = (W + TW - 1) / TW;
        b->lay = L;
        b->dev = gp->dev;
        gp->dev = nullptr;
        b->stats[0] = n_draws; b->stats[1] = L.n_slots; b->stats[2] = L.n_wpairs; b->stats[3] = (size_t)L.wtiles_x * L.wtiles_y;
        b->stats[4] = G.total; b->stats[5] = 0;
        g_geo_counts[0]++;
        g_geo_counts[3] = (uint64_t)((t_done - gp->t_enq) * 1e3);
        g_geo_counts[4] = (uint64_t)((gp->t_built - gp->t_start) * 1e3);
        g_geo_counts[5] = n_tasks;
        return RB_OK;
    }
    if (gp->dev) { cudaFreeAsync(gp->dev, gs); cudaStreamSynchronize(gs); gp->dev = nullptr; }
    g_geo_counts[1]++;
    return RB_GEO_FALLBACK;
}

// Releases a geometry launch nobody will finish (the batch is destroyed or prepared again).
void rb_geo_abandon(rb_batch *b)
{
    while (GeoPending *gp = b->geo) {
        b->geo = gp->next;
        if (gp->dev) { cudaStreamSynchronize(gp->gs); cudaFreeAsync(gp->dev, gp->gs); }
        if (gp->dbg) cudaFree(gp->dbg);
        if (gp->done) cudaEventDestroy(gp->done);
        delete gp;
    }
}

int rb_geo_prepare(rb_batch *b, int32_t n_threads, size_t begin, size_t end)
{
    int st = rb_geo_begin(b, n_threads, begin, end);
    if (st != RB_OK) return st;
    return rb_geo_finish(b);
}
