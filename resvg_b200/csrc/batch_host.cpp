// batch_host.cpp — host half of a draw batch: recording (tiny-skia's fill_path / stroke_path arguments), edge
// building on host threads, tile binning, and the layout of the single block the device consumes.
//
// The reference does this work inside every PixmapMut::fill_path call (tiny-skia painter.rs → scan/path*.rs:
// transform, clip, build edges, sort) right before walking the edges; here it is done for a whole batch at once,
// in parallel over draws, and the result is shipped to the GPU in one copy.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "batch.h"
#include "batch_geo.h"
#include "fill_core.h"
#include "hairline.h"
#include "rb_internal.h"

extern "C" int rb_path_stroke(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, float width,
                              float miter_limit, int32_t cap, int32_t join, float res_scale, uint8_t **out_verbs,
                              int32_t *out_n_verbs, float **out_points, int32_t *out_n_points);
extern "C" void rb_path_free(void *p);
int rb_path_dash_into(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, const float *dash_array,
                      int32_t n_dash, float dash_offset, float res_scale, std::vector<uint8_t> &ov, std::vector<float> &op,
                      bool *spec_valid);
int rb_path_stroke_view(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, float width,
                        float miter_limit, int32_t cap, int32_t join, float res_scale, const uint8_t **out_verbs,
                        int32_t *out_n_verbs, const float **out_points, int32_t *out_n_points);

namespace {

using rbh::DevPaint;
using rbh::DevStop;

constexpr int kMaxDim = 8191;    // tiny-skia DrawTiler::MAX_DIMENSIONS
constexpr size_t kChunk = 128;   // draws per work item

typedef std::chrono::steady_clock Clock;
inline uint64_t us_since(Clock::time_point t0)
{
    return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - t0).count();
}

// Persistent worker threads: a large submit runs a few dozen short parallel phases, and creating 16 threads for each
// of them costs more than some of the phases themselves.  The pool is created on first use and never torn down
// (its threads sleep on a condition variable and die with the process).
class Pool {
public:
    static Pool &get()
    {
        static Pool *p = new Pool();
        return *p;
    }
    int size() const { return (int)threads_.size(); }
    // Runs body(t) for t in [0, nt) on nt pool threads and waits for all of them.
    void run(int nt, const std::function<void(int)> &body)
    {
        std::lock_guard<std::mutex> one_job(run_mu_);
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = &body;
            job_threads_ = nt;
            remaining_ = nt;
            generation_++;
        }
        cv_job_.notify_all();
        std::unique_lock<std::mutex> g(mu_);
        cv_done_.wait(g, [&] { return remaining_ == 0; });
        job_ = nullptr;
    }

private:
    Pool()
    {
        int n = (int)std::thread::hardware_concurrency();
        if (n < 1) n = 1;
        for (int t = 0; t < n; t++) {
            threads_.emplace_back([this, t] { loop(t); });
            threads_.back().detach();
        }
    }
    void loop(int t)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)> *job = nullptr;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_job_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                if (t < job_threads_) job = job_;
            }
            if (!job) continue;
            (*job)(t);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--remaining_ == 0) cv_done_.notify_all();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_job_, cv_done_;
    const std::function<void(int)> *job_ = nullptr;
    int job_threads_ = 0, remaining_ = 0;
    uint64_t generation_ = 0;
};

// Runs fn(item, worker) for item in [0, n) on `nt` threads with dynamic scheduling.
template <class F>
void parallel_for(int nt, size_t n, F fn)
{
    if (nt <= 1 || n <= 1) {
        for (size_t i = 0; i < n; i++) fn(i, 0);
        return;
    }
    Pool &pool = Pool::get();
    nt = std::min(nt, pool.size());
    std::atomic<size_t> next(0);
    pool.run(nt, [&](int t) {
        for (;;) {
            size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n) break;
            fn(i, t);
        }
    });
}

// Upper bound of |winding| for a draw: edges simultaneously active on one scanline.  Chains (an edge plus its
// curve continuations) never overlap themselves in y, so the chain count bounds it; only when that is not tight
// enough is the exact maximum swept.
bool draw_may_exceed_packed_winding(const rbh::Edge *e, size_t n)
{
    if (n < 128) return false;
    size_t chains = 0;
    for (size_t i = 0; i < n; i++) chains += e[i].prev < 0 ? 1 : 0;
    if (chains < 128) return false;
    std::vector<int32_t> ends;
    ends.reserve(n);
    for (size_t i = 0; i < n; i++) ends.push_back(e[i].last_y);
    std::sort(ends.begin(), ends.end());
    size_t active = 0, j = 0, worst = 0;
    for (size_t i = 0; i < n; i++) { // e is sorted by first_y
        while (j < n && ends[j] < e[i].first_y) { j++; active--; }
        active++;
        worst = std::max(worst, active);
    }
    return worst >= 128;
}

// painter.rs stroke_path: PathStroker::compute_resolution_scale(ts)
float resolution_scale(const rbh::Xform &t)
{
    float sx = sqrtf(t.sx * t.sx + t.kx * t.kx), sy = sqrtf(t.ky * t.ky + t.sy * t.sy);
    if (std::isfinite(sx) && std::isfinite(sy)) {
        float s = std::max(sx, sy);
        if (s > 0) return s;
    }
    return 1.0f;
}

// Per-worker output; the big vectors are recycled across builds (their pages stay mapped).
struct Worker {
    std::vector<DevEdge> edges;
    std::vector<DevDraw> draws;   // edge_off relative to the chunk's first edge
    std::vector<DevPaint> paints; // stop_off relative to the chunk's first stop
    std::vector<DevStop> stops;
    std::vector<rbh::CurveRec> curves; // item mode: recorded curves, `item` = draw-relative slot of the first segment
    std::vector<rbh::Edge> scratch;
    std::vector<rbh::CurveRec> cscratch;
    std::vector<int32_t> ends;
    std::vector<rbh::Pt> tmp, spts, bpts;
    std::vector<uint8_t> dverbs;
    std::vector<float> dpts, hstops;
    std::vector<rbh::HairBlit> hblits;
    std::vector<uint32_t> hrank;
    std::vector<uint8_t> sverbs;
    bool wide = false, skipped_hair = false, has_hair = false;
    bool too_large = false; // a draw the device structures cannot index (reported as RB_ERR_UNSUPPORTED, never dropped)
    void reset() { edges.clear(); draws.clear(); paints.clear(); stops.clear(); curves.clear(); wide = false; skipped_hair = false; has_hair = false; too_large = false; }
};

std::mutex g_pool_mu;
std::vector<std::unique_ptr<Worker>> g_pool;

std::unique_ptr<Worker> borrow_worker()
{
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (g_pool.empty()) return std::unique_ptr<Worker>(new Worker());
    std::unique_ptr<Worker> w = std::move(g_pool.back());
    g_pool.pop_back();
    return w;
}
void return_worker(std::unique_ptr<Worker> w)
{
    w->reset();
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (g_pool.size() < 256) g_pool.push_back(std::move(w));
}

struct ChunkInfo {
    int worker = 0;
    size_t e0 = 0, ne = 0, d0 = 0, nd = 0, p0 = 0, np = 0, s0 = 0, ns = 0; // ranges in the worker's vectors
    size_t ge = 0, gd = 0, gp = 0, gs = 0;                                  // global bases
    // warp-tile path: totals of this chunk and their global bases
    size_t n_list = 0, n_row_off = 0, n_row_ent = 0, n_wpairs = 0, g_list = 0, g_row_off = 0;
    // item mode: recorded curves and the slots of the device-side edge array
    size_t c0 = 0, nc = 0, gc = 0, n_slots = 0, g_slots = 0;
};

bool g_force_wide = false;
bool g_host_expand = false;

inline DevEdge pack_edge(const rbh::Edge &e)
{
    DevEdge d;
    d.x = e.x;
    d.dx = e.dx;
    d.ypack = ((uint32_t)e.first_y & 0xffffu) | ((uint32_t)e.last_y << 16);
    d.meta = (e.winding < 0 ? 1u : 0u) | (e.prev >= 0 ? 2u : 0u) | (e.before ? 4u : 0u) | (e.prev >= 0 ? ((uint32_t)e.prev << 4) : 0u);
    return d;
}

#ifdef RB_HOST_PROFILE
#include <x86intrin.h>
std::atomic<uint64_t> g_prof[4];
} namespace rbh { extern std::atomic<uint64_t> g_bd_prof[4]; } using rbh::g_bd_prof; namespace {
struct Tick { uint64_t &acc; uint64_t t0; Tick(uint64_t &a) : acc(a), t0(__rdtsc()) {} ~Tick() { acc += __rdtsc() - t0; } };
#define PROF(i) Tick tick__(prof_local[i])
#else
#define PROF(i)
#endif

} // namespace

// painter.rs treat_as_hairline
float rb_hairline_coverage(const rb_paint &paint, const rb_stroke &stroke, const rbh::Xform &ctm)
{
    auto fast_len = [](float x, float y) { x = fabsf(x); y = fabsf(y); if (x < y) std::swap(x, y); return x + y * 0.5f; };
    const float w = stroke.width;
    if (w == 0.0f) return 1.0f;
    if (!paint.anti_alias) return -1.0f;
    // the two stroke-width vectors mapped by the transform (translation ignored)
    const float len0 = fast_len(ctm.sx * w, ctm.ky * w), len1 = fast_len(ctm.kx * w, ctm.sy * w);
    if (len0 <= 1.0f && len1 <= 1.0f) return (len0 + len1) * 0.5f;
    return -1.0f;
}

namespace {

struct DrawRef {
    const uint8_t *verbs;
    const rbh::Pt *pts;
    int n_verbs, n_pts;
    rb_paint paint;
    const float *stops;
    rbh::Xform ctm;
    bool is_stroke;
    rb_stroke stroke; // dash_array resolved
    int32_t vp_x, vp_y, vp_w, vp_h;
};
// The recorded form of draw i (fill points of bulk segments are NOT transformed here).
DrawRef resolve_draw(const rb_batch *b, size_t i)
{
    size_t lo = 0, hi = b->spans.size() - 1;
    while (lo < hi) { // last span with start <= i
        const size_t mid = (lo + hi + 1) >> 1;
        if (b->spans[mid].start <= i) lo = mid; else hi = mid - 1;
    }
    const DrawSpan &sp = b->spans[lo];
    DrawRef d;
    d.vp_x = sp.vp_x; d.vp_y = sp.vp_y; d.vp_w = sp.vp_w; d.vp_h = sp.vp_h;
    if (sp.bulk < 0) {
        const RecordedDraw &r = b->recs[sp.first + (i - sp.start)];
        d.verbs = b->verbs.data() + r.verb_off;
        d.pts = b->pts.data() + r.pt_off;
        d.n_verbs = (int)r.n_verbs;
        d.n_pts = (int)r.n_pts;
        d.paint = r.paint;
        d.stops = r.n_stops ? b->stops.data() + r.stop_off : nullptr;
        d.ctm = r.ctm;
        d.is_stroke = r.is_stroke;
        d.stroke = r.stroke;
        d.stroke.dash_array = r.stroke.n_dash > 0 ? b->dashes.data() + r.dash_off : nullptr;
    } else {
        const BulkSeg &bs = b->bulk[(size_t)sp.bulk];
        const size_t k = i - sp.start;
        d.verbs = bs.verbs + bs.verb_off[k];
        d.pts = reinterpret_cast<const rbh::Pt *>(bs.points) + bs.point_off[k];
        d.n_verbs = (int)(bs.verb_off[k + 1] - bs.verb_off[k]);
        d.n_pts = (int)(bs.point_off[k + 1] - bs.point_off[k]);
        d.paint = bs.paints[k];
        d.stops = (d.paint.stops && d.paint.n_stops > 0) ? d.paint.stops : nullptr;
        d.ctm = bs.ctm;
        d.is_stroke = bs.strokes && bs.strokes[k].width > 0.0f;
        if (d.is_stroke) d.stroke = bs.strokes[k];
        else memset(&d.stroke, 0, sizeof(d.stroke));
    }
    return d;
}

} // namespace

// painter.rs stroke_path, hairline branch: coverage < 1 is folded into the paint's alpha when the blend mode pre-scales
// coverage ("the old technique"): scale = (coverage * 256) as i32; alpha' = (255 * scale) >> 8;
// shader.apply_opacity(alpha' / 255).  `scaled` receives the modified gradient stops when there are any.
static void hairline_modulate_paint(rb_paint *paint, float coverage, std::vector<float> &scaled)
{
    if (coverage == 1.0f) return;
    const int m = paint->blend_mode;
    const bool pre_scales = m == 2 || m == 4 || m == 12 || m == 8 || m == 9 || m == 3 || m == 11;
    if (!pre_scales) return;
    const int scale = (int)(coverage * 256.0f);
    const float opacity = (float)((255 * scale) >> 8) / 255.0f;
    auto mul = [&](float a) { return std::min(std::max(a * opacity, 0.0f), 1.0f); };
    if (paint->shader == 0) paint->color[3] = mul(paint->color[3]);
    else if (paint->shader == 3) paint->opacity = mul(paint->opacity);
    else if (paint->stops && paint->n_stops > 0) {
        scaled.assign(paint->stops, paint->stops + (size_t)paint->n_stops * 5);
        for (int k = 0; k < paint->n_stops; k++) scaled[(size_t)k * 5 + 4] = mul(scaled[(size_t)k * 5 + 4]);
        paint->stops = scaled.data();
    }
}

bool rb_batch_draw_is_hairline(const rb_batch *b, size_t i)
{
    if (b->n_hair == 0) return false;
    const DrawRef d = resolve_draw(b, i);
    return d.is_stroke && rb_hairline_coverage(d.paint, d.stroke, d.ctm) >= 0.0f;
}

// stroke_path for a hairline (tiny-skia painter.rs): dash, transform into device space, modulate the paint's alpha by
// the hairline coverage, walk every DrawTiler tile, then group the blits by pixel keeping their order.
int rb_batch_hair_build(const rb_batch *b, size_t begin, size_t end, int W, int H, HairBuilt *out)
{
    struct Raw { uint32_t x, y, alpha, paint; int32_t ox, oy; };
    std::vector<Raw> raw;
    std::vector<rbh::HairBlit> blits;
    std::vector<uint8_t> dverbs;
    std::vector<float> dpts;
    std::vector<rbh::Pt> dev, tmp;
    std::vector<float> stops_scaled;
    for (size_t i = begin; i < end; i++) {
        DrawRef d = resolve_draw(b, i);
        if (!d.is_stroke) continue;
        const float coverage = rb_hairline_coverage(d.paint, d.stroke, d.ctm);
        if (coverage < 0.0f) continue;
        const uint8_t *verbs = d.verbs;
        const rbh::Pt *pts = d.pts;
        int n_verbs = d.n_verbs, n_pts = d.n_pts;
        if (d.stroke.n_dash > 0 && d.stroke.dash_array) {
            bool valid = false;
            float sx = sqrtf(d.ctm.sx * d.ctm.sx + d.ctm.kx * d.ctm.kx), sy = sqrtf(d.ctm.ky * d.ctm.ky + d.ctm.sy * d.ctm.sy);
            float res = (std::isfinite(sx) && std::isfinite(sy) && std::max(sx, sy) > 0) ? std::max(sx, sy) : 1.0f;
            int st = rb_path_dash_into(verbs, n_verbs, &pts[0].x, n_pts, d.stroke.dash_array, d.stroke.n_dash, d.stroke.dash_offset, res,
                                       dverbs, dpts, &valid);
            if (valid) {
                if (st != RB_OK) continue;
                verbs = dverbs.data();
                n_verbs = (int)dverbs.size();
                pts = reinterpret_cast<const rbh::Pt *>(dpts.data());
                n_pts = (int)(dpts.size() / 2);
            }
        }
        dev.assign(pts, pts + n_pts);
        rbh::map_points(d.ctm, dev.data(), n_pts);
        rb_paint paint = d.paint;
        paint.stops = d.stops;
        hairline_modulate_paint(&paint, coverage, stops_scaled);
        // the viewport is the document's own pixmap; it may reach beyond the target (only what falls inside is drawn)
        int VX = 0, VY = 0, VW = W, VH = H;
        if (d.vp_w > 0) {
            VX = d.vp_x; VY = d.vp_y; VW = d.vp_w; VH = d.vp_h;
            if (VX >= W || VY >= H || (int64_t)VX + VW <= 0 || (int64_t)VY + VH <= 0) continue;
        }
        for (int ty = 0; ty < VH; ty += kMaxDim) {
            for (int tx = 0; tx < VW; tx += kMaxDim) {
                const int tw = std::min(VW - tx, kMaxDim), th = std::min(VH - ty, kMaxDim);
                const rbh::Pt *p = dev.data();
                rbh::Xform ctm = d.ctm;
                if (tx || ty) {
                    tmp = dev;
                    rbh::Xform tr;
                    tr.tx = -(float)tx;
                    tr.ty = -(float)ty;
                    rbh::map_points(tr, tmp.data(), n_pts);
                    p = tmp.data();
                    ctm = rbh::post_concat(ctm, tr);
                }
                blits.clear();
                rbh::hairline_blits(verbs, n_verbs, &p[0].x, n_pts, d.stroke.cap, tw, th, blits);
                if (blits.empty()) continue;
                rbh::DevPaint P;
                if (!rbh::prepare_paint(&paint, ctm, &P, out->stops)) continue;
                const uint32_t pi = (uint32_t)out->paints.size();
                out->paints.push_back(P);
                for (const rbh::HairBlit &hb : blits) {
                    const int lx = hb.x + VX + tx, ly = hb.y + VY + ty;
                    if (lx < 0 || ly < 0 || lx >= W || ly >= H) continue; // outside the target
                    raw.push_back(Raw{(uint32_t)lx, (uint32_t)ly, hb.alpha, pi, VX + tx, VY + ty});
                }
            }
        }
    }
    // group by pixel, keeping the blit order inside each group
    std::vector<uint32_t> order(raw.size());
    for (size_t i = 0; i < raw.size(); i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) {
        const uint64_t ka = (uint64_t)raw[a].y * (uint64_t)W + raw[a].x, kc = (uint64_t)raw[c].y * (uint64_t)W + raw[c].x;
        return ka < kc;
    });
    out->blits.reserve(raw.size());
    for (size_t k = 0; k < order.size(); k++) {
        const Raw &r = raw[order[k]];
        if (out->groups.empty() || out->groups.back().x != r.x || out->groups.back().y != r.y)
            out->groups.push_back(HairGroup{r.x, r.y, (uint32_t)out->blits.size(), 0});
        out->groups.back().count++;
        out->blits.push_back(HairDevBlit{r.alpha, r.paint, r.ox, r.oy});
    }
    return RB_OK;
}

namespace {

// Upper bound of simultaneously active edges from the y ranges of the chains (a line, or a whole curve): used in item
// mode, where the curve segments do not exist on the host.  `iv` holds (first_y, last_y) pairs.
bool chains_may_exceed_packed_winding(std::vector<int32_t> &iv)
{
    const size_t n = iv.size() / 2;
    if (n < 128) return false;
    std::vector<std::pair<int32_t, int32_t>> ev;
    ev.reserve(2 * n);
    for (size_t i = 0; i < n; i++) { ev.emplace_back(iv[2 * i], 1); ev.emplace_back(iv[2 * i + 1] + 1, -1); }
    std::sort(ev.begin(), ev.end());
    int active = 0, worst = 0;
    for (auto &e : ev) { active += e.second; worst = std::max(worst, active); }
    return worst >= 128;
}

void build_chunk(const rb_batch *b, size_t begin, size_t end, int W, int H, bool mask_target, bool items, bool hair_inline, Worker *out,
                 ChunkInfo *ci)
{
    ci->c0 = out->curves.size();
    ci->e0 = out->edges.size();
    ci->d0 = out->draws.size();
    ci->p0 = out->paints.size();
    ci->s0 = out->stops.size();
#ifdef RB_HOST_PROFILE
    uint64_t prof_local[4] = {0, 0, 0, 0};
    struct Flush { uint64_t *p; ~Flush() { for (int i = 0; i < 4; i++) g_prof[i] += p[i]; } } flush__{prof_local};
#endif
    size_t span_i = 0;
    RecordedDraw bulk_rec;
    for (size_t i = begin; i < end; i++) {
        // locate the draw: an individually recorded one, or entry k of a bulk segment (used in place)
        while (i >= b->spans[span_i].start + b->spans[span_i].count) span_i++;
        while (i < b->spans[span_i].start) span_i--;
        const DrawSpan &sp = b->spans[span_i];
        // the rectangle of the target this draw is rendered into, as if it were a pixmap of its own
        // (it may reach beyond the target: the draw is built against the whole viewport, so curves are clipped and
        // flattened exactly as in a render of the whole document, and only the part inside the target is binned)
        int VX = 0, VY = 0, VW = W, VH = H;
        if (sp.vp_w > 0) {
            VX = sp.vp_x; VY = sp.vp_y; VW = sp.vp_w; VH = sp.vp_h;
            if (VX >= W || VY >= H || (int64_t)VX + VW <= 0 || (int64_t)VY + VH <= 0) continue;
        }
        const RecordedDraw *rp;
        const uint8_t *verbs;
        const rbh::Pt *rpts;
        const float *stops_src;
        if (sp.bulk < 0) {
            rp = &b->recs[sp.first + (i - sp.start)];
            verbs = b->verbs.data() + rp->verb_off;
            rpts = b->pts.data() + rp->pt_off;
            stops_src = rp->n_stops ? b->stops.data() + rp->stop_off : nullptr;
        } else {
            const BulkSeg &bs = b->bulk[(size_t)sp.bulk];
            const size_t k = i - sp.start;
            RecordedDraw &r = bulk_rec;
            r.verb_off = 0; r.pt_off = 0; r.stop_off = 0;
            r.n_verbs = bs.verb_off[k + 1] - bs.verb_off[k];
            r.n_pts = bs.point_off[k + 1] - bs.point_off[k];
            r.paint = bs.paints[k];
            r.n_stops = (r.paint.stops && r.paint.n_stops > 0) ? (uint32_t)r.paint.n_stops : 0u;
            stops_src = r.n_stops ? r.paint.stops : nullptr;
            r.ctm = bs.ctm;
            r.is_stroke = bs.strokes && bs.strokes[k].width > 0.0f;
            if (r.is_stroke) r.stroke = bs.strokes[k];
            r.rule = r.is_stroke ? 0 : (bs.fill_rules[k] ? 1 : 0);
            verbs = bs.verbs + bs.verb_off[k];
            rpts = reinterpret_cast<const rbh::Pt *>(bs.points) + bs.point_off[k];
            if (!r.is_stroke && !bs.ctm.is_identity()) { // painter.rs: path.transform(ts)
                out->bpts.assign(rpts, rpts + r.n_pts);
                rbh::map_points(bs.ctm, out->bpts.data(), (int)r.n_pts);
                rpts = out->bpts.data();
            }
            rp = &r;
        }
        const RecordedDraw &r = *rp;
        int n_verbs = (int)r.n_verbs, n_pts = (int)r.n_pts, rule = r.rule;
        {
            // Cull draws that cannot touch the target before any geometry work (stroking dominates the host build): the
            // path lies inside the hull of its control points; a stroke reaches at most half its width times
            // max(miter limit, sqrt 2) beyond it (miter tips / square caps); anti-aliasing adds at most a pixel.
            float bx0 = INFINITY, by0 = INFINITY, bx1 = -INFINITY, by1 = -INFINITY;
            bool finite = n_pts > 0;
            for (int k = 0; k < n_pts; k++) {
                const float x = rpts[k].x, y = rpts[k].y;
                finite = finite && std::isfinite(x) && std::isfinite(y);
                bx0 = std::min(bx0, x); bx1 = std::max(bx1, x);
                by0 = std::min(by0, y); by1 = std::max(by1, y);
            }
            if (finite) {
                if (r.is_stroke) { // stroke points are in local space: inflate there, then map the box
                    const float reach = 0.5f * r.stroke.width * std::max(r.stroke.miter_limit, 1.4143f);
                    if (std::isfinite(reach)) {
                        rbh::Pt c[4] = {{bx0 - reach, by0 - reach}, {bx1 + reach, by0 - reach}, {bx0 - reach, by1 + reach}, {bx1 + reach, by1 + reach}};
                        rbh::map_points(r.ctm, c, 4);
                        bx0 = by0 = INFINITY; bx1 = by1 = -INFINITY;
                        for (const rbh::Pt &q : c) {
                            finite = finite && std::isfinite(q.x) && std::isfinite(q.y);
                            bx0 = std::min(bx0, q.x); bx1 = std::max(bx1, q.x);
                            by0 = std::min(by0, q.y); by1 = std::max(by1, q.y);
                        }
                    } else {
                        finite = false;
                    }
                }
                const float pad = 2.0f; // viewport-local window of the target: [-VX, W - VX) x [-VY, H - VY)
                if (finite && (bx1 < (float)(-VX) - pad || by1 < (float)(-VY) - pad || bx0 > (float)(W - VX) + pad || by0 > (float)(H - VY) + pad))
                    continue;
            }
        }
        if (r.is_stroke && b->n_hair) {
            const float coverage = rb_hairline_coverage(r.paint, r.stroke, r.ctm);
            if (coverage >= 0.0f) {
                PROF(3);
                // A hairline stroke: not scan-converted.  Its ordered blits become the draw's "edge list" (k_row_lists keeps
                // their order per tile row through the rank stored with each), and the tile kernel applies them one by one.
                if (mask_target) continue;
                if (!hair_inline) { out->skipped_hair = true; continue; } // the caller draws them in separate passes
                if (r.stroke.n_dash > 0) {
                    const float *da = sp.bulk < 0 ? b->dashes.data() + r.dash_off : r.stroke.dash_array;
                    bool valid = false;
                    int dst_ = da ? rb_path_dash_into(verbs, n_verbs, &rpts[0].x, n_pts, da, r.stroke.n_dash, r.stroke.dash_offset,
                                                      resolution_scale(r.ctm), out->dverbs, out->dpts, &valid)
                                  : RB_ERR_INVALID;
                    if (valid) {
                        if (dst_ != RB_OK) continue;
                        verbs = out->dverbs.data();
                        n_verbs = (int)out->dverbs.size();
                        rpts = reinterpret_cast<const rbh::Pt *>(out->dpts.data());
                        n_pts = (int)(out->dpts.size() / 2);
                    }
                }
                out->spts.assign(rpts, rpts + n_pts);
                rbh::map_points(r.ctm, out->spts.data(), n_pts);
                rb_paint hp = r.paint;
                hp.stops = stops_src;
                hairline_modulate_paint(&hp, coverage, out->hstops);
                for (int ty = 0; ty < VH; ty += kMaxDim) {
                    for (int tx = 0; tx < VW; tx += kMaxDim) {
                        const int tw = std::min(VW - tx, kMaxDim), th = std::min(VH - ty, kMaxDim);
                        const int ox = VX + tx, oy = VY + ty; // tile origin inside the target
                        const rbh::Pt *p = out->spts.data();
                        rbh::Xform ctm = r.ctm;
                        if (tx || ty) {
                            out->tmp = out->spts;
                            rbh::Xform tr;
                            tr.tx = -(float)tx;
                            tr.ty = -(float)ty;
                            rbh::map_points(tr, out->tmp.data(), n_pts);
                            p = out->tmp.data();
                            ctm = rbh::post_concat(ctm, tr);
                        }
                        if (ox >= W || oy >= H || ox + tw <= 0 || oy + th <= 0) continue; // tile wholly outside the target
                        out->hblits.clear();
                        rbh::hairline_blits(verbs, n_verbs, &p[0].x, n_pts, r.stroke.cap, tw, th, out->hblits);
                        if (ox < 0 || oy < 0 || ox + tw > W || oy + th > H) { // keep the blits that land inside the target
                            size_t keep = 0;
                            for (const rbh::HairBlit &hb : out->hblits)
                                if (hb.x + ox >= 0 && hb.y + oy >= 0 && hb.x + ox < W && hb.y + oy < H) out->hblits[keep++] = hb;
                            out->hblits.resize(keep);
                        }
                        const size_t nb = out->hblits.size();
                        if (nb == 0) continue;
                        // blits carry the layer pixel as 16 + 16 bits and their rank in 28
                        if (nb >= (1u << 28) || W > 65536 || H > 65536) { out->too_large = true; continue; }
                        int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
                        for (const rbh::HairBlit &hb : out->hblits) {
                            x0 = std::min(x0, hb.x); x1 = std::max(x1, hb.x);
                            y0 = std::min(y0, hb.y); y1 = std::max(y1, hb.y);
                        }
                        DevDraw d;
                        memset(&d, 0, sizeof(d));
                        d.ox = ox; d.oy = oy;
                        d.sx = x0; d.sy = y0; d.sw = x1 - x0 + 1; d.sh = y1 - y0 + 1;
                        d.shift = 2;
                        d.rule = 2; // hairline
                        DevPaint P;
                        const size_t s_before = out->stops.size();
                        if (!rbh::prepare_paint(&hp, ctm, &P, out->stops)) continue;
                        if (out->stops.size() != s_before) P.stop_off = (uint32_t)(P.stop_off - ci->s0);
                        d.paint = (uint32_t)(out->paints.size() - ci->p0);
                        out->paints.push_back(P);
                        const int r0 = (oy + y0) >> 3, r1 = (oy + y1) >> 3, nr = r1 - r0 + 1;
                        const int c0 = (ox + x0) / 32, c1 = (ox + x1) / 32;
                        // the draw's "tile rows" are its warp-tile CELLS (row-major over its bounding box), so that a tile
                        // finds exactly its own blits; rank = order of the blit inside its cell
                        const int ncols = c1 - c0 + 1;
                        out->hrank.assign((size_t)nr * (size_t)ncols, 0u);
                        const size_t eo = out->edges.size();
                        out->edges.resize(eo + nb);
                        DevEdge *dst = out->edges.data() + eo;
                        for (size_t k = 0; k < nb; k++) {
                            const rbh::HairBlit &hb = out->hblits[k];
                            const size_t cell = (size_t)(((oy + hb.y) >> 3) - r0) * (size_t)ncols + (size_t)(((ox + hb.x) >> 5) - c0);
                            DevEdge e; // a blit in an edge-sized record: layer pixel, coverage, rank inside its cell, cell
                            e.x = (int32_t)((uint32_t)(hb.x + ox) | ((uint32_t)(hb.y + oy) << 16));
                            e.dx = (int32_t)hb.alpha;
                            e.ypack = out->hrank[cell]++;
                            e.meta = (uint32_t)cell;
                            dst[k] = e;
                        }
                        d.curve_off = (uint32_t)c0; // hairline draws: first cell column / cells per row
                        d.curve_cnt = (uint32_t)ncols;
                        d.edge_off = items ? (uint32_t)ci->n_slots : (uint32_t)(eo - ci->e0);
                        d.edge_cnt = 0;
                        d.line_off = (uint32_t)(eo - ci->e0);
                        d.line_cnt = (uint32_t)nb;
                        d.r0 = (uint32_t)r0;
                        d.n_rows = (uint32_t)nr;
                        d.list_off = (uint32_t)ci->n_list;
                        d.row_base = (uint32_t)ci->n_row_off;
                        d.list_cap = (uint32_t)nb;
                        ci->n_list += nb;
                        ci->n_row_off += (size_t)nr * (size_t)ncols + 1;
                        ci->n_row_ent += (size_t)nr;
                        ci->n_wpairs += (size_t)nr * (size_t)(c1 - c0 + 1);
                        out->draws.push_back(d);
                        out->has_hair = true;
                    }
                }
                continue;
            }
        }
        if (r.is_stroke) {
            // stroke_path: the outline is computed in local coordinates, then filled (Winding) under the transform
            const uint8_t *ov = nullptr;
            const float *op = nullptr;
            int32_t nv = 0, np = 0;
            int sst;
            if (r.stroke.n_dash > 0) {
                // painter.rs stroke_path: the path is dashed first; a rejected dash specification leaves it solid
                const float *da = sp.bulk < 0 ? b->dashes.data() + r.dash_off : r.stroke.dash_array;
                bool valid = false;
                int dst_ = da ? rb_path_dash_into(verbs, n_verbs, &rpts[0].x, n_pts, da, r.stroke.n_dash, r.stroke.dash_offset,
                                                  resolution_scale(r.ctm), out->dverbs, out->dpts, &valid)
                              : RB_ERR_INVALID;
                if (valid) {
                    if (dst_ != RB_OK) continue; // "path dashing failed": nothing is drawn
                    verbs = out->dverbs.data();
                    n_verbs = (int)out->dverbs.size();
                    rpts = reinterpret_cast<const rbh::Pt *>(out->dpts.data());
                    n_pts = (int)(out->dpts.size() / 2);
                }
            }
            {
                PROF(0);
                sst = rb_path_stroke_view(verbs, n_verbs, &rpts[0].x, n_pts, r.stroke.width, r.stroke.miter_limit, r.stroke.cap,
                                          r.stroke.join, resolution_scale(r.ctm), &ov, &nv, &op, &np);
            }
            if (sst != RB_OK) continue;
            out->sverbs.assign(ov, ov + nv);
            out->spts.resize((size_t)np);
            memcpy(out->spts.data(), op, sizeof(float) * 2 * (size_t)np);
            rbh::map_points(r.ctm, out->spts.data(), np);
            verbs = out->sverbs.data();
            rpts = out->spts.data();
            n_verbs = nv;
            n_pts = np;
            rule = 0;
        }
        const bool aa = r.paint.anti_alias != 0;
        // DrawTiler: tiles of at most 8191x8191 in row-major order; a single tile for ordinary canvases.
        for (int ty = 0; ty < VH; ty += kMaxDim) {
            for (int tx = 0; tx < VW; tx += kMaxDim) {
                const int tw = std::min(VW - tx, kMaxDim), th = std::min(VH - ty, kMaxDim);
                const int ox = VX + tx, oy = VY + ty; // tile origin inside the target
                const rbh::Pt *pts = rpts;
                rbh::Xform ctm = r.ctm;
                if (tx || ty) {
                    out->tmp.assign(rpts, rpts + n_pts);
                    rbh::Xform tr;
                    tr.tx = -(float)tx;
                    tr.ty = -(float)ty;
                    rbh::map_points(tr, out->tmp.data(), n_pts);
                    pts = out->tmp.data();
                    ctm = rbh::post_concat(ctm, tr);
                }
                rbh::DrawGeom g;
                out->scratch.clear();
                out->cscratch.clear();
                bool ok;
                {
                    PROF(1);
                    ok = items ? rbh::build_draw_items(verbs, n_verbs, pts, n_pts, aa, tw, th, out->scratch, out->cscratch, &g)
                               : rbh::build_draw(verbs, n_verbs, pts, n_pts, aa, tw, th, out->scratch, &g);
                }
                if (!ok) continue;
                if (ox < 0 || oy < 0 || ox + tw > W || oy + th > H) { // blitter rectangle ∩ target (tile-local coordinates)
                    const int cx0 = std::max(g.sect.x, -ox), cy0 = std::max(g.sect.y, -oy);
                    const int cx1 = std::min(g.sect.x + g.sect.w, W - ox), cy1 = std::min(g.sect.y + g.sect.h, H - oy);
                    if (cx1 <= cx0 || cy1 <= cy0) continue;
                    g.sect.x = cx0; g.sect.y = cy0; g.sect.w = cx1 - cx0; g.sect.h = cy1 - cy0;
                }
                const size_t ne = out->scratch.size(), ncv = out->cscratch.size();
                DevDraw d;
                memset(&d, 0, sizeof(d));
                d.ox = ox; d.oy = oy;
                d.sx = g.sect.x; d.sy = g.sect.y; d.sw = g.sect.w; d.sh = g.sect.h;
                d.shift = g.shift;
                d.rule = rule;
                if (!mask_target) {
                    DevPaint p;
                    rb_paint rp = r.paint;
                    rp.stops = stops_src;
                    const size_t s_before = out->stops.size();
                    if (!rbh::prepare_paint(&rp, ctm, &p, out->stops)) continue;
                    if (out->stops.size() != s_before) p.stop_off = (uint32_t)(p.stop_off - ci->s0);
                    d.paint = (uint32_t)(out->paints.size() - ci->p0);
                    out->paints.push_back(p);
                }
                PROF(2);
                // warp-tile rows of the draw (layer pixel rows / 8) and the size of its tile-row edge lists
                const int r0 = (oy + g.sect.y) >> 3, r1 = (oy + g.sect.y + g.sect.h - 1) >> 3, nr = r1 - r0 + 1;
                const int c0 = (ox + g.sect.x) / 32, c1 = (ox + g.sect.x + g.sect.w - 1) / 32;
                auto row_of = [&](int32_t suby) { return std::min(std::max((((suby >> g.shift) + oy) >> 3) - r0, 0), nr - 1); };
                size_t n_list = 0;
                const size_t eo = out->edges.size();
                out->edges.resize(eo + ne);
                DevEdge *dst = out->edges.data() + eo;
                const rbh::Edge *src = out->scratch.data();
                d.edge_off = (uint32_t)(eo - ci->e0);
                if (!items) {
                    if (draw_may_exceed_packed_winding(src, ne)) out->wide = true;
                    d.edge_cnt = (uint32_t)ne;
                    for (size_t k = 0; k < ne; k++) {
                        dst[k] = pack_edge(src[k]);
                        n_list += (size_t)(row_of(src[k].last_y) - row_of(src[k].first_y) + 1);
                    }
                } else {
                    // slots of the device-side edge array in emission order: one per line, 2^shift per curve
                    const size_t co = out->curves.size();
                    out->curves.resize(co + ncv);
                    out->ends.clear();
                    struct Ends { std::vector<int32_t> *v; void operator()(int32_t a, int32_t b) { v->push_back(a); v->push_back(b); } } ends{&out->ends};
                    const geo::fl::Packed po = geo::fl::pack_items(src, ne, out->cscratch.data(), ncv, dst, out->curves.data() + co, g.shift, oy, r0, nr, ends);
                    n_list = po.n_list;
                    const uint32_t slot = po.slots;
                    if (po.too_large) { out->too_large = true; continue; }
                    if (chains_may_exceed_packed_winding(out->ends)) out->wide = true;
                    d.edge_cnt = slot;                     // slots of this draw in the device edge array
                    d.edge_off = (uint32_t)ci->n_slots;    // chunk-relative slot base
                    d.line_off = (uint32_t)(eo - ci->e0);
                    d.line_cnt = (uint32_t)ne;
                    d.curve_off = (uint32_t)(co - ci->c0);
                    d.curve_cnt = (uint32_t)ncv;
                    ci->n_slots += slot;
                }
                d.r0 = (uint32_t)r0;
                d.n_rows = (uint32_t)nr;
                d.list_off = (uint32_t)ci->n_list;   // chunk-relative until the pack phase
                d.row_base = (uint32_t)ci->n_row_off;
                d.list_cap = (uint32_t)n_list;
                ci->n_list += n_list;
                ci->n_row_off += (size_t)nr + 1;
                ci->n_row_ent += (size_t)nr;
                ci->n_wpairs += (size_t)nr * (size_t)(c1 - c0 + 1);
                out->draws.push_back(d);
            }
        }
    }
    ci->nc = out->curves.size() - ci->c0;
    ci->ne = out->edges.size() - ci->e0;
    ci->nd = out->draws.size() - ci->d0;
    ci->np = out->paints.size() - ci->p0;
    ci->ns = out->stops.size() - ci->s0;
}

inline void tile_range(const DevDraw &d, int &x0, int &x1, int &y0, int &y1)
{
    x0 = (d.ox + d.sx) / TW; x1 = (d.ox + d.sx + d.sw - 1) / TW;
    y0 = (d.oy + d.sy) / TH; y1 = (d.oy + d.sy + d.sh - 1) / TH;
}

} // namespace

int rb_batch_host_build(rb_batch *b, int W, int H, bool mask_target, int n_threads, rb_stage_alloc alloc, void *user,
                        void **block, size_t begin, size_t end)
{
    memset(b->stats, 0, sizeof(b->stats));
    memset(b->phases, 0, sizeof(b->phases));
    b->lay = BatchLayout();
    *block = nullptr;
    if (end == 0 || end > b->n_total) end = b->n_total;
    if (begin >= end) return RB_OK;
    const size_t n = end - begin;
    const auto t0 = Clock::now();

    // ---- 1. edges + paints on host threads (dynamic chunks; painter's order = chunk order) ------------------
    // Item mode first (lines final, curves recorded for the device to expand); when a draw could exceed the packed
    // winding range — or a debug hook asks for it — everything is rebuilt with the curves expanded on the host.
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min<int>(nt, (int)n_chunks));
    std::vector<std::unique_ptr<Worker>> workers;
    for (int t = 0; t < nt; t++) workers.push_back(borrow_worker());
    struct Return {
        std::vector<std::unique_ptr<Worker>> &w;
        ~Return() { for (auto &x : w) if (x) return_worker(std::move(x)); }
    } give_back{workers};
    std::vector<ChunkInfo> chunks(n_chunks);
    bool items = !g_force_wide && !g_host_expand;
    bool any_wide = false;
    for (;;) {
        for (auto &c : chunks) c = ChunkInfo();
        for (auto &w : workers) w->reset();
        parallel_for(nt, n_chunks, [&](size_t c, int t) {
            chunks[c].worker = t;
            build_chunk(b, begin + c * kChunk, begin + std::min(n, (c + 1) * kChunk), W, H, mask_target, items, /*hair_inline=*/items,
                        workers[(size_t)t].get(), &chunks[c]);
        });
        any_wide = false;
        for (auto &w : workers) any_wide = any_wide || w->wide;
        if (items && any_wide) { items = false; continue; }
        break;
    }
    // hairline strokes are only drawn inline by the tile kernel fed with items; with the fallback builder the caller has
    // to cut the batch into fill runs and hairline runs (rb_batch_submit does)
    for (auto &w : workers) if (w->too_large) return RB_ERR_UNSUPPORTED;
    for (auto &w : workers) if (w->skipped_hair) return RB_NEEDS_RUN_SPLIT;
    b->phases[0] = us_since(t0);
#ifdef RB_HOST_PROFILE
    fprintf(stderr, "[host profile] Mcycles: stroke %.0f build_draw %.0f pack %.0f hairline %.0f\n", g_prof[0].load() / 1e6, g_prof[1].load() / 1e6, g_prof[2].load() / 1e6, g_prof[3].load() / 1e6);
    for (auto &g : g_prof) g = 0;
#endif

    // ---- 2. layout --------------------------------------------------------------------------------------------
    const auto t1 = Clock::now();
    BatchLayout L;
    L.items = items;
    for (auto &c : chunks) {
        c.ge = L.n_edges; c.gd = L.n_draws; c.gp = L.n_paints; c.gs = L.n_stops;
        c.g_list = L.n_list; c.g_row_off = L.n_row_off; c.gc = L.n_curves; c.g_slots = L.n_slots;
        L.n_edges += c.ne; L.n_draws += c.nd; L.n_paints += c.np; L.n_stops += c.ns;
        L.n_list += c.n_list; L.n_row_off += c.n_row_off; L.n_row_ent += c.n_row_ent; L.n_wpairs += c.n_wpairs;
        L.n_curves += c.nc; L.n_slots += c.n_slots;
    }
    L.wide = any_wide || g_force_wide;
    for (auto &w : workers) L.has_hair = L.has_hair || w->has_hair;
    if (L.n_slots > 0xfffffff0ull) return RB_ERR_UNSUPPORTED;
    if (L.n_draws == 0) return RB_OK;
    if (L.n_edges > 0xfffffff0ull || L.n_list > 0xfffffff0ull || L.n_wpairs > 0xfffffff0ull) return RB_ERR_UNSUPPORTED;
    L.wtiles_x = (W + 31) / 32;
    L.wtiles_y = (H + 7) / 8;
    L.tiles_x = (W + TW - 1) / TW;
    const int tiles_y = L.wide ? (H + TH - 1) / TH : 0; // the 64x16 bins are only used by k_raster_tiles_wide
    L.n_tiles = (size_t)L.tiles_x * tiles_y;

    // tile ranges of every draw, computed from the workers' draws in painter's order; counts per tile
    struct TR { uint16_t x0, x1, y0, y1; };
    std::vector<TR> tr(L.wide ? L.n_draws : 0);
    if (L.wide) parallel_for(nt, n_chunks, [&](size_t ci, int) {
        const ChunkInfo &c = chunks[ci];
        const DevDraw *src = workers[(size_t)c.worker]->draws.data() + c.d0;
        for (size_t k = 0; k < c.nd; k++) {
            int x0, x1, y0, y1;
            tile_range(src[k], x0, x1, y0, y1);
            tr[c.gd + k] = TR{(uint16_t)x0, (uint16_t)x1, (uint16_t)y0, (uint16_t)y1};
        }
    });
    // binning is parallel over bands of tile rows: a band owns its tiles' counters and cursors
    constexpr int kBand = 8;
    const size_t n_bands = (size_t)(tiles_y + kBand - 1) / kBand;
    std::vector<uint32_t> tile_cnt(L.n_tiles + 1, 0);
    parallel_for(nt, n_bands, [&](size_t band, int) {
        const int by0 = (int)band * kBand, by1 = std::min(tiles_y, by0 + kBand) - 1;
        for (size_t di = 0; di < L.n_draws; di++) {
            const TR r = tr[di];
            if (r.y1 < by0 || r.y0 > by1) continue;
            for (int y = std::max<int>(r.y0, by0); y <= std::min<int>(r.y1, by1); y++) {
                uint32_t *row = tile_cnt.data() + (size_t)y * L.tiles_x + 1;
                for (int x = r.x0; x <= r.x1; x++) row[x]++;
            }
        }
    });
    for (size_t i = 0; i < L.n_tiles; i++) tile_cnt[i + 1] += tile_cnt[i];
    L.n_pairs = tile_cnt[L.n_tiles];
    // non-empty tiles, heaviest first (longest-processing-time-first over the CTA slots)
    std::vector<uint32_t> tile_ids;
    tile_ids.reserve(L.n_tiles);
    for (size_t i = 0; i < L.n_tiles; i++) if (tile_cnt[i + 1] > tile_cnt[i]) tile_ids.push_back((uint32_t)i);
    std::stable_sort(tile_ids.begin(), tile_ids.end(), [&](uint32_t a, uint32_t c) {
        return tile_cnt[a + 1] - tile_cnt[a] > tile_cnt[c + 1] - tile_cnt[c];
    });
    L.n_tile_ids = tile_ids.size();

    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    L.o_draws = off;  off += al(L.n_draws * sizeof(DevDraw));
    L.o_paints = off; off += al(std::max<size_t>(L.n_paints, 1) * sizeof(DevPaint));
    L.o_stops = off;  off += al(std::max<size_t>(L.n_stops, 1) * sizeof(DevStop));
    L.o_toff = off;   off += al((L.n_tiles + 1) * 4);
    L.o_tids = off;   off += al(L.n_tile_ids * 4);
    L.o_tdraws = off; off += al(L.n_pairs * 4);
    L.o_curves = off; off += al(std::max<size_t>(L.n_curves, 1) * sizeof(rbh::CurveRec));
    L.o_edges = off;  off += al(L.n_edges * sizeof(DevEdge));
    L.total = off;
    b->phases[1] = us_since(t1);

    const auto t2 = Clock::now();
    uint8_t *blk = (uint8_t *)alloc(user, L.total);
    if (!blk) return RB_ERR_OOM;
    b->phases[4] = us_since(t2);

    // ---- 3. pack into the block (threads) ------------------------------------------------------------------------
    const auto t3 = Clock::now();
    DevEdge *o_edges = (DevEdge *)(blk + L.o_edges);
    DevDraw *o_draws = (DevDraw *)(blk + L.o_draws);
    DevPaint *o_paints = (DevPaint *)(blk + L.o_paints);
    DevStop *o_stops = (DevStop *)(blk + L.o_stops);
    rbh::CurveRec *o_curves = (rbh::CurveRec *)(blk + L.o_curves);
    parallel_for(nt, n_chunks, [&](size_t ci, int) {
        const ChunkInfo &c = chunks[ci];
        const Worker &w = *workers[(size_t)c.worker];
        if (c.ne) memcpy(o_edges + c.ge, w.edges.data() + c.e0, c.ne * sizeof(DevEdge));
        if (c.nc) memcpy(o_curves + c.gc, w.curves.data() + c.c0, c.nc * sizeof(rbh::CurveRec));
        if (c.ns) memcpy(o_stops + c.gs, w.stops.data() + c.s0, c.ns * sizeof(DevStop));
        for (size_t k = 0; k < c.np; k++) {
            DevPaint p = w.paints[c.p0 + k];
            p.stop_off += (uint32_t)c.gs;
            o_paints[c.gp + k] = p;
        }
        for (size_t k = 0; k < c.nd; k++) {
            DevDraw d = w.draws[c.d0 + k];
            if (L.items) {
                d.edge_off += (uint32_t)c.g_slots;
                d.line_off += (uint32_t)c.ge;
                if (d.rule != 2) d.curve_off += (uint32_t)c.gc; // hairline draws keep their first cell column there
            } else {
                d.edge_off += (uint32_t)c.ge;
                d.line_off += (uint32_t)c.ge; // hairline draws keep their blits in the edge array in this mode too
            }
            d.paint += (uint32_t)c.gp;
            d.list_off += (uint32_t)c.g_list;
            d.row_base += (uint32_t)c.g_row_off;
            o_draws[c.gd + k] = d;
        }
    });
    memcpy(blk + L.o_toff, tile_cnt.data(), (L.n_tiles + 1) * 4);
    memcpy(blk + L.o_tids, tile_ids.data(), L.n_tile_ids * 4);
    b->phases[2] = us_since(t3);

    // ---- 4. fill the per-tile draw lists (counting sort keeps painter's order inside each tile) -------------------
    const auto t4 = Clock::now();
    uint32_t *tile_draws = (uint32_t *)(blk + L.o_tdraws);
    parallel_for(nt, n_bands, [&](size_t band, int) {
        const int by0 = (int)band * kBand, by1 = std::min(tiles_y, by0 + kBand) - 1;
        const size_t first = (size_t)by0 * L.tiles_x, cnt = (size_t)(by1 - by0 + 1) * L.tiles_x;
        std::vector<uint32_t> cursor(tile_cnt.begin() + first, tile_cnt.begin() + first + cnt);
        for (size_t di = 0; di < L.n_draws; di++) {
            const TR r = tr[di];
            if (r.y1 < by0 || r.y0 > by1) continue;
            for (int y = std::max<int>(r.y0, by0); y <= std::min<int>(r.y1, by1); y++) {
                uint32_t *row = cursor.data() + (size_t)(y - by0) * L.tiles_x;
                for (int x = r.x0; x <= r.x1; x++) tile_draws[row[x]++] = (uint32_t)di;
            }
        }
    });
    b->phases[3] = us_since(t4);

    b->lay = L;
    *block = blk;
    b->stats[0] = L.n_draws;
    b->stats[1] = L.items ? L.n_slots : L.n_edges; // line edges (item mode: upper bound = slots of the device edge array)
    b->stats[2] = L.wide ? L.n_pairs : L.n_wpairs;
    b->stats[3] = L.wide ? L.n_tile_ids : (size_t)L.wtiles_x * L.wtiles_y;
    b->stats[4] = L.total;
    b->stats[5] = b->phases[5] = us_since(t0);
    return RB_OK;
}

// ---- device path geometry: the host half (batch_geo.h) ---------------------------------------------------------------------
static int geo_mode_from_env()
{
    const char *e = getenv("RB_GEO_MODE");
    return e ? atoi(e) : 0;
}
int g_geo_mode = geo_mode_from_env();
std::atomic<uint64_t> g_geo_counts[6];
extern "C" void rb_debug_geo_mode(int mode) { g_geo_mode = mode; }
extern "C" void rb_debug_geo_counts(uint64_t out[6])
{
    if (out) for (int i = 0; i < 6; i++) out[i] = g_geo_counts[i].load();
}
// the debug hooks that select a builder variant only the host has
bool rb_debug_host_only_builder() { return g_force_wide || g_host_expand; }

namespace {

struct GeoWorker {
    std::vector<GeoTask> tasks;
    std::vector<uint8_t> verbs;
    std::vector<rbh::Pt> pts;
    std::vector<float> dashes, hstops;
    std::vector<DevPaint> paints;
    std::vector<DevStop> stops;
    void reset() { tasks.clear(); verbs.clear(); pts.clear(); dashes.clear(); paints.clear(); stops.clear(); }
};
struct GeoChunk {
    int worker = 0;
    size_t t0 = 0, nt = 0, v0 = 0, nv = 0, p0 = 0, np = 0, d0 = 0, nd = 0, pa0 = 0, npa = 0, s0 = 0, ns = 0; // ranges in the worker's vectors
    size_t n_draws = 0, max_units = 0;                                                                         // DevDraw entries / unit bound of this chunk
    size_t gt = 0, gv = 0, gp = 0, gd = 0, gpa = 0, gs = 0, gdraw = 0;                                       // global bases
};

std::mutex g_geo_pool_mu;
std::vector<std::unique_ptr<GeoWorker>> g_geo_pool;

// Cost class of a task for the heaviest-first task lists (log2 of the expected number of edge items).
inline int cost_class(uint32_t hint)
{
    int c = 0;
    while (hint > 1 && c < 23) { hint >>= 1; c++; }
    return c;
}

void geo_build_chunk(const rb_batch *b, size_t begin, size_t end, int W, int H, GeoWorker *out, GeoChunk *ci)
{
    ci->t0 = out->tasks.size(); ci->v0 = out->verbs.size(); ci->p0 = out->pts.size(); ci->d0 = out->dashes.size();
    ci->pa0 = out->paints.size(); ci->s0 = out->stops.size();
    size_t span_i = 0;
    for (size_t i = begin; i < end; i++) {
        while (i >= b->spans[span_i].start + b->spans[span_i].count) span_i++;
        while (i < b->spans[span_i].start) span_i--;
        const DrawSpan &sp = b->spans[span_i];
        int VX = 0, VY = 0, VW = W, VH = H;
        if (sp.vp_w > 0) {
            VX = sp.vp_x; VY = sp.vp_y; VW = sp.vp_w; VH = sp.vp_h;
            if (VX >= W || VY >= H || (int64_t)VX + VW <= 0 || (int64_t)VY + VH <= 0) continue;
        }
        // the draw's recorded form
        const uint8_t *verbs;
        const rbh::Pt *rpts;
        const float *stops_src, *dash_src = nullptr;
        uint32_t n_verbs, n_pts;
        rb_paint paint;
        rbh::Xform ctm;
        bool is_stroke, map_fill = false;
        rb_stroke stroke;
        int rule;
        memset(&stroke, 0, sizeof(stroke));
        if (sp.bulk < 0) {
            const RecordedDraw &r = b->recs[sp.first + (i - sp.start)];
            verbs = b->verbs.data() + r.verb_off;
            rpts = b->pts.data() + r.pt_off;
            n_verbs = r.n_verbs; n_pts = r.n_pts;
            paint = r.paint;
            stops_src = r.n_stops ? b->stops.data() + r.stop_off : nullptr;
            ctm = r.ctm;
            is_stroke = r.is_stroke;
            if (is_stroke) { stroke = r.stroke; dash_src = r.stroke.n_dash > 0 ? b->dashes.data() + r.dash_off : nullptr; }
            rule = r.rule;
        } else {
            const BulkSeg &bs = b->bulk[(size_t)sp.bulk];
            const size_t k = i - sp.start;
            verbs = bs.verbs + bs.verb_off[k];
            rpts = reinterpret_cast<const rbh::Pt *>(bs.points) + bs.point_off[k];
            n_verbs = bs.verb_off[k + 1] - bs.verb_off[k];
            n_pts = bs.point_off[k + 1] - bs.point_off[k];
            paint = bs.paints[k];
            stops_src = (paint.stops && paint.n_stops > 0) ? paint.stops : nullptr;
            ctm = bs.ctm;
            is_stroke = bs.strokes && bs.strokes[k].width > 0.0f;
            if (is_stroke) { stroke = bs.strokes[k]; dash_src = stroke.n_dash > 0 ? stroke.dash_array : nullptr; }
            rule = is_stroke ? 0 : (bs.fill_rules[k] ? 1 : 0);
            map_fill = !is_stroke && !bs.ctm.is_identity(); // painter.rs: path.transform(ts), done by the device
        }
        if (stroke.n_dash > 0 && !dash_src) stroke.n_dash = 0;
        // bounds of the draw in viewport coordinates (control-point hull, inflated by the stroke's reach): culls against
        // the target and, per DrawTiler tile, against the tile
        float bx0 = INFINITY, by0 = INFINITY, bx1 = -INFINITY, by1 = -INFINITY;
        bool finite = n_pts > 0;
        for (uint32_t k = 0; k < n_pts; k++) {
            const float x = rpts[k].x, y = rpts[k].y;
            finite = finite && std::isfinite(x) && std::isfinite(y);
            bx0 = std::min(bx0, x); bx1 = std::max(bx1, x);
            by0 = std::min(by0, y); by1 = std::max(by1, y);
        }
        if (finite && (is_stroke || map_fill)) {
            const float reach = is_stroke ? 0.5f * stroke.width * std::max(stroke.miter_limit, 1.4143f) : 0.0f;
            if (std::isfinite(reach)) {
                rbh::Pt c[4] = {{bx0 - reach, by0 - reach}, {bx1 + reach, by0 - reach}, {bx0 - reach, by1 + reach}, {bx1 + reach, by1 + reach}};
                rbh::map_points(ctm, c, 4);
                bx0 = by0 = INFINITY; bx1 = by1 = -INFINITY;
                for (const rbh::Pt &q : c) {
                    finite = finite && std::isfinite(q.x) && std::isfinite(q.y);
                    bx0 = std::min(bx0, q.x); bx1 = std::max(bx1, q.x);
                    by0 = std::min(by0, q.y); by1 = std::max(by1, q.y);
                }
            } else finite = false;
        }
        const float pad = 2.0f;
        if (finite && (bx1 < (float)(-VX) - pad || by1 < (float)(-VY) - pad || bx0 > (float)(W - VX) + pad || by0 > (float)(H - VY) + pad)) continue;
        float coverage = -1.0f;
        if (is_stroke && b->n_hair) coverage = rb_hairline_coverage(paint, stroke, ctm);
        const bool hair = coverage >= 0.0f;
        rb_paint pp = paint;
        pp.stops = stops_src;
        if (hair) hairline_modulate_paint(&pp, coverage, out->hstops);
        // the draw's path data is copied once, however many tiles it touches
        uint32_t voff = 0, poff = 0, doff = 0;
        bool copied = false;
        // expected edge items: a handful per segment; a stroke outline has two sides plus joins, a dashed one is cut up
        uint32_t hint = n_verbs * 3 + 8;
        if (is_stroke && !hair) {
            hint = n_verbs * 24 + 16;
            if (stroke.n_dash > 0) hint *= 4;
        }
        // A dashed stroke is built dash by dash on the device (one thread per output contour).  Bound of the contours
        // dashing can produce: every contour is at most as long as its control polygon (plus the closing line, itself at
        // most that long), and yields at most one dash per period of the pattern plus two.
        uint32_t max_units = 0;
        if (stroke.n_dash >= 2 && !(stroke.n_dash & 1)) {
            double poly = 0.0, period = 0.0;
            for (uint32_t k = 1; k < n_pts; k++) poly += hypot((double)rpts[k].x - rpts[k - 1].x, (double)rpts[k].y - rpts[k - 1].y);
            for (int k = 0; k < stroke.n_dash; k++) period += dash_src[k];
            if (std::isfinite(poly) && period > 0.0 && std::isfinite(period)) {
                const double u = 2.0 * poly * (double)(stroke.n_dash >> 1) / period * 1.001 + 2.0 * (double)n_verbs + 4.0;
                if (u < 8192.0) max_units = (uint32_t)u;
            }
        }
        for (int ty = 0; ty < VH; ty += kMaxDim) {
            for (int tx = 0; tx < VW; tx += kMaxDim) {
                const int tw = std::min(VW - tx, kMaxDim), th = std::min(VH - ty, kMaxDim);
                const int ox = VX + tx, oy = VY + ty; // tile origin inside the target
                if (ox >= W || oy >= H || ox + tw <= 0 || oy + th <= 0) continue; // tile wholly outside the target
                if (finite && (bx1 < (float)tx - pad || by1 < (float)ty - pad || bx0 > (float)(tx + tw) + pad || by0 > (float)(ty + th) + pad)) continue;
                rbh::Xform tctm = ctm;
                if (tx || ty) {
                    rbh::Xform tr;
                    tr.tx = -(float)tx;
                    tr.ty = -(float)ty;
                    tctm = rbh::post_concat(ctm, tr);
                }
                DevPaint P;
                const size_t s_before = out->stops.size();
                if (!rbh::prepare_paint(&pp, tctm, &P, out->stops)) continue;
                if (out->stops.size() != s_before) P.stop_off = (uint32_t)(P.stop_off - ci->s0);
                if (!copied) {
                    voff = (uint32_t)(out->verbs.size() - ci->v0);
                    poff = (uint32_t)(out->pts.size() - ci->p0);
                    out->verbs.insert(out->verbs.end(), verbs, verbs + n_verbs);
                    out->pts.insert(out->pts.end(), rpts, rpts + n_pts);
                    if (stroke.n_dash > 0) {
                        doff = (uint32_t)(out->dashes.size() - ci->d0);
                        out->dashes.insert(out->dashes.end(), dash_src, dash_src + stroke.n_dash);
                    }
                    copied = true;
                }
                GeoTask t;
                memset(&t, 0, sizeof(t));
                t.verb_off = voff; t.n_verbs = n_verbs; t.pt_off = poff; t.n_pts = n_pts;
                memcpy(t.ctm, &ctm, sizeof(float) * 6);
                t.tile_tx = -(float)tx; t.tile_ty = -(float)ty;
                t.tw = tw; t.th = th; t.ox = ox; t.oy = oy;
                t.paint = (uint32_t)(out->paints.size() - ci->pa0);
                out->paints.push_back(P);
                t.flags = (is_stroke ? GT_STROKE : 0u) | (hair ? GT_HAIR : 0u) | (paint.anti_alias ? GT_AA : 0u) | (rule ? GT_EVENODD : 0u)
                          | (stroke.n_dash > 0 ? GT_DASH : 0u) | ((is_stroke ? !ctm.is_identity() : map_fill) ? GT_MAP : 0u) | ((tx || ty) ? GT_TILE : 0u)
                          | ((uint32_t)stroke.cap << GT_CAP_SHIFT) | ((uint32_t)stroke.join << GT_JOIN_SHIFT);
                if (!is_stroke && !map_fill) { const rbh::Xform id; memcpy(t.ctm, &id, sizeof(float) * 6); }
                t.width = stroke.width; t.miter = stroke.miter_limit; t.res_scale = is_stroke ? resolution_scale(ctm) : 1.0f;
                t.dash_offset = stroke.dash_offset;
                t.dash_off = doff; t.n_dash = (uint32_t)std::max(stroke.n_dash, 0);
                t.hint = hint;
                t.sub = -1;
                t.n_draws = 1;
                static const int units_sel = getenv("RB_GEO_UNITS") ? atoi(getenv("RB_GEO_UNITS")) : 3; // tests: 1 strokes only, 2 hairlines only, 0 none
                if (max_units && (units_sel & (hair ? 2 : 1))) {
                    t.flags |= GT_UNITS;
                    t.max_units = max_units;
                    if (hair) t.n_draws = max_units; // every dash of a hairline is a draw of its own
                }
                if (hair && stroke.n_dash <= 0) {
                    // a hairline's blits are those of its verbs, one after the other: every verb becomes a draw of its own
                    // (same paint), so that a long path is walked by as many threads as it has segments
                    t.hint = 64;
                    for (uint32_t vi = 0; vi < n_verbs; vi++) {
                        if (verbs[vi] == RB_VERB_MOVE) continue;
                        t.sub = (int32_t)vi;
                        t.draw = (uint32_t)ci->n_draws++;
                        out->tasks.push_back(t);
                    }
                    continue;
                }
                t.draw = (uint32_t)ci->n_draws;
                ci->n_draws += t.n_draws;
                ci->max_units += t.max_units;
                out->tasks.push_back(t);
            }
        }
    }
    ci->nt = out->tasks.size() - ci->t0; ci->nv = out->verbs.size() - ci->v0; ci->np = out->pts.size() - ci->p0;
    ci->nd = out->dashes.size() - ci->d0; ci->npa = out->paints.size() - ci->pa0; ci->ns = out->stops.size() - ci->s0;
}

} // namespace

int rb_geo_host_build(rb_batch *b, int W, int H, int n_threads, rb_stage_alloc alloc, void *user, void **block, GeoBlock *gb,
                      size_t begin, size_t end)
{
    *block = nullptr;
    *gb = GeoBlock();
    if (end == 0 || end > b->n_total) end = b->n_total;
    if (begin >= end) return RB_OK;
    const size_t n = end - begin;
    const auto t0g = Clock::now();
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min<int>(nt, (int)n_chunks));
    std::vector<std::unique_ptr<GeoWorker>> workers;
    {
        std::lock_guard<std::mutex> g(g_geo_pool_mu);
        for (int t = 0; t < nt; t++) {
            if (g_geo_pool.empty()) workers.emplace_back(new GeoWorker());
            else { workers.push_back(std::move(g_geo_pool.back())); g_geo_pool.pop_back(); }
        }
    }
    struct Return {
        std::vector<std::unique_ptr<GeoWorker>> &w;
        ~Return()
        {
            std::lock_guard<std::mutex> g(g_geo_pool_mu);
            for (auto &x : w) if (x) { x->reset(); if (g_geo_pool.size() < 256) g_geo_pool.push_back(std::move(x)); }
        }
    } give_back{workers};
    std::vector<GeoChunk> chunks(n_chunks);
    parallel_for(nt, n_chunks, [&](size_t c, int t) {
        chunks[c].worker = t;
        geo_build_chunk(b, begin + c * kChunk, begin + std::min(n, (c + 1) * kChunk), W, H, workers[(size_t)t].get(), &chunks[c]);
    });
    const auto t_chunks = Clock::now();
    GeoBlock G;
    for (auto &c : chunks) {
        c.gt = G.n_tasks; c.gv = G.n_verbs; c.gp = G.n_pts; c.gd = G.n_dashes; c.gpa = G.n_paints; c.gs = G.n_stops; c.gdraw = G.n_draws;
        G.n_draws += c.n_draws; G.max_units += c.max_units;
        G.n_tasks += c.nt; G.n_verbs += c.nv; G.n_pts += c.np; G.n_dashes += c.nd; G.n_paints += c.npa; G.n_stops += c.ns;
    }
    if (G.n_tasks == 0) return RB_OK;
    if (G.n_draws > 0x7ffffff0ull || G.max_units > 0x7ffffff0ull || G.n_tasks > 0x7ffffff0ull || G.n_verbs > 0xfffffff0ull || G.n_pts > 0xfffffff0ull) return RB_ERR_UNSUPPORTED;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    G.o_tasks = off;  off += al(G.n_tasks * sizeof(GeoTask));
    G.o_paints = off; off += al(std::max<size_t>(G.n_paints, 1) * sizeof(DevPaint));
    G.o_stops = off;  off += al(std::max<size_t>(G.n_stops, 1) * sizeof(DevStop));
    G.o_lists = off;  off += al(G.n_tasks * 3 * sizeof(uint32_t)); // a dashed stroke is on three lists: dash, stroke, fill
    G.o_dashes = off; off += al(std::max<size_t>(G.n_dashes, 1) * sizeof(float));
    G.o_pts = off;    off += al(G.n_pts * sizeof(rbh::Pt));
    G.o_verbs = off;  off += al(G.n_verbs);
    G.total = off;
    uint8_t *blk = (uint8_t *)alloc(user, G.total);
    if (!blk) return RB_ERR_OOM;
    GeoTask *o_tasks = (GeoTask *)(blk + G.o_tasks);
    DevPaint *o_paints = (DevPaint *)(blk + G.o_paints);
    DevStop *o_stops = (DevStop *)(blk + G.o_stops);
    parallel_for(nt, n_chunks, [&](size_t ci, int) {
        const GeoChunk &c = chunks[ci];
        const GeoWorker &w = *workers[(size_t)c.worker];
        if (c.nv) memcpy(blk + G.o_verbs + c.gv, w.verbs.data() + c.v0, c.nv);
        if (c.np) memcpy(blk + G.o_pts + c.gp * sizeof(rbh::Pt), w.pts.data() + c.p0, c.np * sizeof(rbh::Pt));
        if (c.nd) memcpy(blk + G.o_dashes + c.gd * sizeof(float), w.dashes.data() + c.d0, c.nd * sizeof(float));
        if (c.ns) memcpy(o_stops + c.gs, w.stops.data() + c.s0, c.ns * sizeof(DevStop));
        for (size_t k = 0; k < c.npa; k++) {
            DevPaint p = w.paints[c.pa0 + k];
            p.stop_off += (uint32_t)c.gs;
            o_paints[c.gpa + k] = p;
        }
        for (size_t k = 0; k < c.nt; k++) {
            GeoTask t = w.tasks[c.t0 + k];
            t.verb_off += (uint32_t)c.gv; t.pt_off += (uint32_t)c.gp; t.dash_off += (uint32_t)c.gd; t.paint += (uint32_t)c.gpa; t.draw += (uint32_t)c.gdraw;
            o_tasks[c.gt + k] = t;
        }
    });
    const auto t_copy = Clock::now();
    // task lists per kernel, heaviest first (counting sort over cost classes; ties keep painter's order)
    {
        uint32_t *lists = (uint32_t *)(blk + G.o_lists);
        constexpr int NC = 24, NL = 6; // lists: dash, stroke, hair, plain fills, the dashed strokes built in units, outline fills
        size_t cnt[NL][NC];
        memset(cnt, 0, sizeof(cnt));
        auto kinds_of = [](const GeoTask &t, int k[3]) {
            int m = 0;
            if (t.flags & GT_UNITS) { k[m++] = 4; return m; }
            if (t.flags & GT_DASH) k[m++] = 0;
            k[m++] = (t.flags & GT_HAIR) ? 2 : ((t.flags & GT_STROKE) ? 1 : 3);
            if ((t.flags & (GT_STROKE | GT_HAIR)) == GT_STROKE) k[m++] = 5; // an outline is filled
            return m;
        };
        for (size_t i = 0; i < G.n_tasks; i++) {
            int k[3];
            const int m = kinds_of(o_tasks[i], k);
            for (int j = 0; j < m; j++) cnt[k[j]][NC - 1 - cost_class(o_tasks[i].hint)]++;
        }
        size_t base[NL][NC], run = 0;
        size_t first[NL + 1];
        for (int k = 0; k < NL; k++) {
            first[k] = run;
            for (int c = 0; c < NC; c++) { base[k][c] = run; run += cnt[k][c]; }
        }
        first[NL] = run;
        for (size_t i = 0; i < G.n_tasks; i++) {
            int k[3];
            const int m = kinds_of(o_tasks[i], k);
            for (int j = 0; j < m; j++) lists[base[k[j]][NC - 1 - cost_class(o_tasks[i].hint)]++] = (uint32_t)i;
        }
        G.n_dash_l = first[1] - first[0]; G.n_stroke_l = first[2] - first[1]; G.n_hair_l = first[3] - first[2]; G.n_fill_l = first[4] - first[3];
        G.n_units_l = first[5] - first[4];
        G.n_outline_l = first[6] - first[5];
        G.has_hair = false;
        for (size_t i = 0; i < G.n_tasks && !G.has_hair; i++) G.has_hair = (o_tasks[i].flags & GT_HAIR) != 0;
    }
    const auto t_lists = Clock::now();
    // heap the geometry is expected to need: ~48 bytes per expected edge item plus the builders' chunks
    size_t hint_bytes = 0;
    for (size_t i = 0; i < G.n_tasks; i++) hint_bytes += (o_tasks[i].flags & GT_UNITS) ? (size_t)o_tasks[i].max_units * 6144 + 8192 : (size_t)o_tasks[i].hint * 96 + 2048;
    G.heap_hint = hint_bytes;
    if (getenv("RB_GEO_HOST_DIAG")) fprintf(stderr, "[geo host] chunks %.2f ms, layout + copy %.2f ms, lists %.2f ms, heap hint %.2f ms\n", (double)std::chrono::duration_cast<std::chrono::microseconds>(t_chunks - t0g).count() / 1e3, (double)std::chrono::duration_cast<std::chrono::microseconds>(t_copy - t_chunks).count() / 1e3, (double)std::chrono::duration_cast<std::chrono::microseconds>(t_lists - t_copy).count() / 1e3, (double)us_since(t_lists) / 1e3);
    *gb = G;
    *block = blk;
    return RB_OK;
}

// Host-only batches: runs the host half of the device geometry path (classification, culling, paints, task lists) and
// reports out[0..7] = tasks, dashed, stroked, hairline, fill-list entries, uploaded bytes, verbs, points.  No device work.
extern "C" int rb_debug_geo_host_stats(rb_batch *b, uint64_t out[8])
{
    if (!b || !out || b->host_w <= 0) return RB_ERR_INVALID;
    void *blk = nullptr;
    GeoBlock G;
    int st = rb_geo_host_build(b, b->host_w, b->host_h, 0, [](void *, size_t bytes) { return malloc(bytes); }, nullptr, &blk, &G, 0, 0);
    free(blk);
    if (st != RB_OK) return st;
    out[0] = G.n_tasks; out[1] = G.n_dash_l + G.n_units_l; out[2] = G.n_stroke_l; out[3] = G.n_hair_l; out[4] = G.n_fill_l + G.n_outline_l; out[5] = G.total;
    out[6] = G.n_verbs; out[7] = G.n_pts;
    return RB_OK;
}

// ---- recording --------------------------------------------------------------------------------------------------------
int rb_batch_record(rb_batch *b, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                    const rb_paint *paint, int32_t rule, const float ts[6])
{
    rb_prof_scope prof__(RB_T_RECORD);
    if (!b || !verbs || !points || n_verbs <= 0 || n_points <= 0 || !paint) return RB_ERR_INVALID;
    // a pattern's source layer must hold its final pixels before this draw runs: execute its pending immediate draws now,
    // on the caller's thread (rb_layer_device_ptr flushes)
    if (paint->shader == 3 && paint->pattern) (void)rb_layer_device_ptr(const_cast<rb_layer *>(paint->pattern));
    // validate the verb/point bookkeeping so the builder never reads past the arrays
    int need = 0;
    for (int i = 0; i < n_verbs; i++) {
        switch (verbs[i]) {
        case 0: case 1: need += 1; break;
        case 2: need += 2; break;
        case 3: need += 3; break;
        case 4: break;
        default: return RB_ERR_INVALID;
        }
    }
    if (need != n_points || verbs[0] != 0) return RB_ERR_INVALID;
    if (b->verbs.size() + (size_t)n_verbs > 0xfffffff0ull || b->pts.size() + (size_t)n_points > 0xfffffff0ull) return RB_ERR_UNSUPPORTED;
    RecordedDraw r;
    memset(&r, 0, sizeof(r));
    r.verb_off = (uint32_t)b->verbs.size();
    r.n_verbs = (uint32_t)n_verbs;
    r.pt_off = (uint32_t)b->pts.size();
    r.n_pts = (uint32_t)n_points;
    b->verbs.insert(b->verbs.end(), verbs, verbs + n_verbs);
    b->pts.resize(b->pts.size() + (size_t)n_points);
    memcpy(b->pts.data() + r.pt_off, points, sizeof(float) * 2 * (size_t)n_points);
    r.ctm = ts ? rbh::Xform::from(ts) : rbh::Xform();
    rbh::map_points(r.ctm, b->pts.data() + r.pt_off, n_points); // painter.rs: path.transform(ts), shader.transform(ts)
    r.paint = *paint;
    if (paint->stops && paint->n_stops > 0) {
        r.stop_off = (uint32_t)b->stops.size();
        r.n_stops = (uint32_t)paint->n_stops;
        b->stops.insert(b->stops.end(), paint->stops, paint->stops + (size_t)paint->n_stops * 5);
    }
    r.paint.stops = nullptr;
    r.rule = rule ? 1 : 0;
    r.is_stroke = false;
    b->recs.push_back(r);
    if (!b->spans.empty() && b->spans.back().bulk < 0 && b->spans.back().vp_x == b->vp_x && b->spans.back().vp_y == b->vp_y
        && b->spans.back().vp_w == b->vp_w && b->spans.back().vp_h == b->vp_h)
        b->spans.back().count++;
    else b->spans.push_back(DrawSpan{b->n_total, 1, -1, b->recs.size() - 1, b->vp_x, b->vp_y, b->vp_w, b->vp_h});
    b->n_total++;
    return RB_OK;
}

extern "C" int rb_batch_fill_path(rb_batch *b, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                                  const rb_paint *paint, int32_t fill_rule, const float ts[6])
{
    if (paint && (paint->shader < 0 || paint->shader > 3 || paint->blend_mode < 0 || paint->blend_mode > 28)) return RB_ERR_INVALID;
    if (paint && (paint->shader == 1 || paint->shader == 2) && paint->n_stops > rbh::kMaxStops) return RB_ERR_UNSUPPORTED;
    // a pattern that IS the target would have to be flushed (and so destroy this very batch) to be read
    if (paint && b && paint->shader == 3 && (!paint->pattern || paint->pattern == b->layer)) return RB_ERR_INVALID;
    return rb_batch_record(b, verbs, n_verbs, points, n_points, paint, fill_rule, ts);
}

// PixmapMut::stroke_path(path, paint, stroke, transform, None) — path.rs:113.  Thin anti-aliased strokes that
// tiny-skia draws as hairlines (both transformed stroke-width vectors no longer than 1 px) are walked by hairline.cpp
// and applied by the tile kernel as ordered blits (targets up to 65536 px per side).
extern "C" int rb_batch_stroke_path(rb_batch *b, const uint8_t *verbs, int32_t n_verbs, const float *points,
                                    int32_t n_points, const rb_paint *paint, const rb_stroke *stroke, const float ts[6])
{
    if (!stroke || !paint || !b) return RB_ERR_INVALID;
    if (stroke->width < 0.0f) return RB_OK;
    if (stroke->cap < 0 || stroke->cap > 2 || stroke->join < 0 || stroke->join > 3) return RB_ERR_INVALID;
    const rbh::Xform ctm = ts ? rbh::Xform::from(ts) : rbh::Xform();
    const bool hair = rb_hairline_coverage(*paint, *stroke, ctm) >= 0.0f;
    static const float ident[6] = {1, 0, 0, 1, 0, 0};
    int st = rb_batch_fill_path(b, verbs, n_verbs, points, n_points, paint, 0, ident); // keep local coordinates
    if (st != RB_OK) return st;
    RecordedDraw &r = b->recs.back();
    r.ctm = ctm;
    r.is_stroke = true;
    r.stroke = *stroke;
    r.stroke.dash_array = nullptr;
    r.dash_off = (uint32_t)b->dashes.size();
    if (stroke->dash_array && stroke->n_dash > 0) b->dashes.insert(b->dashes.end(), stroke->dash_array, stroke->dash_array + stroke->n_dash);
    else r.stroke.n_dash = 0;
    if (hair) b->n_hair++;
    return RB_OK;
}

// Bulk recording: n_paths paths in packed arrays (verb_off / point_off have n_paths + 1 entries).  Same checks and
// semantics as n calls of rb_batch_fill_path / rb_batch_stroke_path, but BY REFERENCE: the arrays (including the
// gradient stops the paints point to) are read in place by rb_batch_prepare / submit and must stay valid and unchanged
// until that call has returned.
extern "C" int rb_batch_draw_paths(rb_batch *b, int32_t n_paths, const uint32_t *verb_off, const uint32_t *point_off,
                                   const uint8_t *verbs, const float *points, const rb_paint *paints,
                                   const uint8_t *fill_rules, const rb_stroke *strokes, const float ts[6])
{
    if (!b || n_paths < 0 || !verb_off || !point_off || !verbs || !points || !paints || !fill_rules) return RB_ERR_INVALID;
    if (n_paths == 0) return RB_OK;
    const rbh::Xform ctm = ts ? rbh::Xform::from(ts) : rbh::Xform();
    // validation (the builders trust the verb / point bookkeeping) — on the worker threads for large calls: 100 000 paths are
    // a million verbs, 3 ms on one thread
    std::atomic<size_t> n_hair_total(0);
    std::atomic<int> first_error(RB_OK);
    auto check_range = [&](int32_t lo, int32_t hi) {
        size_t n_hair = 0;
        for (int32_t i = lo; i < hi; i++) {
            const uint32_t va = verb_off[i], vb = verb_off[i + 1], pa = point_off[i], pb = point_off[i + 1];
            const rb_paint &paint = paints[i];
            int err = RB_OK;
            if (vb <= va || pb <= pa) err = RB_ERR_INVALID;
            else if (paint.shader < 0 || paint.shader > 3 || paint.blend_mode < 0 || paint.blend_mode > 28) err = RB_ERR_INVALID;
            else if ((paint.shader == 1 || paint.shader == 2) && paint.n_stops > rbh::kMaxStops) err = RB_ERR_UNSUPPORTED;
            if (err == RB_OK) {
                // the verb/point bookkeeping must be right or the builder would read past the arrays
                uint32_t need = 0;
                bool bad = false;
                for (uint32_t k = va; k < vb; k++) {
                    const uint8_t v = verbs[k];
                    if (v > 4) { bad = true; break; }
                    need += v == 4 ? 0u : (v <= 1 ? 1u : v);
                }
                if (bad || need != pb - pa || verbs[va] != 0) err = RB_ERR_INVALID;
            }
            if (err == RB_OK && strokes && strokes[i].width > 0.0f) {
                const rb_stroke &sk = strokes[i];
                if (sk.cap < 0 || sk.cap > 2 || sk.join < 0 || sk.join > 3) err = RB_ERR_INVALID;
                else if (rb_hairline_coverage(paint, sk, ctm) >= 0.0f) n_hair++;
            }
            if (err != RB_OK) { int expected = RB_OK; first_error.compare_exchange_strong(expected, err); return; }
        }
        n_hair_total += n_hair;
    };
    // a pattern's source layer must hold its final pixels before this draw runs (flushes on the caller's thread)
    for (int32_t i = 0; i < n_paths; i++)
        if (paints[i].shader == 3 && paints[i].pattern) (void)rb_layer_device_ptr(const_cast<rb_layer *>(paints[i].pattern));
    if (n_paths >= 16384) {
        const size_t n_chunks = ((size_t)n_paths + 4095) / 4096;
        parallel_for((int)std::thread::hardware_concurrency(), n_chunks, [&](size_t c, int) {
            check_range((int32_t)(c * 4096), (int32_t)std::min<size_t>((size_t)n_paths, (c + 1) * 4096));
        });
    } else {
        check_range(0, n_paths);
    }
    if (first_error.load() != RB_OK) return first_error.load();
    const size_t n_hair = n_hair_total.load();
    b->n_hair += n_hair;
    BulkSeg bs;
    bs.n = n_paths;
    bs.verb_off = verb_off; bs.point_off = point_off; bs.verbs = verbs; bs.points = points;
    bs.paints = paints; bs.fill_rules = fill_rules; bs.strokes = strokes;
    bs.ctm = ctm;
    b->bulk.push_back(bs);
    b->spans.push_back(DrawSpan{b->n_total, (size_t)n_paths, (int)b->bulk.size() - 1, 0, b->vp_x, b->vp_y, b->vp_w, b->vp_h});
    b->n_total += (size_t)n_paths;
    return RB_OK;
}

extern "C" int rb_batch_fill_paths(rb_batch *b, int32_t n_paths, const uint32_t *verb_off, const uint32_t *point_off,
                                   const uint8_t *verbs, const float *points, const rb_paint *paints,
                                   const uint8_t *fill_rules, const float ts[6])
{
    return rb_batch_draw_paths(b, n_paths, verb_off, point_off, verbs, points, paints, fill_rules, nullptr, ts);
}

extern "C" int rb_batch_stats(rb_batch *b, uint64_t stats[6])
{
    if (!b || !stats) return RB_ERR_INVALID;
    memcpy(stats, b->stats, sizeof(b->stats));
    return RB_OK;
}

// The draws recorded after this call are rendered as if the rectangle (x, y, w, h) of the target were a pixmap of its
// own: coordinates are relative to its origin and nothing is drawn outside it.  This is how many small documents are
// rendered into one atlas layer by a single batch (document-parallel thumbnailing).  w == 0 restores the whole target.
// The rectangle may reach beyond the target (negative x / y, or larger than it): the document is still built against
// its whole pixmap and only what falls inside the target is drawn — this is how one rank renders a strip of a large
// canvas with exactly the pixels of the whole-canvas render (canvas-strip sharding).
extern "C" int rb_batch_set_viewport(rb_batch *b, int32_t x, int32_t y, uint32_t w, uint32_t h)
{
    if (!b || w > 0x3fffffffu || h > 0x3fffffffu || ((w == 0) != (h == 0)) || x < -0x3fffffff || y < -0x3fffffff) return RB_ERR_INVALID;
    b->vp_x = x; b->vp_y = y; b->vp_w = (int32_t)w; b->vp_h = (int32_t)h;
    return RB_OK;
}

// Bulk form of { rb_batch_set_viewport; rb_batch_draw_paths } per document: document k owns doc_count[k] paths starting at
// doc_first[k] of the packed arrays and is rendered into viewports[4k .. 4k + 3] = (x, y, w, h).  The batch's current
// viewport is restored afterwards.
extern "C" int rb_batch_draw_documents(rb_batch *b, int32_t n_docs, const int32_t *viewports, const uint32_t *doc_first,
                                       const uint32_t *doc_count, const uint32_t *verb_off, const uint32_t *point_off,
                                       const uint8_t *verbs, const float *points, const rb_paint *paints,
                                       const uint8_t *fill_rules, const rb_stroke *strokes, const float ts[6])
{
    if (!b || n_docs < 0 || (n_docs > 0 && (!viewports || !doc_first || !doc_count))) return RB_ERR_INVALID;
    const int32_t ox = b->vp_x, oy = b->vp_y, ow = b->vp_w, oh = b->vp_h;
    int st = RB_OK;
    for (int32_t k = 0; k < n_docs && st == RB_OK; k++) {
        const uint32_t a = doc_first[k], n = doc_count[k];
        if (n > 0x7fffffffu || viewports[4 * k + 2] <= 0 || viewports[4 * k + 3] <= 0) { st = RB_ERR_INVALID; break; }
        if (n == 0) continue;
        st = rb_batch_set_viewport(b, viewports[4 * k], viewports[4 * k + 1], (uint32_t)viewports[4 * k + 2], (uint32_t)viewports[4 * k + 3]);
        if (st == RB_OK)
            st = rb_batch_draw_paths(b, (int32_t)n, verb_off + a, point_off + a, verbs, points, paints + a, fill_rules + a,
                                     strokes ? strokes + a : nullptr, ts);
    }
    b->vp_x = ox; b->vp_y = oy; b->vp_w = ow; b->vp_h = oh;
    return st;
}

// Test hook: route every batch prepared from now on through the any-winding fallback kernel (k_raster_tiles_wide).
extern "C" void rb_debug_force_wide_kernel(int on) { g_force_wide = on != 0; }

// Test hook: expand curves on the host (the device then only bins and rasterises), to cross-check the device expansion.
extern "C" void rb_debug_host_expand(int on) { g_host_expand = on != 0; }

// ---- host-only batches (CPU test-suite / host-build profiling; no device work) ------------------------------------------
extern "C" int rb_debug_batch_begin_host(uint32_t width, uint32_t height, rb_batch **out)
{
    if (!out || width == 0 || height == 0) return RB_ERR_INVALID;
    rb_batch *b = new rb_batch();
    b->host_w = (int)width;
    b->host_h = (int)height;
    *out = b;
    return RB_OK;
}

extern "C" int rb_debug_batch_phases(rb_batch *b, uint64_t phases[RB_PHASES])
{
    if (!b || !phases) return RB_ERR_INVALID;
    memcpy(phases, b->phases, sizeof(b->phases));
    return RB_OK;
}

// Copies the arrays of a host-only batch's block out for inspection: which = 0 draws (DevDraw), 1 tile offsets,
// 2 tile draw lists, 3 tile ids, 4 edges (DevEdge).  Returns the element count; copies at most max_bytes.
extern "C" int64_t rb_debug_batch_block(rb_batch *b, int32_t which, void *out, uint64_t max_bytes)
{
    if (!b || !b->host_block) return -1;
    const BatchLayout &L = b->lay;
    const uint8_t *blk = (const uint8_t *)b->host_block;
    size_t off, cnt, esz;
    switch (which) {
    case 0: off = L.o_draws; cnt = L.n_draws; esz = sizeof(DevDraw); break;
    case 1: off = L.o_toff; cnt = L.n_tiles + 1; esz = 4; break;
    case 2: off = L.o_tdraws; cnt = L.n_pairs; esz = 4; break;
    case 3: off = L.o_tids; cnt = L.n_tile_ids; esz = 4; break;
    case 4: off = L.o_edges; cnt = L.n_edges; esz = sizeof(DevEdge); break;
    default: return -1;
    }
    if (out) memcpy(out, blk + off, std::min<uint64_t>(max_bytes, (uint64_t)cnt * esz));
    return (int64_t)cnt;
}
