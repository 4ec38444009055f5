// render.cpp — the render traversal behind rb_render / rb_render_node / rb_submit: resvg's crates/resvg/src/
// {lib,render,path,clip,mask,image,geom}.rs restated over this library's own layer operations, so that a whole
// usvg::Tree is drawn by ONE call across the C ABI (SURVEY §8(b) `rb_submit`).
//
// The reference walks the tree and issues a tiny-skia call per node; so does this file, against device-resident
// layers: fills and strokes are recorded lazily per layer (rb_fill_path / rb_stroke_path) and run as one tile-binned
// batch when the layer is next read, isolated groups get a layer the size of their bounding box, and nothing crosses
// PCIe until the caller downloads the target.  Function names and order follow the reference files cited on each.
#include "render.h"

#include <math.h>
#include <string.h>

namespace rbr {

using rbh::Xform;
using rbt::IntRect;
using rbt::Rect;

// ---------------------------------------------------------------------------------------------------------------------
// tiny-skia-path geometry helpers (rect.rs, transform.rs) in f32, with Rust `as` cast semantics
// ---------------------------------------------------------------------------------------------------------------------
static inline int32_t f2i(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static inline uint32_t f2u(float v)
{
    if (!(v > 0.0f)) return 0; // NaN, negative
    if (v >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)v;
}

Xform translate(float tx, float ty) { Xform t; t.tx = tx; t.ty = ty; return t; }
Xform scale(float sx, float sy) { Xform t; t.sx = sx; t.sy = sy; return t; }

void get_scale(const Xform &t, float *sx, float *sy) // Transform::get_scale
{
    *sx = sqrtf(t.sx * t.sx + t.kx * t.kx);
    *sy = sqrtf(t.ky * t.ky + t.sy * t.sy);
}

bool int_rect_from_xywh(int64_t x, int64_t y, uint64_t w, uint64_t h, IntRect *out) // IntRect::from_xywh
{
    if (w == 0 || h == 0 || w > (uint64_t)INT32_MAX || h > (uint64_t)INT32_MAX) return false;
    if (x < INT32_MIN || x > INT32_MAX || y < INT32_MIN || y > INT32_MAX) return false;
    if (x + (int64_t)w > INT32_MAX || y + (int64_t)h > INT32_MAX) return false; // checked_add
    out->x = (int32_t)x; out->y = (int32_t)y; out->w = (uint32_t)w; out->h = (uint32_t)h;
    return true;
}

bool int_rect_from_ltrb(int64_t l, int64_t t, int64_t r, int64_t b, IntRect *out)
{
    if (r <= l || b <= t) return false;
    return int_rect_from_xywh(l, t, (uint64_t)(r - l), (uint64_t)(b - t), out);
}

// geom.rs:5-30
bool fit_to_rect(const IntRect &r, const IntRect &bounds, IntRect *out)
{
    int32_t left = r.x < bounds.x ? bounds.x : r.x;
    int32_t top = r.y < bounds.y ? bounds.y : r.y;
    int32_t right = r.right() > bounds.right() ? bounds.right() : r.right();
    int32_t bottom = r.bottom() > bounds.bottom() ? bounds.bottom() : r.bottom();
    return int_rect_from_ltrb(left, top, right, bottom, out);
}

// Rect::transform / NonZeroRect::transform: the bounding box of the mapped corners; `non_zero` additionally requires a
// positive extent.  False = the reference's None.
bool rect_transform(const Rect &r, const Xform &ts, bool non_zero, Rect *out)
{
    Rect o = r;
    if (!ts.is_identity()) {
        const float right = r.x + r.w, bottom = r.y + r.h;
        rbh::Pt p[4] = {{r.x, r.y}, {right, r.y}, {right, bottom}, {r.x, bottom}};
        rbh::map_points(ts, p, 4);
        float l = p[0].x, t = p[0].y, rr = p[0].x, bb = p[0].y;
        for (int i = 1; i < 4; i++) {
            l = fminf(l, p[i].x); rr = fmaxf(rr, p[i].x);
            t = fminf(t, p[i].y); bb = fmaxf(bb, p[i].y);
        }
        if (!(std::isfinite(l) && std::isfinite(t) && std::isfinite(rr) && std::isfinite(bb))) return false;
        o.x = l; o.y = t; o.w = rr - l; o.h = bb - t;
        if (!(std::isfinite(o.w) && std::isfinite(o.h))) return false;
    }
    if (non_zero && !(o.w > 0.0f && o.h > 0.0f)) return false;
    *out = o;
    return true;
}

// NonZeroRect::to_int_rect: floor the origin, ceil the extent (at least 1)
bool to_int_rect(const Rect &r, IntRect *out)
{
    uint32_t w = f2u(ceilf(r.w)), h = f2u(ceilf(r.h));
    return int_rect_from_xywh(f2i(floorf(r.x)), f2i(floorf(r.y)), w < 1 ? 1 : w, h < 1 ? 1 : h, out);
}

// strict_num NormalizedF32::to_u8
uint8_t opacity_to_u8(float o) { return (uint8_t)f2u(ceilf(o * 255.0f)); }
static inline float clamp01(float v) { return v > 1.0f ? 1.0f : (v >= 0.0f ? v : 0.0f); } // NormalizedF32::new_clamped

// tiny_skia Color{r,g,b,a}.premultiply().to_color_u8()
void premultiplied_u8(float r, float g, float b, float a, uint8_t out[4])
{
    float c[4] = {r, g, b, a};
    if (a != 1.0f) for (int i = 0; i < 3; i++) c[i] = clamp01(c[i] * a);
    for (int i = 0; i < 4; i++) out[i] = (uint8_t)f2u(c[i] * 255.0f + 0.5f);
}

// lib.rs:86-97
IntRect max_filter_bbox(uint32_t width, uint32_t height)
{
    const int64_t w = width > (uint32_t)INT32_MAX ? INT32_MAX : (int32_t)width, h = height > (uint32_t)INT32_MAX ? INT32_MAX : (int32_t)height;
    int64_t x = w * -2, y = h * -2; // saturating_mul
    if (x < INT32_MIN) x = INT32_MIN;
    if (y < INT32_MIN) y = INT32_MIN;
    uint64_t ww = (uint64_t)width * 5, hh = (uint64_t)height * 5;
    if (ww > UINT32_MAX) ww = UINT32_MAX;
    if (hh > UINT32_MAX) hh = UINT32_MAX;
    IntRect r;
    if (int_rect_from_xywh(x, y, ww, hh, &r)) return r;
    int_rect_from_ltrb(INT32_MIN / 2, INT32_MIN / 2, INT32_MAX / 2, INT32_MAX / 2, &r);
    return r;
}

// render.rs:145-164
static const int kBlendMap[16] = {RB_BLEND_SOURCE_OVER, RB_BLEND_MULTIPLY, RB_BLEND_SCREEN, RB_BLEND_OVERLAY, RB_BLEND_DARKEN,
                                  RB_BLEND_LIGHTEN, RB_BLEND_COLOR_DODGE, RB_BLEND_COLOR_BURN, RB_BLEND_HARD_LIGHT,
                                  RB_BLEND_SOFT_LIGHT, RB_BLEND_DIFFERENCE, RB_BLEND_EXCLUSION, RB_BLEND_HUE,
                                  RB_BLEND_SATURATION, RB_BLEND_COLOR, RB_BLEND_LUMINOSITY};
int convert_blend_mode(int usvg_mode) { return kBlendMap[usvg_mode & 15]; }

// ---------------------------------------------------------------------------------------------------------------------
// Renderer
// ---------------------------------------------------------------------------------------------------------------------
int Renderer::new_layer(uint32_t w, uint32_t h, Layer *out)
{
    rb_layer *l = nullptr;
    int st = rb_layer_create(rb, w, h, &l);
    if (st != RB_OK) return st;
    out->reset(l);
    return RB_OK;
}

// rows [y0, y0 + height) of a virtual full_w x full_h pixmap; an ordinary layer is the window of itself
struct Win { bool on; int32_t y0; uint32_t full_w, full_h; };
static Win window_of(const rb_layer *l)
{
    if (l->vp_w > 0) return Win{true, -l->vp_y, (uint32_t)l->vp_w, (uint32_t)l->vp_h};
    return Win{false, 0, l->w, l->h};
}

int Renderer::new_layer_like(const rb_layer *like, Layer *out)
{
    int st = new_layer(like->w, like->h, out);
    if (st != RB_OK) return st;
    (*out)->vp_x = like->vp_x; (*out)->vp_y = like->vp_y; (*out)->vp_w = like->vp_w; (*out)->vp_h = like->vp_h;
    return RB_OK;
}

// A failed device call ends the traversal: the status is kept and every later step is skipped.
#define RBR_TRY(call)                          \
    do {                                       \
        int st__ = (call);                     \
        if (st__ != RB_OK) { fail(st__); return; } \
    } while (0)

void Renderer::fail(int st)
{
    if (status == RB_OK) status = st;
}

// render.rs:10-19
void Renderer::render_nodes(const rbt::Group &parent, const Context &ctx, const Xform &ts, rb_layer *pixmap)
{
    for (const rbt::Node &n : parent.children) {
        if (status != RB_OK) return;
        render_node(n, ctx, ts, pixmap);
    }
}

// render.rs:21-47 (usvg::Node::Text arrives as its flattened() group)
void Renderer::render_node(const rbt::Node &node, const Context &ctx, const Xform &ts, rb_layer *pixmap)
{
    switch (node.kind) {
    case 0: render_group(*node.group, ctx, ts, pixmap); break;
    case 1: render_path(*node.path, RB_BLEND_SOURCE_OVER, ctx, ts, pixmap); break;
    default: render_image(*node.image, ts, pixmap);
    }
}

// render.rs:49-143
void Renderer::render_group(const rbt::Group &group, const Context &ctx, const Xform &transform_in, rb_layer *pixmap)
{
    const Xform transform = rbh::pre_concat(transform_in, group.ts);
    if (!group.should_isolate()) {
        render_nodes(group, ctx, transform, pixmap);
        return;
    }
    Rect bbox;
    if (!rect_transform(group.layer_bbox, transform, true, &bbox)) return;
    IntRect ibbox;
    if (group.filters.empty()) {
        // the group's bbox as integers, each side moved outwards by 2 px so anti-aliased pixels are not clipped
        const int64_t x = (int64_t)f2i(floorf(bbox.x)) - 2, y = (int64_t)f2i(floorf(bbox.y)) - 2;
        if (x < INT32_MIN || y < INT32_MIN) return; // checked_sub
        const uint64_t w = (uint64_t)f2u(ceilf(bbox.w)) + 4, h = (uint64_t)f2u(ceilf(bbox.h)) + 4;
        if (w > UINT32_MAX || h > UINT32_MAX) return; // checked_add
        if (!int_rect_from_xywh(x, y, w, h, &ibbox)) return;
        if (!fit_to_rect(ibbox, ctx.max_bbox, &ibbox)) return; // no layer larger than 4x the canvas
    } else {
        // a filter region already is a clipping region: not expanded
        const float cw = fmaxf(ceilf(bbox.w), 1.0f), ch = fmaxf(ceilf(bbox.h), 1.0f);
        IntRect r;
        if (!int_rect_from_xywh(f2i(floorf(bbox.x)), f2i(floorf(bbox.y)), f2u(cw), f2u(ch), &r)) return;
        if (!fit_to_rect(r, ctx.max_bbox, &ibbox)) return;
    }
    // canvas strips: inside a window only the rows of the group's layer that can reach it are rendered
    const Win win = window_of(pixmap);
    int32_t sub_top = ibbox.y;
    uint32_t sub_rows = ibbox.h;
    if (win.on) {
        const int64_t t = std::max<int64_t>(ibbox.y, win.y0), b = std::min<int64_t>((int64_t)ibbox.y + ibbox.h, (int64_t)win.y0 + rb_layer_height(pixmap));
        if (b <= t) return; // the group misses the window
        if (group.filters.empty()) { sub_top = (int32_t)t; sub_rows = (uint32_t)(b - t); } // a filter reads beyond its rows: whole layer
    }
    // keep the sub-pixel phase of the layer (render.rs:94-106)
    float dx = bbox.x, dy = bbox.y;
    dx -= bbox.x - (float)ibbox.x;
    dy -= bbox.y - (float)ibbox.y;
    const Xform ts = rbh::pre_concat(translate(-dx, -dy), transform);

    Layer sub;
    {
        int st = new_layer(ibbox.w, sub_rows, &sub);
        if (st == RB_ERR_OOM || st == RB_ERR_INVALID) return; // "Failed to allocate a group layer": the group is skipped
        if (st != RB_OK) { fail(st); return; }
        if (sub_rows != ibbox.h) { sub->vp_x = 0; sub->vp_y = -(sub_top - ibbox.y); sub->vp_w = (int32_t)ibbox.w; sub->vp_h = (int32_t)ibbox.h; }
    }
    render_nodes(group, ctx, ts, sub.get());
    for (const rbt::Filter &f : group.filters) {
        if (status != RB_OK) return;
        apply_filter(f, ts, sub.get());
    }
    if (group.clip_path) clip_apply(*group.clip_path, ts, sub.get());
    if (group.mask) mask_apply(*group.mask, ctx, ts, sub.get());
    if (status != RB_OK) return;
    RBR_TRY(rb_draw_layer(pixmap, sub.get(), ibbox.x, sub_top - win.y0, group.opacity, convert_blend_mode(group.blend_mode)));
}

// ---- path.rs ---------------------------------------------------------------------------------------------------------

// path.rs:6-24
void Renderer::render_path(const rbt::Path &path, int blend_mode, const Context &ctx, const Xform &ts, rb_layer *pixmap)
{
    if (!path.visible) return;
    if (path.paint_order == 0) {
        fill_path(path, blend_mode, ctx, ts, pixmap);
        stroke_path(path, blend_mode, ctx, ts, pixmap);
    } else {
        stroke_path(path, blend_mode, ctx, ts, pixmap);
        fill_path(path, blend_mode, ctx, ts, pixmap);
    }
}

// path.rs:45-71 / 89-111 + convert_linear_gradient / convert_radial_gradient / convert_base_gradient (path.rs:118-177):
// usvg::Paint + opacity -> tiny_skia::Paint.  `stops` and `pattern` keep what the rb_paint points to alive.
bool Renderer::convert_paint(const rbt::Paint &p, float opacity, bool anti_alias, int blend_mode, const Context &ctx, const Xform &ts,
                             rb_paint *out, std::vector<float> *stops, Layer *pattern)
{
    memset(out, 0, sizeof(*out));
    out->opacity = 1.0f;
    out->ts[0] = out->ts[3] = 1.0f;
    out->anti_alias = anti_alias ? 1 : 0;
    out->blend_mode = blend_mode;
    switch (p.kind) {
    case 0: // paint.set_color_rgba8(c.red, c.green, c.blue, opacity.to_u8())
        out->shader = RB_SHADER_SOLID;
        out->color[0] = (float)p.r / 255.0f; out->color[1] = (float)p.g / 255.0f; out->color[2] = (float)p.b / 255.0f;
        out->color[3] = (float)opacity_to_u8(opacity) / 255.0f;
        return true;
    case 1:
    case 2: {
        out->shader = p.kind == 1 ? RB_SHADER_LINEAR : RB_SHADER_RADIAL;
        out->spread = p.spread; // usvg SpreadMethod and tiny-skia SpreadMode share Pad, Reflect, Repeat
        stops->clear();
        for (const rbt::Stop &s : p.stops) {
            const float alpha = clamp01(s.opacity * opacity); // stop.opacity() * opacity
            stops->push_back(s.offset);
            stops->push_back((float)s.r / 255.0f); stops->push_back((float)s.g / 255.0f); stops->push_back((float)s.b / 255.0f);
            stops->push_back((float)opacity_to_u8(alpha) / 255.0f);
        }
        out->n_stops = (int32_t)p.stops.size();
        out->stops = stops->data();
        if (p.kind == 1) { out->x0 = p.x1; out->y0 = p.y1; out->x1 = p.x2; out->y1 = p.y2; }
        else { out->x0 = p.fx; out->y0 = p.fy; out->r0 = p.fr; out->x1 = p.cx; out->y1 = p.cy; out->r1 = p.rr; }
        const Xform &g = p.ts;
        out->ts[0] = g.sx; out->ts[1] = g.ky; out->ts[2] = g.kx; out->ts[3] = g.sy; out->ts[4] = g.tx; out->ts[5] = g.ty;
        return true;
    }
    default: {
        Xform pts;
        if (!render_pattern_pixmap(p, ctx, ts, pattern, &pts)) return false;
        out->shader = RB_SHADER_PATTERN;
        out->pattern = pattern->get();
        out->spread = RB_SPREAD_REPEAT;
        out->quality = RB_QUALITY_BICUBIC;
        out->opacity = opacity;
        out->ts[0] = pts.sx; out->ts[1] = pts.ky; out->ts[2] = pts.kx; out->ts[3] = pts.sy; out->ts[4] = pts.tx; out->ts[5] = pts.ty;
        return true;
    }
    }
}

static void xf_to_array(const Xform &t, float a[6]) { a[0] = t.sx; a[1] = t.ky; a[2] = t.kx; a[3] = t.sy; a[4] = t.tx; a[5] = t.ty; }

// path.rs:26-75
void Renderer::fill_path(const rbt::Path &path, int blend_mode, const Context &ctx, const Xform &ts, rb_layer *pixmap)
{
    if (!path.fill) return;
    if (path.bounds_w == 0.0f || path.bounds_h == 0.0f) return; // horizontal and vertical lines cannot be filled
    rb_paint paint;
    std::vector<float> stops;
    Layer pattern;
    if (!convert_paint(path.fill->paint, path.fill->opacity, path.anti_alias, blend_mode, ctx, ts, &paint, &stops, &pattern)) return;
    float t6[6];
    xf_to_array(ts, t6);
    int st = rb_fill_path(pixmap, path.verbs.data(), (int32_t)path.verbs.size(), path.pts.data(), (int32_t)(path.pts.size() / 2), &paint,
                          path.fill->rule ? RB_FILL_EVENODD : RB_FILL_WINDING, t6);
    if (st != RB_OK && st != RB_ERR_INVALID) fail(st); // INVALID = a shader tiny-skia refuses to build (Option::None): skipped
}

// path.rs:77-116; Stroke::to_tiny_skia tree/mod.rs:638-664
void Renderer::stroke_path(const rbt::Path &path, int blend_mode, const Context &ctx, const Xform &ts, rb_layer *pixmap)
{
    if (!path.stroke) return;
    const rbt::Stroke &s = *path.stroke;
    rb_paint paint;
    std::vector<float> stops;
    Layer pattern;
    if (!convert_paint(s.paint, s.opacity, path.anti_alias, blend_mode, ctx, ts, &paint, &stops, &pattern)) return;
    rb_stroke sk;
    memset(&sk, 0, sizeof(sk));
    sk.width = s.width;
    sk.miter_limit = s.miterlimit;
    sk.cap = s.linecap;
    sk.join = s.linejoin;
    sk.dash_array = s.dasharray.empty() ? nullptr : s.dasharray.data();
    sk.n_dash = (int32_t)s.dasharray.size();
    sk.dash_offset = s.dashoffset;
    float t6[6];
    xf_to_array(ts, t6);
    int st = rb_stroke_path(pixmap, path.verbs.data(), (int32_t)path.verbs.size(), path.pts.data(), (int32_t)(path.pts.size() / 2), &paint,
                            &sk, t6);
    if (st != RB_OK && st != RB_ERR_INVALID) fail(st);
}

// path.rs:179-205
bool Renderer::render_pattern_pixmap(const rbt::Paint &pattern, const Context &ctx, const Xform &transform, Layer *out, Xform *out_ts)
{
    float sx, sy;
    get_scale(rbh::pre_concat(transform, pattern.ts), &sx, &sy);
    const Rect &rect = pattern.rect;
    const uint32_t iw = f2u(roundf(rect.w * sx)), ih = f2u(roundf(rect.h * sy));
    if (iw == 0 || ih == 0) return false; // IntSize::from_wh
    int st = new_layer(iw, ih, out);
    if (st == RB_ERR_OOM || st == RB_ERR_INVALID) return false;
    if (st != RB_OK) { fail(st); return false; }
    render_nodes(*pattern.root, ctx, scale(sx, sy), out->get());
    if (status != RB_OK) return false;
    Xform ts;
    ts = rbh::pre_concat(ts, pattern.ts);
    ts = rbh::pre_concat(ts, translate(rect.x, rect.y));
    ts = rbh::pre_concat(ts, scale(1.0f / sx, 1.0f / sy));
    *out_ts = ts;
    return true;
}

// ---- clip.rs ---------------------------------------------------------------------------------------------------------

// clip.rs:6-28
void Renderer::clip_apply(const rbt::ClipPath &clip, const Xform &transform, rb_layer *pixmap)
{
    if (status != RB_OK) return;
    Layer clip_pixmap;
    RBR_TRY(new_layer_like(pixmap, &clip_pixmap));
    RBR_TRY(rb_layer_fill(clip_pixmap.get(), 0, 0, 0, 255)); // Color::BLACK
    clip_draw_children(*clip.root, RB_BLEND_CLEAR, rbh::pre_concat(transform, clip.ts), clip_pixmap.get());
    if (clip.clip_path) clip_apply(*clip.clip_path, transform, pixmap);
    if (status != RB_OK) return;
    // Mask::from_pixmap(clip_pixmap, Alpha); mask.invert(); pixmap.apply_mask(&mask) — one pass over both layers
    RBR_TRY(rb_layer_apply_clip_layer(pixmap, clip_pixmap.get()));
}

// clip.rs:30-68
void Renderer::clip_draw_children(const rbt::Group &parent, int mode, const Xform &transform, rb_layer *pixmap)
{
    const Context ctx{IntRect{0, 0, 1, 1}}; // "We could use any values here. They will not be used anyway."
    for (const rbt::Node &child : parent.children) {
        if (status != RB_OK) return;
        if (child.kind == 1) {
            if (!child.path->visible) continue;
            fill_path(*child.path, mode, ctx, transform, pixmap);
        } else if (child.kind == 0) {
            const rbt::Group &group = *child.group;
            const Xform ts = rbh::pre_concat(transform, group.ts);
            // a clipPath child with a clip-path of its own is drawn on a new canvas, clipped, then drawn onto the clipPath
            if (group.clip_path) clip_group(group, *group.clip_path, ts, pixmap);
            else clip_draw_children(group, mode, ts, pixmap);
        }
    }
}

// clip.rs:70-97
void Renderer::clip_group(const rbt::Group &children, const rbt::ClipPath &clip, const Xform &transform, rb_layer *pixmap)
{
    Layer clip_pixmap;
    RBR_TRY(new_layer_like(pixmap, &clip_pixmap));
    clip_draw_children(children, RB_BLEND_SOURCE_OVER, transform, clip_pixmap.get());
    clip_apply(clip, transform, clip_pixmap.get());
    if (status != RB_OK) return;
    RBR_TRY(rb_draw_layer(pixmap, clip_pixmap.get(), 0, 0, 1.0f, RB_BLEND_XOR));
}

// ---- mask.rs ---------------------------------------------------------------------------------------------------------

// mask.rs:6-46
void Renderer::mask_apply(const rbt::Mask &mask, const Context &ctx, const Xform &transform, rb_layer *pixmap)
{
    if (status != RB_OK) return;
    if (mask.root->children.empty()) {
        RBR_TRY(rb_layer_fill(pixmap, 0, 0, 0, 0));
        return;
    }
    const uint32_t w = rb_layer_width(pixmap), h = rb_layer_height(pixmap);
    Layer mask_pixmap;
    RBR_TRY(new_layer_like(pixmap, &mask_pixmap));
    {
        // the mask content is clipped by mask.rect()
        MaskHolder alpha_mask;
        {
            rb_mask *m = nullptr;
            RBR_TRY(rb_mask_create(rb, w, h, &m));
            alpha_mask.reset(m);
            m->vp_x = pixmap->vp_x; m->vp_y = pixmap->vp_y; m->vp_w = pixmap->vp_w; m->vp_h = pixmap->vp_h;
        }
        const Rect &r = mask.rect;
        const float right = r.x + r.w, bottom = r.y + r.h; // to_rect()
        const uint8_t verbs[5] = {RB_VERB_MOVE, RB_VERB_LINE, RB_VERB_LINE, RB_VERB_LINE, RB_VERB_CLOSE};
        const float pts[8] = {r.x, r.y, right, r.y, right, bottom, r.x, bottom}; // PathBuilder::from_rect
        float t6[6];
        xf_to_array(transform, t6);
        int st = rb_mask_fill_path(alpha_mask.get(), verbs, 5, pts, 4, RB_FILL_WINDING, 1, t6);
        if (st != RB_OK && st != RB_ERR_INVALID) { fail(st); return; }
        render_nodes(*mask.root, ctx, transform, mask_pixmap.get());
        if (status != RB_OK) return;
        RBR_TRY(rb_layer_apply_mask(mask_pixmap.get(), alpha_mask.get()));
    }
    if (mask.mask) mask_apply(*mask.mask, ctx, transform, pixmap);
    if (status != RB_OK) return;
    // Mask::from_pixmap(mask_pixmap, kind); pixmap.apply_mask(&mask) — one pass over both layers
    RBR_TRY(rb_layer_apply_mask_layer(pixmap, mask_pixmap.get(), mask.kind == 0 ? 1 : 0));
}

// ---- image.rs --------------------------------------------------------------------------------------------------------

// image.rs:4-35
void Renderer::render_image(const rbt::Image &image, const Xform &transform, rb_layer *pixmap)
{
    if (!image.visible) return;
    if (image.kind == 0) render_vector(*image.tree, transform, pixmap);
    else render_raster(image, transform, pixmap);
}

// image.rs:37-54
void Renderer::render_vector(const rbt::Tree &tree, const Xform &transform, rb_layer *pixmap)
{
    Layer sub;
    const Win win = window_of(pixmap); // inside a window the nested document is still rendered on a layer of the whole size
    RBR_TRY(new_layer(win.full_w, win.full_h, &sub));
    render_tree(tree, transform, sub.get());
    if (status != RB_OK) return;
    RBR_TRY(rb_draw_layer(pixmap, sub.get(), 0, -win.y0, 1.0f, RB_BLEND_SOURCE_OVER));
}

// image.rs:173-206 (the decoders of image.rs:62-170 stay on the host: the stream carries premultiplied RGBA8)
void Renderer::render_raster(const rbt::Image &image, const Xform &transform, rb_layer *pixmap)
{
    Layer raster;
    RBR_TRY(new_layer(image.w, image.h, &raster));
    RBR_TRY(rb_layer_upload(raster.get(), image.pixels.data()));
    rb_paint paint;
    memset(&paint, 0, sizeof(paint));
    paint.shader = RB_SHADER_PATTERN;
    paint.pattern = raster.get();
    paint.spread = RB_SPREAD_PAD;
    paint.quality = image.quality;
    paint.opacity = 1.0f;
    paint.ts[0] = paint.ts[3] = 1.0f;
    paint.blend_mode = RB_BLEND_SOURCE_OVER;
    paint.anti_alias = 1; // Paint::default()
    float t6[6];
    xf_to_array(transform, t6);
    int st = rb_fill_rect(pixmap, 0.0f, 0.0f, (float)image.w, (float)image.h, &paint, t6);
    if (st != RB_OK && st != RB_ERR_INVALID) fail(st);
}

// lib.rs:34-43
void Renderer::render_tree(const rbt::Tree &tree, const Xform &transform, rb_layer *pixmap)
{
    const Win win = window_of(pixmap);
    const Context ctx{max_filter_bbox(win.full_w, win.full_h)};
    render_nodes(tree.root, ctx, transform, pixmap);
}

// ---- node lookup for render_node (lib.rs:55-70) ----------------------------------------------------------------------
static bool find_in_group(const rbt::Group &g, const char *id, NodeRef *out);

static bool find_in_node(const rbt::Node &n, const char *id, NodeRef *out)
{
    if (n.kind == 0) {
        if (n.group->id == id) { out->node = &n; return true; }
        return find_in_group(*n.group, id, out);
    }
    if (n.kind == 1 && n.path->id == id) { out->node = &n; return true; }
    if (n.kind == 2 && n.image->id == id) { out->node = &n; return true; }
    return false;
}

static bool find_in_group(const rbt::Group &g, const char *id, NodeRef *out)
{
    for (const rbt::Node &n : g.children)
        if (find_in_node(n, id, out)) return true;
    return false;
}

// usvg Tree::node_by_id: depth-first over the rendered tree (an empty id never matches)
bool find_node(const rbt::Tree &tree, const char *id, NodeRef *out)
{
    if (!id || !*id) return false;
    return find_in_group(tree.root, id, out);
}

// Node::abs_layer_bounding_box (tree/mod.rs:984-992): NonZeroRect, i.e. None for a zero-sized node
bool node_abs_layer_bbox(const rbt::Node &n, Rect *out)
{
    const Rect *r = nullptr;
    if (n.kind == 0) r = &n.group->abs_layer_bbox;
    else if (n.kind == 1) { if (!n.path->has_abs_bbox) return false; r = &n.path->abs_layer_bbox; }
    else { if (!n.image->has_abs_bbox) return false; r = &n.image->abs_layer_bbox; }
    if (!(r->w > 0.0f && r->h > 0.0f)) return false;
    *out = *r;
    return true;
}

} // namespace rbr

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int rb_tree_parse(const void *stream, size_t len, rb_tree **out)
{
    if (!out) return RB_ERR_INVALID;
    *out = nullptr;
    std::string err;
    std::unique_ptr<rbt::Tree> t = rbt::parse(stream, len, &err);
    if (!t) return RB_ERR_INVALID;
    rb_tree *h = new rb_tree();
    h->t = std::move(t);
    *out = h;
    return RB_OK;
}

extern "C" void rb_tree_destroy(rb_tree *tree) { delete tree; }

extern "C" int rb_tree_size(const rb_tree *tree, float *width, float *height)
{
    if (!tree || !width || !height) return RB_ERR_INVALID;
    *width = tree->t->width;
    *height = tree->t->height;
    return RB_OK;
}

extern "C" int rb_tree_node_bbox(const rb_tree *tree, const char *id, float out_xywh[4])
{
    if (!tree || !out_xywh) return RB_ERR_INVALID;
    rbr::NodeRef ref;
    if (!rbr::find_node(*tree->t, id, &ref)) return RB_ERR_INVALID;
    rbt::Rect r;
    if (!rbr::node_abs_layer_bbox(*ref.node, &r)) return RB_ERR_INVALID;
    out_xywh[0] = r.x; out_xywh[1] = r.y; out_xywh[2] = r.w; out_xywh[3] = r.h;
    return RB_OK;
}

extern "C" int rb_render(rb_ctx *ctx, const rb_tree *tree, const float ts[6], rb_layer *target)
{
    if (!ctx || !tree || !target) return RB_ERR_INVALID;
    rbr::Renderer r(ctx);
    r.render_tree(*tree->t, ts ? rbh::Xform::from(ts) : rbh::Xform(), target);
    return r.status;
}

// Canvas-strip sharding of a tree: rows [y0, y0 + height of target) of the canvas_w x canvas_h render.
extern "C" int rb_render_strip(rb_ctx *ctx, const rb_tree *tree, const float ts[6], uint32_t canvas_w, uint32_t canvas_h, int32_t y0, rb_layer *target)
{
    if (!ctx || !tree || !target || canvas_w == 0 || canvas_h == 0 || y0 < 0 || rb_layer_width(target) != canvas_w
        || (uint64_t)y0 + rb_layer_height(target) > canvas_h)
        return RB_ERR_INVALID;
    int st = rb_layer_flush(target);
    if (st != RB_OK) return st;
    rbr::Renderer r(ctx);
    target->vp_x = 0; target->vp_y = -y0; target->vp_w = (int32_t)canvas_w; target->vp_h = (int32_t)canvas_h;
    r.render_tree(*tree->t, ts ? rbh::Xform::from(ts) : rbh::Xform(), target);
    st = rb_layer_flush(target); // the draws recorded with the strip's viewport
    target->vp_x = target->vp_y = target->vp_w = target->vp_h = 0;
    return r.status != RB_OK ? r.status : st;
}

extern "C" int rb_render_node(rb_ctx *ctx, const rb_tree *tree, const char *id, const float ts[6], rb_layer *target)
{
    if (!ctx || !tree || !target) return RB_ERR_INVALID;
    rbr::NodeRef ref;
    if (!rbr::find_node(*tree->t, id, &ref)) return RB_ERR_INVALID;
    rbt::Rect bbox;
    if (!rbr::node_abs_layer_bbox(*ref.node, &bbox)) return RB_ERR_INVALID; // render_node's None: a zero-sized node
    rbr::Renderer r(ctx);
    const rbr::Context c{rbr::max_filter_bbox(rb_layer_width(target), rb_layer_height(target))};
    rbh::Xform transform = ts ? rbh::Xform::from(ts) : rbh::Xform();
    transform = rbh::pre_concat(transform, rbr::translate(-bbox.x, -bbox.y)); // pre_translate
    r.render_node(*ref.node, c, transform, target);
    return r.status;
}

extern "C" int rb_submit(rb_ctx *ctx, const void *stream, size_t len, const float ts[6], rb_layer *target)
{
    rb_tree *t = nullptr;
    int st = rb_tree_parse(stream, len, &t);
    if (st != RB_OK) return st;
    st = rb_render(ctx, t, ts, target);
    rb_tree_destroy(t);
    return st;
}

// c-api resvg_render (crates/c-api/lib.rs:875-893): the caller's host pixmap is the canvas — drawn over, not cleared.
extern "C" int rb_render_to_host(rb_ctx *ctx, const rb_tree *tree, const float ts[6], uint32_t width, uint32_t height, uint8_t *pixmap)
{
    if (!ctx || !tree || !pixmap || width == 0 || height == 0) return RB_ERR_INVALID;
    rb_layer *l = nullptr;
    int st = rb_layer_create(ctx, width, height, &l);
    if (st != RB_OK) return st;
    st = rb_layer_upload(l, pixmap);
    if (st == RB_OK) st = rb_render(ctx, tree, ts, l);
    if (st == RB_OK) st = rb_layer_download(l, pixmap);
    rb_layer_destroy(l);
    return st;
}
