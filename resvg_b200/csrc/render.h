// render.h — the render traversal (render.cpp) and the filter-graph executor (filter_exec.cpp) behind rb_render.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "rb_internal.h"
#include "tree.h"

namespace rbr {

using rbh::Xform;

struct LayerDeleter { void operator()(rb_layer *l) const { rb_layer_destroy(l); } };
struct MaskDeleter { void operator()(rb_mask *m) const { rb_mask_destroy(m); } };
using Layer = std::unique_ptr<rb_layer, LayerDeleter>;
using MaskHolder = std::unique_ptr<rb_mask, MaskDeleter>;

// render.rs:6-8
struct Context { rbt::IntRect max_bbox; };

struct NodeRef { const rbt::Node *node = nullptr; };

Xform translate(float tx, float ty);
Xform scale(float sx, float sy);
void get_scale(const Xform &t, float *sx, float *sy);
bool int_rect_from_xywh(int64_t x, int64_t y, uint64_t w, uint64_t h, rbt::IntRect *out);
bool int_rect_from_ltrb(int64_t l, int64_t t, int64_t r, int64_t b, rbt::IntRect *out);
bool fit_to_rect(const rbt::IntRect &r, const rbt::IntRect &bounds, rbt::IntRect *out);
bool rect_transform(const rbt::Rect &r, const Xform &ts, bool non_zero, rbt::Rect *out);
bool to_int_rect(const rbt::Rect &r, rbt::IntRect *out);
uint8_t opacity_to_u8(float o);
void premultiplied_u8(float r, float g, float b, float a, uint8_t out[4]);
rbt::IntRect max_filter_bbox(uint32_t width, uint32_t height);
int convert_blend_mode(int usvg_mode);
bool find_node(const rbt::Tree &tree, const char *id, NodeRef *out);
bool node_abs_layer_bbox(const rbt::Node &n, rbt::Rect *out);

struct Renderer {
    rb_ctx *rb;
    int status = RB_OK; // first device error; once set the traversal unwinds without issuing more work
    explicit Renderer(rb_ctx *c) : rb(c) {}
    // Canvas strips (rb_render_strip): a layer may be a WINDOW — rows [y0, y0 + its height) — of a larger
    // virtual pixmap (rb_layer::vp_*).  Draws into a window are built against the whole virtual pixmap (the batch viewport),
    // an isolated group inside a window gets a window of its own layer (the rows that can reach the parent's window), its
    // clip / mask pixmaps the same window; a group with filters is rendered whole (a filter reads beyond its rows) and
    // composited shifted; groups that miss the window are skipped.  Every pixel equals the whole-canvas render's.
    int new_layer_like(const rb_layer *like, Layer *out);

    void fail(int st);
    int new_layer(uint32_t w, uint32_t h, Layer *out);

    // lib.rs / render.rs
    void render_tree(const rbt::Tree &tree, const Xform &transform, rb_layer *pixmap);
    void render_nodes(const rbt::Group &parent, const Context &ctx, const Xform &ts, rb_layer *pixmap);
    void render_node(const rbt::Node &node, const Context &ctx, const Xform &ts, rb_layer *pixmap);
    void render_group(const rbt::Group &group, const Context &ctx, const Xform &ts, rb_layer *pixmap);
    // path.rs
    void render_path(const rbt::Path &path, int blend_mode, const Context &ctx, const Xform &ts, rb_layer *pixmap);
    void fill_path(const rbt::Path &path, int blend_mode, const Context &ctx, const Xform &ts, rb_layer *pixmap);
    void stroke_path(const rbt::Path &path, int blend_mode, const Context &ctx, const Xform &ts, rb_layer *pixmap);
    bool convert_paint(const rbt::Paint &p, float opacity, bool anti_alias, int blend_mode, const Context &ctx, const Xform &ts,
                       rb_paint *out, std::vector<float> *stops, Layer *pattern);
    bool render_pattern_pixmap(const rbt::Paint &pattern, const Context &ctx, const Xform &transform, Layer *out, Xform *out_ts);
    // clip.rs
    void clip_apply(const rbt::ClipPath &clip, const Xform &transform, rb_layer *pixmap);
    void clip_draw_children(const rbt::Group &parent, int mode, const Xform &transform, rb_layer *pixmap);
    void clip_group(const rbt::Group &children, const rbt::ClipPath &clip, const Xform &transform, rb_layer *pixmap);
    // mask.rs
    void mask_apply(const rbt::Mask &mask, const Context &ctx, const Xform &transform, rb_layer *pixmap);
    // image.rs
    void render_image(const rbt::Image &image, const Xform &transform, rb_layer *pixmap);
    void render_vector(const rbt::Tree &tree, const Xform &transform, rb_layer *pixmap);
    void render_raster(const rbt::Image &image, const Xform &transform, rb_layer *pixmap);
    // filter/mod.rs (filter_exec.cpp)
    void apply_filter(const rbt::Filter &filter, const Xform &ts, rb_layer *source);
};

} // namespace rbr
