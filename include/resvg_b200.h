/*
 * resvg_b200.h — C ABI of the B200-native resvg pixel hot path.
 *
 * This is the drop-in seam described in SURVEY.md §8(b): the reference renderer
 * (crates/resvg/src/{render,path,clip,mask}.rs and crates/resvg/src/filter/mod.rs) depends only on
 * (1) a finite list of tiny-skia Pixmap/Mask calls and (2) the filter primitive functions of
 * crates/resvg/src/filter/*.rs.  Every entry point below replaces one of those and cites it.
 *
 * Conventions
 *   - every function returns an rb_status (0 = RB_OK); nothing throws across the boundary;
 *   - all pixel data is tightly packed premultiplied RGBA8888, R first, no stride — the resvg pixmap
 *     contract (crates/c-api/resvg.h:482-483);
 *   - an rb_layer is a device-resident Pixmap, an rb_mask a device-resident tiny-skia Mask (u8 plane);
 *   - work is enqueued on the context's CUDA stream; only *_download / rb_ctx_synchronize block;
 *   - there is no CPU fallback: without a usable CUDA device rb_ctx_create fails with RB_ERR_CUDA.
 */
#ifndef RESVG_B200_H
#define RESVG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_ctx rb_ctx;
typedef struct rb_layer rb_layer;
typedef struct rb_mask rb_mask;

typedef enum {
    RB_OK = 0,
    RB_ERR_INVALID = 1, /* bad argument (null, zero size, size mismatch) — reference: Option::None / warn+skip */
    RB_ERR_CUDA = 2,    /* CUDA runtime error; see rb_last_error() */
    RB_ERR_OOM = 3,
    RB_ERR_UNSUPPORTED = 4
} rb_status;

/* ------------------------------------------------------------------------------------------------
 * Context / memory
 * ---------------------------------------------------------------------------------------------- */
int rb_ctx_create(int device, rb_ctx **out);
void rb_ctx_destroy(rb_ctx *ctx);
int rb_ctx_synchronize(rb_ctx *ctx);
const char *rb_last_error(rb_ctx *ctx);
void *rb_ctx_stream(rb_ctx *ctx); /* cudaStream_t */
int rb_ctx_device(rb_ctx *ctx);
/* CUDA-event timer on the context's stream (bench.py times kernels with it). */
int rb_timer_begin(rb_ctx *ctx);
int rb_timer_end(rb_ctx *ctx, float *elapsed_ms); /* synchronises */
/* Number of kernel launches issued through this context since creation (bench "gpu_launches"). */
uint64_t rb_ctx_launch_count(rb_ctx *ctx);
/* bytes copied host -> device so far by batches (edge / paint blocks) and layer / mask uploads */
uint64_t rb_ctx_h2d_bytes(rb_ctx *ctx);
/* Pinned host memory for upload/download staging. */
int rb_host_alloc(size_t bytes, void **out);
void rb_host_free(void *p);

/* tiny_skia::Pixmap::new — zeroed premultiplied RGBA8 (render.rs:108, filter/mod.rs:101-103). */
int rb_layer_create(rb_ctx *ctx, uint32_t width, uint32_t height, rb_layer **out);
void rb_layer_destroy(rb_layer *layer);
uint32_t rb_layer_width(const rb_layer *layer);
uint32_t rb_layer_height(const rb_layer *layer);
void *rb_layer_device_ptr(rb_layer *layer);
/* PixmapMut::from_bytes / Pixmap::data (c-api/lib.rs:887-890): host <-> device copies. */
int rb_layer_upload(rb_layer *layer, const uint8_t *host_rgba);
int rb_layer_download(rb_layer *layer, uint8_t *host_rgba);
/* Asynchronous form: the copy is ordered after everything enqueued on the layer so far and runs on a copy stream of the
 * context, so the next render (into another layer) overlaps it; host_rgba should be pinned (rb_host_alloc).  The layer must
 * not be written, and host_rgba not read, until rb_layer_download_end has returned. */
int rb_layer_download_begin(rb_layer *layer, uint8_t *host_rgba);
int rb_layer_download_end(rb_layer *layer);
/* Pixmap::fill(color) with an already premultiplied RGBA8 colour (filter/mod.rs:112, 824). */
int rb_layer_fill(rb_layer *layer, uint8_t r, uint8_t g, uint8_t b, uint8_t a);
/* Pixmap::clone (filter/mod.rs:531): dst and src must have equal size. */
int rb_layer_copy(rb_layer *dst, const rb_layer *src);

/* ------------------------------------------------------------------------------------------------
 * Filter helpers — crates/resvg/src/filter/mod.rs
 * ---------------------------------------------------------------------------------------------- */
int rb_layer_multiply_alpha(rb_layer *layer);   /* mod.rs:129-136 */
int rb_layer_demultiply_alpha(rb_layer *layer); /* mod.rs:139-146 */
int rb_layer_into_linear_rgb(rb_layer *layer);  /* mod.rs:120-124, fused demultiply+LUT+multiply */
int rb_layer_into_srgb(rb_layer *layer);        /* mod.rs:114-118 */

/* ------------------------------------------------------------------------------------------------
 * Filter primitives — one per file of crates/resvg/src/filter/
 * ---------------------------------------------------------------------------------------------- */
/* box_blur::apply(sigma_x, sigma_y, src) — box_blur.rs:23 */
int rb_filter_box_blur(rb_layer *layer, double sigma_x, double sigma_y);
/* iir_blur::apply(sigma_x, sigma_y, src) — iir_blur.rs:47: strictly sequential f64 recurrences, bit-identical to the
 * reference.  This is what the renderer calls (filter/mod.rs:605, 661). */
int rb_filter_iir_blur(rb_layer *layer, double sigma_x, double sigma_y);
/* The same cascade in f32 on register-resident line segments with halos: within 1/255 of the reference on the stage output
 * (the tolerance BASELINE.json grants the IIR blur), ~10x faster.  Opt-in: a 1-level difference in a flat linearRGB area
 * becomes many levels after the conversion to sRGB, so corpus parity needs the exact form. */
int rb_filter_iir_blur_fast(rb_layer *layer, double sigma_x, double sigma_y);
/* morphology::apply(operator, rx, ry, src) — morphology.rs:15.  op: 0 erode, 1 dilate */
int rb_filter_morphology(rb_layer *layer, int op, float rx, float ry);
/* convolve_matrix::apply(matrix, src) — convolve_matrix.rs:15.
 * kernel is row-major (usvg ConvolveMatrixData::get(x,y) = data[y*columns+x]); edge_mode 0 none,
 * 1 duplicate, 2 wrap. */
int rb_filter_convolve_matrix(rb_layer *layer, const float *kernel, uint32_t columns, uint32_t rows,
                              uint32_t target_x, uint32_t target_y, float divisor, float bias,
                              int edge_mode, int preserve_alpha);
/* color_matrix::apply(kind, src) — color_matrix.rs:11.
 * kind 0 matrix (20 floats), 1 saturate (1), 2 hueRotate (1, degrees), 3 luminanceToAlpha. */
int rb_filter_color_matrix(rb_layer *layer, int kind, const float *params);
/* component_transfer::apply(fe, src) — component_transfer.rs:10.
 * type 0 identity, 1 table, 2 discrete, 3 linear, 4 gamma; order r,g,b,a. */
typedef struct {
    int32_t type;
    int32_t n_values;
    const float *values;
    float slope, intercept;
    float amplitude, exponent, offset;
} rb_transfer_fn;
int rb_filter_component_transfer(rb_layer *layer, const rb_transfer_fn funcs[4]);
/* Rows / columns the box blur of this standard deviation reads beyond a pixel (sum of the five box radii,
 * box_blur.rs:37-71): the halo a strip of a larger image needs for its own rows to equal the blur of the whole image. */
int rb_filter_box_blur_reach(double sigma);
/* box_blur::apply on n sub-pixmaps at once: rectangle i = rects[4i..4i+3] = (x, y, w, h) of the layer is blurred as a
 * pixmap of its own (windows clipped to it) with sigma_x[i], sigma_y[i]; everything else is untouched.  One launch per
 * pass for all rectangles (per-document filters on an atlas).  Rectangles must lie inside the layer and not overlap. */
int rb_filter_box_blur_cells(rb_layer *layer, int32_t n, const int32_t *rects, const double *sigma_x, const double *sigma_y);
/* apply_drop_shadow's flood step, filter/mod.rs:606-617: every pixel := the colour (r, g, b, a) with its opacity scaled
 * by the pixel's alpha / 255, premultiplied (Color::apply_opacity, premultiply, to_color_u8). */
int rb_filter_flood_alpha(rb_layer *layer, uint8_t r, uint8_t g, uint8_t b, uint8_t a);
/* composite::arithmetic(k1..k4, src1, src2, dest) — composite.rs:14 */
int rb_filter_composite_arithmetic(rb_layer *dest, const rb_layer *src1, const rb_layer *src2,
                                   float k1, float k2, float k3, float k4);
/* displacement_map::apply(fe, sx, sy, src, map, dest) — displacement_map.rs:15.  channel 0..3 = R,G,B,A */
int rb_filter_displacement_map(rb_layer *dest, const rb_layer *src, const rb_layer *map,
                               int x_channel, int y_channel, float scale, float sx, float sy);
/* lighting — lighting.rs:132 / :175.  The light source is the one already transformed by
 * transform_light_source (filter/mod.rs:1063-1098). */
typedef struct {
    int32_t kind;              /* 0 distant, 1 point, 2 spot */
    float azimuth, elevation;  /* distant, degrees */
    float x, y, z;             /* point / spot */
    float points_at_x, points_at_y, points_at_z;
    float specular_exponent;   /* spot */
    int32_t has_cone;
    float limiting_cone_angle; /* degrees */
} rb_light_source;
int rb_filter_diffuse_lighting(rb_layer *dest, const rb_layer *src, float surface_scale,
                               float diffuse_constant, uint8_t r, uint8_t g, uint8_t b,
                               const rb_light_source *light);
int rb_filter_specular_lighting(rb_layer *dest, const rb_layer *src, float surface_scale,
                                float specular_constant, float specular_exponent,
                                uint8_t r, uint8_t g, uint8_t b, const rb_light_source *light);
/* turbulence::apply(...) — turbulence.rs:33 (result is unpremultiplied; caller multiplies alpha as
 * filter/mod.rs:1015 does). */
int rb_filter_turbulence(rb_layer *dest, double offset_x, double offset_y, double sx, double sy,
                         double base_frequency_x, double base_frequency_y, uint32_t num_octaves,
                         int32_t seed, int stitch_tiles, int fractal_noise);

/* ------------------------------------------------------------------------------------------------
 * Rasteriser — the tiny-skia calls made by crates/resvg/src/{path,render,clip,mask}.rs
 * ---------------------------------------------------------------------------------------------- */
/* path verbs (tiny_skia_path::PathVerb) */
enum { RB_VERB_MOVE = 0, RB_VERB_LINE = 1, RB_VERB_QUAD = 2, RB_VERB_CUBIC = 3, RB_VERB_CLOSE = 4 };
/* tiny_skia::BlendMode in declaration order (render.rs:145-164 maps usvg::BlendMode onto it) */
enum {
    RB_BLEND_CLEAR = 0, RB_BLEND_SOURCE, RB_BLEND_DESTINATION, RB_BLEND_SOURCE_OVER, RB_BLEND_DESTINATION_OVER,
    RB_BLEND_SOURCE_IN, RB_BLEND_DESTINATION_IN, RB_BLEND_SOURCE_OUT, RB_BLEND_DESTINATION_OUT,
    RB_BLEND_SOURCE_ATOP, RB_BLEND_DESTINATION_ATOP, RB_BLEND_XOR, RB_BLEND_PLUS, RB_BLEND_MODULATE,
    RB_BLEND_SCREEN, RB_BLEND_OVERLAY, RB_BLEND_DARKEN, RB_BLEND_LIGHTEN, RB_BLEND_COLOR_DODGE,
    RB_BLEND_COLOR_BURN, RB_BLEND_HARD_LIGHT, RB_BLEND_SOFT_LIGHT, RB_BLEND_DIFFERENCE, RB_BLEND_EXCLUSION,
    RB_BLEND_MULTIPLY, RB_BLEND_HUE, RB_BLEND_SATURATION, RB_BLEND_COLOR, RB_BLEND_LUMINOSITY
};
enum { RB_SHADER_SOLID = 0, RB_SHADER_LINEAR = 1, RB_SHADER_RADIAL = 2, RB_SHADER_PATTERN = 3 };
enum { RB_SPREAD_PAD = 0, RB_SPREAD_REFLECT = 1, RB_SPREAD_REPEAT = 2 };
enum { RB_QUALITY_NEAREST = 0, RB_QUALITY_BILINEAR = 1, RB_QUALITY_BICUBIC = 2 };
enum { RB_FILL_WINDING = 0, RB_FILL_EVENODD = 1 };

/* Transform layout everywhere: ts[6] = {sx, ky, kx, sy, tx, ty} = tiny_skia::Transform::from_row order, i.e.
 * resvg_transform {a,b,c,d,e,f} (crates/c-api/lib.rs:67-81). NULL means identity. */

/* tiny_skia::Paint as path.rs builds it (path.rs:45-71, 118-177): shader + blend mode + anti_alias. */
typedef struct {
    int32_t shader;                 /* RB_SHADER_* */
    float color[4];                 /* solid: non-premultiplied r,g,b,a = Color::from_rgba8 (c / 255) */
    float x0, y0, r0, x1, y1, r1;   /* LinearGradient::new(start,end) / RadialGradient::new(start,r0,end,r1) */
    int32_t n_stops;
    const float *stops;             /* n_stops x {offset, r, g, b, a}, non-premultiplied */
    int32_t spread;                 /* RB_SPREAD_* */
    float ts[6];                    /* gradient.transform() / pattern transform */
    const rb_layer *pattern;        /* Pattern::new(pixmap, spread, quality, opacity, ts) */
    int32_t quality;
    float opacity;
    int32_t blend_mode;             /* RB_BLEND_* */
    int32_t anti_alias;
    int32_t force_hq;
} rb_paint;

/* PixmapMut::fill_path(path, paint, rule, transform, None) — path.rs:73.  Host: transform, chop, clip, edge
 * build; device: coverage + shade + blend.  The draw is validated and recorded at once but executed lazily: consecutive
 * rb_fill_path calls on a layer are collected and run as ONE batch (painter's order = call order) the next time any
 * entry point reads or writes that layer (composite, filter, mask, copy, download, rb_layer_device_ptr, an explicit
 * batch on it, rb_ctx_synchronize, rb_timer_end), so a traversal that issues fill after fill and then composites the
 * layer pays one tile-kernel launch.  Draws with a pattern paint are executed immediately.  The path, paint and
 * stops are copied; nothing the caller passed needs to outlive the call. */
int rb_fill_path(rb_layer *layer, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                 const rb_paint *paint, int32_t fill_rule, const float ts[6]);

/* Batched drawing: many fill_path calls against one layer executed by ONE tile-binned kernel launch that
 * keeps every destination tile on chip across all the draws that touch it (painter's order preserved). */
typedef struct rb_batch rb_batch;
int rb_batch_begin(rb_layer *target, rb_batch **out);
int rb_batch_fill_path(rb_batch *batch, const uint8_t *verbs, int32_t n_verbs, const float *points,
                       int32_t n_points, const rb_paint *paint, int32_t fill_rule, const float ts[6]);
/* tiny_skia::Stroke as usvg fills it (tree/mod.rs:638-664).  dash_array = StrokeDash::new(array, offset) arguments
 * (NULL / n_dash 0: no dashing; an array StrokeDash::new rejects — odd or < 2 entries, a negative entry, sum <= 0 —
 * leaves the stroke solid, as usvg does). */
typedef struct {
    float width, miter_limit;
    int32_t cap;  /* 0 butt, 1 round, 2 square */
    int32_t join; /* 0 miter, 1 miter-clip, 2 round, 3 bevel */
    const float *dash_array;
    int32_t n_dash;
    float dash_offset;
} rb_stroke;
/* PixmapMut::stroke_path(path, paint, stroke, transform, None) — path.rs:113: dash, then outline on the host
 * (rb_path_stroke) and a Winding fill; or, when tiny-skia's treat_as_hairline says so (width 0, or an anti-aliased
 * stroke whose transformed width is at most one pixel), the anti-aliased hairline walker (scan/hairline_aa.rs) with
 * the paint's alpha modulated by the width.  Batches holding hairlines must be executed with rb_batch_submit. */
int rb_batch_stroke_path(rb_batch *batch, const uint8_t *verbs, int32_t n_verbs, const float *points,
                         int32_t n_points, const rb_paint *paint, const rb_stroke *stroke, const float ts[6]);
/* Bulk recording of fills and strokes: strokes may be NULL; entry i is stroked when strokes[i].width > 0, filled
 * with fill_rules[i] otherwise.  Recorded BY REFERENCE: every array (and the stops the paints point to) must stay
 * valid and unchanged until the next rb_batch_prepare / rb_batch_submit on this batch has returned. */
int rb_batch_draw_paths(rb_batch *batch, int32_t n_paths, const uint32_t *verb_off, const uint32_t *point_off,
                        const uint8_t *verbs, const float *points, const rb_paint *paints, const uint8_t *fill_rules,
                        const rb_stroke *strokes, const float ts[6]);
/* Bulk form of rb_batch_fill_path: n_paths paths in packed arrays; verb_off / point_off hold n_paths + 1 prefix
 * offsets into verbs / points (points counted in x,y pairs); one paint and one fill rule per path. */
int rb_batch_fill_paths(rb_batch *batch, int32_t n_paths, const uint32_t *verb_off, const uint32_t *point_off,
                        const uint8_t *verbs, const float *points, const rb_paint *paints, const uint8_t *fill_rules,
                        const float ts[6]);
/* The draws recorded after this call are rendered as if the rectangle (x, y, w, h) of the target were a pixmap of its
 * own (what resvg::render gets for one document): coordinates are relative to its origin, nothing is drawn outside it.
 * One batch can thus render many small documents into one atlas layer (document-parallel thumbnailing, BASELINE
 * config 5).  The rectangle may reach beyond the target (negative x / y, larger than the target): the document is still
 * clipped and flattened against its whole pixmap and only what falls inside the target is drawn, pixel for pixel what
 * a render of the whole pixmap holds there — one GPU can thus render a strip of a large canvas (canvas-strip sharding,
 * SURVEY 8(e)).  w = h = 0 restores the whole target. */
int rb_batch_set_viewport(rb_batch *batch, int32_t x, int32_t y, uint32_t w, uint32_t h);
/* Bulk form of { rb_batch_set_viewport; rb_batch_draw_paths } per document (BASELINE config 5: one resvg::render per
 * icon): document k owns doc_count[k] paths starting at doc_first[k] of the packed arrays (same layout and lifetime rules
 * as rb_batch_draw_paths) and is rendered into viewports[4k..4k+3] = (x, y, w, h).  The current viewport is unchanged. */
int rb_batch_draw_documents(rb_batch *batch, int32_t n_docs, const int32_t *viewports, const uint32_t *doc_first,
                            const uint32_t *doc_count, const uint32_t *verb_off, const uint32_t *point_off,
                            const uint8_t *verbs, const float *points, const rb_paint *paints, const uint8_t *fill_rules,
                            const rb_stroke *strokes, const float ts[6]);
/* Builds edges on host threads (n_threads <= 0: all cores), uploads, launches, frees the device copy. */
int rb_batch_submit(rb_batch *batch, int32_t n_threads);
/* rb_batch_submit, then rb_layer_download_begin(layer, host) (w * h * 4 bytes; pinned memory from rb_host_alloc lets the
 * copy overlap): what resvg::render does with its host pixmap target (crates/resvg/src/lib.rs:34-53).  The last raster
 * launch of the submit runs in bands of tile rows and every finished band is copied out while the next one is rendered.
 * Returns once everything is enqueued; rb_layer_download_end(layer) waits for the pixels, which are those of
 * rb_batch_submit + rb_layer_download. */
int rb_batch_submit_download(rb_batch *batch, int32_t n_threads, uint8_t *host);
/* Tests: how many rb_batch_submit_download calls overlapped their download with the last raster launch so far. */
uint64_t rb_debug_banded_downloads(void);
/* Split form: prepare = host edge build + binning + upload (device copy stays resident in the batch);
 * run = the kernel launch only, repeatable (e.g. after clearing the layer). */
int rb_batch_prepare(rb_batch *batch, int32_t n_threads);
int rb_batch_run(rb_batch *batch);
/* rb_batch_run with pixel counters (roofline accounting): out[0] = pixels read-modify-written, out[1] = pixels
 * stored without a read (full coverage + opaque solid paint).  Synchronises the stream. */
int rb_batch_run_counting(rb_batch *batch, uint64_t out[2]);
/* Device time (CUDA events on the context's stream) of the last rb_batch_run / submit on this context:
 * ms[0] = binning + edge-list pre-pass kernels, ms[1] = the raster kernel.  Synchronises. */
int rb_ctx_last_run_ms(rb_ctx *ctx, float ms[2]);
void rb_batch_destroy(rb_batch *batch);
/* statistics of the last submit: [0] draws, [1] line edges, [2] (draw,tile) pairs, [3] non-empty tiles,
 * [4] bytes uploaded, [5] host build microseconds */
int rb_batch_stats(rb_batch *batch, uint64_t stats[6]);

/* PixmapMut::draw_pixmap(x, y, src, PixmapPaint{opacity, blend_mode, Nearest}, identity, None) —
 * render.rs:133, clip.rs:88, filter/mod.rs (9 call sites): the layer composite. */
int rb_draw_layer(rb_layer *dst, const rb_layer *src, int32_t x, int32_t y, float opacity, int32_t blend_mode);
/* Atlas form of the layer composite (render.rs:108-133 per document; the offset / merge draws of filter/mod.rs): for every
 * rectangle i = rects[4i..4i+3] = (x, y, w, h) of `dst`, the w x h pixels of `src` at src_xy[2i..2i+1] (NULL: the same
 * position) are drawn with opacity[i], as draw_pixmap of that sub-pixmap would; rectangles are clipped to both layers; one
 * launch for all documents of an atlas (see rb_batch_set_viewport).  Destination rectangles must not overlap. */
int rb_draw_layer_rects(rb_layer *dst, const rb_layer *src, int32_t n, const int32_t *rects, const int32_t *src_xy,
                        const float *opacity, int32_t blend_mode);

/* PixmapMut::stroke_path, immediate form of rb_batch_stroke_path (path.rs:113): recorded lazily like rb_fill_path. */
int rb_stroke_path(rb_layer *layer, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                   const rb_paint *paint, const rb_stroke *stroke, const float ts[6]);
/* PixmapMut::fill_rect(Rect::from_xywh(x, y, w, h), paint, transform, None) — filter/mod.rs:474-497 (clearing outside a
 * primitive subregion), :853 (feTile), image.rs:203 (raster images).  Integer rectangles under the identity and any
 * rectangle under another transform (tiny-skia then fills PathBuilder::from_rect) are drawn; a fractional anti-aliased
 * rectangle under the identity (tiny-skia's fill_rect_aa, never issued by resvg) is RB_ERR_UNSUPPORTED. */
int rb_fill_rect(rb_layer *layer, float x, float y, float w, float h, const rb_paint *paint, const float ts[6]);
/* Pixmap::clone_rect as PixmapExt::copy_region uses it (filter/mod.rs:104-108, feTile :849): the part of the rectangle
 * inside `src` as a new layer; RB_ERR_INVALID when they do not intersect. */
int rb_layer_clone_rect(const rb_layer *src, int32_t x, int32_t y, uint32_t w, uint32_t h, rb_layer **out);

/* tiny_skia::Mask — clip.rs:25-27, mask.rs:17-45 */
int rb_mask_create(rb_ctx *ctx, uint32_t width, uint32_t height, rb_mask **out); /* Mask::new (zeroed) */
void rb_mask_destroy(rb_mask *mask);
int rb_mask_download(rb_mask *mask, uint8_t *host);
int rb_mask_upload(rb_mask *mask, const uint8_t *host);
int rb_mask_from_layer(rb_mask *mask, const rb_layer *layer, int32_t luminance); /* Mask::from_pixmap */
int rb_mask_invert(rb_mask *mask);                                                /* Mask::invert */
int rb_layer_apply_mask(rb_layer *layer, const rb_mask *mask);                    /* Pixmap::apply_mask */
/* Mask::from_pixmap(mask_pixmap, Luminance | Alpha) + Pixmap::apply_mask fused (mask.rs:40-45), and
 * Mask::from_pixmap(clip_pixmap, Alpha) + Mask::invert + Pixmap::apply_mask fused (clip.rs:25-27): the mask value is a
 * function of the source pixel, so the u8 plane is never materialised.  Bit-identical to the three-call sequence. */
int rb_layer_apply_mask_layer(rb_layer *layer, const rb_layer *mask_pixmap, int32_t luminance);
int rb_layer_apply_clip_layer(rb_layer *layer, const rb_layer *clip_pixmap);
/* Mask::fill_path(path, rule, anti_alias, transform) */
int rb_mask_fill_path(rb_mask *mask, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                      int32_t fill_rule, int32_t anti_alias, const float ts[6]);

/* ------------------------------------------------------------------------------------------------
 * Whole-tree rendering — resvg::render / resvg::render_node (crates/resvg/src/lib.rs:34-70) and the C API's resvg_render
 * / resvg_render_node (crates/c-api/lib.rs:875-942).  The traversal of crates/resvg/src/{render,path,clip,mask,image}.rs
 * and the filter executor of filter/mod.rs run inside the library (csrc/render.cpp, csrc/filter_exec.cpp) against
 * device-resident layers: one call across the boundary per document.
 *
 * The usvg::Tree is handed over as a flat little-endian stream of 4-byte words ("RBT1") that the Rust shim writes once
 * per tree by walking usvg's public accessors (INTEGRATION.md shows the writer):
 *
 *   stream := u32 0x31544252 ("RBT1")  TREE
 *   TREE   := f32 width, f32 height (tree.size())  GROUP (tree.root())
 *   STR    := u32 n, n bytes, zero padding to a multiple of 4
 *   XF     := f32 sx, ky, kx, sy, tx, ty          RECT := f32 x, y, width, height        RGB := u32 r | g << 8 | b << 16
 *   GROUP  := STR id, XF transform, f32 opacity, u32 blend_mode (usvg::BlendMode order), u32 isolate,
 *             RECT layer_bounding_box, RECT abs_layer_bounding_box, u32 has_clip_path [CLIP], u32 has_mask [MASK],
 *             u32 n_filters FILTER*, u32 n_children NODE*
 *   NODE   := u32 0 GROUP (Node::Group, and Node::Text as text.flattened()) | u32 1 PATH | u32 2 IMAGE
 *   PATH   := STR id, u32 visible, u32 paint_order (0 FillAndStroke, 1 StrokeAndFill), u32 anti_alias
 *             (rendering_mode().use_shape_antialiasing()), u32 has_bbox, RECT abs_stroke_bounding_box,
 *             u32 has_fill [PAINT, f32 opacity, u32 rule (0 NonZero, 1 EvenOdd)],
 *             u32 has_stroke [PAINT, f32 opacity, f32 width, f32 miterlimit, u32 linecap, u32 linejoin, u32 n_dash, f32 dash*,
 *             f32 dashoffset], u32 n_verbs, verbs (RB_VERB_*, padded to 4), u32 n_points, f32 x,y per point
 *   PAINT  := u32 0 RGB | u32 1 f32 x1 y1 x2 y2 BASE | u32 2 f32 cx cy r fx fy fr BASE | u32 3 RECT rect, XF transform, GROUP root
 *   BASE   := u32 spread_method, XF transform, u32 n_stops, { f32 offset, RGB, f32 opacity }*
 *   IMAGE  := STR id, u32 visible, u32 quality (RB_QUALITY_* of image.rs:180-187), u32 has_bbox, RECT abs bounding box,
 *             u32 0 TREE (ImageKind::SVG) | u32 1 u32 w, u32 h, w*h*4 bytes premultiplied RGBA8 (decoded by the host, image.rs:62-170)
 *   CLIP   := XF transform, u32 has_clip_path [CLIP], GROUP root
 *   MASK   := RECT rect, u32 kind (0 Luminance, 1 Alpha), u32 has_mask [MASK], GROUP root
 *   FILTER := RECT rect, u32 n_primitives, PRIM*
 *   INPUT  := u32 0 (SourceGraphic) | u32 1 (SourceAlpha) | u32 2 STR name (Reference)
 *   PRIM   := RECT rect, u32 color_interpolation (0 sRGB, 1 linearRGB), STR result, u32 kind, then by kind:
 *      0 Blend: u32 mode, INPUT in1, INPUT in2            1 DropShadow: INPUT, f32 dx dy std_dev_x std_dev_y, RGB, f32 opacity
 *      2 Flood: RGB, f32 opacity                          3 GaussianBlur: INPUT, f32 std_dev_x std_dev_y
 *      4 Offset: INPUT, f32 dx dy                         5 Composite: u32 op (over,in,out,atop,xor,arithmetic), f32 k1..k4, INPUT, INPUT
 *      6 Merge: u32 n, INPUT*                             7 Tile: INPUT                     8 Image: GROUP root
 *      9 ComponentTransfer: INPUT, 4 x { u32 type (rb_transfer_fn), u32 n, f32 values*, f32 slope intercept amplitude exponent offset }
 *      10 ColorMatrix: INPUT, u32 kind, u32 n, f32 params*
 *      11 ConvolveMatrix: INPUT, u32 columns rows target_x target_y, f32 divisor bias, u32 edge_mode, u32 preserve_alpha, u32 n, f32 kernel*
 *      12 Morphology: INPUT, u32 op (0 erode, 1 dilate), f32 rx ry
 *      13 DisplacementMap: INPUT in1, INPUT in2, f32 scale, u32 x_channel y_channel
 *      14 Turbulence: f32 base_freq_x base_freq_y, u32 octaves, i32 seed, u32 stitch, u32 fractal_noise
 *      15 DiffuseLighting / 16 SpecularLighting: INPUT, f32 surface_scale, constant, specular_exponent, RGB lighting_color,
 *         LIGHT := u32 kind, f32 azimuth elevation x y z points_at_x points_at_y points_at_z specular_exponent, u32 has_cone, f32 cone
 * ---------------------------------------------------------------------------------------------- */
typedef struct rb_tree rb_tree;
/* Parses and validates a stream into a host-side tree (no device work).  RB_ERR_INVALID: malformed. */
int rb_tree_parse(const void *stream, size_t len, rb_tree **out);
void rb_tree_destroy(rb_tree *tree);
int rb_tree_size(const rb_tree *tree, float *width, float *height); /* resvg_get_image_size */
/* usvg::Tree::node_by_id(id)?.abs_layer_bounding_box() — the pixmap size render_node expects (c-api/lib.rs:795-814);
 * RB_ERR_INVALID: no such node, or a zero-sized one. */
int rb_tree_node_bbox(const rb_tree *tree, const char *id, float out_xywh[4]);
/* resvg::render(tree, transform, pixmap) — lib.rs:34-43.  Draws over the target's current content. */
int rb_render(rb_ctx *ctx, const rb_tree *tree, const float ts[6], rb_layer *target);
/* resvg::render_node(node, transform, pixmap) — lib.rs:55-70: the node is placed at -abs_layer_bounding_box.  RB_ERR_INVALID
 * is the reference's None (unknown id / zero-sized node). */
int rb_render_node(rb_ctx *ctx, const rb_tree *tree, const char *id, const float ts[6], rb_layer *target);
/* Canvas-strip sharding of ONE document across GPUs: renders rows [y0, y0 + height(target)) of the
 * canvas_w x canvas_h render of `tree` into `target` (whose width must be canvas_w).  The strip holds exactly the pixels of
 * the whole-canvas rb_render: draws that reach the target directly are clipped and flattened against the whole canvas,
 * isolated groups are rendered as in the whole render (those that miss the strip are skipped) and composited shifted. */
int rb_render_strip(rb_ctx *ctx, const rb_tree *tree, const float ts[6], uint32_t canvas_w, uint32_t canvas_h, int32_t y0, rb_layer *target);
/* One-shot form: parse + render + free (SURVEY 8(b) `rb_submit`). */
int rb_submit(rb_ctx *ctx, const void *stream, size_t len, const float ts[6], rb_layer *target);
/* resvg_render (c-api/lib.rs:875-893) over a HOST pixmap: upload (the caller's pixels are the canvas), render, download. */
int rb_render_to_host(rb_ctx *ctx, const rb_tree *tree, const float ts[6], uint32_t width, uint32_t height, uint8_t *pixmap);

/* ------------------------------------------------------------------------------------------------
 * Host geometry: tiny_skia_path::Path::stroke(&Stroke, res_scale) — the outline PixmapMut::stroke_path fills
 * (path.rs:113; usvg Stroke::to_tiny_skia tree/mod.rs:638-664).  Pure host code (the north star keeps stroking on the
 * host).  cap: 0 butt, 1 round, 2 square; join: 0 miter, 1 miter-clip, 2 round, 3 bevel.  res_scale =
 * PathStroker::compute_resolution_scale(transform).  Outputs are malloc'ed; release them with rb_path_free.
 * Returns RB_ERR_INVALID when the stroke is empty (the reference's None).
 * ---------------------------------------------------------------------------------------------- */
int rb_path_stroke(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, float width,
                   float miter_limit, int32_t cap, int32_t join, float res_scale, uint8_t **out_verbs,
                   int32_t *out_n_verbs, float **out_points, int32_t *out_n_points);
void rb_path_free(void *p);
/* The ordered blits {x, y, alpha} (int32 triples, malloc'ed) of tiny-skia's anti-aliased hairline walker
 * (scan/hairline_aa.rs via hairline::stroke_path_impl) for a path already in device space, clipped to clip_w x clip_h;
 * cap: 0 butt, 1 round, 2 square.  Host code; the batch uses the same walker internally. */
int rb_path_hairline(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, int32_t cap,
                     int32_t clip_w, int32_t clip_h, int32_t **out_blits, int32_t *out_n);
/* tiny_skia_path::Path::dash(&StrokeDash::new(dash_array, dash_offset)?, res_scale) — the path stroke_path strokes when
 * the stroke is dashed (tiny-skia painter.rs stroke_path).  Host code.  RB_ERR_INVALID: the dash specification is
 * rejected or nothing is left of the path. */
int rb_path_dash(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points, const float *dash_array,
                 int32_t n_dash, float dash_offset, float res_scale, uint8_t **out_verbs, int32_t *out_n_verbs,
                 float **out_points, int32_t *out_n_points);

/* Test hook: route every batch through the any-winding fallback kernel (k_raster_tiles_wide) instead of the packed
 * one, which the host otherwise selects only when a draw could reach |winding| > 127. */
void rb_debug_force_wide_kernel(int on);

/* Host-only introspection (no device work; used by the CPU test-suite): the line edges {x, dx, first_y,
 * last_y, winding} and blitter bounds geom = {sect.x, sect.y, sect.w, sect.h, shift, start_y, stop_y} the device
 * would receive for one path on a cw x ch tile.  Returns the edge count, 0 if nothing is drawn, < 0 on error. */
int rb_debug_build_edges(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                         int32_t anti_alias, int32_t cw, int32_t ch, const float ts[6], int32_t *out_edges,
                         int32_t *out_meta /* optional: {prev segment index | -1, inserts-before flag} per edge */,
                         int32_t max_edges, int32_t geom[7]);

/* Tuning hook: with RB_PROFILE=1 in the environment the library accumulates host-side phase timers (host build, upload,
 * launches, layer create / destroy, recording, composites); this prints them to stderr (reset_only == 0) and clears them. */
void rb_debug_profile(int reset_only);

/* Test hook: expand curves into line edges on the host (as the fallback path does) instead of on the device. */
void rb_debug_host_expand(int on);

/* Path geometry (transform, dash, stroke, hairline walk, monotone chop, clip, edge set-up: path.rs:73,113 down to tiny-skia's
 * edge builder) runs on the device for large batches on layers and on host threads otherwise.  Test / tuning hook:
 * mode 0 = that default, 1 = on the device for every eligible batch, 2 = always on the host (the RB_GEO_MODE environment
 * variable sets the initial mode).  rb_debug_geo_counts: out[0] = batch ranges built by the geometry kernels so far,
 * out[1] = ranges they handed back to the host builder, out[2] = launches repeated with a larger heap; of the last range
 * built on the device: out[3] = microseconds the host waited for the geometry kernels, out[4] = microseconds of host task
 * building, out[5] = geometry tasks. */
void rb_debug_geo_mode(int mode);
void rb_debug_geo_counts(uint64_t out[6]);
/* Host-only batches (rb_debug_batch_begin_host): runs the host half of the device geometry path and reports out[0..7] =
 * tasks, dashed, stroked, hairline, fill-list entries, bytes that would be uploaded, verbs, points. */
int rb_debug_geo_host_stats(rb_batch *batch, uint64_t out[8]);
/* The dash-by-dash decomposition the geometry kernels use for a dashed stroke, run on the host with the same code: the
 * outline of the dashed path built one dash at a time (outputs malloc'ed, rb_path_free).  Tests compare it with
 * rb_path_dash followed by rb_path_stroke. */
int rb_debug_stroke_dashed_in_units(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                                    const float *dash_array, int32_t n_dash, float dash_offset, float width, float miter_limit,
                                    int32_t cap, int32_t join, float res_scale, uint8_t **out_verbs, int32_t *out_n_verbs,
                                    float **out_points, int32_t *out_n_points);

/* Host-only batch (no target, no device work): records like any batch; rb_batch_prepare runs the host build (edges,
 * binning, block layout) for a width x height canvas and keeps the block on the host.  For the CPU test-suite and for
 * profiling the host half.  rb_debug_batch_phases: microseconds of the last host build — [0] edge build, [1] layout +
 * tile counting, [2] pack, [3] tile lists, [4] staging allocation/wait, [5] total.  rb_debug_batch_block copies one
 * array of the block out (which: 0 draws (48 B), 1 tile offsets, 2 tile draw lists, 3 tile ids, 4 edges (16 B)) and
 * returns its element count. */
int rb_debug_batch_begin_host(uint32_t width, uint32_t height, rb_batch **out);
int rb_debug_batch_phases(rb_batch *batch, uint64_t phases[6]);
int64_t rb_debug_batch_block(rb_batch *batch, int32_t which, void *out, uint64_t max_bytes);

#ifdef __cplusplus
}
#endif
#endif /* RESVG_B200_H */
