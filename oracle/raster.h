/*
 * raster.h — CPU oracle for the tiny-skia 0.12.0 raster path that resvg calls into
 * (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * tiny-skia / tiny-skia-path 0.12.0 are crates.io dependencies (Cargo.lock:654-655, 669-670) whose
 * source is NOT under /root/reference.  This file restates their published algorithm (a Rust port of
 * Skia's SkScan_Path / SkScan_AntiPath / SkEdge / SkEdgeBuilder / SkEdgeClipper / SkAlphaRuns /
 * SkRasterPipeline lowp+highp / gradient and image shaders) and is anchored on the reference's call
 * sites (crates/resvg/src/{path,render,clip,mask}.rs, filter/mod.rs) and golden PNGs.
 *
 * Transform layout everywhere: ts[6] = {sx, ky, kx, sy, tx, ty} (tiny_skia::Transform::from_row order,
 * crates/c-api/lib.rs:67-81):  x' = sx*x + kx*y + tx,  y' = ky*x + sy*y + ty.
 */
#ifndef RESVG_B200_ORACLE_RASTER_H
#define RESVG_B200_ORACLE_RASTER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* path verbs */
enum { ORC_MOVE = 0, ORC_LINE = 1, ORC_QUAD = 2, ORC_CUBIC = 3, ORC_CLOSE = 4 };

/* tiny_skia::BlendMode, declaration order */
enum {
    ORC_BLEND_CLEAR = 0, ORC_BLEND_SOURCE, ORC_BLEND_DESTINATION, ORC_BLEND_SOURCE_OVER,
    ORC_BLEND_DESTINATION_OVER, ORC_BLEND_SOURCE_IN, ORC_BLEND_DESTINATION_IN, ORC_BLEND_SOURCE_OUT,
    ORC_BLEND_DESTINATION_OUT, ORC_BLEND_SOURCE_ATOP, ORC_BLEND_DESTINATION_ATOP, ORC_BLEND_XOR,
    ORC_BLEND_PLUS, ORC_BLEND_MODULATE, ORC_BLEND_SCREEN, ORC_BLEND_OVERLAY, ORC_BLEND_DARKEN,
    ORC_BLEND_LIGHTEN, ORC_BLEND_COLOR_DODGE, ORC_BLEND_COLOR_BURN, ORC_BLEND_HARD_LIGHT,
    ORC_BLEND_SOFT_LIGHT, ORC_BLEND_DIFFERENCE, ORC_BLEND_EXCLUSION, ORC_BLEND_MULTIPLY, ORC_BLEND_HUE,
    ORC_BLEND_SATURATION, ORC_BLEND_COLOR, ORC_BLEND_LUMINOSITY
};

enum { ORC_SHADER_SOLID = 0, ORC_SHADER_LINEAR = 1, ORC_SHADER_RADIAL = 2, ORC_SHADER_PATTERN = 3 };
enum { ORC_SPREAD_PAD = 0, ORC_SPREAD_REFLECT = 1, ORC_SPREAD_REPEAT = 2 };
enum { ORC_QUALITY_NEAREST = 0, ORC_QUALITY_BILINEAR = 1, ORC_QUALITY_BICUBIC = 2 };
enum { ORC_FILL_WINDING = 0, ORC_FILL_EVENODD = 1 };

/* tiny_skia::Paint with its Shader, as resvg builds it (crates/resvg/src/path.rs:45-71). */
typedef struct {
    int32_t shader;
    float color[4];                 /* solid: non-premultiplied r,g,b,a (Color::from_rgba8 = c/255) */
    float x0, y0, r0, x1, y1, r1;   /* linear: start/end points; radial: start circle, end circle */
    int32_t n_stops;
    const float *stops;             /* n_stops x {position, r, g, b, a} non-premultiplied */
    int32_t spread;
    float ts[6];                    /* shader local transform (gradient.transform() / pattern ts) */
    const uint8_t *pattern;         /* pattern: premultiplied RGBA8 */
    uint32_t pattern_w, pattern_h;
    int32_t quality;
    float opacity;                  /* pattern only */
    int32_t blend_mode;
    int32_t anti_alias;
    int32_t force_hq;
} orc_paint;

/* PixmapMut::fill_path(path, paint, rule, transform, None) */
int orc_fill_path(uint8_t *px, uint32_t w, uint32_t h, const uint8_t *verbs, int32_t n_verbs, const float *pts,
                  int32_t n_pts, const orc_paint *paint, int32_t fill_rule, const float ts[6]);
/* n_paths fill_path calls in painter's order over packed arrays (verb_off / pt_off: n_paths + 1 offsets) */
int orc_fill_paths(uint8_t *px, uint32_t w, uint32_t h, int32_t n_paths, const uint32_t *verb_off, const uint32_t *pt_off,
                   const uint8_t *verbs, const float *pts, const orc_paint *paints, const uint8_t *rules, const float ts[6]);
/* PixmapMut::fill_rect(rect, paint, transform, None) */
/* n single-pixel coverage blits {x, y, alpha} blended in order (Blitter::blit_anti_h; used for hairline strokes) */
int orc_blit_coverage(uint8_t *px, uint32_t w, uint32_t h, int32_t n, const int32_t *blits, const orc_paint *paint, const float ts[6]);

int orc_fill_rect(uint8_t *px, uint32_t w, uint32_t h, float x, float y, float rw, float rh, const orc_paint *paint,
                  const float ts[6]);
/* PixmapMut::draw_pixmap(x, y, src, PixmapPaint{opacity, blend_mode, quality}, transform, None) */
int orc_draw_pixmap(uint8_t *dst, uint32_t dw, uint32_t dh, int32_t x, int32_t y, const uint8_t *src, uint32_t sw,
                    uint32_t sh, float opacity, int32_t blend_mode, int32_t quality, const float ts[6]);
/* Pixmap::fill(color): color non-premultiplied floats */
void orc_pixmap_fill(uint8_t *px, uint32_t w, uint32_t h, float r, float g, float b, float a);
/* Mask::from_pixmap(pixmap, type): 0 alpha, 1 luminance */
void orc_mask_from_pixmap(const uint8_t *px, uint32_t w, uint32_t h, int32_t type, uint8_t *mask);
void orc_mask_invert(uint8_t *mask, uint32_t w, uint32_t h);
/* Pixmap::apply_mask(mask) */
void orc_apply_mask(uint8_t *px, uint32_t w, uint32_t h, const uint8_t *mask);
/* Mask::fill_path(path, rule, anti_alias, transform) */
int orc_mask_fill_path(uint8_t *mask, uint32_t w, uint32_t h, const uint8_t *verbs, int32_t n_verbs, const float *pts,
                       int32_t n_pts, int32_t fill_rule, int32_t anti_alias, const float ts[6]);
/* Coverage only (what scan::path_aa / scan::path hand to the blitter): u8 plane, 255 = full. */
int orc_path_coverage(uint8_t *cov, uint32_t w, uint32_t h, const uint8_t *verbs, int32_t n_verbs, const float *pts,
                      int32_t n_pts, int32_t fill_rule, int32_t anti_alias, const float ts[6]);

#ifdef __cplusplus
}
#endif
#endif
