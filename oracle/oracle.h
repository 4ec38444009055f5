/*
 * oracle.h — CPU restatement of the resvg pixel hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This library is the parity checker for resvg_b200's CUDA kernels and the timed
 * "cpu_baseline" in bench.py.  Nothing under resvg_b200/ links, loads or calls it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * Filters (filters.c) restate crates/resvg/src/filter/*.rs line by line; that source is
 * in the reference tree, so those functions are exact by construction and are pinned
 * against the reference's golden PNGs through tests/golden/.
 * The rasteriser (raster.c) restates tiny-skia 0.12.0 (Cargo.lock:654-655), whose source
 * is NOT in /root/reference; it is pinned end-to-end against the reference's golden PNGs
 * at the reference's own +-1 criterion (crates/resvg/tests/integration/main.rs:151-226).
 *
 * All buffers are tightly packed RGBA8888, row-major, `width*y + x` indexing
 * (crates/resvg/src/filter/mod.rs:30-84).
 */
#ifndef RESVG_B200_ORACLE_H
#define RESVG_B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- filter/mod.rs helpers ---- */
void orc_multiply_alpha(uint8_t *rgba, size_t npix);   /* mod.rs:129-136 */
void orc_demultiply_alpha(uint8_t *rgba, size_t npix); /* mod.rs:139-146 */
void orc_into_linear_rgb(uint8_t *rgba, size_t npix);  /* mod.rs:120-124 (demul, LUT, premul) */
void orc_into_srgb(uint8_t *rgba, size_t npix);        /* mod.rs:114-118 */
const uint8_t *orc_srgb_to_linear_table(void);         /* mod.rs:162-179 */
const uint8_t *orc_linear_to_srgb_table(void);         /* mod.rs:195-212 */

/* ---- box_blur.rs ---- */
void orc_box_blur(double sigma_x, double sigma_y, uint8_t *rgba, uint32_t w, uint32_t h);
void orc_create_box_gauss(float sigma, int32_t sizes[5]); /* box_blur.rs:37-71 */

/* ---- iir_blur.rs ---- */
void orc_iir_blur(double sigma_x, double sigma_y, uint8_t *rgba, uint32_t w, uint32_t h);

/* ---- morphology.rs ---- op: 0 = erode, 1 = dilate */
void orc_morphology(int op, float rx, float ry, uint8_t *rgba, uint32_t w, uint32_t h);

/* ---- convolve_matrix.rs ---- edge_mode: 0 none, 1 duplicate, 2 wrap */
void orc_convolve_matrix(const float *kernel, uint32_t columns, uint32_t rows,
                         uint32_t target_x, uint32_t target_y, float divisor, float bias,
                         int edge_mode, int preserve_alpha,
                         uint8_t *rgba, uint32_t w, uint32_t h);

/* ---- color_matrix.rs ---- kind: 0 matrix(20), 1 saturate(1), 2 hueRotate(1, degrees), 3 luminanceToAlpha */
void orc_color_matrix(int kind, const float *params, uint8_t *rgba, size_t npix);

/* ---- component_transfer.rs ---- type: 0 identity, 1 table, 2 discrete, 3 linear, 4 gamma */
typedef struct {
    int32_t type;
    int32_t n_values;        /* table / discrete */
    const float *values;
    float slope, intercept;  /* linear */
    float amplitude, exponent, offset; /* gamma */
} orc_transfer_fn;
void orc_component_transfer(const orc_transfer_fn funcs[4] /* r,g,b,a */, uint8_t *rgba, size_t npix);
uint8_t orc_transfer(const orc_transfer_fn *f, uint8_t c);

/* ---- composite.rs ---- */
void orc_composite_arithmetic(float k1, float k2, float k3, float k4,
                              const uint8_t *src1, const uint8_t *src2, uint8_t *dest, size_t npix);

/* ---- displacement_map.rs ---- channel: 0 R, 1 G, 2 B, 3 A */
void orc_displacement_map(int x_channel, int y_channel, float scale, float sx, float sy,
                          const uint8_t *src, const uint8_t *map, uint8_t *dest,
                          uint32_t w, uint32_t h);

/* ---- lighting.rs ---- */
typedef struct {
    int32_t kind;              /* 0 distant, 1 point, 2 spot */
    float azimuth, elevation;  /* distant (degrees) */
    float x, y, z;             /* point / spot */
    float points_at_x, points_at_y, points_at_z; /* spot */
    float specular_exponent;   /* spot */
    int32_t has_cone;          /* spot */
    float limiting_cone_angle; /* spot (degrees) */
} orc_light_source;

void orc_diffuse_lighting(float surface_scale, float diffuse_constant,
                          uint8_t lr, uint8_t lg, uint8_t lb, const orc_light_source *light,
                          const uint8_t *src, uint8_t *dest, uint32_t w, uint32_t h);
void orc_specular_lighting(float surface_scale, float specular_constant, float specular_exponent,
                           uint8_t lr, uint8_t lg, uint8_t lb, const orc_light_source *light,
                           const uint8_t *src, uint8_t *dest, uint32_t w, uint32_t h);

/* ---- turbulence.rs ---- */
void orc_turbulence(double offset_x, double offset_y, double sx, double sy,
                    double base_frequency_x, double base_frequency_y, uint32_t num_octaves,
                    int32_t seed, int stitch_tiles, int fractal_noise,
                    uint8_t *dest, uint32_t w, uint32_t h);
/* lattice (514 x int32) and gradient (4 x 514 x 2 doubles) as built by turbulence.rs:93-140 */
void orc_turbulence_init(int32_t seed, int32_t *lattice, double *gradient);

#ifdef __cplusplus
}
#endif
#endif
