/*
 * filters.c — CPU oracle for crates/resvg/src/filter/*.rs  (TEST INFRASTRUCTURE ONLY; see oracle.h)
 *
 * Every function restates the Rust arithmetic of the cited file in the same operation order.
 * Compile with -ffp-contract=off (Rust never contracts a*b+c into an FMA) and without
 * -ffast-math.  Rust float->int `as` casts truncate toward zero, saturate, and map NaN to 0
 *  — see f2u8 / f2i32 below.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- Rust cast semantics ---- */
static inline uint8_t f2u8(float v)
{
    if (!(v > 0.0f)) return 0; /* NaN, negatives, -0 */
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
static inline uint8_t d2u8(double v)
{
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}
static inline int32_t d2i32(double v)
{
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int32_t)v;
}
static inline int64_t d2i64(double v)
{
    if (v != v) return 0;
    if (v >= 9223372036854775807.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)v;
}
static inline int32_t f2i32(float v) { return d2i32((double)v); }
static inline uint32_t f2u32(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

/* float-cmp ApproxEqUlps for f32 (usvg/src/tree/geom.rs:8-18 -> strict-num -> float-cmp) */
static inline int approx_eq_ulps_f32(float a, float b, int32_t ulps)
{
    if (a == b) return 1;
    if ((signbit(a) != 0) != (signbit(b) != 0)) return 0;
    int32_t ai, bi;
    memcpy(&ai, &a, 4);
    memcpy(&bi, &b, 4);
    int32_t diff = (int32_t)((uint32_t)ai - (uint32_t)bi);
    return diff >= -ulps && diff <= ulps;
}
static inline int approx_zero_ulps_f32(float a) { return approx_eq_ulps_f32(a, 0.0f, 4); }
static inline int approx_eq_ulps_f64(double a, double b, int64_t ulps)
{
    if (a == b) return 1;
    if ((signbit(a) != 0) != (signbit(b) != 0)) return 0;
    int64_t ai, bi;
    memcpy(&ai, &a, 8);
    memcpy(&bi, &b, 8);
    int64_t diff = (int64_t)((uint64_t)ai - (uint64_t)bi);
    return diff >= -ulps && diff <= ulps;
}

/* filter/mod.rs:242-254 */
static inline float f32_bound(float min, float val, float max)
{
    if (val > max) return max;
    else if (val >= min) return val;
    else return min;
}

/* ------------------------------------------------------------------------------------------
 * filter/mod.rs helpers
 * ---------------------------------------------------------------------------------------- */

/* mod.rs:129-136 */
void orc_multiply_alpha(uint8_t *d, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        uint8_t *p = d + i * 4;
        float a = (float)p[3] / 255.0f;
        p[2] = f2u8((float)p[2] * a + 0.5f);
        p[1] = f2u8((float)p[1] * a + 0.5f);
        p[0] = f2u8((float)p[0] * a + 0.5f);
    }
}

/* mod.rs:139-146 */
void orc_demultiply_alpha(uint8_t *d, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        uint8_t *p = d + i * 4;
        float a = (float)p[3] / 255.0f;
        p[2] = f2u8((float)p[2] / a + 0.5f);
        p[1] = f2u8((float)p[1] / a + 0.5f);
        p[0] = f2u8((float)p[0] / a + 0.5f);
    }
}

/* mod.rs:162-179 */
static const uint8_t SRGB_TO_LINEAR[256] = {
    0,   0,   0,   0,   0,   0,  0,    1,   1,   1,   1,   1,   1,   1,   1,   1,
    1,   1,   2,   2,   2,   2,  2,    2,   2,   2,   3,   3,   3,   3,   3,   3,
    4,   4,   4,   4,   4,   5,  5,    5,   5,   6,   6,   6,   6,   7,   7,   7,
    8,   8,   8,   8,   9,   9,  9,   10,  10,  10,  11,  11,  12,  12,  12,  13,
    13,  13,  14,  14,  15,  15,  16,  16,  17,  17,  17,  18,  18,  19,  19,  20,
    20,  21,  22,  22,  23,  23,  24,  24,  25,  25,  26,  27,  27,  28,  29,  29,
    30,  30,  31,  32,  32,  33,  34,  35,  35,  36,  37,  37,  38,  39,  40,  41,
    41,  42,  43,  44,  45,  45,  46,  47,  48,  49,  50,  51,  51,  52,  53,  54,
    55,  56,  57,  58,  59,  60,  61,  62,  63,  64,  65,  66,  67,  68,  69,  70,
    71,  72,  73,  74,  76,  77,  78,  79,  80,  81,  82,  84,  85,  86,  87,  88,
    90,  91,  92,  93,  95,  96,  97,  99, 100, 101, 103, 104, 105, 107, 108, 109,
    111, 112, 114, 115, 116, 118, 119, 121, 122, 124, 125, 127, 128, 130, 131, 133,
    134, 136, 138, 139, 141, 142, 144, 146, 147, 149, 151, 152, 154, 156, 157, 159,
    161, 163, 164, 166, 168, 170, 171, 173, 175, 177, 179, 181, 183, 184, 186, 188,
    190, 192, 194, 196, 198, 200, 202, 204, 206, 208, 210, 212, 214, 216, 218, 220,
    222, 224, 226, 229, 231, 233, 235, 237, 239, 242, 244, 246, 248, 250, 253, 255,
};

/* mod.rs:195-212 */
static const uint8_t LINEAR_TO_SRGB[256] = {
    0,  13,  22,  28,  34,  38,  42,  46,  50,  53,  56,  59,  61,  64,  66,  69,
    71,  73,  75,  77,  79,  81,  83,  85,  86,  88,  90,  92,  93,  95,  96,  98,
    99, 101, 102, 104, 105, 106, 108, 109, 110, 112, 113, 114, 115, 117, 118, 119,
    120, 121, 122, 124, 125, 126, 127, 128, 129, 130, 131, 132, 133, 134, 135, 136,
    137, 138, 139, 140, 141, 142, 143, 144, 145, 146, 147, 148, 148, 149, 150, 151,
    152, 153, 154, 155, 155, 156, 157, 158, 159, 159, 160, 161, 162, 163, 163, 164,
    165, 166, 167, 167, 168, 169, 170, 170, 171, 172, 173, 173, 174, 175, 175, 176,
    177, 178, 178, 179, 180, 180, 181, 182, 182, 183, 184, 185, 185, 186, 187, 187,
    188, 189, 189, 190, 190, 191, 192, 192, 193, 194, 194, 195, 196, 196, 197, 197,
    198, 199, 199, 200, 200, 201, 202, 202, 203, 203, 204, 205, 205, 206, 206, 207,
    208, 208, 209, 209, 210, 210, 211, 212, 212, 213, 213, 214, 214, 215, 215, 216,
    216, 217, 218, 218, 219, 219, 220, 220, 221, 221, 222, 222, 223, 223, 224, 224,
    225, 226, 226, 227, 227, 228, 228, 229, 229, 230, 230, 231, 231, 232, 232, 233,
    233, 234, 234, 235, 235, 236, 236, 237, 237, 238, 238, 238, 239, 239, 240, 240,
    241, 241, 242, 242, 243, 243, 244, 244, 245, 245, 246, 246, 246, 247, 247, 248,
    248, 249, 249, 250, 250, 251, 251, 251, 252, 252, 253, 253, 254, 254, 255, 255,
};

const uint8_t *orc_srgb_to_linear_table(void) { return SRGB_TO_LINEAR; }
const uint8_t *orc_linear_to_srgb_table(void) { return LINEAR_TO_SRGB; }

static void apply_lut_rgb(const uint8_t *lut, uint8_t *d, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        d[i * 4 + 0] = lut[d[i * 4 + 0]];
        d[i * 4 + 1] = lut[d[i * 4 + 1]];
        d[i * 4 + 2] = lut[d[i * 4 + 2]];
    }
}

/* mod.rs:120-124 */
void orc_into_linear_rgb(uint8_t *d, size_t n)
{
    orc_demultiply_alpha(d, n);
    apply_lut_rgb(SRGB_TO_LINEAR, d, n);
    orc_multiply_alpha(d, n);
}

/* mod.rs:114-118 */
void orc_into_srgb(uint8_t *d, size_t n)
{
    orc_demultiply_alpha(d, n);
    apply_lut_rgb(LINEAR_TO_SRGB, d, n);
    orc_multiply_alpha(d, n);
}

/* ------------------------------------------------------------------------------------------
 * box_blur.rs
 * ---------------------------------------------------------------------------------------- */

/* box_blur.rs:37-71 */
void orc_create_box_gauss(float sigma, int32_t sizes[5])
{
    if (sigma > 0.0f) {
        float n_float = 5.0f;
        float w_ideal = sqrtf(12.0f * sigma * sigma / n_float) + 1.0f;
        int32_t wl = f2i32(floorf(w_ideal));
        if (wl % 2 == 0) wl -= 1;
        int32_t wu = wl + 2;
        float wl_float = (float)wl;
        float m_ideal = (12.0f * sigma * sigma - n_float * wl_float * wl_float
                         - 4.0f * n_float * wl_float - 3.0f * n_float)
                        / (-4.0f * wl_float - 4.0f);
        /* f32::round = half away from zero; `as usize` saturates at 0 */
        float mr = roundf(m_ideal);
        size_t m = (!(mr > 0.0f)) ? 0 : (mr >= 1.8e19f ? (size_t)-1 : (size_t)mr);
        for (size_t i = 0; i < 5; i++) sizes[i] = (i < m) ? wl : wu;
    } else {
        for (int i = 0; i < 5; i++) sizes[i] = 1;
    }
}

/* box_blur.rs:327-331 */
static inline float magic_round(float x)
{
    volatile float t = x + 12582912.0f;
    return t - 12582912.0f;
}

/* One axis pass over `count` lines; element j of line l is at base(l) + j*stride (in pixels).
 * Restates box_blur_vert (box_blur.rs:85-199) / box_blur_horz (:202-320): a sliding window
 * whose out-of-range samples are RGBA8::default() (fv = lv = 0). */
static void box_pass(size_t radius, const uint8_t *src, uint8_t *dst, size_t lines, size_t len,
                     size_t line_step, size_t stride)
{
    if (radius == 0) {
        /* copy_from_slice */
        if (src != dst) memcpy(dst, src, lines * len * 4);
        return;
    }
    float iarr = 1.0f / (float)(radius + radius + 1);
    for (size_t l = 0; l < lines; l++) {
        size_t start = l * line_step;
        size_t ti = start, li = start, ri = start + radius * stride;
        size_t end = start + stride * (len - 1); /* inclusive */
        int64_t val[4] = {0, 0, 0, 0};
        size_t lim = radius < len ? radius : len;
        for (size_t j = 0; j < lim; j++)
            for (int c = 0; c < 4; c++) val[c] += src[(ti + j * stride) * 4 + c];
        /* blur_radius > len: val += (radius - len) * lv, lv = 0 */
        size_t n1 = len < radius + 1 ? len : radius + 1;
        for (size_t k = 0; k < n1; k++) {
            for (int c = 0; c < 4; c++) {
                int64_t bb = (ri > end) ? 0 : src[ri * 4 + c];
                val[c] += bb; /* sub(bb, fv) */
                dst[ti * 4 + c] = f2u8(magic_round((float)val[c] * iarr));
            }
            ri += stride;
            ti += stride;
        }
        if (len <= radius) continue;
        for (size_t k = radius + 1; k < len - radius; k++) {
            for (int c = 0; c < 4; c++) {
                val[c] += (int64_t)src[ri * 4 + c] - (int64_t)src[li * 4 + c];
                dst[ti * 4 + c] = f2u8(magic_round((float)val[c] * iarr));
            }
            ri += stride;
            li += stride;
            ti += stride;
        }
        size_t n3 = (len - radius - 1) < radius ? (len - radius - 1) : radius;
        for (size_t k = 0; k < n3; k++) {
            for (int c = 0; c < 4; c++) {
                val[c] += 0 - (int64_t)src[li * 4 + c]; /* sub(lv, bb); li >= start always */
                dst[ti * 4 + c] = f2u8(magic_round((float)val[c] * iarr));
            }
            li += stride;
            ti += stride;
        }
    }
}

/* box_blur.rs:23-34, 74-82 */
void orc_box_blur(double sigma_x, double sigma_y, uint8_t *rgba, uint32_t w, uint32_t h)
{
    int32_t bh[5], bv[5];
    orc_create_box_gauss((float)sigma_x, bh);
    orc_create_box_gauss((float)sigma_y, bv);
    size_t n = (size_t)w * h;
    if (n == 0) return;
    uint8_t *back = (uint8_t *)malloc(n * 4);
    memcpy(back, rgba, n * 4);
    for (int i = 0; i < 5; i++) {
        size_t rh = (size_t)((bh[i] - 1) / 2);
        size_t rv = (size_t)((bv[i] - 1) / 2);
        /* box_blur_vert(rv, frontbuf=src -> backbuf) */
        box_pass(rv, rgba, back, w, h, 1, w);
        /* box_blur_horz(rh, backbuf -> frontbuf=src) */
        box_pass(rh, back, rgba, h, w, w, 1);
    }
    free(back);
}

/* ------------------------------------------------------------------------------------------
 * iir_blur.rs
 * ---------------------------------------------------------------------------------------- */

/* iir_blur.rs:142-146 */
static void gen_coefficients(double sigma, size_t steps, double *lambda, double *dnu)
{
    *lambda = (sigma * sigma) / (2.0 * (double)steps);
    *dnu = (1.0 + 2.0 * *lambda - sqrt(1.0 + 4.0 * *lambda)) / (2.0 * *lambda);
}

/* f64::powi(n): LLVM powi expands to repeated multiplication (compiler-rt __powidf2). */
static double powi_f64(double a, int b)
{
    int recip = b < 0;
    double r = 1.0;
    while (1) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

/* iir_blur.rs:79-140 */
static void gaussianiir2d(size_t width, size_t height, double sigma_x, double sigma_y, size_t steps,
                          double *buf)
{
    size_t len = width * height;
    double lambda_x = 1.0, dnu_x = 1.0, lambda_y = 1.0, dnu_y = 1.0;
    if (sigma_x > 0.0) {
        gen_coefficients(sigma_x, steps, &lambda_x, &dnu_x);
        double dnu = dnu_x;
        for (size_t y = 0; y < height; y++) {
            for (size_t s = 0; s < steps; s++) {
                size_t idx = width * y;
                for (size_t x = 1; x < width; x++) buf[idx + x] += dnu * buf[idx + x - 1];
                size_t x = width - 1;
                while (x > 0) {
                    buf[idx + x - 1] += dnu * buf[idx + x];
                    x -= 1;
                }
            }
        }
    }
    if (sigma_y > 0.0) {
        gen_coefficients(sigma_y, steps, &lambda_y, &dnu_y);
        double dnu = dnu_y;
        for (size_t x = 0; x < width; x++) {
            for (size_t s = 0; s < steps; s++) {
                size_t idx = x;
                size_t y = width;
                while (y < len) {
                    buf[idx + y] += dnu * buf[idx + y - width];
                    y += width;
                }
                y = len - width;
                while (y > 0) {
                    buf[idx + y - width] += dnu * buf[idx + y];
                    y -= width;
                }
            }
        }
    }
    double post_scale = powi_f64(sqrt(dnu_x * dnu_y) / sqrt(lambda_x * lambda_y), 2 * (int)steps);
    for (size_t i = 0; i < len; i++) buf[i] *= post_scale;
}

/* iir_blur.rs:47-77 */
void orc_iir_blur(double sigma_x, double sigma_y, uint8_t *rgba, uint32_t w, uint32_t h)
{
    size_t n = (size_t)w * h;
    if (n == 0) return;
    double *buf = (double *)malloc(n * sizeof(double));
    for (int ch = 0; ch < 4; ch++) {
        for (size_t i = 0; i < n; i++) buf[i] = (double)rgba[i * 4 + ch] / 255.0;
        gaussianiir2d(w, h, sigma_x, sigma_y, 4, buf);
        for (size_t i = 0; i < n; i++) rgba[i * 4 + ch] = d2u8(buf[i] * 255.0);
    }
    free(buf);
}

/* ------------------------------------------------------------------------------------------
 * morphology.rs:15-73
 * ---------------------------------------------------------------------------------------- */
void orc_morphology(int op, float rx, float ry, uint8_t *rgba, uint32_t w, uint32_t h)
{
    uint32_t cx = f2u32(ceilf(rx)) * 2u, cy = f2u32(ceilf(ry)) * 2u;
    uint32_t columns = cx < w ? cx : w;
    uint32_t rows = cy < h ? cy : h;
    uint32_t target_x = f2u32(floorf((float)columns / 2.0f));
    uint32_t target_y = f2u32(floorf((float)rows / 2.0f));
    int32_t width_max = (int32_t)w - 1, height_max = (int32_t)h - 1;
    size_t n = (size_t)w * h;
    if (n == 0) return;
    uint8_t *buf = (uint8_t *)calloc(n, 4);
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            uint8_t np[4];
            memset(np, op == 0 ? 255 : 0, 4);
            for (uint32_t oy = 0; oy < rows; oy++) {
                for (uint32_t ox = 0; ox < columns; ox++) {
                    int32_t tx = (int32_t)x - (int32_t)target_x + (int32_t)ox;
                    int32_t ty = (int32_t)y - (int32_t)target_y + (int32_t)oy;
                    if (tx < 0 || tx > width_max || ty < 0 || ty > height_max) continue;
                    const uint8_t *p = rgba + ((size_t)w * ty + tx) * 4;
                    for (int c = 0; c < 4; c++) {
                        if (op == 0) np[c] = p[c] < np[c] ? p[c] : np[c];
                        else np[c] = p[c] > np[c] ? p[c] : np[c];
                    }
                }
            }
            memcpy(buf + ((size_t)w * y + x) * 4, np, 4);
        }
    }
    memcpy(rgba, buf, n * 4);
    free(buf);
}

/* ------------------------------------------------------------------------------------------
 * convolve_matrix.rs:15-111
 * `kernel` is usvg's ConvolveMatrixData row-major: get(x, y) = data[y*columns + x].
 * ---------------------------------------------------------------------------------------- */
void orc_convolve_matrix(const float *kernel, uint32_t columns, uint32_t rows,
                         uint32_t target_x, uint32_t target_y, float divisor, float bias,
                         int edge_mode, int preserve_alpha,
                         uint8_t *rgba, uint32_t w, uint32_t h)
{
    int32_t width_max = (int32_t)w - 1, height_max = (int32_t)h - 1;
    size_t n = (size_t)w * h;
    if (n == 0) return;
    uint8_t *buf = (uint8_t *)calloc(n, 4);
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            const uint8_t *in_p = rgba + ((size_t)w * y + x) * 4;
            float new_r = 0.0f, new_g = 0.0f, new_b = 0.0f, new_a = 0.0f;
            for (uint32_t oy = 0; oy < rows; oy++) {
                for (uint32_t ox = 0; ox < columns; ox++) {
                    int32_t tx = (int32_t)x - (int32_t)target_x + (int32_t)ox;
                    int32_t ty = (int32_t)y - (int32_t)target_y + (int32_t)oy;
                    if (edge_mode == 0) {
                        if (tx < 0 || tx > width_max || ty < 0 || ty > height_max) continue;
                    } else if (edge_mode == 1) {
                        tx = tx < 0 ? 0 : (tx > width_max ? width_max : tx);
                        ty = ty < 0 ? 0 : (ty > height_max ? height_max : ty);
                    } else {
                        while (tx < 0) tx += (int32_t)w;
                        tx %= (int32_t)w;
                        while (ty < 0) ty += (int32_t)h;
                        ty %= (int32_t)h;
                    }
                    float k = kernel[(size_t)(rows - oy - 1) * columns + (columns - ox - 1)];
                    const uint8_t *p = rgba + ((size_t)w * ty + tx) * 4;
                    new_r += (float)p[0] / 255.0f * k;
                    new_g += (float)p[1] / 255.0f * k;
                    new_b += (float)p[2] / 255.0f * k;
                    if (!preserve_alpha) new_a += (float)p[3] / 255.0f * k;
                }
            }
            if (preserve_alpha) new_a = (float)in_p[3] / 255.0f;
            else new_a = new_a / divisor + bias;
            float bounded_new_a = f32_bound(0.0f, new_a, 1.0f);
            float chans[3] = {new_r, new_g, new_b};
            uint8_t *out = buf + ((size_t)w * y + x) * 4;
            for (int c = 0; c < 3; c++) {
                float v = chans[c] / divisor + bias * new_a;
                if (preserve_alpha) v = f32_bound(0.0f, v, 1.0f) * bounded_new_a;
                else v = f32_bound(0.0f, v, bounded_new_a);
                out[c] = f2u8(v * 255.0f + 0.5f);
            }
            out[3] = f2u8(bounded_new_a * 255.0f + 0.5f);
        }
    }
    memcpy(rgba, buf, n * 4);
    free(buf);
}

/* ------------------------------------------------------------------------------------------
 * color_matrix.rs:11-110
 * ---------------------------------------------------------------------------------------- */
static inline uint8_t from_normalized(float c) { return f2u8(f32_bound(0.0f, c, 1.0f) * 255.0f); }

void orc_color_matrix(int kind, const float *params, uint8_t *d, size_t n)
{
    if (kind == 0) {
        const float *m = params;
        for (size_t i = 0; i < n; i++) {
            uint8_t *p = d + i * 4;
            float r = (float)p[0] / 255.0f, g = (float)p[1] / 255.0f, b = (float)p[2] / 255.0f,
                  a = (float)p[3] / 255.0f;
            float nr = r * m[0] + g * m[1] + b * m[2] + a * m[3] + m[4];
            float ng = r * m[5] + g * m[6] + b * m[7] + a * m[8] + m[9];
            float nb = r * m[10] + g * m[11] + b * m[12] + a * m[13] + m[14];
            float na = r * m[15] + g * m[16] + b * m[17] + a * m[18] + m[19];
            p[0] = from_normalized(nr);
            p[1] = from_normalized(ng);
            p[2] = from_normalized(nb);
            p[3] = from_normalized(na);
        }
    } else if (kind == 1 || kind == 2) {
        float m[9];
        if (kind == 1) {
            float v = params[0];
            v = v > 0.0f ? v : 0.0f; /* f32::max(0.0); PositiveF32 is never NaN */
            m[0] = 0.213f + 0.787f * v; m[1] = 0.715f - 0.715f * v; m[2] = 0.072f - 0.072f * v;
            m[3] = 0.213f - 0.213f * v; m[4] = 0.715f + 0.285f * v; m[5] = 0.072f - 0.072f * v;
            m[6] = 0.213f - 0.213f * v; m[7] = 0.715f - 0.715f * v; m[8] = 0.072f + 0.928f * v;
        } else {
            /* f32::to_radians: x * (PI / 180) with the constant rounded to f32 */
            float angle = params[0] * 0.017453292519943295769236907684886f;
            float a1 = cosf(angle), a2 = sinf(angle);
            m[0] = 0.213f + 0.787f * a1 - 0.213f * a2;
            m[1] = 0.715f - 0.715f * a1 - 0.715f * a2;
            m[2] = 0.072f - 0.072f * a1 + 0.928f * a2;
            m[3] = 0.213f - 0.213f * a1 + 0.143f * a2;
            m[4] = 0.715f + 0.285f * a1 + 0.140f * a2;
            m[5] = 0.072f - 0.072f * a1 - 0.283f * a2;
            m[6] = 0.213f - 0.213f * a1 - 0.787f * a2;
            m[7] = 0.715f - 0.715f * a1 + 0.715f * a2;
            m[8] = 0.072f + 0.928f * a1 + 0.072f * a2;
        }
        for (size_t i = 0; i < n; i++) {
            uint8_t *p = d + i * 4;
            float r = (float)p[0] / 255.0f, g = (float)p[1] / 255.0f, b = (float)p[2] / 255.0f;
            float nr = r * m[0] + g * m[1] + b * m[2];
            float ng = r * m[3] + g * m[4] + b * m[5];
            float nb = r * m[6] + g * m[7] + b * m[8];
            p[0] = from_normalized(nr);
            p[1] = from_normalized(ng);
            p[2] = from_normalized(nb);
        }
    } else {
        for (size_t i = 0; i < n; i++) {
            uint8_t *p = d + i * 4;
            float r = (float)p[0] / 255.0f, g = (float)p[1] / 255.0f, b = (float)p[2] / 255.0f;
            float na = r * 0.2125f + g * 0.7154f + b * 0.0721f;
            p[0] = 0; p[1] = 0; p[2] = 0;
            p[3] = from_normalized(na);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * component_transfer.rs:10-72
 * ---------------------------------------------------------------------------------------- */
static int is_dummy(const orc_transfer_fn *f)
{
    switch (f->type) {
    case 0: return 1;
    case 1: case 2: return f->n_values == 0;
    default: return 0;
    }
}

static inline size_t f2usize(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 1.8e19f) return (size_t)-1;
    return (size_t)v;
}

uint8_t orc_transfer(const orc_transfer_fn *f, uint8_t cu)
{
    float c = (float)cu / 255.0f;
    switch (f->type) {
    case 0: break;
    case 1: {
        size_t n = (size_t)f->n_values - 1;
        size_t k = f2usize(floorf(c * (float)n));
        if (k > n) k = n;
        if (k == n) c = f->values[k];
        else {
            float vk = f->values[k], vk1 = f->values[k + 1];
            float kf = (float)k, nf = (float)n;
            c = vk + (c - kf / nf) * nf * (vk1 - vk);
        }
        break;
    }
    case 2: {
        size_t n = (size_t)f->n_values;
        size_t k = f2usize(floorf(c * (float)n));
        c = f->values[k < n - 1 ? k : n - 1];
        break;
    }
    case 3: c = f->slope * c + f->intercept; break;
    case 4: c = f->amplitude * powf(c, f->exponent) + f->offset; break;
    }
    return f2u8(f32_bound(0.0f, c, 1.0f) * 255.0f);
}

void orc_component_transfer(const orc_transfer_fn funcs[4], uint8_t *d, size_t n)
{
    /* Channels are independent, so the reference's R,B,G,A visiting order is immaterial. */
    for (int ch = 0; ch < 4; ch++) {
        if (is_dummy(&funcs[ch])) continue;
        uint8_t lut[256];
        for (int v = 0; v < 256; v++) lut[v] = orc_transfer(&funcs[ch], (uint8_t)v);
        for (size_t i = 0; i < n; i++) d[i * 4 + ch] = lut[d[i * 4 + ch]];
    }
}

/* ------------------------------------------------------------------------------------------
 * composite.rs:14-50
 * ---------------------------------------------------------------------------------------- */
static inline float arith_calc(float k1, float k2, float k3, float k4, uint8_t c1, uint8_t c2, float max)
{
    float i1 = (float)c1 / 255.0f;
    float i2 = (float)c2 / 255.0f;
    float result = k1 * i1 * i2 + k2 * i1 + k3 * i2 + k4;
    return f32_bound(0.0f, result, max);
}

void orc_composite_arithmetic(float k1, float k2, float k3, float k4,
                              const uint8_t *s1, const uint8_t *s2, uint8_t *dest, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        const uint8_t *c1 = s1 + i * 4, *c2 = s2 + i * 4;
        float a = arith_calc(k1, k2, k3, k4, c1[3], c2[3], 1.0f);
        if (approx_zero_ulps_f32(a)) continue;
        uint8_t r = f2u8(arith_calc(k1, k2, k3, k4, c1[0], c2[0], a) * 255.0f);
        uint8_t g = f2u8(arith_calc(k1, k2, k3, k4, c1[1], c2[1], a) * 255.0f);
        uint8_t b = f2u8(arith_calc(k1, k2, k3, k4, c1[2], c2[2], a) * 255.0f);
        dest[i * 4 + 0] = r;
        dest[i * 4 + 1] = g;
        dest[i * 4 + 2] = b;
        dest[i * 4 + 3] = f2u8(a * 255.0f);
    }
}

/* ------------------------------------------------------------------------------------------
 * displacement_map.rs:15-62
 * ---------------------------------------------------------------------------------------- */
void orc_displacement_map(int xch, int ych, float scale, float sx, float sy,
                          const uint8_t *src, const uint8_t *map, uint8_t *dest,
                          uint32_t w, uint32_t h)
{
    int32_t wi = (int32_t)w, hi = (int32_t)h;
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            const uint8_t *p = map + ((size_t)w * y + x) * 4;
            float dx = (float)p[xch] / 255.0f - 0.5f;
            float dy = (float)p[ych] / 255.0f - 0.5f;
            int32_t ox = f2i32(roundf((float)x + dx * sx * scale));
            int32_t oy = f2i32(roundf((float)y + dy * sy * scale));
            if (ox >= 0 && ox < wi && oy >= 0 && oy < hi) {
                size_t idx = (size_t)oy * w + (size_t)ox;
                memcpy(dest + ((size_t)w * y + x) * 4, src + idx * 4, 4);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * lighting.rs
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } vec3;

static inline float v3dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float v3len(vec3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* lighting.rs:69-81 + unwrap_or(v) */
static inline vec3 v3norm_or_self(vec3 a)
{
    float l = v3len(a);
    if (!approx_zero_ulps_f32(l)) {
        vec3 r = {a.x / l, a.y / l, a.z / l};
        return r;
    }
    return a;
}

typedef struct { float fx, fy, nx, ny; } normal_t;

#define TO_RAD 0.017453292519943295769236907684886f

typedef struct {
    int specular;
    float surface_scale, constant, exponent;
    uint8_t lr, lg, lb;
    const orc_light_source *light;
    const uint8_t *src;
    uint8_t *dest;
    uint32_t w, h;
    vec3 light_vector;
} light_ctx;

static inline int16_t alpha_at(const light_ctx *c, uint32_t x, uint32_t y)
{
    return (int16_t)c->src[((size_t)c->w * y + x) * 4 + 3];
}

/* lighting.rs:340-476 — nine positional variants.  bx/by: 0 = first, 1 = interior, 2 = last. */
static normal_t compute_normal(const light_ctx *c, uint32_t x, uint32_t y)
{
    uint32_t w = c->w, h = c->h;
    int bx = (x == 0) ? 0 : (x == w - 1 ? 2 : 1);
    int by = (y == 0) ? 0 : (y == h - 1 ? 2 : 1);
    const float F12 = 1.0f / 2.0f, F13 = 1.0f / 3.0f, F14 = 1.0f / 4.0f, F23 = 2.0f / 3.0f;
    normal_t n;
    int nx, ny; /* i16 arithmetic cannot overflow: |sum| <= 4*255*... < 32767 */
#define A(dx, dy) ((int)alpha_at(c, (uint32_t)((int)x + (dx)), (uint32_t)((int)y + (dy))))
    if (bx == 0 && by == 0) { /* top_left :340-353 */
        int center = A(0, 0), right = A(1, 0), bottom = A(0, 1), bottom_right = A(1, 1);
        n.fx = F23; n.fy = F23;
        nx = -2 * center + 2 * right - bottom + bottom_right;
        ny = -2 * center - right + 2 * bottom + bottom_right;
    } else if (bx == 2 && by == 0) { /* top_right :355-368 */
        int left = A(-1, 0), center = A(0, 0), bottom_left = A(-1, 1), bottom = A(0, 1);
        n.fx = F23; n.fy = F23;
        nx = -2 * left + 2 * center - bottom_left + bottom;
        ny = -left - 2 * center + bottom_left + 2 * bottom;
    } else if (bx == 0 && by == 2) { /* bottom_left :370-383 */
        int top = A(0, -1), top_right = A(1, -1), center = A(0, 0), right = A(1, 0);
        n.fx = F23; n.fy = F23;
        nx = -top + top_right - 2 * center + 2 * right;
        ny = -2 * top - top_right + 2 * center + right;
    } else if (bx == 2 && by == 2) { /* bottom_right :385-398 */
        int top_left = A(-1, -1), top = A(0, -1), left = A(-1, 0), center = A(0, 0);
        n.fx = F23; n.fy = F23;
        nx = -top_left + top - 2 * left + 2 * center;
        ny = -top_left - 2 * top + left + 2 * center;
    } else if (by == 0) { /* top_row :400-415 */
        int left = A(-1, 0), center = A(0, 0), right = A(1, 0);
        int bottom_left = A(-1, 1), bottom = A(0, 1), bottom_right = A(1, 1);
        n.fx = F13; n.fy = F12;
        nx = -2 * left + 2 * right - bottom_left + bottom_right;
        ny = -left - 2 * center - right + bottom_left + 2 * bottom + bottom_right;
    } else if (by == 2) { /* bottom_row :417-432 */
        int top_left = A(-1, -1), top = A(0, -1), top_right = A(1, -1);
        int left = A(-1, 0), center = A(0, 0), right = A(1, 0);
        n.fx = F13; n.fy = F12;
        nx = -top_left + top_right - 2 * left + 2 * right;
        ny = -top_left - 2 * top - top_right + left + 2 * center + right;
    } else if (bx == 0) { /* left_column :434-449 */
        int top = A(0, -1), top_right = A(1, -1), center = A(0, 0), right = A(1, 0);
        int bottom = A(0, 1), bottom_right = A(1, 1);
        n.fx = F12; n.fy = F13;
        nx = -top + top_right - 2 * center + 2 * right - bottom + bottom_right;
        ny = -2 * top - top_right + 2 * bottom + bottom_right;
    } else if (bx == 2) { /* right_column :451-466 */
        int top_left = A(-1, -1), top = A(0, -1), left = A(-1, 0), center = A(0, 0);
        int bottom_left = A(-1, 1), bottom = A(0, 1);
        n.fx = F12; n.fy = F13;
        nx = -top_left + top - 2 * left + 2 * center - bottom_left + bottom;
        ny = -top_left - 2 * top + bottom_left + 2 * bottom;
    } else { /* interior :468-485 */
        int top_left = A(-1, -1), top = A(0, -1), top_right = A(1, -1);
        int left = A(-1, 0), right = A(1, 0);
        int bottom_left = A(-1, 1), bottom = A(0, 1), bottom_right = A(1, 1);
        n.fx = F14; n.fy = F14;
        nx = -top_left + top_right - 2 * left + 2 * right - bottom_left + bottom_right;
        ny = -top_left - 2 * top - top_right + bottom_left + 2 * bottom + bottom_right;
    }
#undef A
    /* Normal::new :123-129: Vector2::new(-nx as f32, -ny as f32) */
    n.nx = (float)(-nx);
    n.ny = (float)(-ny);
    return n;
}

/* lighting.rs:309-338; colour returned as u8 triple */
static void light_color(const light_ctx *c, vec3 lv, uint8_t out[3])
{
    const orc_light_source *L = c->light;
    out[0] = c->lr; out[1] = c->lg; out[2] = c->lb;
    if (L->kind != 2) return;
    vec3 origin = {L->x, L->y, L->z};
    vec3 direction = {L->points_at_x, L->points_at_y, L->points_at_z};
    direction.x -= origin.x; direction.y -= origin.y; direction.z -= origin.z;
    direction = v3norm_or_self(direction);
    float minus_l_dot_s = -v3dot(lv, direction);
    if (minus_l_dot_s <= 0.0f) { out[0] = out[1] = out[2] = 0; return; }
    if (L->has_cone) {
        if (minus_l_dot_s < cosf(L->limiting_cone_angle * TO_RAD)) { out[0] = out[1] = out[2] = 0; return; }
    }
    float factor = powf(minus_l_dot_s, L->specular_exponent);
    out[0] = f2u8(f32_bound(0.0f, (float)c->lr * factor, 255.0f) + 0.5f);
    out[1] = f2u8(f32_bound(0.0f, (float)c->lg * factor, 255.0f) + 0.5f);
    out[2] = f2u8(f32_bound(0.0f, (float)c->lb * factor, 255.0f) + 0.5f);
}

/* lighting.rs:141-153 (diffuse) and :185-219 (specular) */
static float light_factor(const light_ctx *c, normal_t normal, vec3 lv)
{
    int nzero = approx_zero_ulps_f32(normal.nx) && approx_zero_ulps_f32(normal.ny);
    if (!c->specular) {
        float k;
        if (nzero) k = lv.z;
        else {
            float s = c->surface_scale / 255.0f;
            float nx = normal.nx * s, ny = normal.ny * s;
            nx *= normal.fx;
            ny *= normal.fy;
            vec3 n = {nx, ny, 1.0f};
            k = v3dot(n, lv) / v3len(n);
        }
        return c->constant * k;
    } else {
        vec3 hv = {lv.x + 0.0f, lv.y + 0.0f, lv.z + 1.0f};
        float h_length = v3len(hv);
        if (approx_zero_ulps_f32(h_length)) return 0.0f;
        int exp_is_one = approx_eq_ulps_f32(c->exponent, 1.0f, 4);
        float k;
        if (nzero) {
            float n_dot_h = hv.z / h_length;
            k = exp_is_one ? n_dot_h : powf(n_dot_h, c->exponent);
        } else {
            float s = c->surface_scale / 255.0f;
            float nx = normal.nx * s, ny = normal.ny * s;
            nx *= normal.fx;
            ny *= normal.fy;
            vec3 n = {nx, ny, 1.0f};
            float n_dot_h = v3dot(n, hv) / v3len(n) / h_length;
            k = exp_is_one ? n_dot_h : powf(n_dot_h, c->exponent);
        }
        return c->constant * k;
    }
}

/* the `calc` closure, lighting.rs:257-285 */
static void light_calc(light_ctx *c, uint32_t x, uint32_t y)
{
    const orc_light_source *L = c->light;
    normal_t normal = compute_normal(c, x, y);
    if (L->kind != 0) {
        float nz = (float)alpha_at(c, x, y) / 255.0f * c->surface_scale;
        vec3 v = {L->x - (float)x, L->y - (float)y, L->z - nz};
        c->light_vector = v3norm_or_self(v);
    }
    uint8_t lc[3];
    light_color(c, c->light_vector, lc);
    float factor = light_factor(c, normal, c->light_vector);
    uint8_t r = f2u8(f32_bound(0.0f, (float)lc[0] * factor, 255.0f) + 0.5f);
    uint8_t g = f2u8(f32_bound(0.0f, (float)lc[1] * factor, 255.0f) + 0.5f);
    uint8_t b = f2u8(f32_bound(0.0f, (float)lc[2] * factor, 255.0f) + 0.5f);
    uint8_t a;
    if (!c->specular) a = 255;
    else { a = r > g ? r : g; a = a > b ? a : b; }
    uint8_t *o = c->dest + ((size_t)c->w * y + x) * 4;
    o[0] = r; o[1] = g; o[2] = b; o[3] = a;
}

/* lighting.rs:227-307.  Every pixel is computed independently, so raster order is used here
 * instead of the reference's corners/edges/interior order. */
static void lighting_apply(light_ctx *c)
{
    if (c->w < 3 || c->h < 3) return;
    const orc_light_source *L = c->light;
    if (L->kind == 0) {
        float az = L->azimuth * TO_RAD, el = L->elevation * TO_RAD;
        c->light_vector.x = cosf(az) * cosf(el);
        c->light_vector.y = sinf(az) * cosf(el);
        c->light_vector.z = sinf(el);
    } else {
        c->light_vector.x = c->light_vector.y = c->light_vector.z = 1.0f;
    }
    for (uint32_t y = 0; y < c->h; y++)
        for (uint32_t x = 0; x < c->w; x++) light_calc(c, x, y);
}

void orc_diffuse_lighting(float surface_scale, float diffuse_constant,
                          uint8_t lr, uint8_t lg, uint8_t lb, const orc_light_source *light,
                          const uint8_t *src, uint8_t *dest, uint32_t w, uint32_t h)
{
    light_ctx c = {0, surface_scale, diffuse_constant, 1.0f, lr, lg, lb, light, src, dest, w, h, {0, 0, 0}};
    lighting_apply(&c);
}

void orc_specular_lighting(float surface_scale, float specular_constant, float specular_exponent,
                           uint8_t lr, uint8_t lg, uint8_t lb, const orc_light_source *light,
                           const uint8_t *src, uint8_t *dest, uint32_t w, uint32_t h)
{
    light_ctx c = {1, surface_scale, specular_constant, specular_exponent, lr, lg, lb, light, src, dest, w, h, {0, 0, 0}};
    lighting_apply(&c);
}

/* ------------------------------------------------------------------------------------------
 * turbulence.rs
 * ---------------------------------------------------------------------------------------- */
#define RAND_M 2147483647
#define RAND_A 16807
#define RAND_Q 127773
#define RAND_R 2836
#define B_SIZE 0x100
#define B_LEN (B_SIZE + B_SIZE + 2)
#define BM 0xff
#define PERLIN_N 0x1000

/* turbulence.rs:287-294 (wrapping i32 arithmetic) */
static int32_t tb_random(int32_t seed)
{
    int32_t result = (int32_t)((uint32_t)RAND_A * (uint32_t)(seed % RAND_Q)
                               - (uint32_t)RAND_R * (uint32_t)(seed / RAND_Q));
    if (result <= 0) result = (int32_t)((uint32_t)result + (uint32_t)RAND_M);
    return result;
}

/* turbulence.rs:93-140; gradient laid out [k][i][j] = gradient[(k*B_LEN + i)*2 + j] */
void orc_turbulence_init(int32_t seed, int32_t *lattice, double *gradient)
{
    if (seed <= 0) seed = -(seed) % (RAND_M - 1) + 1; /* seed == i32::MIN would overflow in Rust too */
    if (seed > RAND_M - 1) seed = RAND_M - 1;
    memset(gradient, 0, sizeof(double) * 4 * B_LEN * 2);
    memset(lattice, 0, sizeof(int32_t) * B_LEN);
    for (int k = 0; k < 4; k++) {
        for (int i = 0; i < B_SIZE; i++) {
            lattice[i] = i;
            double *g = gradient + ((size_t)k * B_LEN + i) * 2;
            for (int j = 0; j < 2; j++) {
                seed = tb_random(seed);
                g[j] = (double)((seed % (B_SIZE + B_SIZE)) - B_SIZE) / (double)B_SIZE;
            }
            double s = sqrt(g[0] * g[0] + g[1] * g[1]);
            g[0] /= s;
            g[1] /= s;
        }
    }
    for (int i = B_SIZE - 1; i >= 1; i--) {
        int32_t k = lattice[i];
        seed = tb_random(seed);
        int j = seed % B_SIZE;
        lattice[i] = lattice[j];
        lattice[j] = k;
    }
    for (int i = 0; i < B_SIZE + 2; i++) {
        lattice[B_SIZE + i] = lattice[i];
        for (int k = 0; k < 4; k++)
            for (int j = 0; j < 2; j++)
                gradient[((size_t)k * B_LEN + B_SIZE + i) * 2 + j] = gradient[((size_t)k * B_LEN + i) * 2 + j];
    }
}

typedef struct { int has; int32_t width, height, wrap_x, wrap_y; } stitch_info;

static inline double s_curve(double t) { return t * t * (3.0 - 2.0 * t); }
static inline double lerp64(double t, double a, double b) { return a + t * (b - a); }

/* turbulence.rs:224-285 */
static double noise2(int ch, double x, double y, const int32_t *lat, const double *grad, stitch_info st)
{
    double t = x + (double)PERLIN_N;
    int32_t bx0 = d2i32(t);
    int32_t bx1 = (int32_t)((uint32_t)bx0 + 1u);
    double rx0 = t - (double)d2i64(t);
    double rx1 = rx0 - 1.0;
    t = y + (double)PERLIN_N;
    int32_t by0 = d2i32(t);
    int32_t by1 = (int32_t)((uint32_t)by0 + 1u);
    double ry0 = t - (double)d2i64(t);
    double ry1 = ry0 - 1.0;
    if (st.has) {
        if (bx0 >= st.wrap_x) bx0 = (int32_t)((uint32_t)bx0 - (uint32_t)st.width);
        if (bx1 >= st.wrap_x) bx1 = (int32_t)((uint32_t)bx1 - (uint32_t)st.width);
        if (by0 >= st.wrap_y) by0 = (int32_t)((uint32_t)by0 - (uint32_t)st.height);
        if (by1 >= st.wrap_y) by1 = (int32_t)((uint32_t)by1 - (uint32_t)st.height);
    }
    bx0 &= BM; bx1 &= BM; by0 &= BM; by1 &= BM;
    int32_t i = lat[bx0], j = lat[bx1];
    int32_t b00 = lat[i + by0], b10 = lat[j + by0], b01 = lat[i + by1], b11 = lat[j + by1];
    double sx = s_curve(rx0), sy = s_curve(ry0);
    const double *g = grad + (size_t)ch * B_LEN * 2;
    const double *q = g + (size_t)b00 * 2;
    double u = rx0 * q[0] + ry0 * q[1];
    q = g + (size_t)b10 * 2;
    double v = rx1 * q[0] + ry0 * q[1];
    double a = lerp64(sx, u, v);
    q = g + (size_t)b01 * 2;
    u = rx0 * q[0] + ry1 * q[1];
    q = g + (size_t)b11 * 2;
    v = rx1 * q[0] + ry1 * q[1];
    double b = lerp64(sx, u, v);
    return lerp64(sy, a, b);
}

/* turbulence.rs:142-222 */
static double turbulence_at(int ch, double x, double y, double tile_x, double tile_y, double tile_w,
                            double tile_h, double bfx, double bfy, uint32_t octaves, int fractal,
                            int stitching, const int32_t *lat, const double *grad)
{
    stitch_info st = {0, 0, 0, 0, 0};
    if (stitching) {
        if (!approx_eq_ulps_f64(bfx, 0.0, 4)) {
            double lo = floor(tile_w * bfx) / tile_w;
            double hi = ceil(tile_w * bfx) / tile_w;
            if (bfx / lo < hi / bfx) bfx = lo; else bfx = hi;
        }
        if (!approx_eq_ulps_f64(bfy, 0.0, 4)) {
            double lo = floor(tile_h * bfy) / tile_h;
            double hi = ceil(tile_h * bfy) / tile_h;
            if (bfy / lo < hi / bfy) bfy = lo; else bfy = hi;
        }
        st.has = 1;
        st.width = d2i32(tile_w * bfx + 0.5);
        st.height = d2i32(tile_h * bfy + 0.5);
        st.wrap_x = d2i32(tile_x * bfx + (double)PERLIN_N + (double)st.width);
        st.wrap_y = d2i32(tile_y * bfy + (double)PERLIN_N + (double)st.height);
    }
    double sum = 0.0;
    x *= bfx;
    y *= bfy;
    double ratio = 1.0;
    for (uint32_t o = 0; o < octaves; o++) {
        double n = noise2(ch, x, y, lat, grad, st);
        if (fractal) sum += n / ratio;
        else sum += fabs(n) / ratio;
        x *= 2.0;
        y *= 2.0;
        ratio *= 2.0;
        if (st.has) {
            st.width = (int32_t)((uint32_t)st.width * 2u);
            st.wrap_x = (int32_t)(2u * (uint32_t)st.wrap_x - (uint32_t)PERLIN_N);
            st.height = (int32_t)((uint32_t)st.height * 2u);
            st.wrap_y = (int32_t)(2u * (uint32_t)st.wrap_y - (uint32_t)PERLIN_N);
        }
    }
    return sum;
}

/* turbulence.rs:33-91 */
void orc_turbulence(double offset_x, double offset_y, double sx, double sy,
                    double bfx, double bfy, uint32_t num_octaves,
                    int32_t seed, int stitch_tiles, int fractal_noise,
                    uint8_t *dest, uint32_t w, uint32_t h)
{
    int32_t *lat = (int32_t *)malloc(sizeof(int32_t) * B_LEN);
    double *grad = (double *)malloc(sizeof(double) * 4 * B_LEN * 2);
    orc_turbulence_init(seed, lat, grad);
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            uint8_t *p = dest + ((size_t)w * y + x) * 4;
            for (int ch = 0; ch < 4; ch++) {
                double tx = ((double)x + offset_x) / sx, ty = ((double)y + offset_y) / sy;
                double n = turbulence_at(ch, tx, ty, (double)x, (double)y, (double)w, (double)h,
                                         bfx, bfy, num_octaves, fractal_noise, stitch_tiles, lat, grad);
                if (fractal_noise) n = (n * 255.0 + 255.0) / 2.0;
                else n = n * 255.0;
                p[ch] = f2u8(f32_bound(0.0f, (float)n, 255.0f) + 0.5f);
            }
        }
    }
    free(lat);
    free(grad);
}
