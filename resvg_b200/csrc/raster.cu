// raster.cu — device half of the fill path: the reference's pipeline arithmetic, the any-winding fallback kernel, the
// batch upload / launch glue, layer compositing and masks.  The main tile kernel lives in raster_warp.cuh.
//
// Replaces tiny-skia's per-scanline edge walk + SuperBlitter + RasterPipelineBlitter
// (scan/path.rs, scan/path_aa.rs, alpha_runs.rs, pipeline/{blitter,lowp,highp}.rs), reached from
// crates/resvg/src/path.rs:73.  B200 design:
//   * the layer is cut into tiles (32x8 px per WARP in k_raster_warp, 64x16 px per CTA in the fallback); the owner
//     keeps the tile's destination pixels in REGISTERS for a whole batch of draws, so a tile is read from HBM once
//     and written once no matter how many paths cover it (the reference re-reads/re-writes it for every path);
//   * per draw, edge crossings are scattered into shared memory and the signed winding of every one of the 4x4
//     sub-samples follows by a prefix sum along the row, from which the reference's coverage value follows exactly
//     (coverage rules below), including its 64/64/64/63 full-pixel rule and the abutting-span exception;
//   * shading and blending are straight-line per-pixel code (u16 "lowp" or f32 "highp" arithmetic).
//
// Coverage rules restated from tiny-skia SuperBlitter::blit_h / AlphaRuns::add: on sub-scanline s a
// sub-sample c is inside iff the winding accumulated over all crossings with round(x) <= c is non-zero
// (Winding) / odd (EvenOdd).  A pixel whose 4 sub-samples are all inside receives 64 on sub-rows 0..2 and
// 63 on sub-row 3 when one span covers it, but 4*16 = 64 when two spans abut strictly inside it; partially
// covered pixels receive 16 per sub-sample; the four sub-rows are summed and 256 is folded to 255.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "batch.h"
#include "edge_math.h"
#include "common.cuh"
#include "raster_host.h"
#include "rb_internal.h"

using rbh::DevPaint;
using rbh::DevStop;

constexpr int RT_THREADS = 256;   // 8 warps; warp w owns pixel rows w and w+8, lane l owns columns 2l, 2l+1
constexpr int ROW_POS = TW * 4;   // sub-sample positions per row
constexpr int CNT_ROWS = TH * 4;
constexpr int CNT_BYTES = CNT_ROWS * ROW_POS * 4;

__host__ __device__ __forceinline__ int edge_winding(uint32_t meta) { return (meta & 1u) ? -1 : 1; }

// =================================================================================================
// pipeline arithmetic (tiny-skia pipeline/lowp.rs, highp.rs; Skia SkRasterPipeline_opts.h)
// =================================================================================================
struct P16 { uint32_t r, g, b, a; };
struct PF { float r, g, b, a; };

__device__ __forceinline__ uint32_t div255(uint32_t v) { return (v + 255u) >> 8; }

__device__ __forceinline__ bool blend_pre_scales(int m) { return m == 2 || m == 4 || m == 12 || m == 8 || m == 9 || m == 3 || m == 11; }
__device__ __forceinline__ bool blend_alpha_srcover(int m) { return (m >= 15 && m <= 23) || m >= 25; }

__device__ __forceinline__ uint32_t blend_ch16(int m, uint32_t s, uint32_t d, uint32_t sa, uint32_t da)
{
    switch (m) {
    case 0: return 0;
    case 1: return s;
    case 2: return d;
    case 3: return s + div255(d * (255 - sa));
    case 4: return d + div255(s * (255 - da));
    case 5: return div255(s * da);
    case 6: return div255(d * sa);
    case 7: return div255(s * (255 - da));
    case 8: return div255(d * (255 - sa));
    case 9: return div255(s * da + d * (255 - sa));
    case 10: return div255(d * sa + s * (255 - da));
    case 11: return div255(s * (255 - da) + d * (255 - sa));
    case 12: return min(s + d, 255u);
    case 13: return div255(s * d);
    case 14: return s + d - div255(s * d);
    case 24: return div255(s * (255 - da) + d * (255 - sa) + s * d);
    case 16: return s + d - div255(max(s * da, d * sa));
    case 17: return s + d - div255(min(s * da, d * sa));
    case 22: return s + d - 2 * div255(min(s * da, d * sa));
    case 23: return s + d - 2 * div255(s * d);
    case 20: {
        uint32_t t = (2 * s <= sa) ? 2 * s * d : sa * da - 2 * (sa - s) * (da - d);
        return div255(s * (255 - da) + d * (255 - sa) + t);
    }
    case 15: {
        uint32_t t = (2 * d <= da) ? 2 * s * d : sa * da - 2 * (sa - s) * (da - d);
        return div255(s * (255 - da) + d * (255 - sa) + t);
    }
    default: return s;
    }
}

__device__ __forceinline__ P16 blend16(int m, P16 s, P16 d)
{
    P16 o;
    o.r = blend_ch16(m, s.r, d.r, s.a, d.a) & 0xffffu;
    o.g = blend_ch16(m, s.g, d.g, s.a, d.a) & 0xffffu;
    o.b = blend_ch16(m, s.b, d.b, s.a, d.a) & 0xffffu;
    if (blend_alpha_srcover(m)) o.a = s.a + div255(d.a * (255 - s.a));
    else o.a = blend_ch16(m, s.a, d.a, s.a, d.a) & 0xffffu;
    return o;
}

__device__ __forceinline__ float finv(float v) { return 1.0f - v; }
__device__ __forceinline__ float two(float v) { return v + v; }
__device__ __forceinline__ float mad(float f, float m, float a) { return f * m + a; } // -fmad=false: two roundings
__device__ __forceinline__ float lum(float r, float g, float b) { return r * 0.30f + g * 0.59f + b * 0.11f; }

__device__ __forceinline__ float blend_chf(int m, float s, float d, float sa, float da)
{
    switch (m) {
    case 0: return 0.0f;
    case 1: return s;
    case 2: return d;
    case 3: return mad(d, finv(sa), s);
    case 4: return mad(s, finv(da), d);
    case 5: return s * da;
    case 6: return d * sa;
    case 7: return s * finv(da);
    case 8: return d * finv(sa);
    case 9: return s * da + d * finv(sa);
    case 10: return d * sa + s * finv(da);
    case 11: return s * finv(da) + d * finv(sa);
    case 12: return fminf(s + d, 1.0f);
    case 13: return s * d;
    case 14: return s + d - s * d;
    case 24: return s * finv(da) + d * finv(sa) + s * d;
    case 16: return s + d - fmaxf(s * da, d * sa);
    case 17: return s + d - fminf(s * da, d * sa);
    case 22: return s + d - two(fminf(s * da, d * sa));
    case 23: return s + d - two(s * d);
    case 19:
        if (d == da) return d + s * finv(da);
        if (s == 0.0f) return d * finv(sa);
        return sa * (da - fminf(da, (da - d) * sa * __fdiv_rn(1.0f, s))) + s * finv(da) + d * finv(sa);
    case 18:
        if (d == 0.0f) return s * finv(da);
        if (s == sa) return s + d * finv(sa);
        return sa * fminf(da, (d * sa) * __fdiv_rn(1.0f, sa - s)) + s * finv(da) + d * finv(sa);
    case 20: return s * finv(da) + d * finv(sa) + (two(s) <= sa ? two(s * d) : sa * da - two((da - d) * (sa - s)));
    case 15: return s * finv(da) + d * finv(sa) + (two(d) <= da ? two(s * d) : sa * da - two((da - d) * (sa - s)));
    case 21: {
        float mm = da > 0.0f ? __fdiv_rn(d, da) : 0.0f, s2 = two(s), m4 = two(two(mm));
        float dark_src = d * (sa + (s2 - sa) * (1.0f - mm));
        float dark_dst = (m4 * m4 + m4) * (mm - 1.0f) + 7.0f * mm;
        float lite_dst = __fsqrt_rn(mm) - mm;
        float lite_src = d * sa + da * (s2 - sa) * (two(two(d)) <= da ? dark_dst : lite_dst);
        return s * finv(da) + d * finv(sa) + (s2 <= sa ? dark_src : lite_src);
    }
    default: return s;
    }
}

__device__ __forceinline__ void set_sat(float &r, float &g, float &b, float s)
{
    float mn = fminf(r, fminf(g, b)), mx = fmaxf(r, fmaxf(g, b)), st = mx - mn;
    r = st == 0.0f ? 0.0f : __fdiv_rn((r - mn) * s, st);
    g = st == 0.0f ? 0.0f : __fdiv_rn((g - mn) * s, st);
    b = st == 0.0f ? 0.0f : __fdiv_rn((b - mn) * s, st);
}
__device__ __forceinline__ void set_lum(float &r, float &g, float &b, float l)
{
    float diff = l - lum(r, g, b);
    r += diff; g += diff; b += diff;
}
__device__ __forceinline__ float clip_ch(float c, float mn, float mx, float l, float a)
{
    if (!(mx >= 0.0f)) c = l + __fdiv_rn((c - l) * l, l - mn); // tiny-skia tests mx (pinned by mix-blend-mode goldens)
    if (mx > a) c = l + __fdiv_rn((c - l) * (a - l), mx - l);
    return fmaxf(c, 0.0f);
}

__device__ __forceinline__ PF blendf(int m, PF s, PF d)
{
    PF o;
    if (m >= 25) {
        float R, G, B;
        if (m == 25) {
            R = s.r * s.a; G = s.g * s.a; B = s.b * s.a;
            set_sat(R, G, B, (fmaxf(d.r, fmaxf(d.g, d.b)) - fminf(d.r, fminf(d.g, d.b))) * s.a);
            set_lum(R, G, B, lum(d.r, d.g, d.b) * s.a);
        } else if (m == 26) {
            R = d.r * s.a; G = d.g * s.a; B = d.b * s.a;
            set_sat(R, G, B, (fmaxf(s.r, fmaxf(s.g, s.b)) - fminf(s.r, fminf(s.g, s.b))) * d.a);
            set_lum(R, G, B, lum(d.r, d.g, d.b) * s.a);
        } else if (m == 27) {
            R = s.r * d.a; G = s.g * d.a; B = s.b * d.a;
            set_lum(R, G, B, lum(d.r, d.g, d.b) * s.a);
        } else {
            R = d.r * s.a; G = d.g * s.a; B = d.b * s.a;
            set_lum(R, G, B, lum(s.r, s.g, s.b) * d.a);
        }
        float mn = fminf(R, fminf(G, B)), mx = fmaxf(R, fmaxf(G, B)), l = lum(R, G, B), a = s.a * d.a;
        R = clip_ch(R, mn, mx, l, a);
        G = clip_ch(G, mn, mx, l, a);
        B = clip_ch(B, mn, mx, l, a);
        o.r = s.r * finv(d.a) + d.r * finv(s.a) + R;
        o.g = s.g * finv(d.a) + d.g * finv(s.a) + G;
        o.b = s.b * finv(d.a) + d.b * finv(s.a) + B;
        o.a = s.a + d.a - s.a * d.a;
        return o;
    }
    o.r = blend_chf(m, s.r, d.r, s.a, d.a);
    o.g = blend_chf(m, s.g, d.g, s.a, d.a);
    o.b = blend_chf(m, s.b, d.b, s.a, d.a);
    o.a = blend_alpha_srcover(m) ? mad(d.a, finv(s.a), s.a) : blend_chf(m, s.a, d.a, s.a, d.a);
    return o;
}

// highp store: round-to-nearest-even of clamp(c, 0, 1) * 255
__device__ __forceinline__ uint32_t unnorm(float v) { return (uint32_t)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); }
__device__ __forceinline__ PF load_pf(uint32_t p)
{
    const float k = 1.0f / 255.0f;
    PF c = {(float)RB_R(p) * k, (float)RB_G(p) * k, (float)RB_B(p) * k, (float)RB_A(p) * k};
    return c;
}
__device__ __forceinline__ uint32_t store_pf(PF o) { return rb_pack(unnorm(o.r), unnorm(o.g), unnorm(o.b), unnorm(o.a)); }

// =================================================================================================
// shaders (tiny-skia shaders/*.rs)
// =================================================================================================
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// What the two-point-conical and tiling stages need of a paint, read once (a row of pixels shares it).
struct GradGeom {
    int geom, spread;
    uint32_t fl; // 1 focal_on_circle, 2 well_behaved, 4 smaller, 8 negate_x, 16 !natively_focal, 32 swapped, 64 pad_x1
    float p0, p1, conc_scale, conc_bias;
};
__device__ __forceinline__ GradGeom grad_geom(const DevPaint &P)
{
    GradGeom g;
    g.geom = P.geom; g.spread = P.spread;
    g.fl = (P.focal_on_circle ? 1u : 0u) | (P.well_behaved ? 2u : 0u) | (P.smaller ? 4u : 0u) | (P.negate_x ? 8u : 0u) |
           (P.natively_focal ? 0u : 16u) | (P.swapped ? 32u : 0u) | (P.pad_x1 ? 64u : 0u);
    g.p0 = P.p0; g.p1 = P.p1; g.conc_scale = P.conc_scale; g.conc_bias = P.conc_bias;
    return g;
}

// t of a gradient at the (already transformed) point (x, y): the two-point-conical stage, then the tiling stage
__device__ __forceinline__ float gradient_t_at(const GradGeom &G, float x, float y, bool &masked)
{
    masked = false;
    float t = x;
    switch (G.geom) {
    case 0: break;
    case 1: t = __fsqrt_rn(x * x + y * y); break;
    case 4: t = __fsqrt_rn(x * x + y * y); t = t * G.conc_scale + G.conc_bias; break;
    case 3:
        t = x + __fsqrt_rn(G.p0 - y * y);
        if (t != t) { masked = true; t = 0.0f; }
        break;
    default:
        if (G.fl & 1u) t = x + __fdiv_rn(y * y, x);
        else if (G.fl & 2u) t = __fsqrt_rn(x * x + y * y) - x * G.p0;
        else if (G.fl & 4u) t = -__fsqrt_rn(x * x - y * y) - x * G.p0;
        else t = __fsqrt_rn(x * x - y * y) - x * G.p0;
        if (!(G.fl & 2u) && (t <= 0.0f || t != t)) { masked = true; t = 0.0f; }
        if (G.fl & 8u) t = -t;
        if (G.fl & 16u) t = t + G.p1;
        if (G.fl & 32u) t = 1.0f - t;
        break;
    }
    if (G.spread == 1) {
        float v = (t - 1.0f) - two(floorf((t - 1.0f) * 0.5f)) - 1.0f;
        t = clamp01(fabsf(v));
    } else if (G.spread == 2) {
        t = clamp01(t - floorf(t));
    } else if (G.fl & 64u) {
        t = clamp01(t);
    }
    return t;
}

__device__ __noinline__ float gradient_t(const DevPaint &P, int px, int py, bool &masked)
{
    float x = (float)px + 0.5f, y = (float)py + 0.5f;
    if (P.has_ts) {
        float nx = mad(x, P.ts[0], mad(y, P.ts[2], P.ts[4]));
        float ny = mad(x, P.ts[1], mad(y, P.ts[3], P.ts[5]));
        x = nx; y = ny;
    }
    const GradGeom G = grad_geom(P);
    return gradient_t_at(G, x, y, masked);
}

// colour(t): interval search over the first `len` thresholds (t0s holds +inf beyond the paint's own count; len <= 1: one
// interval), then t * f + b of that interval
__device__ __forceinline__ PF gradient_color_at(const float *__restrict__ t0s, const DevStop *__restrict__ st, int len, float t)
{
    int idx = 0;
    if (len > 1) {
        if (len <= 12) {
            const float4 a = *reinterpret_cast<const float4 *>(t0s);
            idx = (t >= a.y) + (t >= a.z) + (t >= a.w);
            if (len > 4) {
                const float4 b = *reinterpret_cast<const float4 *>(t0s + 4);
                idx += (t >= b.x) + (t >= b.y) + (t >= b.z) + (t >= b.w);
                if (len > 8) {
                    const float4 c = *reinterpret_cast<const float4 *>(t0s + 8);
                    idx += (t >= c.x) + (t >= c.y) + (t >= c.z) + (t >= c.w);
                }
            }
        } else {
#pragma unroll 1
            for (int i = 1; i < len; i++) idx += (t >= st[i].t0) ? 1 : 0;
        }
    }
    const float4 f = *reinterpret_cast<const float4 *>(st[idx].f);
    const float4 b = *reinterpret_cast<const float4 *>(st[idx].b);
    PF c = {mad(t, f.x, b.x), mad(t, f.y, b.y), mad(t, f.z, b.z), mad(t, f.w, b.w)};
    return c;
}
__device__ __forceinline__ PF gradient_color(const DevPaint &P, const DevStop *__restrict__ stops, float t)
{
    return gradient_color_at(P.t0s, stops + P.stop_off, P.two_stop ? 1 : P.len, t);
}

__device__ __forceinline__ float ulp_sub(float v) { return __uint_as_float(__float_as_uint(v) - 1u); }
__device__ __forceinline__ float tile_coord(float v, int mode, float limit, float inv_limit)
{
    if (mode == 2) return v - floorf(v * inv_limit) * limit;
    if (mode == 1) return fabsf((v - limit) - (limit + limit) * floorf((v - limit) * (inv_limit * 0.5f)) - limit);
    return v;
}
__device__ __forceinline__ PF gather(const DevPaint &P, float x, float y)
{
    float w = ulp_sub((float)P.pw), h = ulp_sub((float)P.ph);
    x = fminf(fmaxf(x, 0.0f), w);
    y = fminf(fmaxf(y, 0.0f), h);
    int ix = __float2int_rz(x), iy = __float2int_rz(y);
    return load_pf(__ldg(reinterpret_cast<const uint32_t *>(P.pix) + (size_t)iy * P.pw + ix));
}
__device__ __forceinline__ float bicubic_near(float t) { return mad(t, mad(t, mad(-21.0f / 18.0f, t, 27.0f / 18.0f), 9.0f / 18.0f), 1.0f / 18.0f); }
__device__ __forceinline__ float bicubic_far(float t) { return (t * t) * mad(7.0f / 18.0f, t, -6.0f / 18.0f); }

__device__ __noinline__ PF shade_pattern(const DevPaint &P, int px, int py)
{
    float x = (float)px + 0.5f, y = (float)py + 0.5f;
    if (P.has_ts) {
        float nx = mad(x, P.ts[0], mad(y, P.ts[2], P.ts[4]));
        float ny = mad(x, P.ts[1], mad(y, P.ts[3], P.ts[5]));
        x = nx; y = ny;
    }
    float fw = (float)P.pw, fh = (float)P.ph, iw = __fdiv_rn(1.0f, fw), ih = __fdiv_rn(1.0f, fh);
    PF c;
    if (P.quality == 0) {
        c = gather(P, tile_coord(x, P.spread, fw, iw), tile_coord(y, P.spread, fh, ih));
    } else {
        int n = P.quality == 1 ? 2 : 4;
        float fx = (x + 0.5f) - floorf(x + 0.5f), fy = (y + 0.5f) - floorf(y + 0.5f);
        float wx[4], wy[4];
        if (n == 2) {
            wx[0] = 1.0f - fx; wx[1] = fx; wy[0] = 1.0f - fy; wy[1] = fy;
            wx[2] = wx[3] = wy[2] = wy[3] = 0.0f;
        } else {
            wx[0] = bicubic_far(1.0f - fx); wx[1] = bicubic_near(1.0f - fx); wx[2] = bicubic_near(fx); wx[3] = bicubic_far(fx);
            wy[0] = bicubic_far(1.0f - fy); wy[1] = bicubic_near(1.0f - fy); wy[2] = bicubic_near(fy); wy[3] = bicubic_far(fy);
        }
        float start = -0.5f * (float)(n - 1);
        c.r = c.g = c.b = c.a = 0.0f;
        float yy = y + start;
        for (int j = 0; j < n; j++) {
            float xx = x + start;
            for (int i = 0; i < n; i++) {
                PF s = gather(P, tile_coord(xx, P.spread, fw, iw), tile_coord(yy, P.spread, fh, ih));
                float w = wx[i] * wy[j];
                c.r = mad(w, s.r, c.r); c.g = mad(w, s.g, c.g); c.b = mad(w, s.b, c.b); c.a = mad(w, s.a, c.a);
                xx = xx + 1.0f;
            }
            yy = yy + 1.0f;
        }
        if (n == 4) {
            c.r = fmaxf(c.r, 0.0f); c.g = fmaxf(c.g, 0.0f); c.b = fmaxf(c.b, 0.0f); c.a = fmaxf(c.a, 0.0f);
            c.a = fminf(c.a, 1.0f);
            c.r = fminf(c.r, c.a); c.g = fminf(c.g, c.a); c.b = fminf(c.b, c.a);
        }
    }
    if (P.opacity != 1.0f) { c.r *= P.opacity; c.g *= P.opacity; c.b *= P.opacity; c.a *= P.opacity; }
    return c;
}

__device__ __noinline__ P16 shade16_gradient(const DevPaint &P, const DevStop *__restrict__ stops, int x, int y)
{
    P16 o;
    bool masked;
    float t = gradient_t(P, x, y, masked);
    PF c = gradient_color(P, stops, t);
    o.r = min(__float2uint_rz(clamp01(c.r) * 255.0f + 0.5f), 65535u);
    o.g = min(__float2uint_rz(clamp01(c.g) * 255.0f + 0.5f), 65535u);
    o.b = min(__float2uint_rz(clamp01(c.b) * 255.0f + 0.5f), 65535u);
    o.a = min(__float2uint_rz(clamp01(c.a) * 255.0f + 0.5f), 65535u);
    if (P.premul_after) { o.r = div255(o.r * o.a); o.g = div255(o.g * o.a); o.b = div255(o.b * o.a); }
    return o;
}

__device__ __forceinline__ P16 shade16(const DevPaint &P, const DevStop *__restrict__ stops, int x, int y)
{
    if (P.kind == 0) { P16 o = {P.solid16[0], P.solid16[1], P.solid16[2], P.solid16[3]}; return o; }
    return shade16_gradient(P, stops, x, y);
}

__device__ __noinline__ PF shadef_gradient(const DevPaint &P, const DevStop *__restrict__ stops, int x, int y)
{
    bool masked;
    float t = gradient_t(P, x, y, masked);
    PF c = gradient_color(P, stops, t);
    if (P.premul_after) { c.r *= c.a; c.g *= c.a; c.b *= c.a; }
    if (masked) c.r = c.g = c.b = c.a = 0.0f;
    return c;
}

__device__ __forceinline__ PF shadef(const DevPaint &P, const DevStop *__restrict__ stops, int x, int y)
{
    if (P.kind == 0) { PF c = {P.premul[0], P.premul[1], P.premul[2], P.premul[3]}; return c; }
    if (P.kind == 2) return shade_pattern(P, x, y);
    return shadef_gradient(P, stops, x, y);
}

// k_raster_warp hands the pixels of a lane to blend_tile_gradient (raster_warp.cuh) in this form
struct Px8 { uint32_t v[8]; };

// RasterPipelineBlitter: full-coverage pixels run the blit_rect program, others blit_anti_h.
__device__ __noinline__ uint32_t blend_pixel(const DevPaint &P, const DevStop *__restrict__ stops, uint32_t dst, uint32_t cov,
                                                int x, int y)
{
    if (cov == 255 && P.has_memset) return P.memset_color;
    const int m = P.blend;
    if (P.lowp) {
        P16 d = {RB_R(dst), RB_G(dst), RB_B(dst), RB_A(dst)};
        P16 s = shade16(P, stops, x, y), o;
        if (cov == 255) {
            o = m == 1 ? s : blend16(m, s, d);
        } else if (blend_pre_scales(m)) {
            s.r = div255(s.r * cov); s.g = div255(s.g * cov); s.b = div255(s.b * cov); s.a = div255(s.a * cov);
            o = blend16(m, s, d);
        } else {
            P16 t = blend16(m, s, d);
            o.r = div255(d.r * (255 - cov) + t.r * cov);
            o.g = div255(d.g * (255 - cov) + t.g * cov);
            o.b = div255(d.b * (255 - cov) + t.b * cov);
            o.a = div255(d.a * (255 - cov) + t.a * cov);
        }
        return rb_pack(o.r & 0xffu, o.g & 0xffu, o.b & 0xffu, o.a & 0xffu);
    }
    PF d = load_pf(dst);
    PF s = shadef(P, stops, x, y), o;
    if (cov == 255) {
        o = m == 1 ? s : blendf(m, s, d);
    } else {
        float cf = (float)cov * (1.0f / 255.0f);
        if (blend_pre_scales(m)) {
            s.r *= cf; s.g *= cf; s.b *= cf; s.a *= cf;
            o = blendf(m, s, d);
        } else {
            PF t = blendf(m, s, d);
            o.r = mad(t.r - d.r, cf, d.r); o.g = mad(t.g - d.g, cf, d.g);
            o.b = mad(t.b - d.b, cf, d.b); o.a = mad(t.a - d.a, cf, d.a);
        }
    }
    return store_pf(o);
}

// =================================================================================================
// coverage
// =================================================================================================

// Rare path: crossings of both directions share one sub-sample position strictly inside a fully covered
// pixel on the 4th sub-row, so whether the span breaks there depends on the scanline walker's list order
// (tiny-skia scan/path.rs walk_edges).  One lane replays just that position.  The walker's order is: ascending
// FDot16 x; among equal x, edges that were already active keep the order they had on the previous scanline
// (i.e. ascending previous x — a curve's next segment inherits its curve's slot), and an edge that becomes
// active on this scanline goes after them, or before them when insert_new_edges' forward scan stops at them
// (DevEdge meta bit 2).  Two new edges keep their sorted order.
__device__ __forceinline__ int edge_x_at(const DevEdge &E, int y)
{
    return (int)((uint32_t)E.x + (uint32_t)(y - (int)(E.ypack & 0xffffu)) * (uint32_t)E.dx);
}

// Walker list order of two edges that cross sub-scanline y.  kind: 0 = was active on y-1 ("survivor", keyed by
// its previous x), 1 = new edge placed after equal-x survivors, 2 = new edge placed before them.
__device__ bool walker_less(const DevEdge *__restrict__ edges, const DevEdge &A, uint32_t ia, const DevEdge &B, uint32_t ib, int y)
{
    for (int depth = 0; depth < 2; depth++) {
        const int xa = edge_x_at(A, y), xb = edge_x_at(B, y);
        if (xa != xb) return xa < xb;
        const int fya = (int)(A.ypack & 0xffffu), fyb = (int)(B.ypack & 0xffffu);
        int ka, kb, pa = 0, pb = 0;
        if (y > fya) { ka = 0; pa = (int)((uint32_t)xa - (uint32_t)A.dx); }
        else if (A.meta & 2u) { const DevEdge P = edges[A.meta >> 4]; ka = 0; pa = edge_x_at(P, (int)(P.ypack >> 16)); }
        else ka = (A.meta & 4u) ? 2 : 1;
        if (y > fyb) { kb = 0; pb = (int)((uint32_t)xb - (uint32_t)B.dx); }
        else if (B.meta & 2u) { const DevEdge P = edges[B.meta >> 4]; kb = 0; pb = edge_x_at(P, (int)(P.ypack >> 16)); }
        else kb = (B.meta & 4u) ? 2 : 1;
        if (ka == 0 && kb == 0) {
            if (pa != pb) return pa < pb;
            // coincident lines: their order was fixed on the scanline where the later one joined the list
            if (y > fya && y > fyb) { y = max(fya, fyb); continue; }
            return ia < ib;
        }
        if (ka == 0) return kb == 1;
        if (kb == 0) return ka == 2;
        return ia < ib;
    }
    return ia < ib;
}

__device__ __noinline__ bool exact_span_break(const DevEdge *__restrict__ edges, uint32_t n, int y, int target_r, int w_before)
{
    uint32_t c[12];
    int cnt = 0;
    for (uint32_t e = 0; e < n; e++) {
        const DevEdge E = edges[e];
        int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
        if (fy > y) break; // sorted by first_y
        if (ly < y) continue;
        int x = edge_x_at(E, y);
        int r = (int)((uint32_t)x + 0x8000u) >> 16;
        if (r != target_r) continue;
        if (cnt == 12) return true;
        int j = cnt++;
        while (j > 0 && walker_less(edges, E, e, edges[c[j - 1]], c[j - 1], y)) { c[j] = c[j - 1]; j--; }
        c[j] = e;
    }
    int w = w_before;
    for (int i = 0; i < cnt; i++) {
        w += edge_winding(edges[c[i]].meta);
        if (w == 0) return true;
    }
    return false;
}

// ---- fallback kernel: one 32-bit counter pair per sub-sample (any winding magnitude) ----------------------------
template <bool MASK>
__global__ void __launch_bounds__(RT_THREADS)
k_raster_tiles_wide(void *__restrict__ target, int W, int H, int tiles_x, const uint32_t *__restrict__ tile_ids,
               const uint32_t *__restrict__ tile_off, const uint32_t *__restrict__ tile_draws,
               const DevDraw *__restrict__ draws, const DevEdge *__restrict__ edges, const DevPaint *__restrict__ paints,
               const DevStop *__restrict__ stops, unsigned long long *__restrict__ px_stats)
{
    extern __shared__ int cnt[]; // [CNT_ROWS][ROW_POS]: low 16 bits = downward (+1) crossings, high 16 = upward
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t tile = tile_ids[blockIdx.x];
    const int X0 = (int)(tile % (uint32_t)tiles_x) * TW, Y0 = (int)(tile / (uint32_t)tiles_x) * TH;

    // destination pixels owned by this thread: rows wid, wid+8; columns 2*lane, 2*lane+1
    uint32_t dst[2][2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int gy = Y0 + wid + 8 * i, gx = X0 + 2 * lane;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            uint32_t v = 0;
            if (gy < H && gx + j < W) {
                size_t o = (size_t)gy * W + gx + j;
                v = MASK ? (uint32_t) reinterpret_cast<const uint8_t *>(target)[o] : reinterpret_cast<const uint32_t *>(target)[o];
            }
            dst[i][j] = v;
        }
    }
    for (int i = tid; i < CNT_ROWS * ROW_POS / 4; i += RT_THREADS) reinterpret_cast<int4 *>(cnt)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();

    uint32_t n_partial = 0, n_full = 0; // pixels blended / pixels stored at full coverage (for the roofline model)
    const uint32_t d_begin = tile_off[tile], d_end = tile_off[tile + 1];
    for (uint32_t di = d_begin; di < d_end; di++) {
        const DevDraw D = draws[tile_draws[di]];
        const int tlx = X0 - D.ox, tly = Y0 - D.oy; // tile origin in DrawTiler-tile-local pixels
        const int py0 = max(0, D.sy - tly), py1 = min(TH, D.sy + D.sh - tly);
        const int pxa = max(0, D.sx - tlx), pxb = min(TW, D.sx + D.sw - tlx);
        if (py0 >= py1 || pxa >= pxb) continue; // uniform
        const int sh = D.shift;
        const int lo_pos = pxa << sh, hi_pos = pxb << sh;
        const int sub_top = (tly + py0) << sh, sub_bot = (tly + py1) << sh; // [sub_top, sub_bot) in draw sub-scanlines
        const int row0 = tly << sh, col0 = tlx << sh;
        const DevEdge *E0 = edges + D.edge_off;

        // ---- edge pass: scatter crossings --------------------------------------------------------
        for (uint32_t e = tid; e < D.edge_cnt; e += RT_THREADS) {
            const DevEdge E = E0[e];
            const int fy = (int)(E.ypack & 0xffffu), ly = (int)(E.ypack >> 16);
            if (fy >= sub_bot) break; // edges are sorted by first_y
            const int ys = max(fy, sub_top), ye = min(ly, sub_bot - 1);
            if (ys > ye) continue;
            uint32_t x = (uint32_t)E.x + (uint32_t)(ys - fy) * (uint32_t)E.dx;
            const int inc = (E.meta & 1u) ? 0x10000 : 1;
            int *row = cnt + (ys - row0) * ROW_POS;
            for (int y = ys; y <= ye; y++) {
                int r = (int)(x + 0x8000u) >> 16;
                int pos = max(r - col0, lo_pos);
                if (pos < hi_pos) atomicAdd(row + pos, inc);
                x += (uint32_t)E.dx;
                row += ROW_POS;
            }
        }
        __syncthreads();

        // ---- scan pass: winding prefix sums -> coverage ------------------------------------------
        uint32_t cov[2][2] = {{0, 0}, {0, 0}};
        const bool evenodd = D.rule != 0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int prow = wid + 8 * i;
            if (prow < py0 || prow >= py1) continue; // warp-uniform
            if (sh == 2) {
                uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
                for (int sub = 0; sub < 4; sub++) {
                    int4 *rp = reinterpret_cast<int4 *>(cnt + (prow * 4 + sub) * ROW_POS + lane * 8);
                    int4 a = rp[0], b = rp[1];
                    int c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    int any = a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w;
                    if (__any_sync(0xffffffffu, any != 0)) {
                        int wv[8];
                        int run = 0;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            run += (c[k] & 0xffff) - (int)((uint32_t)c[k] >> 16);
                            wv[k] = run;
                        }
                        int incl = run;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            int v = __shfl_up_sync(0xffffffffu, incl, d);
                            if (lane >= d) incl += v;
                        }
                        const int base = incl - run;
                        uint32_t bits = 0;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            wv[k] += base;
                            bool in = evenodd ? (wv[k] & 1) : (wv[k] != 0);
                            bits |= (in ? 1u : 0u) << k;
                        }
                        uint32_t add0 = 16u * __popc(bits & 0xfu), add1 = 16u * __popc(bits >> 4);
                        if (sub == 3) {
#pragma unroll
                            for (int p = 0; p < 2; p++) {
                                if (((bits >> (4 * p)) & 0xfu) != 0xfu) continue;
                                bool brk = false;
#pragma unroll
                                for (int k = 1; k < 4; k++) {
                                    int ck = c[4 * p + k];
                                    if (ck == 0) continue;
                                    if (evenodd) { brk = true; continue; }
                                    int pc = ck & 0xffff, nc = (int)((uint32_t)ck >> 16);
                                    int before = wv[4 * p + k - 1], after = wv[4 * p + k];
                                    if (pc && nc) {
                                        brk = brk || exact_span_break(E0, D.edge_cnt, row0 + prow * 4 + 3,
                                                                      col0 + lane * 8 + 4 * p + k, before);
                                    } else if ((before ^ after) < 0) brk = true;
                                }
                                if (p == 0) add0 = brk ? 64u : 63u; else add1 = brk ? 64u : 63u;
                            }
                        }
                        acc0 += add0;
                        acc1 += add1;
                        if (any) { rp[0] = make_int4(0, 0, 0, 0); rp[1] = make_int4(0, 0, 0, 0); }
                    }
                }
                cov[i][0] = min(acc0, 255u);
                cov[i][1] = min(acc1, 255u);
            } else {
                int2 *rp = reinterpret_cast<int2 *>(cnt + prow * ROW_POS + lane * 2);
                int2 a = *rp;
                int any = a.x | a.y;
                if (__any_sync(0xffffffffu, any != 0)) {
                    int w0 = (a.x & 0xffff) - (int)((uint32_t)a.x >> 16);
                    int w1 = w0 + (a.y & 0xffff) - (int)((uint32_t)a.y >> 16);
                    int incl = w1;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        int v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    int base = incl - w1;
                    w0 += base; w1 += base;
                    cov[i][0] = (evenodd ? (w0 & 1) : (w0 != 0)) ? 255u : 0u;
                    cov[i][1] = (evenodd ? (w1 & 1) : (w1 != 0)) ? 255u : 0u;
                    if (any) *rp = make_int2(0, 0);
                }
            }
        }

        // spans are clamped to the draw's bounds (SuperBlitter: x + width <= bounds.width)
        if (2 * lane >= pxb) { cov[0][0] = 0; cov[1][0] = 0; }
        if (2 * lane + 1 >= pxb) { cov[0][1] = 0; cov[1][1] = 0; }

        // ---- blend pass ---------------------------------------------------------------------------
        if (cov[0][0] | cov[0][1] | cov[1][0] | cov[1][1]) {
            if (MASK) {
                // RasterPipelineBlitter::new_mask: white, lerp by coverage
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        uint32_t c = cov[i][j];
                        if (c == 255) dst[i][j] = 255;
                        else if (c) dst[i][j] = div255(dst[i][j] * (255 - c) + 255u * c);
                    }
            } else {
                const DevPaint &P = paints[D.paint];
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        uint32_t c = cov[i][j];
                        if (c) {
                            dst[i][j] = blend_pixel(P, stops, dst[i][j], c, tlx + 2 * lane + j, tly + wid + 8 * i);
                            if (c == 255 && P.has_memset) n_full++; else n_partial++;
                        }
                    }
            }
        }
        __syncthreads(); // histogram rows are zero again before the next draw scatters into them
    }

    if (px_stats) {
        n_partial = __reduce_add_sync(0xffffffffu, n_partial);
        n_full = __reduce_add_sync(0xffffffffu, n_full);
        if (lane == 0) {
            atomicAdd(px_stats, (unsigned long long)n_partial);
            atomicAdd(px_stats + 1, (unsigned long long)n_full);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int gy = Y0 + wid + 8 * i, gx = X0 + 2 * lane;
        if (gy >= H) continue;
        size_t o = (size_t)gy * W + gx;
        if (MASK) {
            uint8_t *t = reinterpret_cast<uint8_t *>(target);
            if (gx < W) t[o] = (uint8_t)dst[i][0];
            if (gx + 1 < W) t[o + 1] = (uint8_t)dst[i][1];
        } else {
            uint32_t *t = reinterpret_cast<uint32_t *>(target);
            if (gx + 1 < W && (W & 1) == 0) *reinterpret_cast<uint2 *>(t + o) = make_uint2(dst[i][0], dst[i][1]);
            else {
                if (gx < W) t[o] = dst[i][0];
                if (gx + 1 < W) t[o + 1] = dst[i][1];
            }
        }
    }
}


#include "raster_warp.cuh"
#include "batch_geo.h"
int rb_geo_prepare(rb_batch *b, int32_t n_threads, size_t begin, size_t end);
int rb_geo_begin(rb_batch *b, int32_t n_threads, size_t begin, size_t end);
int rb_geo_finish(rb_batch *b);
void rb_geo_abandon(rb_batch *b);
bool rb_debug_host_only_builder();

// =================================================================================================
// batch: upload + launch (recording, edge building and binning live in batch_host.cpp)
// =================================================================================================
static rb_ctx *batch_ctx(const rb_batch *b) { return b->ctx ? b->ctx : (b->mask ? b->mask->ctx : (b->layer ? b->layer->ctx : nullptr)); }

extern "C" int rb_batch_begin(rb_layer *target, rb_batch **out)
{
    rb_enter(target ? target->ctx : nullptr);
    if (!target || !out) return RB_ERR_INVALID;
    rb_batch *b = new rb_batch();
    b->layer = target;
    b->ctx = target->ctx;
    rb_ctx_retain(target->ctx);
    *out = b;
    return RB_OK;
}

static void batch_release(rb_batch *b)
{
    if (b->dev) {
        cudaFreeAsync(b->dev, batch_ctx(b)->stream);
        b->dev = nullptr;
    }
    if (b->dev_scratch && b->scratch_owned) cudaFreeAsync(b->dev_scratch, batch_ctx(b)->stream);
    b->dev_scratch = nullptr; // otherwise carved out of the same allocation as `dev`
    b->scratch_owned = false;
    if (b->host_block) {
        free(b->host_block);
        b->host_block = nullptr;
    }
    b->lay = BatchLayout();
}

extern "C" void rb_batch_destroy(rb_batch *b)
{
    if (!b) return;
    rb_ctx *ctx = batch_ctx(b);
    if (b->geo) rb_geo_abandon(b);
    batch_release(b);
    delete b;
    if (ctx) rb_ctx_release(ctx);
}

// Device-built structures of the warp-tile path (sizes come from the host build).
struct WarpScratch { size_t o_row_off, o_row_cols, o_row_edges, o_boxes, o_row_cnt, o_row_draws, o_tile_off, o_tile_pairs, o_edges, o_flag, o_scan, total; };
static WarpScratch warp_scratch_layout(const BatchLayout &L)
{
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    WarpScratch w;
    size_t off = 0;
    w.o_row_off = off;    off += al((L.n_row_off + 1) * 4);
    w.o_row_cols = off;   off += al((L.n_row_off + 1) * 4);
    w.o_row_edges = off;  off += al((L.n_list + 1) * sizeof(DevEdge));
    w.o_boxes = off;      off += al(L.n_draws * sizeof(DrawBox));
    w.o_row_cnt = off;    off += al(((size_t)L.wtiles_y + 2) * 4);
    w.o_row_draws = off;  off += al((L.n_row_ent + 1) * sizeof(RowEnt));
    w.o_tile_off = off;   off += al(((size_t)L.wtiles_x * L.wtiles_y + 2) * 4);
    w.o_tile_pairs = off; off += al((L.n_wpairs + 1) * 4);
    w.o_edges = off;      off += al(((L.items ? L.n_slots : 0) + 1) * sizeof(DevEdge));
    w.o_flag = off;       off += 256;
    w.o_scan = off;       off += al((((size_t)L.wtiles_x * L.wtiles_y + SCAN_PER_CTA) / SCAN_PER_CTA + 1) * 4);
    w.total = off;
    return w;
}

struct StageReq { rb_ctx *ctx; int status; };
static void *stage_pinned(void *user, size_t bytes)
{
    StageReq *r = (StageReq *)user;
    void *p = nullptr;
    r->status = rb_staging(r->ctx, bytes, &p);
    return r->status == RB_OK ? p : nullptr;
}
static void *stage_malloc(void *, size_t bytes) { return malloc(bytes); }

// Host build (threads) of draws [begin, end) into the context's pinned staging block, then ONE host-to-device copy.
static int batch_prepare_range(rb_batch *b, int32_t n_threads, size_t begin, size_t end, bool allow_geo = true)
{
    if (!b) return RB_ERR_INVALID;
    batch_release(b);
    void *blk = nullptr;
    if (!b->layer && !b->mask) { // host-only batch (rb_debug_batch_begin_host)
        int st = rb_batch_host_build(b, b->host_w, b->host_h, false, n_threads, stage_malloc, nullptr, &blk, begin, end);
        b->host_block = blk;
        return st;
    }
    const bool mask_target = b->mask != nullptr;
    rb_ctx *ctx = batch_ctx(b);
    const int W = (int)(mask_target ? b->mask->w : b->layer->w), H = (int)(mask_target ? b->mask->h : b->layer->h);
    cudaSetDevice(ctx->device);
    // Device path geometry (geo.cu): large batches on layers are dashed / stroked / chopped / clipped by the geometry kernels
    // from the raw paths; the host builder below remains for small batches (a tree traversal's handful of draws per layer
    // is not worth a round trip), masks, and ranges the device hands back (RB_GEO_FALLBACK).
    {
        size_t n_range = (end == 0 || end > b->n_total ? b->n_total : end) - begin;
        size_t geo_from = 4096;
        if (const char *e = getenv("RB_GEO_FROM")) geo_from = (size_t)std::max(1, atoi(e));
        const bool eligible = !mask_target && W <= 65536 && H <= 65536 && !rb_debug_host_only_builder();
        // Which builder: the geometry kernels take ~0.35 us per draw of the 100 000-path scene (B200) and leave the host idle; the
        // host builder takes ~8 core-us per draw but overlaps the raster kernel part by part.  With more than 16 host threads
        // for this GPU the host builder finishes first; with fewer (several GPUs sharing the box's cores: one process per GPU)
        // the device does.  RB_GEO_MODE / rb_debug_geo_mode override.
        const int host_threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        int geo_max_threads = 16;
        if (const char *e = getenv("RB_GEO_MAX_HOST_THREADS")) geo_max_threads = atoi(e);
        if (allow_geo && eligible && g_geo_mode != 2 && (g_geo_mode == 1 || (n_range >= geo_from && host_threads <= geo_max_threads))) {
            int gst = rb_geo_prepare(b, n_threads, begin, end);
            if (gst == RB_OK) {
                if (!b->dev || b->lay.n_draws == 0) return RB_OK;
                const size_t scratch_bytes = warp_scratch_layout(b->lay).total;
                RB_CUDA(ctx, cudaMallocAsync((void **)&b->dev_scratch, scratch_bytes, ctx->stream));
                b->scratch_owned = true;
                return RB_OK;
            }
            if (gst != RB_GEO_FALLBACK) return gst;
        }
    }
    StageReq req{ctx, RB_OK};
    int st;
    { rb_prof_scope prof__(RB_T_BUILD); st = rb_batch_host_build(b, W, H, mask_target, n_threads, stage_pinned, &req, &blk, begin, end); }
    if (req.status != RB_OK) return req.status;
    rb_prof_scope prof_up__(RB_T_UPLOAD);
    if (st == RB_NEEDS_RUN_SPLIT) return st;
    if (st != RB_OK) return rb_fail(ctx, st, "batch host build failed");
    if (!blk || b->lay.n_draws == 0) return RB_OK;
    // one stream-ordered allocation holds the uploaded block and, behind it, the device-built lists and bins
    const size_t block_bytes = (b->lay.total + 255) & ~(size_t)255;
    const size_t scratch_bytes = b->lay.wide ? 0 : warp_scratch_layout(b->lay).total;
    uint8_t *dev = nullptr;
    RB_CUDA(ctx, cudaMallocAsync((void **)&dev, block_bytes + scratch_bytes, ctx->stream));
    b->dev = dev;
    RB_CUDA(ctx, cudaMemcpyAsync(dev, blk, b->lay.total, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += b->lay.total;
    { int st__ = rb_staging_mark(ctx); if (st__ != RB_OK) return st__; }
    if (!b->lay.wide) b->dev_scratch = dev + block_bytes;
    return RB_OK;
}

extern "C" int rb_batch_prepare(rb_batch *b, int32_t n_threads)
{
    rb_enter(b ? b->ctx : nullptr);
    if (b && b->layer && b->layer->pending != b) RB_SYNC_LAYER(b->layer);
    if (b && b->geo) rb_geo_abandon(b);
    int st = batch_prepare_range(b, n_threads, 0, 0);
    // hairline strokes + the fallback builder: only rb_batch_submit can interleave the two kinds of passes
    if (st == RB_NEEDS_RUN_SPLIT)
        return rb_fail(batch_ctx(b), RB_ERR_UNSUPPORTED, "rb_batch_prepare: hairline strokes in a batch that needs the fallback builder, use rb_batch_submit");
    return st;
}

static std::atomic<uint64_t> g_banded_downloads{0};
static int batch_run(rb_batch *b, unsigned long long *px_stats);
extern "C" int rb_batch_run(rb_batch *b)
{
    rb_enter(b ? b->ctx : nullptr);
    if (b && b->layer && b->layer->pending != b) RB_SYNC_LAYER(b->layer);
    return batch_run(b, nullptr);
}

// Runs the batch once with the coverage counters on: out[0] = pixels read-modify-written (partial coverage or
// non-opaque paint), out[1] = pixels stored without reading (full coverage, opaque solid).  Synchronises.
extern "C" int rb_batch_run_counting(rb_batch *b, uint64_t out[2])
{
    rb_enter(b ? b->ctx : nullptr);
    if (!b || !out) return RB_ERR_INVALID;
    rb_ctx *ctx = batch_ctx(b);
    if (!ctx) return RB_ERR_INVALID;
    unsigned long long *d = nullptr;
    RB_CUDA(ctx, cudaMallocAsync((void **)&d, 128, ctx->stream));
    RB_CUDA(ctx, cudaMemsetAsync(d, 0, 128, ctx->stream));
    int st = batch_run(b, d);
    unsigned long long h[16] = {0};
    RB_CUDA(ctx, cudaMemcpyAsync(h, d, 128, cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RB_CUDA(ctx, cudaFreeAsync(d, ctx->stream));
    out[0] = h[0];
    out[1] = h[1];
    if (getenv("RB_RASTER_DIAG"))
        fprintf(stderr, "[raster diag] pairs %llu, skipped(bounds/empty) %llu, skipped(no span) %llu, with coverage %llu, list entries %llu, crossings %llu, "
                        "blended px %llu, stored px %llu, pairs without a crossing edge %llu\n", h[2], h[3], h[4], h[5], h[6], h[7], h[0], h[1], h[8]);
    if (st == RB_OK && b->dev_scratch && !b->lay.wide) {
        // the list builder counts entries it had to drop (the host's capacity bound makes that impossible)
        if (rb_check_flags(ctx) != RB_OK) return RB_ERR_CUDA;
    }
    return st;
}

static int batch_run(rb_batch *b, unsigned long long *px_stats)
{
    rb_prof_scope prof__(RB_T_RUN);
    if (!b) return RB_ERR_INVALID;
    if (!b->dev || b->lay.n_draws == 0) return RB_OK; // nothing to draw
    const bool mask_target = b->mask != nullptr;
    rb_ctx *ctx = mask_target ? b->mask->ctx : b->layer->ctx;
    const int W = (int)(mask_target ? b->mask->w : b->layer->w), H = (int)(mask_target ? b->mask->h : b->layer->h);
    void *target = mask_target ? (void *)b->mask->d : (void *)b->layer->d;
    if (!(ctx->attr_bits & RB_ATTR_WIDE)) {
        RB_CUDA(ctx, cudaFuncSetAttribute(k_raster_tiles_wide<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CNT_BYTES));
        RB_CUDA(ctx, cudaFuncSetAttribute(k_raster_tiles_wide<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CNT_BYTES));
        ctx->attr_bits |= RB_ATTR_WIDE;
    }
    const BatchLayout &L = b->lay;
    if (!L.wide) {
        // warp-tile path: bin the draws, build the per-draw tile-row edge lists, rasterise
        const WarpScratch ws = warp_scratch_layout(L);
        uint8_t *sc = b->dev_scratch;
        const DevDraw *d_draws = (const DevDraw *)(b->dev + L.o_draws);
        // item mode: o_edges holds the uploaded line edges; the edge array proper is expanded into the scratch block
        const DevEdge *d_lines = (const DevEdge *)(b->dev + L.o_edges);
        DevEdge *d_edges = L.items ? (DevEdge *)(sc + ws.o_edges) : (DevEdge *)(b->dev + L.o_edges);
        unsigned int *d_flag = ctx->d_flags; // sticky, host-mapped: surfaced by rb_check_flags at every sync point
        uint32_t *row_off = (uint32_t *)(sc + ws.o_row_off), *row_cnt = (uint32_t *)(sc + ws.o_row_cnt);
        uint32_t *row_cols = (uint32_t *)(sc + ws.o_row_cols);
        uint32_t *tile_off = (uint32_t *)(sc + ws.o_tile_off), *tile_pairs = (uint32_t *)(sc + ws.o_tile_pairs);
        DevEdge *row_edges = (DevEdge *)(sc + ws.o_row_edges);
        DrawBox *boxes = (DrawBox *)(sc + ws.o_boxes);
        RowEnt *row_draws = (RowEnt *)(sc + ws.o_row_draws);
        const uint32_t n_draws = (uint32_t)L.n_draws, n_wtiles = (uint32_t)((size_t)L.wtiles_x * L.wtiles_y);
        const bool time_run = n_draws > 32; // rb_ctx_last_run_ms is about the large batches; small ones skip the event records
        if (time_run) RB_CUDA(ctx, cudaEventRecord(ctx->ev_run[0], ctx->stream));
        // Small batches (a tree traversal records a handful of draws per layer) skip the binning pre-pass: the raster
        // kernel then tests every draw of the batch against the tile itself (direct mode), which saves two memsets and
        // eight launches per batch.
        const bool direct = n_draws <= 32 && !getenv("RB_NO_DIRECT");
        if (direct) {
            k_row_lists<<<n_draws, RL_THREADS, 0, ctx->stream>>>(d_draws, d_lines, (const rbh::CurveRec *)(b->dev + L.o_curves), d_edges, row_off,
                                                                  row_edges, row_cols, L.items ? 1 : 0, d_flag, L.wtiles_x, nullptr, nullptr, nullptr);
            RB_LAUNCHED(ctx, "row_lists");
            tile_off = nullptr;
            tile_pairs = nullptr;
        } else {
        RB_CUDA(ctx, cudaMemsetAsync(row_cnt, 0, ((size_t)L.wtiles_y + 2) * 4, ctx->stream));
        RB_CUDA(ctx, cudaMemsetAsync(tile_off, 0, ((size_t)n_wtiles + 2) * 4, ctx->stream));
        k_row_lists<<<n_draws, RL_THREADS, 0, ctx->stream>>>(d_draws, d_lines, (const rbh::CurveRec *)(b->dev + L.o_curves), d_edges, row_off,
                                                              row_edges, row_cols, L.items ? 1 : 0, d_flag, L.wtiles_x, boxes, row_cnt, tile_off);
        RB_LAUNCHED(ctx, "row_lists");
        uint32_t *scan_tmp = (uint32_t *)(sc + ws.o_scan);
        auto exclusive_scan = [&](uint32_t *a, uint32_t n) -> int {
            const uint32_t nb = (n + SCAN_PER_CTA - 1) / SCAN_PER_CTA;
            if (nb > (uint32_t)SCAN_PER_CTA) return rb_fail(ctx, RB_ERR_UNSUPPORTED, "layer too large for the tile table scan");
            k_scan_local<<<nb, SCAN_THREADS, 0, ctx->stream>>>(a, n, scan_tmp);
            RB_LAUNCHED(ctx, "scan_local");
            k_scan_sums<<<1, SCAN_THREADS, 0, ctx->stream>>>(scan_tmp, nb, a, n);
            RB_LAUNCHED(ctx, "scan_sums");
            if (nb > 1) {
                k_scan_add<<<nb, SCAN_THREADS, 0, ctx->stream>>>(a, n, scan_tmp);
                RB_LAUNCHED(ctx, "scan_add");
            }
            return RB_OK;
        };
        { int st = exclusive_scan(row_cnt, (uint32_t)L.wtiles_y); if (st != RB_OK) return st; }
        { int st = exclusive_scan(tile_off, n_wtiles); if (st != RB_OK) return st; }
        k_bin_rows<<<L.wtiles_y, 256, 0, ctx->stream>>>(boxes, n_draws, row_cols, row_cnt, row_draws);
        RB_LAUNCHED(ctx, "bin_rows");
        k_bin_tiles<<<(n_wtiles + 7) / 8, 256, 0, ctx->stream>>>(row_draws, row_cnt, tile_off, L.wtiles_x, n_wtiles, tile_pairs);
        RB_LAUNCHED(ctx, "bin_tiles");
        }
        if (time_run) RB_CUDA(ctx, cudaEventRecord(ctx->ev_run[1], ctx->stream));
        unsigned grid = (n_wtiles + WT_WARPS - 1) / WT_WARPS;
        if (direct) grid = std::min(grid, (unsigned)(ctx->sm_count * RW_MIN_CTAS * 2)); // warps stride over the tiles
#define RB_WARP_ARGS target, W, H, L.wtiles_x, n_wtiles, tile_off, tile_pairs, d_draws, row_off, row_edges, d_edges,                   \
    (const DevPaint *)(b->dev + L.o_paints), (const DevStop *)(b->dev + L.o_stops), px_stats, row_cols, n_draws, tile0
        uint32_t tile0 = 0;
        if (b->dl_arm && b->dl_host && !direct && !mask_target && !px_stats && b->layer) {
            // The submit's last launch, the layer is wanted on the host: render it in bands of tile rows, top to bottom, and
            // copy every finished band out on the copy stream while the next one is rendered (the tiles of a band are
            // complete once its launch is: every draw of the batch that touches them is in their lists).
            b->dl_arm = false;
            int bands = 6;
            if (const char *e = getenv("RB_DL_BANDS")) bands = std::max(1, std::min(32, atoi(e)));
            bands = std::min(bands, L.wtiles_y);
            const bool diag = getenv("RB_DL_DIAG") != nullptr;
            std::vector<cudaEvent_t> evs; // diag: kernel start/end and copy start/end per band
            auto mark = [&](cudaStream_t s) { if (!diag) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); evs.push_back(e); };
            mark(ctx->stream);
            for (int k = 0; k < bands; k++) {
                const int ty0 = (int)((int64_t)L.wtiles_y * k / bands), ty1 = (int)((int64_t)L.wtiles_y * (k + 1) / bands);
                if (ty0 >= ty1) continue;
                tile0 = (uint32_t)ty0 * (uint32_t)L.wtiles_x;
                const unsigned bgrid = (unsigned)(ty1 - ty0) * (unsigned)L.wtiles_x;
                if (L.has_hair) k_raster_warp<false, true, false><<<bgrid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
                else k_raster_warp<false, false, false><<<bgrid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
                RB_LAUNCHED(ctx, "raster_warp");
                const int y0 = ty0 * WT_H, y1 = std::min(H, ty1 * WT_H);
                RB_CUDA(ctx, cudaEventRecord(ctx->ev_band, ctx->stream));
                mark(ctx->stream);
                RB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_band, 0));
                mark(ctx->copy_stream);
                RB_CUDA(ctx, cudaMemcpyAsync(b->dl_host + (size_t)y0 * W * 4, (const uint8_t *)target + (size_t)y0 * W * 4, (size_t)(y1 - y0) * W * 4,
                                             cudaMemcpyDeviceToHost, ctx->copy_stream));
                mark(ctx->copy_stream);
            }
            if (diag) {
                cudaStreamSynchronize(ctx->copy_stream);
                cudaStreamSynchronize(ctx->stream);
                fprintf(stderr, "[dl diag] per band: kernel end, copy start, copy end (ms after the first launch)\n");
                for (size_t i = 1; i + 2 < evs.size() + 1; i += 3) {
                    float a = 0, c0 = 0, c1 = 0;
                    cudaEventElapsedTime(&a, evs[0], evs[i]); cudaEventElapsedTime(&c0, evs[0], evs[i + 1]); cudaEventElapsedTime(&c1, evs[0], evs[i + 2]);
                    fprintf(stderr, "[dl diag]   %.2f  %.2f  %.2f\n", a, c0, c1);
                }
                for (auto e : evs) cudaEventDestroy(e);
            }
            b->dl_done = true;
            g_banded_downloads++;
            if (time_run) RB_CUDA(ctx, cudaEventRecord(ctx->ev_run[2], ctx->stream));
            return RB_OK;
        }
        if (direct) {
            if (mask_target) k_raster_warp<true, false, true><<<grid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
            else if (L.has_hair) k_raster_warp<false, true, true><<<grid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
            else k_raster_warp<false, false, true><<<grid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
        } else {
            if (mask_target) k_raster_warp<true, false, false><<<grid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
            else if (L.has_hair) k_raster_warp<false, true, false><<<grid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
            else k_raster_warp<false, false, false><<<grid, WT_WARPS * 32, 0, ctx->stream>>>(RB_WARP_ARGS);
        }
#undef RB_WARP_ARGS
        RB_LAUNCHED(ctx, "raster_warp");
        if (time_run) RB_CUDA(ctx, cudaEventRecord(ctx->ev_run[2], ctx->stream));
        return RB_OK;
    }
    const unsigned n_tile_ids = (unsigned)L.n_tile_ids;
#define RB_RASTER_ARGS target, W, H, L.tiles_x, (const uint32_t *)(b->dev + L.o_tids), (const uint32_t *)(b->dev + L.o_toff), \
    (const uint32_t *)(b->dev + L.o_tdraws), (const DevDraw *)(b->dev + L.o_draws), (const DevEdge *)(b->dev + L.o_edges),    \
    (const DevPaint *)(b->dev + L.o_paints), (const DevStop *)(b->dev + L.o_stops), px_stats
    // the any-winding fallback: CTA per 64x16 tile, host-expanded edges, host-built bins
    if (mask_target) k_raster_tiles_wide<true><<<n_tile_ids, RT_THREADS, CNT_BYTES, ctx->stream>>>(RB_RASTER_ARGS);
    else k_raster_tiles_wide<false><<<n_tile_ids, RT_THREADS, CNT_BYTES, ctx->stream>>>(RB_RASTER_ARGS);
#undef RB_RASTER_ARGS
    RB_LAUNCHED(ctx, "raster_tiles");
    return RB_OK;
}

// Device time of the context's last batch run: ms[0] = binning + edge-list pre-pass, ms[1] = the raster kernel.
extern "C" int rb_ctx_last_run_ms(rb_ctx *ctx, float ms[2])
{
    rb_enter(ctx);
    if (!ctx || !ms) return RB_ERR_INVALID;
    RB_CUDA(ctx, cudaEventSynchronize(ctx->ev_run[2]));
    RB_CUDA(ctx, cudaEventElapsedTime(&ms[0], ctx->ev_run[0], ctx->ev_run[1]));
    RB_CUDA(ctx, cudaEventElapsedTime(&ms[1], ctx->ev_run[1], ctx->ev_run[2]));
    return RB_OK;
}

// ---- hairline strokes: one thread per touched pixel applies that pixel's blits in the order the walker produced them ----
__global__ void __launch_bounds__(128)
k_hair_blits(uint32_t *__restrict__ px, int W, const HairGroup *__restrict__ groups, uint32_t n_groups, const HairDevBlit *__restrict__ blits,
             const DevPaint *__restrict__ paints, const DevStop *__restrict__ stops)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const HairGroup G = groups[g];
    uint32_t *p = px + (size_t)G.y * W + G.x;
    uint32_t d = *p;
    for (uint32_t k = 0; k < G.count; k++) {
        const HairDevBlit B = blits[G.first + k];
        d = blend_pixel(paints[B.paint], stops, d, B.alpha, (int)G.x - B.ox, (int)G.y - B.oy);
    }
    *p = d;
}

static int hair_run(rb_batch *b, size_t lo, size_t hi)
{
    if (!b->layer) return RB_OK; // Mask::fill_path never strokes
    rb_ctx *ctx = b->layer->ctx;
    HairBuilt hb;
    int st = rb_batch_hair_build(b, lo, hi, (int)b->layer->w, (int)b->layer->h, &hb);
    if (st != RB_OK) return rb_fail(ctx, st, "hairline build failed");
    if (hb.groups.empty()) return RB_OK;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_groups = 0, o_blits = o_groups + al(hb.groups.size() * sizeof(HairGroup));
    const size_t o_paints = o_blits + al(hb.blits.size() * sizeof(HairDevBlit));
    const size_t o_stops = o_paints + al(hb.paints.size() * sizeof(DevPaint));
    const size_t total = o_stops + al(std::max<size_t>(hb.stops.size(), 1) * sizeof(DevStop));
    void *stage = nullptr;
    st = rb_staging(ctx, total, &stage);
    if (st != RB_OK) return st;
    uint8_t *h = (uint8_t *)stage;
    memcpy(h + o_groups, hb.groups.data(), hb.groups.size() * sizeof(HairGroup));
    memcpy(h + o_blits, hb.blits.data(), hb.blits.size() * sizeof(HairDevBlit));
    memcpy(h + o_paints, hb.paints.data(), hb.paints.size() * sizeof(DevPaint));
    if (!hb.stops.empty()) memcpy(h + o_stops, hb.stops.data(), hb.stops.size() * sizeof(DevStop));
    uint8_t *dev = nullptr;
    cudaSetDevice(ctx->device);
    RB_CUDA(ctx, cudaMallocAsync((void **)&dev, total, ctx->stream));
    RB_CUDA(ctx, cudaMemcpyAsync(dev, h, total, cudaMemcpyHostToDevice, ctx->stream));
    { int st__ = rb_staging_mark(ctx); if (st__ != RB_OK) return st__; }
    const uint32_t n = (uint32_t)hb.groups.size();
    k_hair_blits<<<(n + 127) / 128, 128, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(b->layer->d), (int)b->layer->w, (const HairGroup *)(dev + o_groups), n,
                                                        (const HairDevBlit *)(dev + o_blits), (const DevPaint *)(dev + o_paints),
                                                        (const DevStop *)(dev + o_stops));
    RB_LAUNCHED(ctx, "hair_blits");
    RB_CUDA(ctx, cudaFreeAsync(dev, ctx->stream));
    return RB_OK;
}

// Large batches are submitted in a few consecutive parts (painter's order is kept: part k + 1 is rasterised after part
// k on the same stream), so the GPU works on one part while the host threads build the edges of the next.
static int submit_fill_run(rb_batch *b, int32_t n_threads, size_t first, size_t last, uint64_t total[6], size_t *stopped_at);

// Hairline strokes are not scan-converted: the batch is cut into runs of ordinary draws (tile kernel) and runs of
// hairlines (k_hair_blits), executed in painter's order on the context's stream.
extern "C" int rb_batch_submit(rb_batch *b, int32_t n_threads)
{
    rb_enter(b ? b->ctx : nullptr);
    if (!b) return RB_ERR_INVALID;
    if (b->layer && b->layer->pending != b) RB_SYNC_LAYER(b->layer); // immediate draws issued before this batch come first
    uint64_t total[6] = {0, 0, 0, 0, 0, 0};
    int st = RB_OK;
    // Hairline strokes ride along as draws of their own kind (their blits are applied by the tile kernel), except with
    // the any-winding fallback kernel, which does not know them: then the batch is cut into fill runs and hairline runs.
    size_t resume = 0;
    st = submit_fill_run(b, n_threads, 0, b->n_total, total, &resume);
    if (st == RB_NEEDS_RUN_SPLIT) {
        st = RB_OK;
        b->dl_host = nullptr; // fill runs and hairline runs alternate from here on: the caller downloads afterwards
        size_t i = resume; // everything before was drawn already
        while (i < b->n_total && st == RB_OK) {
            const bool hair = rb_batch_draw_is_hairline(b, i);
            size_t j = i + 1;
            while (j < b->n_total && rb_batch_draw_is_hairline(b, j) == hair) j++;
            st = hair ? hair_run(b, i, j) : submit_fill_run(b, n_threads, i, j, total, &resume);
            i = j;
        }
    }
    memcpy(b->stats, total, sizeof(total));
    return st;
}

// how many rb_batch_submit_download calls took the banded path so far (tests)
extern "C" uint64_t rb_debug_banded_downloads(void) { return g_banded_downloads.load(); }

// rb_batch_submit followed by rb_layer_download_begin(layer, host), with the download of the finished bands of the layer
// overlapping the rendering of the rest (resvg::render's target is a host pixmap: crates/resvg/src/lib.rs:34).  `host`:
// w * h * 4 bytes, pinned (rb_host_alloc) for the overlap to happen.  Returns once everything is enqueued;
// rb_layer_download_end(layer) waits for the pixels.
extern "C" int rb_batch_submit_download(rb_batch *b, int32_t n_threads, uint8_t *host)
{
    rb_enter(b ? b->ctx : nullptr);
    if (!b || !host || !b->layer) return RB_ERR_INVALID;
    rb_layer *l = b->layer;
    rb_ctx *ctx = l->ctx;
    if (l->dl_pending) { RB_CUDA(ctx, cudaEventSynchronize(l->dl_done)); l->dl_pending = false; }
    b->dl_host = host;
    b->dl_arm = false;
    b->dl_done = false;
    int st = rb_batch_submit(b, n_threads);
    const bool banded = b->dl_done;
    b->dl_host = nullptr;
    b->dl_done = false;
    if (st != RB_OK) {
        if (banded) cudaStreamSynchronize(ctx->copy_stream); // nothing may still be writing into `host` once the error is out
        return st;
    }
    if (!banded) return rb_layer_download_begin(l, host); // nothing to draw, a small (direct) batch, the any-winding fallback
    if (!l->dl_ready) {
        RB_CUDA(ctx, cudaEventCreateWithFlags(&l->dl_ready, cudaEventDisableTiming));
        RB_CUDA(ctx, cudaEventCreateWithFlags(&l->dl_done, cudaEventDisableTiming));
    }
    RB_CUDA(ctx, cudaEventRecord(l->dl_done, ctx->copy_stream)); // after the last band's copy
    l->dl_pending = true;
    return RB_OK;
}

// Runs draws [lo, hi) through the host builder in `parts` consecutive parts (the GPU rasterises part k while the host
// threads build part k + 1).
static int submit_host_parts(rb_batch *b, int32_t n_threads, size_t lo0, size_t hi0, size_t parts, uint64_t total[6], size_t *stopped_at,
                             bool tail = false) // tail: the range ends the whole submit (see rb_batch::dl_arm)
{
    const size_t n = hi0 - lo0;
    int st = RB_OK;
    for (size_t k = 0; k < parts && st == RB_OK; k++) {
        const size_t lo = lo0 + n * k / parts, hi = lo0 + n * (k + 1) / parts;
        if (lo >= hi) continue;
        *stopped_at = lo;
        st = batch_prepare_range(b, n_threads, lo, hi, /*allow_geo=*/false);
        b->dl_arm = tail && hi == hi0 && b->dl_host && st == RB_OK;
        if (st == RB_OK) st = rb_batch_run(b);
        b->dl_arm = false;
        for (int i = 0; i < 6; i++) total[i] += b->stats[i];
        batch_release(b);
    }
    return st;
}

static int submit_fill_run(rb_batch *b, int32_t n_threads, size_t first, size_t last, uint64_t total[6], size_t *stopped_at)
{
    const size_t n = last - first;
    *stopped_at = first;
    if (n == 0) return RB_OK;
    size_t parts = 1, split_from = 32768;
    if (const char *e = getenv("RB_SUBMIT_SPLIT_FROM")) split_from = (size_t)std::max(1, atoi(e)); // tests
    if (n >= split_from && (b->layer || b->mask)) {
        parts = 8;
        if (const char *e = getenv("RB_SUBMIT_PARTS")) parts = (size_t)std::max(1, atoi(e));
    }
    // Who builds the geometry.  The geometry kernels (geo.cu) take ~0.35 us per draw of the 100 000-path scene on a B200 and
    // pay the latency of their longest draws once per launch; the host builder takes ~8 core-us per draw and overlaps the
    // raster kernel part by part.  Both at once: the host threads build (and the GPU rasterises) the FIRST draws of the
    // batch while the geometry kernels, on their own stream, build the rest in one launch; the shares follow the two rates,
    // so a box with many cores per GPU gives the host more and eight processes sharing the box's cores give it next to
    // nothing.  RB_GEO_MODE / rb_debug_geo_mode: 1 = everything on the device, 2 = everything on the host.
    // a geometry launch costs a few milliseconds whatever its size (a dozen kernels, the latency of the longest draw, one
    // round trip for the totals): with host threads to spare only large batches are worth it
    const int host_threads_avail = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    size_t geo_from = host_threads_avail >= 8 ? 32768 : 4096;
    if (const char *e = getenv("RB_GEO_FROM")) geo_from = (size_t)std::max(1, atoi(e));
    const bool geo_ok = b->layer && !b->mask && b->layer->w <= 65536 && b->layer->h <= 65536 && !rb_debug_host_only_builder() && g_geo_mode != 2
                        && (g_geo_mode == 1 || n >= geo_from);
    size_t split = last; // draws [split, last) go to the geometry kernels
    if (geo_ok) {
        const int host_threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        double host_share = (host_threads / 8.3) / (host_threads / 8.3 + 2.9);
        if (const char *e = getenv("RB_GEO_HOST_SHARE")) host_share = atof(e);
        if (g_geo_mode == 1 || host_share < 0.12) host_share = 0.0; // not worth a second pipeline
        host_share = std::min(host_share, 1.0);
        split = first + (size_t)((double)n * host_share);
        if (last - split < geo_from && g_geo_mode != 1) split = last;
    }
    // The device share goes out in one launch (RB_GEO_SUBRANGES: in a few consecutive ones, each enqueued as soon as its tasks
    // exist — kept for experiments, see below).
    int st = RB_OK;
    size_t subs = 1;
    {
        const int host_threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        (void)host_threads; // measured with 4 host threads: 1 / 2 / 3 / 4 launches = 65 / 66 / 75 / 84 ms per scene — every launch
                            // pays the latency of its longest draws, which outweighs starting the GPU earlier
        if (const char *e = getenv("RB_GEO_SUBRANGES")) subs = (size_t)std::max(1, std::min(8, atoi(e)));
    }
    const bool whole_tail = b->dl_host != nullptr && last == b->n_total; // this run ends the submit: its last launch is banded
    struct Sub { size_t lo, hi; bool pending; };
    std::vector<Sub> sub_ranges;
    for (size_t k = 0; k < subs && split < last; k++) {
        Sub sr{split + (last - split) * k / subs, split + (last - split) * (k + 1) / subs, false};
        if (sr.lo >= sr.hi) continue;
        st = rb_geo_begin(b, n_threads, sr.lo, sr.hi);
        if (st == RB_OK) sr.pending = true;
        else if (st != RB_GEO_FALLBACK) { rb_geo_abandon(b); return st; }
        sub_ranges.push_back(sr);
    }
    st = RB_OK;
    if (split > first) {
        const size_t host_parts = std::max<size_t>(1, parts * (split - first) / n);
        st = submit_host_parts(b, n_threads, first, split, host_parts, total, stopped_at, whole_tail && sub_ranges.empty());
        if (st != RB_OK) { rb_geo_abandon(b); return st; }
    }
    for (const Sub &sr : sub_ranges) {
        *stopped_at = sr.lo;
        const bool sub_tail = whole_tail && &sr == &sub_ranges.back();
        int gst = RB_GEO_FALLBACK;
        if (sr.pending) {
            batch_release(b);
            gst = rb_geo_finish(b);
            if (gst == RB_OK) {
                if (b->dev && b->lay.n_draws) {
                    rb_ctx *ctx = batch_ctx(b);
                    const size_t scratch_bytes = warp_scratch_layout(b->lay).total;
                    RB_CUDA(ctx, cudaMallocAsync((void **)&b->dev_scratch, scratch_bytes, ctx->stream));
                    b->scratch_owned = true;
                    b->dl_arm = sub_tail;
                    gst = rb_batch_run(b);
                    b->dl_arm = false;
                }
                for (int i = 0; i < 6; i++) total[i] += b->stats[i];
                batch_release(b);
            }
        }
        if (gst == RB_GEO_FALLBACK)
            gst = submit_host_parts(b, n_threads, sr.lo, sr.hi, std::max<size_t>(1, parts * (sr.hi - sr.lo) / n), total, stopped_at, sub_tail);
        if (gst != RB_OK) { rb_geo_abandon(b); return gst; }
    }
    return RB_OK;
}

// Immediate draws are collected per layer and executed as one batch by rb_layer_flush, which every entry point that
// reads or writes a layer calls first (RB_SYNC_LAYER): a traversal that issues fill_path after fill_path on a layer and
// then composites it pays one tile-kernel launch, as if it had recorded an explicit batch.  Painter's order is the call
// order.  Draws with a pattern paint are executed at once (their source layer may be gone or changed by the next call).
constexpr uint32_t RB_PENDING_MAX = 16384;

int rb_layer_flush(rb_layer *l)
{
    if (!l || !l->pending) return RB_OK;
    rb_batch *b = l->pending;
    const uint32_t n = l->pending_n;
    l->pending = nullptr;
    l->pending_n = 0;
    {
        auto &dv = l->ctx->dirty;
        dv.erase(std::remove(dv.begin(), dv.end(), l), dv.end());
    }
    int st = rb_batch_submit(b, n > 64 ? 0 : 1);
    rb_batch_destroy(b);
    return st;
}

extern "C" int rb_fill_path(rb_layer *layer, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                            const rb_paint *paint, int32_t fill_rule, const float ts[6])
{
    rb_enter(layer ? layer->ctx : nullptr);
    if (!layer || !paint) return RB_ERR_INVALID;
    const bool pattern = paint->shader == 3;
    if (pattern) RB_SYNC_LAYER(paint->pattern);
    if (!layer->pending) {
        int st = rb_batch_begin(layer, &layer->pending);
        if (st != RB_OK) return st;
        layer->pending_n = 0;
        layer->ctx->dirty.push_back(layer);
        if (layer->vp_w > 0) (void)rb_batch_set_viewport(layer->pending, layer->vp_x, layer->vp_y, (uint32_t)layer->vp_w, (uint32_t)layer->vp_h);
    }
    int st = rb_batch_fill_path(layer->pending, verbs, n_verbs, points, n_points, paint, fill_rule, ts);
    if (st != RB_OK) return st; // nothing was recorded
    layer->pending_n++;
    if (pattern || layer->pending_n >= RB_PENDING_MAX) return rb_layer_flush(layer);
    return RB_OK;
}

// PixmapMut::stroke_path, immediate form: recorded into the layer's pending batch like rb_fill_path (painter's order =
// call order); rb_batch_stroke_path holds the reference's logic (dash, hairline test, stroker).
extern "C" int rb_stroke_path(rb_layer *layer, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                              const rb_paint *paint, const rb_stroke *stroke, const float ts[6])
{
    rb_enter(layer ? layer->ctx : nullptr);
    if (!layer || !paint || !stroke) return RB_ERR_INVALID;
    const bool pattern = paint->shader == 3;
    if (pattern) RB_SYNC_LAYER(paint->pattern);
    if (!layer->pending) {
        int st = rb_batch_begin(layer, &layer->pending);
        if (st != RB_OK) return st;
        layer->pending_n = 0;
        layer->ctx->dirty.push_back(layer);
        if (layer->vp_w > 0) (void)rb_batch_set_viewport(layer->pending, layer->vp_x, layer->vp_y, (uint32_t)layer->vp_w, (uint32_t)layer->vp_h);
    }
    const size_t before = layer->pending->n_total;
    int st = rb_batch_stroke_path(layer->pending, verbs, n_verbs, points, n_points, paint, stroke, ts);
    if (st != RB_OK) return st;
    if (layer->pending->n_total != before) layer->pending_n++;
    if (pattern || layer->pending_n >= RB_PENDING_MAX) return rb_layer_flush(layer);
    return RB_OK;
}

// PixmapMut::fill_rect(rect, paint, transform, None) (tiny-skia painter.rs).  With the identity transform tiny-skia
// blits the rectangle directly (scan::fill_rect / fill_rect_aa); otherwise it fills PathBuilder::from_rect(rect).  For
// a rectangle with integer edges the direct blit and the path fill cover exactly the same pixels at full coverage, and
// resvg only ever issues integer rectangles under the identity (filter/mod.rs:474-497, 853; image.rs:203) — a fractional
// anti-aliased rectangle under the identity would need fill_rect_aa's own edge coverage and is reported as unsupported
// rather than drawn with the path rasteriser's 4x4 coverage.
extern "C" int rb_fill_rect(rb_layer *layer, float x, float y, float w, float h, const rb_paint *paint, const float ts[6])
{
    if (!layer || !paint) return RB_ERR_INVALID;
    const float r = x + w, b = y + h; // Rect::from_xywh
    if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(r) && std::isfinite(b)) || !(x <= r && y <= b)) return RB_ERR_INVALID;
    const bool identity = !ts || (ts[0] == 1 && ts[1] == 0 && ts[2] == 0 && ts[3] == 1 && ts[4] == 0 && ts[5] == 0);
    const bool integral = x == floorf(x) && y == floorf(y) && r == floorf(r) && b == floorf(b);
    if (identity && paint->anti_alias && !integral && layer->w <= 8191 && layer->h <= 8191) return RB_ERR_UNSUPPORTED;
    if (!(w > 0.0f && h > 0.0f)) return RB_OK; // an empty rectangle blits nothing
    const uint8_t verbs[5] = {RB_VERB_MOVE, RB_VERB_LINE, RB_VERB_LINE, RB_VERB_LINE, RB_VERB_CLOSE};
    const float pts[8] = {x, y, r, y, r, b, x, b};
    return rb_fill_path(layer, verbs, 5, pts, 4, paint, RB_FILL_WINDING, ts);
}

extern "C" int rb_mask_fill_path(rb_mask *mask, const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                                 int32_t fill_rule, int32_t anti_alias, const float ts[6])
{
    rb_enter(mask ? mask->ctx : nullptr);
    if (!mask) return RB_ERR_INVALID;
    rb_batch b;
    b.mask = mask;
    rb_paint paint;
    memset(&paint, 0, sizeof(paint));
    paint.anti_alias = anti_alias;
    paint.blend_mode = RB_BLEND_SOURCE_OVER;
    if (mask->vp_w > 0) { b.vp_x = mask->vp_x; b.vp_y = mask->vp_y; b.vp_w = mask->vp_w; b.vp_h = mask->vp_h; }
    int st = rb_batch_record(&b, verbs, n_verbs, points, n_points, &paint, fill_rule, ts);
    if (st != RB_OK) return st;
    return rb_batch_submit(&b, 1);
}

// =================================================================================================
// layer composite — PixmapMut::draw_pixmap(x, y, src, {opacity, blend, Nearest}): highp pipeline
// (gather is highp-only): s = src/255 [* opacity]; blend with dst/255; store round(clamp*255).
// 12 B/px: src read, dst read, dst write.
// =================================================================================================
// SO: the blend mode is known to be SourceOver at compile time (every group composite of render.rs:133 with the default
// mix-blend-mode), which folds blendf's mode dispatch away.
template <bool SO = false>
__device__ __forceinline__ uint32_t draw_layer_px(uint32_t sp, uint32_t dp, float opacity, int blend)
{
    PF s = load_pf(sp);
    if (opacity != 1.0f) { s.r *= opacity; s.g *= opacity; s.b *= opacity; s.a *= opacity; }
    return store_pf(blendf(SO ? (int)RB_BLEND_SOURCE_OVER : blend, s, load_pf(dp)));
}

// grid.y strides over the rows of the clipped rectangle, threads run along x.  VEC: source and destination rows are
// 16-byte aligned at x0 and the width is a multiple of 4 (the whole-layer composite of render.rs:133): 4 px per thread.
template <bool VEC, bool SO>
__global__ void __launch_bounds__(256)
k_draw_layer(uint32_t *__restrict__ dst, int dw, const uint32_t *__restrict__ src, int sw, int x0, int y0, int x1, int y1,
             int ox, int oy, float opacity, int blend)
{
    const int w = x1 - x0;
    for (int y = y0 + blockIdx.y; y < y1; y += gridDim.y) {
        const uint32_t *srow = src + (size_t)(y - oy) * sw + (x0 - ox);
        uint32_t *drow = dst + (size_t)y * dw + x0;
        if (VEC) {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (w >> 2); i += gridDim.x * blockDim.x) {
                const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(srow) + i);
                uint4 d4 = reinterpret_cast<uint4 *>(drow)[i];
                d4.x = draw_layer_px<SO>(s4.x, d4.x, opacity, blend);
                d4.y = draw_layer_px<SO>(s4.y, d4.y, opacity, blend);
                d4.z = draw_layer_px<SO>(s4.z, d4.z, opacity, blend);
                d4.w = draw_layer_px<SO>(s4.w, d4.w, opacity, blend);
                reinterpret_cast<uint4 *>(drow)[i] = d4;
            }
        } else {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w; i += gridDim.x * blockDim.x)
                drow[i] = draw_layer_px<SO>(__ldg(srow + i), drow[i], opacity, blend);
        }
    }
}

extern "C" int rb_draw_layer(rb_layer *dst, const rb_layer *src, int32_t x, int32_t y, float opacity, int32_t blend_mode)
{
    rb_prof_scope prof__(RB_T_COMPOSITE);
    rb_enter(dst ? dst->ctx : nullptr);
    RB_SYNC_LAYER(dst);
    RB_SYNC_LAYER(src);
    if (!dst || !src || blend_mode < 0 || blend_mode > 28 || dst->d == src->d) return RB_ERR_INVALID;
    if (blend_mode == RB_BLEND_DESTINATION) return RB_OK;
    rb_ctx *ctx = dst->ctx;
    // fill_rect(rect(x, y, sw, sh)) clipped to the destination
    int64_t x0 = std::max<int64_t>(x, 0), y0 = std::max<int64_t>(y, 0);
    int64_t x1 = std::min<int64_t>((int64_t)x + src->w, dst->w), y1 = std::min<int64_t>((int64_t)y + src->h, dst->h);
    if (x1 <= x0 || y1 <= y0) return RB_OK;
    size_t n = (size_t)(x1 - x0) * (size_t)(y1 - y0);
    int blend = blend_mode == RB_BLEND_CLEAR ? RB_BLEND_CLEAR : blend_mode;
    (void)n;
    const int cw = (int)(x1 - x0), ch = (int)(y1 - y0);
    const bool vec = (cw & 3) == 0 && (x0 & 3) == 0 && ((x0 - x) & 3) == 0 && (dst->w & 3) == 0 && (src->w & 3) == 0;
    const int per_row = vec ? cw / 4 : cw;
    dim3 grid((unsigned)std::min(std::max((per_row + 255) / 256, 1), 64), (unsigned)std::min(ch, 16384));
#define RB_DL_ARGS reinterpret_cast<uint32_t *>(dst->d), (int)dst->w, reinterpret_cast<const uint32_t *>(src->d), (int)src->w, \
    (int)x0, (int)y0, (int)x1, (int)y1, x, y, opacity, blend
    const bool so = blend == RB_BLEND_SOURCE_OVER;
    if (vec && so) k_draw_layer<true, true><<<grid, 256, 0, ctx->stream>>>(RB_DL_ARGS);
    else if (vec) k_draw_layer<true, false><<<grid, 256, 0, ctx->stream>>>(RB_DL_ARGS);
    else if (so) k_draw_layer<false, true><<<grid, 256, 0, ctx->stream>>>(RB_DL_ARGS);
    else k_draw_layer<false, false><<<grid, 256, 0, ctx->stream>>>(RB_DL_ARGS);
#undef RB_DL_ARGS
    RB_LAUNCHED(ctx, "draw_layer");
    return RB_OK;
}

// Region-wise composite for atlases (many small documents in one layer, rb_batch_set_viewport): for every rectangle i,
// the w x h pixels of `src` at src_xy[i] are drawn onto `dst` at (x, y) with their own opacity — what render.rs:108-133
// does per document when a group needs a layer of its own (sub-pixmap + draw_pixmap with the group's opacity), and what
// filter/mod.rs does when it offsets / merges sub-images; one launch for all documents.
struct LayerRect {
    int32_t x, y, w, h, sx, sy;
    float opacity;
};
template <bool SO>
__global__ void __launch_bounds__(256)
k_draw_layer_rects(uint32_t *__restrict__ dst, int dw, const uint32_t *__restrict__ src, int sw, const LayerRect *__restrict__ rects,
                   int blend)
{
    const LayerRect r = rects[blockIdx.z];
    for (int y = blockIdx.y; y < r.h; y += gridDim.y) {
        uint32_t *drow = dst + (size_t)(r.y + y) * dw + r.x;
        const uint32_t *srow = src + (size_t)(r.sy + y) * sw + r.sx;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r.w; i += gridDim.x * blockDim.x)
            drow[i] = draw_layer_px<SO>(__ldg(srow + i), drow[i], r.opacity, blend);
    }
}

extern "C" int rb_draw_layer_rects(rb_layer *dst, const rb_layer *src, int32_t n, const int32_t *rects, const int32_t *src_xy,
                                   const float *opacity, int32_t blend_mode)
{
    rb_enter(dst ? dst->ctx : nullptr);
    RB_SYNC_LAYER(dst);
    RB_SYNC_LAYER(src);
    if (!dst || !src || n < 0 || (n > 0 && (!rects || !opacity)) || blend_mode < 0 || blend_mode > 28 || dst->d == src->d)
        return RB_ERR_INVALID;
    if (n == 0 || blend_mode == RB_BLEND_DESTINATION) return RB_OK;
    rb_ctx *ctx = dst->ctx;
    std::vector<LayerRect> host;
    host.reserve((size_t)n);
    int max_w = 0, max_h = 0;
    for (int32_t i = 0; i < n; i++) { // clip every rectangle to both layers
        int64_t x = rects[4 * i], y = rects[4 * i + 1], w = rects[4 * i + 2], h = rects[4 * i + 3];
        int64_t sx = src_xy ? src_xy[2 * i] : x, sy = src_xy ? src_xy[2 * i + 1] : y;
        if (w <= 0 || h <= 0) continue;
        int64_t cut = std::max<int64_t>(std::max<int64_t>(-x, -sx), 0);
        x += cut; sx += cut; w -= cut;
        cut = std::max<int64_t>(std::max<int64_t>(-y, -sy), 0);
        y += cut; sy += cut; h -= cut;
        w = std::min<int64_t>(w, std::min<int64_t>((int64_t)dst->w - x, (int64_t)src->w - sx));
        h = std::min<int64_t>(h, std::min<int64_t>((int64_t)dst->h - y, (int64_t)src->h - sy));
        if (w <= 0 || h <= 0) continue;
        host.push_back(LayerRect{(int32_t)x, (int32_t)y, (int32_t)w, (int32_t)h, (int32_t)sx, (int32_t)sy, opacity[i]});
        max_w = std::max(max_w, (int)w);
        max_h = std::max(max_h, (int)h);
    }
    if (host.empty()) return RB_OK;
    LayerRect *dev = nullptr;
    RB_CUDA(ctx, cudaMallocAsync((void **)&dev, host.size() * sizeof(LayerRect), ctx->stream));
    RB_CUDA(ctx, cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(LayerRect), cudaMemcpyHostToDevice, ctx->stream));
    for (size_t first = 0; first < host.size(); first += 65535) {
        const unsigned cnt = (unsigned)std::min<size_t>(65535, host.size() - first);
        dim3 grid((unsigned)std::min((max_w + 255) / 256, 16), (unsigned)std::min(max_h, 64), cnt);
        if (blend_mode == RB_BLEND_SOURCE_OVER)
            k_draw_layer_rects<true><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(dst->d), (int)dst->w,
                                                                    reinterpret_cast<const uint32_t *>(src->d), (int)src->w, dev + first, blend_mode);
        else
            k_draw_layer_rects<false><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(dst->d), (int)dst->w,
                                                                     reinterpret_cast<const uint32_t *>(src->d), (int)src->w, dev + first, blend_mode);
        RB_LAUNCHED(ctx, "draw_layer_rects");
    }
    RB_CUDA(ctx, cudaFreeAsync(dev, ctx->stream));
    return RB_OK;
}

// =================================================================================================
// masks — tiny-skia mask.rs
// =================================================================================================
extern "C" int rb_mask_create(rb_ctx *ctx, uint32_t w, uint32_t h, rb_mask **out)
{
    rb_enter(ctx);
    if (!ctx || !out || w == 0 || h == 0) return RB_ERR_INVALID;
    void *d = nullptr;
    cudaSetDevice(ctx->device);
    RB_CUDA(ctx, cudaMallocAsync(&d, (size_t)w * h, ctx->stream));
    RB_CUDA(ctx, cudaMemsetAsync(d, 0, (size_t)w * h, ctx->stream));
    *out = new rb_mask{ctx, w, h, (uint8_t *)d};
    rb_ctx_retain(ctx);
    return RB_OK;
}
extern "C" void rb_mask_destroy(rb_mask *m)
{
    if (!m) return;
    cudaFreeAsync(m->d, m->ctx->stream);
    rb_ctx_release(m->ctx);
    delete m;
}
extern "C" int rb_mask_download(rb_mask *m, uint8_t *host)
{
    rb_enter(m ? m->ctx : nullptr);
    if (!m || !host) return RB_ERR_INVALID;
    RB_CUDA(m->ctx, cudaMemcpyAsync(host, m->d, (size_t)m->w * m->h, cudaMemcpyDeviceToHost, m->ctx->stream));
    RB_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
    return RB_OK;
}
extern "C" int rb_mask_upload(rb_mask *m, const uint8_t *host)
{
    rb_enter(m ? m->ctx : nullptr);
    if (!m || !host) return RB_ERR_INVALID;
    RB_CUDA(m->ctx, cudaMemcpyAsync(m->d, host, (size_t)m->w * m->h, cudaMemcpyHostToDevice, m->ctx->stream));
    m->ctx->h2d_bytes += (size_t)m->w * m->h;
    return RB_OK;
}

// Mask::from_pixmap: 5 B/px
// div255[c] = c / 255.0f (IEEE division, as Rust's `c as f32 / 255.0`), tabulated per block
__device__ __forceinline__ uint32_t mask_px(uint32_t p, int luminance, const float *div255)
{
    const uint32_t av = RB_A(p);
    if (!luminance) return av;
    float r = div255[RB_R(p)], g = div255[RB_G(p)], b = div255[RB_B(p)];
    const float a = div255[av];
    if (av != 0 && av != 255) { r = __fdiv_rn(r, a); g = __fdiv_rn(g, a); b = __fdiv_rn(b, a); } // x / 1.0 == x
    const float luma = r * 0.2126f + g * 0.7152f + b * 0.0722f; // Rec. 709 (pinned by masking/mask goldens)
    float v = (luma * a) * 255.0f;
    v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v); // f32::clamp
    return rb_f2u8(ceilf(v));
}
// 4 pixels per thread: one 16-byte load, one 4-byte store (layers and masks are 256-byte aligned).
__global__ void __launch_bounds__(256) k_mask_from_layer(const uint32_t *__restrict__ px, uint8_t *__restrict__ m, size_t n, int luminance)
{
    __shared__ float div255[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) div255[i] = __fdiv_rn((float)i, 255.0f);
    __syncthreads();
    const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uint4 p = reinterpret_cast<const uint4 *>(px)[i];
        reinterpret_cast<uint32_t *>(m)[i] = mask_px(p.x, luminance, div255) | (mask_px(p.y, luminance, div255) << 8)
                                             | (mask_px(p.z, luminance, div255) << 16) | (mask_px(p.w, luminance, div255) << 24);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        m[i] = (uint8_t)mask_px(px[i], luminance, div255);
    }
}
__global__ void __launch_bounds__(256) k_mask_invert(uint8_t *__restrict__ m, size_t n)
{
    const size_t n16 = n >> 4, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 v = reinterpret_cast<uint4 *>(m)[i];
        v.x = ~v.x; v.y = ~v.y; v.z = ~v.z; v.w = ~v.w;
        reinterpret_cast<uint4 *>(m)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 15)) {
        const size_t i = (n16 << 4) + threadIdx.x;
        m[i] = (uint8_t)(255 - m[i]);
    }
}
// LoadMaskU8, LoadDestination, DestinationIn, Store (lowp): c' = div255(c * m).  9 B/px.
__device__ __forceinline__ uint32_t apply_mask_px(uint32_t p, uint32_t k)
{
    return rb_pack(div255(RB_R(p) * k), div255(RB_G(p) * k), div255(RB_B(p) * k), div255(RB_A(p) * k));
}
__global__ void __launch_bounds__(256) k_apply_mask(uint32_t *__restrict__ px, const uint8_t *__restrict__ m, size_t n)
{
    const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        uint4 p = reinterpret_cast<uint4 *>(px)[i];
        const uint32_t k = reinterpret_cast<const uint32_t *>(m)[i];
        p.x = apply_mask_px(p.x, k & 0xffu);
        p.y = apply_mask_px(p.y, (k >> 8) & 0xffu);
        p.z = apply_mask_px(p.z, (k >> 16) & 0xffu);
        p.w = apply_mask_px(p.w, k >> 24);
        reinterpret_cast<uint4 *>(px)[i] = p;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        px[i] = apply_mask_px(px[i], m[i]);
    }
}

extern "C" int rb_mask_from_layer(rb_mask *m, const rb_layer *l, int32_t luminance)
{
    rb_enter(m ? m->ctx : nullptr);
    RB_SYNC_LAYER(l);
    if (!m || !l || m->w != l->w || m->h != l->h) return RB_ERR_INVALID;
    rb_ctx *ctx = m->ctx;
    size_t n = (size_t)m->w * m->h;
    k_mask_from_layer<<<rb_grid_1d(ctx, (n + 3) / 4, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t *>(l->d), m->d, n, luminance);
    RB_LAUNCHED(ctx, "mask_from_layer");
    return RB_OK;
}
extern "C" int rb_mask_invert(rb_mask *m)
{
    rb_enter(m ? m->ctx : nullptr);
    if (!m) return RB_ERR_INVALID;
    rb_ctx *ctx = m->ctx;
    size_t n = (size_t)m->w * m->h;
    k_mask_invert<<<rb_grid_1d(ctx, (n + 15) / 16, 256), 256, 0, ctx->stream>>>(m->d, n);
    RB_LAUNCHED(ctx, "mask_invert");
    return RB_OK;
}
extern "C" int rb_layer_apply_mask(rb_layer *l, const rb_mask *m)
{
    rb_enter(l ? l->ctx : nullptr);
    RB_SYNC_LAYER(l);
    if (!m || !l) return RB_ERR_INVALID;
    if (m->w != l->w || m->h != l->h) return RB_OK; // tiny-skia: warn and return
    rb_ctx *ctx = l->ctx;
    size_t n = (size_t)m->w * m->h;
    k_apply_mask<<<rb_grid_1d(ctx, (n + 3) / 4, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(l->d), m->d, n);
    RB_LAUNCHED(ctx, "apply_mask");
    return RB_OK;
}

// Mask::from_pixmap(src, kind) [+ Mask::invert()] + Pixmap::apply_mask in ONE pass: the mask value is a function of the
// source pixel alone, so the u8 plane never has to exist.  12 B/px instead of 5 + (2) + 9.
//   mode 0: alpha mask, 1: luminance mask (mask.rs:40-45), 2: inverted alpha mask (clip.rs:25-27)
__global__ void __launch_bounds__(256) k_apply_layer_as_mask(uint32_t *__restrict__ px, const uint32_t *__restrict__ src, size_t n, int mode)
{
    __shared__ float div255[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) div255[i] = __fdiv_rn((float)i, 255.0f);
    __syncthreads();
    const int lum = mode == 1;
    const uint32_t inv = mode == 2 ? 255u : 0u;
    const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        uint4 p = reinterpret_cast<uint4 *>(px)[i];
        const uint4 s = reinterpret_cast<const uint4 *>(src)[i];
        p.x = apply_mask_px(p.x, mask_px(s.x, lum, div255) ^ inv);
        p.y = apply_mask_px(p.y, mask_px(s.y, lum, div255) ^ inv);
        p.z = apply_mask_px(p.z, mask_px(s.z, lum, div255) ^ inv);
        p.w = apply_mask_px(p.w, mask_px(s.w, lum, div255) ^ inv);
        reinterpret_cast<uint4 *>(px)[i] = p;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        px[i] = apply_mask_px(px[i], mask_px(src[i], lum, div255) ^ inv);
    }
}

static int apply_layer_as_mask(rb_layer *l, const rb_layer *src, int mode)
{
    rb_enter(l ? l->ctx : nullptr);
    RB_SYNC_LAYER(l);
    RB_SYNC_LAYER(src);
    if (!l || !src || l == src) return RB_ERR_INVALID;
    if (src->w != l->w || src->h != l->h) return RB_OK; // Pixmap::apply_mask: size mismatch is a warning and a no-op
    rb_ctx *ctx = l->ctx;
    const size_t n = (size_t)l->w * l->h;
    k_apply_layer_as_mask<<<rb_grid_1d(ctx, (n + 3) / 4, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t *>(l->d),
                                                                                     reinterpret_cast<const uint32_t *>(src->d), n, mode);
    RB_LAUNCHED(ctx, "apply_layer_as_mask");
    return RB_OK;
}
extern "C" int rb_layer_apply_mask_layer(rb_layer *layer, const rb_layer *mask_pixmap, int32_t luminance)
{
    return apply_layer_as_mask(layer, mask_pixmap, luminance ? 1 : 0);
}
extern "C" int rb_layer_apply_clip_layer(rb_layer *layer, const rb_layer *clip_pixmap)
{
    return apply_layer_as_mask(layer, clip_pixmap, 2);
}

// =================================================================================================
// host-only introspection used by the CPU test-suite (no device work): the line-edge list and blitter
// bounds the device would receive for one path on a (cw x ch) DrawTiler tile.
// =================================================================================================
extern "C" int rb_debug_build_edges(const uint8_t *verbs, int32_t n_verbs, const float *points, int32_t n_points,
                                    int32_t anti_alias, int32_t cw, int32_t ch, const float ts[6], int32_t *out_edges,
                                    int32_t *out_meta, int32_t max_edges, int32_t geom[7])
{
    if (!verbs || !points || !out_edges || !geom) return -1;
    std::vector<rbh::Pt> pts((size_t)n_points);
    memcpy(pts.data(), points, sizeof(float) * 2 * (size_t)n_points);
    if (ts) rbh::map_points(rbh::Xform::from(ts), pts.data(), n_points);
    std::vector<rbh::Edge> edges;
    rbh::DrawGeom g;
    if (!rbh::build_draw(verbs, n_verbs, pts.data(), n_points, anti_alias != 0, cw, ch, edges, &g)) return 0;
    if ((int64_t)edges.size() > max_edges) return -2;
    for (size_t i = 0; i < edges.size(); i++) {
        out_edges[i * 5 + 0] = edges[i].x;
        out_edges[i * 5 + 1] = edges[i].dx;
        out_edges[i * 5 + 2] = edges[i].first_y;
        out_edges[i * 5 + 3] = edges[i].last_y;
        out_edges[i * 5 + 4] = edges[i].winding;
        if (out_meta) { out_meta[i * 2 + 0] = edges[i].prev; out_meta[i * 2 + 1] = edges[i].before; }
    }
    geom[0] = g.sect.x; geom[1] = g.sect.y; geom[2] = g.sect.w; geom[3] = g.sect.h;
    geom[4] = g.shift; geom[5] = g.start_y; geom[6] = g.stop_y;
    return (int)edges.size();
}
