// edge_math.h — fixed-point edge setup shared by the host edge builder (raster_host.cpp) and the device kernel that
// expands curves (raster_warp.cuh), so both produce bit-identical line edges.
//
// Restates tiny-skia 0.12.0 edge.rs (a Rust port of Skia's SkEdge.cpp): LineEdge::new, QuadraticEdge::new / update,
// CubicEdge::new / update, with fixed_point.rs arithmetic (FDot6 = 26.6, FDot16 = 16.16, wrapping i32).  A curve is
// forward-differenced into the line edges its update() calls would hand to the scanline walker, one after another.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD static inline
#endif

namespace rbe {

RB_HD int32_t shl(int32_t v, int s) { return (int32_t)((uint32_t)v << s); }
RB_HD int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
RB_HD int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
RB_HD int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
RB_HD int32_t fdot6_round(int32_t n) { return wadd(n, 32) >> 6; }
RB_HD int32_t fdot16_mul(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 16); }
RB_HD int32_t fdot6_div(int32_t a, int32_t b)
{
    if (a >= -32768 && a <= 32767) return shl(a, 16) / b;
    int64_t v = ((int64_t)a * 65536) / (int64_t)b;
    if (v < INT32_MIN) v = INT32_MIN;
    if (v > INT32_MAX) v = INT32_MAX;
    return (int32_t)v;
}
RB_HD int clz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// One line edge: FDot16 x at first_y, FDot16 slope, inclusive scanline range.
struct RawEdge { int32_t x, dx, first_y, last_y; };

// LineEdge::new (after the y swap) and the tail of every curve update(): FDot6 end points with y0 <= y1.
RB_HD bool line_edge(int32_t x0, int32_t y0, int32_t x1, int32_t y1, RawEdge *e)
{
    const int32_t top = fdot6_round(y0), bot = fdot6_round(y1);
    if (top == bot) return false;
    const int32_t slope = fdot6_div(wsub(x1, x0), wsub(y1, y0));
    const int32_t dy = wsub(wadd(shl(top, 6), 32), y0);
    e->x = shl(wadd(x0, fdot16_mul(slope, dy)), 10);
    e->dx = slope;
    e->first_y = top;
    e->last_y = bot - 1;
    return true;
}

RB_HD int32_t cheap_distance(int32_t dx, int32_t dy)
{
    dx = dx < 0 ? -dx : dx;
    dy = dy < 0 ? -dy : dy;
    return dx > dy ? dx + (dy >> 1) : dy + (dx >> 1);
}
RB_HD int diff_to_shift(int32_t dx, int32_t dy, int shift_aa)
{
    const int32_t dist = (cheap_distance(dx, dy) + 16) >> (3 + shift_aa);
    return (32 - clz32((uint32_t)dist)) / 2;
}

// Number of subdivisions (log2) QuadraticEdge::new picks; FDot6 points.
RB_HD int quad_shift(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int shift_aa)
{
    int sh = diff_to_shift((shl(x1, 1) - x0 - x2) >> 2, (shl(y1, 1) - y0 - y2) >> 2, shift_aa);
    if (sh == 0) sh = 1;
    else if (sh > 6) sh = 6;
    return sh;
}

// QuadraticEdge::new + every update(): calls emit(RawEdge) for each non-degenerate segment, top to bottom.
// Points are FDot6 with y0 <= y2 (the caller swapped them and dropped curves with round(y0) == round(y2)).
template <class F>
RB_HD void quad_expand(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int sh, F emit)
{
    int count = 1 << sh;
    const int cshift = sh - 1;
    int32_t a = shl(x0 - x1 - x1 + x2, 9), b = shl(x1 - x0, 10);
    int32_t qx = shl(x0, 10), qdx = wadd(b, a >> sh);
    const int32_t qddx = a >> (sh - 1);
    a = shl(y0 - y1 - y1 + y2, 9);
    b = shl(y1 - y0, 10);
    int32_t qy = shl(y0, 10), qdy = wadd(b, a >> sh);
    const int32_t qddy = a >> (sh - 1);
    const int32_t lastx = shl(x2, 10), lasty = shl(y2, 10);
    while (count > 0) {
        int32_t nx, ny;
        if (--count > 0) {
            nx = wadd(qx, qdx >> cshift);
            qdx = wadd(qdx, qddx);
            ny = wadd(qy, qdy >> cshift);
            qdy = wadd(qdy, qddy);
        } else { nx = lastx; ny = lasty; }
        RawEdge e;
        if (line_edge(qx >> 10, qy >> 10, nx >> 10, ny >> 10, &e)) emit(e);
        qx = nx; qy = ny;
    }
}

RB_HD int32_t cubic_delta(int32_t a, int32_t b, int32_t c, int32_t d)
{
    int32_t one = wmul(a * 8 - b * 15 + 6 * c + d, 19) >> 9;
    int32_t two = wmul(a + 6 * b - c * 15 + d * 8, 19) >> 9;
    one = one < 0 ? -one : one;
    two = two < 0 ? -two : two;
    return one > two ? one : two;
}

// Number of subdivisions (log2) CubicEdge::new picks; FDot6 points.
RB_HD int cubic_shift(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t x3, int32_t y3)
{
    int sh = diff_to_shift(cubic_delta(x0, x1, x2, x3), cubic_delta(y0, y1, y2, y3), 2) + 1;
    if (sh > 6) sh = 6;
    return sh;
}

// CubicEdge::new + every update(); same contract as quad_expand (y0 <= y3).
template <class F>
RB_HD void cubic_expand(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t x3, int32_t y3, int sh, F emit)
{
    int up = 6, down = sh + up - 10;
    if (down < 0) { down = 0; up = 10 - sh; }
    int count = -(1 << sh);
    int32_t b = shl(3 * (x1 - x0), up), c = shl(3 * (x0 - x1 - x1 + x2), up), d = shl(x3 + 3 * (x1 - x2) - x0, up);
    int32_t cx = shl(x0, 10), cdx = wadd(wadd(b, c >> sh), d >> (2 * sh)), cddx = wadd(wmul(2, c), wmul(3, d) >> (sh - 1));
    const int32_t cdddx = wmul(3, d) >> (sh - 1);
    b = shl(3 * (y1 - y0), up);
    c = shl(3 * (y0 - y1 - y1 + y2), up);
    d = shl(y3 + 3 * (y1 - y2) - y0, up);
    int32_t cy = shl(y0, 10), cdy = wadd(wadd(b, c >> sh), d >> (2 * sh)), cddy = wadd(wmul(2, c), wmul(3, d) >> (sh - 1));
    const int32_t cdddy = wmul(3, d) >> (sh - 1);
    const int32_t lastx = shl(x3, 10), lasty = shl(y3, 10);
    while (count < 0) {
        int32_t nx, ny;
        if (++count < 0) {
            nx = wadd(cx, cdx >> down);
            cdx = wadd(cdx, cddx >> sh);
            cddx = wadd(cddx, cdddx);
            ny = wadd(cy, cdy >> down);
            cdy = wadd(cdy, cddy >> sh);
            cddy = wadd(cddy, cdddy);
        } else { nx = lastx; ny = lasty; }
        if (ny < cy) ny = cy;
        RawEdge e;
        if (line_edge(cx >> 10, cy >> 10, nx >> 10, ny >> 10, &e)) emit(e);
        cx = nx; cy = ny;
    }
}

} // namespace rbe
