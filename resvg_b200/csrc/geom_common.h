// geom_common.h — what the path-geometry cores (stroker_core.h, dasher_core.h, hairline_core.h, fill_core.h) share.
//
// Every core is ONE source compiled twice: by g++ into the host builder (batch_host.cpp and the rb_path_* exports, with
// std::vector storage) and by nvcc into the device geometry kernels (geo.cu, with heap-backed DVec storage), so the two
// sides cannot drift apart: the CPU test-suite that pins the host geometry against the independent CPU checker pins the device code too,
// and the GPU tests compare the two builds bit for bit.  Both compilers run without FMA contraction (-fmad=false /
// -ffp-contract=off) and with IEEE division and square root.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "libm_compat.h"

// GEO_HD: a function of the cores.  On the device these are NOT inlined: the geometry code is a large tree of cases
// (clipping, joins, caps, recursive subdivision) and inlining it into every call site produced kernels of more than a
// megabyte of SASS, whose warps — each at a different point of it — spent most of their time waiting for instruction
// fetches (ncu: 50-67 no-instruction stall cycles per issued instruction).  GEO_HDI: the small helpers that are inlined.
#if defined(__CUDACC__)
#define GEO_HD __host__ __device__ __noinline__
#define GEO_HDI __host__ __device__
#else
#define GEO_HD
#define GEO_HDI
#endif

namespace geo {

struct P {
    float x, y;
};
GEO_HDI inline P operator+(P a, P b) { return P{a.x + b.x, a.y + b.y}; }
GEO_HDI inline P operator-(P a, P b) { return P{a.x - b.x, a.y - b.y}; }
GEO_HDI inline P operator-(P a) { return P{-a.x, -a.y}; }
GEO_HDI inline P operator*(P a, float s) { return P{a.x * s, a.y * s}; }
GEO_HDI inline bool operator==(P a, P b) { return a.x == b.x && a.y == b.y; }
GEO_HDI inline bool operator!=(P a, P b) { return !(a == b); }

template <class T> GEO_HDI inline T gmin(T a, T b) { return b < a ? b : a; } // std::min
template <class T> GEO_HDI inline T gmax(T a, T b) { return a < b ? b : a; } // std::max
template <class T> GEO_HDI inline void gswap(T &a, T &b) { T t = a; a = b; b = t; }
GEO_HDI inline bool gfinite(float v) { return fabsf(v) <= 3.402823466e+38f; } // false for NaN and +-inf
GEO_HDI inline bool gfinite(double v) { return fabs(v) <= 1.7976931348623157e+308; }
GEO_HDI inline bool finite(P a) { return gfinite(a.x) && gfinite(a.y); }

// Rust `as i32` casts: truncate, saturate, NaN -> 0
GEO_HDI inline int32_t f2i(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
GEO_HDI inline int32_t d2i(double v)
{
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}

// Transcendentals of the cubic solver (path_geometry.rs solve_cubic_poly; Rust forwards f32::acos / cos / cbrt to the platform
// libm): glibc's algorithms restated operation by operation (libm_compat.h), on BOTH sides — glibc 2.39's acosf / cosf /
// cbrtf are not correctly rounded, so CUDA's own functions (or double evaluation rounded once) give other floats for a few
// per cent of the arguments, which moved hairline cubics by a pixel here and there.
GEO_HDI inline float g_acosf(float v) { return lmc::acosf_(v); }
GEO_HDI inline float g_cosf(float v) { return lmc::cosf_(v); }
GEO_HDI inline float g_cbrtf(float v) { return lmc::cbrtf_(v); }

// ---- storage ---------------------------------------------------------------------------------------------------------------
// The cores are templates over the vector type: std::vector on the host, DVec on the device.  DVec grows like a vector
// but takes its chunks from a bump heap in global memory shared by the whole launch (one atomicAdd per growth; a chunk
// that has been outgrown is simply abandoned).  When the heap runs out the push is dropped and the launch-wide flag is
// raised: the host discards the launch's results and repeats it with a larger heap (geo.cu).  Everything here stays
// memory-safe after a dropped push; the results are garbage that nobody reads.
struct GeoHeap {
    uint8_t *base;
    unsigned long long *cursor; // bytes handed out so far
    unsigned long long size;
    unsigned int *overflow;
};
constexpr uint32_t kHeapAlign = 80; // lcm(sizeof(DevEdge) = 16, sizeof(CurveRec) = 40): element offsets from the base stay integral

#if defined(__CUDACC__)
template <class T> struct DVec {
    T *p;
    uint32_t n, cap;
    GeoHeap *h;
    __device__ __forceinline__ void init(GeoHeap *heap, uint32_t reserve_hint)
    {
        h = heap; p = nullptr; n = 0; cap = 0;
        grow_to(reserve_hint < 8 ? 8 : reserve_hint);
    }
    __device__ __forceinline__ bool ok() const { return p != nullptr; }
    __device__ __noinline__ void grow_to(uint32_t want)
    {
        const unsigned long long bytes = (((unsigned long long)want * sizeof(T) + kHeapAlign - 1) / kHeapAlign) * kHeapAlign;
        const unsigned long long off = atomicAdd(h->cursor, bytes);
        if (off + bytes > h->size) { *(volatile unsigned int *)h->overflow = 1u; return; }
        T *np = reinterpret_cast<T *>(h->base + off);
        for (uint32_t i = 0; i < n; i++) np[i] = p[i];
        p = np;
        cap = want;
    }
    __device__ __forceinline__ void push_back(const T &v)
    {
        if (n == cap) grow_to(cap * 2);
        if (n < cap) p[n++] = v;
    }
    __device__ __forceinline__ void pop_back() { if (n) n--; }
    __device__ __forceinline__ T &back() { return p[n ? n - 1 : 0]; }
    __device__ __forceinline__ const T &back() const { return p[n ? n - 1 : 0]; }
    __device__ __forceinline__ T &operator[](size_t i) { return p[i]; }
    __device__ __forceinline__ const T &operator[](size_t i) const { return p[i]; }
    __device__ __forceinline__ size_t size() const { return n; }
    __device__ __forceinline__ bool empty() const { return n == 0; }
    __device__ __forceinline__ void clear() { n = 0; }
    __device__ void resize(size_t m) // shrink, or grow with zero-filled elements
    {
        if (m > cap) grow_to((uint32_t)m);
        if (m > cap) return;
        for (size_t i = n; i < m; i++) memset(&p[i], 0, sizeof(T));
        n = (uint32_t)m;
    }
    __device__ __forceinline__ T *data() { return p; }
    __device__ __forceinline__ const T *data() const { return p; }
};
#endif

// Path verbs (include/resvg_b200.h RB_VERB_*)
enum { V_MOVE = 0, V_LINE = 1, V_QUAD = 2, V_CUBIC = 3, V_CLOSE = 4 };

} // namespace geo
