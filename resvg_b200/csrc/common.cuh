// common.cuh — device helpers shared by the resvg_b200 kernels.
//
// Parity rules: the reference is Rust, which never contracts a*b+c into an
// FMA, uses IEEE division/sqrt, and whose float->int `as` casts truncate toward zero, saturate and
// map NaN to 0.  All .cu files are compiled with -fmad=false and without -use_fast_math; the helpers
// below restate the cast rules.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// Rust `x as u8` for f32.
__device__ __forceinline__ uint32_t rb_f2u8(float v)
{
    // cvt.rzi.u32.f32 saturates (negatives -> 0, +inf -> 0xFFFFFFFF) and maps NaN to 0.
    return min(__float2uint_rz(v), 255u);
}

// Rust `x as u8` for f64.
__device__ __forceinline__ uint32_t rb_d2u8(double v)
{
    return min(__double2uint_rz(v), 255u);
}

// filter/mod.rs:242-254
__device__ __forceinline__ float rb_f32_bound(float mn, float val, float mx)
{
    if (val > mx) return mx;
    else if (val >= mn) return val;
    else return mn;
}

// float-cmp approx_eq_ulps(&0.0, 4) for f32 (usvg/src/tree/geom.rs:14-18).
__device__ __forceinline__ bool rb_approx_zero_ulps(float a)
{
    if (a == 0.0f) return true;
    int32_t bits = __float_as_int(a);
    if (bits < 0) return false; // sign differs from +0.0
    return bits <= 4;
}

__device__ __forceinline__ bool rb_approx_eq_ulps(float a, float b, int32_t ulps)
{
    if (a == b) return true;
    int32_t ai = __float_as_int(a), bi = __float_as_int(b);
    if ((ai < 0) != (bi < 0)) return false;
    int32_t diff = (int32_t)((uint32_t)ai - (uint32_t)bi);
    return diff >= -ulps && diff <= ulps;
}

__device__ __forceinline__ uint32_t rb_pack(uint32_t r, uint32_t g, uint32_t b, uint32_t a)
{
    return r | (g << 8) | (b << 16) | (a << 24);
}
#define RB_R(p) ((p) & 0xffu)
#define RB_G(p) (((p) >> 8) & 0xffu)
#define RB_B(p) (((p) >> 16) & 0xffu)
#define RB_A(p) ((p) >> 24)
