// libm_compat.h — the three libm functions the path geometry calls (tiny-skia-path path_geometry.rs solve_cubic_poly:
// f32::acos, f32::cos, f32::cbrt, which Rust forwards to the platform libm), restated operation by operation after glibc
// 2.39 (sysdeps/ieee754/flt-32: e_acosf.c — the fdlibm float acos; s_cosf.c / sincosf.h — Szabolcs Nagy's double-polynomial
// cosf; s_cbrtf.c), so that the device computes the same floats as the host's libm does: none of the three is correctly
// rounded in glibc 2.39 (8 % / 1 % / 11 % of random arguments differ from the correctly rounded value), so evaluating in
// double and rounding once does not reproduce them.  Both builds of the geometry cores call THESE, which makes host and
// device identical by construction; tools/libm_compat_check.cpp compares them with the host's libm.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LMC_HD __host__ __device__
#else
#define LMC_HD
#endif

namespace lmc {

LMC_HD inline int32_t float_word(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
LMC_HD inline float word_float(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

// e_acosf.c
LMC_HD inline float acosf_(float x)
{
    const float one = 1.0000000000e+00f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f,
                pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    float z, p, q, r, w, s, c, df;
    const int32_t hx = float_word(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) { // |x| == 1
        if (hx > 0) return 0.0f;
        return pi + 2.0f * pio2_lo;
    } else if (ix > 0x3f800000) {
        return (x - x) / (x - x); // NaN
    }
    if (ix < 0x3f000000) { // |x| < 0.5
        if (ix <= 0x32800000) return pio2_hi + pio2_lo; // |x| <= 2^-26
        z = x * x;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (hx < 0) { // x < -0.5
        z = (one + x) * 0.5f;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        s = sqrtf(z);
        r = p / q;
        w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    } else { // x > 0.5
        z = (one - x) * 0.5f;
        s = sqrtf(z);
        df = word_float(float_word(s) & (int32_t)0xfffff000);
        c = (z - df * df) / (s + df);
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        w = r * s + c;
        return 2.0f * (df + w);
    }
}

// s_cosf.c / sincosf.h for |y| < 120 (the cubic solver's arguments lie in [-2.1, 3.2]); beyond, the caller's libm / device cos.
LMC_HD inline uint32_t abstop12(float x) { return ((uint32_t)float_word(x) >> 20) & 0x7ff; }
LMC_HD inline float sincos_poly(double x, double x2, bool negated, int n)
{
    // __sincosf_table[negated]: c0..c4 change sign, s1..s3 do not
    const double sg = negated ? -1.0 : 1.0;
    const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5, c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = x * x2, s1_ = s2 + x2 * s3;
        const double x7 = x3 * x2, s = x + x3 * s1;
        return (float)(s + x7 * s1_);
    } else {
        const double x4 = x2 * x2, c2_ = c3 + x2 * c4, c1_ = c0 + x2 * c1;
        const double x6 = x4 * x2, c = c1_ + x4 * c2;
        return (float)(c + x6 * c2_);
    }
}
LMC_HD inline float cosf_(float y)
{
    double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) { // |y| < pi/4
        const double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sincos_poly(x, x2, false, 1);
    }
    if (abstop12(y) < abstop12(120.0f)) {
        // reduce_fast: the quadrant from a scaled float-to-int conversion (hpi_inv carries 2^24)
        const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
        const double r = x * hpi_inv;
        const int n = ((int32_t)r + 0x800000) >> 24;
        x = x - n * hpi;
        const double sign[4] = {1.0, -1.0, -1.0, 1.0};
        const double s = sign[n & 3];
        return sincos_poly(x * s, x * x, (n & 2) != 0, n ^ 1);
    }
#if defined(__CUDA_ARCH__)
    return (float)cos((double)y);
#else
    return cosf(y);
#endif
}

// s_cbrtf.c
LMC_HD inline float cbrtf_(float x)
{
    const double CBRT2 = 1.2599210498948731648, SQR_CBRT2 = 1.5874010519681994748;
    const double factor[5] = {1.0 / SQR_CBRT2, 1.0 / CBRT2, 1.0, CBRT2, SQR_CBRT2};
    int xe;
    const float xm = frexpf(fabsf(x), &xe);
    // zero, infinity and NaN: frexp leaves the exponent at 0
    if (xe == 0 && (x == 0.0f || !(fabsf(x) <= 3.402823466e+38f))) return x + x;
    const float u = (float)(0.492659620528969547 + (0.697570460207922770 - 0.191502161678719066 * (double)xm) * (double)xm);
    const float t2 = u * u * u;
    const float ym = (float)((double)u * ((double)t2 + 2.0 * (double)xm) / (2.0 * (double)t2 + (double)xm) * factor[2 + xe % 3]);
    return ldexpf(x > 0.0f ? ym : -ym, xe / 3);
}

} // namespace lmc
