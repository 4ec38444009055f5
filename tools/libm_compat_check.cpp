// Compares resvg_b200/csrc/libm_compat.h with the host's libm: g++ -O2 -ffp-contract=off tools/libm_compat_check.cpp -o /tmp/lmc && /tmp/lmc
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../resvg_b200/csrc/libm_compat.h"

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static double uni() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (double)(rng >> 11) * (1.0 / 9007199254740992.0); }
static bool same(float a, float b) { return memcmp(&a, &b, 4) == 0 || (a != a && b != b); }

int main()
{
    long bad_a = 0, bad_c = 0, bad_r = 0, n = 0;
    // every float in [-1, 1] for acosf would be 2^31 values; a dense sample instead, plus every float near the branch points
    for (uint32_t bits = 0; bits <= 0x3f800000u; bits += 97) {
        for (int sgn = 0; sgn < 2; sgn++) {
            float x; uint32_t b = bits | (sgn ? 0x80000000u : 0u); memcpy(&x, &b, 4);
            if (!same(acosf(x), lmc::acosf_(x))) { if (bad_a < 5) printf("acosf(%a): libm %a ours %a\n", x, acosf(x), lmc::acosf_(x)); bad_a++; }
            n++;
        }
    }
    printf("acosf: %ld mismatches of %ld\n", bad_a, n);
    n = 0;
    for (long i = 0; i < 40000000; i++) {
        float x = (float)(uni() * 8.0 - 4.0);
        if (!same(cosf(x), lmc::cosf_(x))) { if (bad_c < 5) printf("cosf(%a): libm %a ours %a\n", x, cosf(x), lmc::cosf_(x)); bad_c++; }
        n++;
    }
    printf("cosf on [-4, 4]: %ld mismatches of %ld\n", bad_c, n);
    n = 0;
    for (uint32_t bits = 1; bits < 0x7f800000u; bits += 53) {
        float x; memcpy(&x, &bits, 4);
        if (!same(cbrtf(x), lmc::cbrtf_(x))) { if (bad_r < 5) printf("cbrtf(%a): libm %a ours %a\n", x, cbrtf(x), lmc::cbrtf_(x)); bad_r++; }
        if (!same(cbrtf(-x), lmc::cbrtf_(-x))) bad_r++;
        n += 2;
    }
    printf("cbrtf: %ld mismatches of %ld\n", bad_r, n);
    return (bad_a || bad_c || bad_r) ? 1 : 0;
}
