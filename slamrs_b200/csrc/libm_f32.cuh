// Bit-reproducible f32 sine / cosine for the SLAM kernels.
//
// Why this exists: the Rust reference computes every pose and beam endpoint with
// `f32::cos` / `f32::sin` (slamrs/common/src/robot.rs:180-181,
// slamrs/slam/src/grid/map.rs:76-77,121-122). On x86-64 Linux those lower to glibc's
// sinf/cosf, which are NOT correctly rounded (max error ~0.56 ulp), and CUDA's
// sinf/cosf round differently again. A one-ulp difference in an endpoint can move it across a
// cell boundary, which breaks the "ray-cast cell indices bit-exact" requirement.
//
// glibc >= 2.28 evaluates sinf/cosf entirely in IEEE binary64 (range reduction by pi/2 in
// double, then a degree-7/8 polynomial, then one rounding to binary32). Every step is a
// plain IEEE operation, so performing the same operations in the same order on the device
// gives the same bits. The operation order below (which products are fused) follows what
// glibc 2.39's FMA-enabled x86-64 build executes (checked by disassembling libm.so.6;
// tools/check_libm_f32.c compares all 2^32 inputs against the host libm).
//
// Provenance and licence: the algorithm and its constants (the reduction constant 2^24 / pi-ish `hpi_inv`,
// `hpi`, and the two sets of minimax coefficients) are those of glibc's sysdeps/ieee754/flt-32/s_sincosf.h and
// s_sincosf_data.c (contributed by Szabolcs Nagy / Arm, also published in ARM-software/optimized-routines under
// MIT OR Apache-2.0 WITH LLVM-exception); glibc itself is LGPL-2.1-or-later. They are restated here from the
// disassembly of the installed libm.so.6, because bit-equality with that library is the requirement; a
// redistributor should treat this header as derived from optimized-routines' sincosf (MIT).
//
// This file compiles for host (gcc/g++) and device (nvcc).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define SLAMRS_HD __host__ __device__ __forceinline__
#else
#define SLAMRS_HD static inline
#endif

namespace slamrs_libm {

SLAMRS_HD double fma_rn(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
SLAMRS_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
SLAMRS_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

// Polynomial coefficients (minimax fits on [-pi/4, pi/4], published with the algorithm).
// `neg` selects the table that yields -cos (sine coefficients are shared because the sine
// polynomial is odd and the argument carries the sign).
struct Poly {
    double c0, c1, c2, c3, c4, s1, s2, s3;
};

SLAMRS_HD Poly poly_table(bool neg) {
    Poly p;
    const double sg = neg ? -1.0 : 1.0;
    p.c0 = sg * 0x1p0;
    p.c1 = sg * -0x1.ffffffd0c621cp-2;
    p.c2 = sg * 0x1.55553e1068f19p-5;
    p.c3 = sg * -0x1.6c087e89a359dp-10;
    p.c4 = sg * 0x1.99343027bf8c3p-16;
    p.s1 = -0x1.555545995a603p-3;
    p.s2 = 0x1.1107605230bc4p-7;
    p.s3 = -0x1.994eb3774cf24p-13;
    return p;
}

// Evaluate sin (even quadrant) or cos (odd quadrant) of the reduced argument.
// x carries the quadrant sign already, x2 = x_reduced^2.
SLAMRS_HD float eval_poly(double x, double x2, const Poly& p, int n) {
    if ((n & 1) == 0) {
        const double x3 = mul_rn(x, x2);
        const double s1 = fma_rn(x2, p.s3, p.s2);
        const double x7 = mul_rn(x3, x2);
        const double s = fma_rn(x3, p.s1, x);
        return (float)fma_rn(s1, x7, s);
    } else {
        const double x4 = mul_rn(x2, x2);
        const double c2 = fma_rn(x2, p.c4, p.c3);
        const double c1 = fma_rn(x2, p.c1, p.c0);
        const double x6 = mul_rn(x4, x2);
        const double c = fma_rn(x4, p.c2, c1);
        return (float)fma_rn(c2, x6, c);
    }
}

// |x| < 120: quadrant n = round(x * 2/pi) via a 2^24-scaled truncating conversion.
SLAMRS_HD double reduce_fast(double x, int* np) {
    const double hpi_inv_2p24 = 0x1.45f306dc9c883p+23;
    const double hpi = 0x1.921fb54442d18p+0;
    const double r = mul_rn(x, hpi_inv_2p24);
    const int32_t n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma_rn(-(double)n, hpi, x);
}

// 120 <= |x| < inf: multiply the 24-bit mantissa by a 96-bit window of 4/pi.
// Successive 32-bit windows (8-bit stride) of the binary expansion of 4/pi.
#define SLAMRS_INV_PIO4_WORDS                                                                  \
    0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u, \
    0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, \
    0xf534ddc0u, 0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, \
    0x993c4390u, 0x3c439041u
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t k_inv_pio4_dev[24] = {SLAMRS_INV_PIO4_WORDS};
#endif
static const uint32_t k_inv_pio4_host[24] = {SLAMRS_INV_PIO4_WORDS};

SLAMRS_HD double reduce_large(uint32_t xi, int* np) {
    const double pi63 = 0x1.921fb54442d18p-62;
#if defined(__CUDA_ARCH__)
    const uint32_t* arr = &k_inv_pio4_dev[(xi >> 26) & 15];
#else
    const uint32_t* arr = &k_inv_pio4_host[(xi >> 26) & 15];
#endif
    const int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    res0 = (uint32_t)(xi * arr[0]);
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    const double x = (double)(int64_t)res0;
    *np = (int)n;
    return mul_rn(x, pi63);
}

SLAMRS_HD double quadrant_sign(int n) {
    // sign of sine in quadrants 0..3: +, -, -, +
    const int q = n & 3;
    return (q == 1 || q == 2) ? -1.0 : 1.0;
}

SLAMRS_HD float sinf_exact(float y) {
    const uint32_t yi = f32_bits(y);
    const uint32_t top = (yi >> 20) & 0x7ffu;
    double x = (double)y;
    int n;
    if (top < 0x3f4u) {  // |y| < pi/4
        if (top < 0x398u) return y;  // |y| < 2^-12
        return eval_poly(x, mul_rn(x, x), poly_table(false), 0);
    } else if (top < 0x42fu) {  // |y| < 120
        x = reduce_fast(x, &n);
        const double s = quadrant_sign(n);
        return eval_poly(mul_rn(x, s), mul_rn(x, x), poly_table((n & 2) != 0), n);
    } else if (top < 0x7f8u) {
        const int sign = (int)(yi >> 31);
        x = reduce_large(yi, &n);
        const double s = quadrant_sign(n + sign);
        return eval_poly(mul_rn(x, s), mul_rn(x, x), poly_table(((n + sign) & 2) != 0), n);
    }
    return y - y;  // inf/nan -> nan
}

SLAMRS_HD float cosf_exact(float y) {
    const uint32_t yi = f32_bits(y);
    const uint32_t top = (yi >> 20) & 0x7ffu;
    double x = (double)y;
    int n;
    if (top < 0x3f4u) {
        if (top < 0x398u) return 1.0f;
        return eval_poly(x, mul_rn(x, x), poly_table(false), 1);
    } else if (top < 0x42fu) {
        x = reduce_fast(x, &n);
        const double s = quadrant_sign(n);
        return eval_poly(mul_rn(x, s), mul_rn(x, x), poly_table((n & 2) != 0), n ^ 1);
    } else if (top < 0x7f8u) {
        const int sign = (int)(yi >> 31);
        x = reduce_large(yi, &n);
        const double s = quadrant_sign(n + sign);
        return eval_poly(mul_rn(x, s), mul_rn(x, x), poly_table(((n + sign) & 2) != 0), n ^ 1);
    }
    return y - y;
}

// sinf and cosf of the same argument with one shared range reduction (what glibc's sincosf does;
// the results are bit-identical to the two separate calls above: same reduced argument, same
// polynomials, the quadrant parity only decides which polynomial feeds which output).
SLAMRS_HD void sincosf_exact(float y, float* sn, float* cs) {
    const uint32_t yi = f32_bits(y);
    const uint32_t top = (yi >> 20) & 0x7ffu;
    double x = (double)y;
    int n = 0, tsel = 0;
    double sg = 1.0;
    if (top < 0x3f4u) {  // |y| < pi/4
        if (top < 0x398u) { *sn = y; *cs = 1.0f; return; }  // |y| < 2^-12
    } else if (top < 0x42fu) {  // |y| < 120
        x = reduce_fast(x, &n);
        sg = quadrant_sign(n);
        tsel = n;
    } else if (top < 0x7f8u) {
        const int sign = (int)(yi >> 31);
        x = reduce_large(yi, &n);
        sg = quadrant_sign(n + sign);
        tsel = n + sign;
    } else {
        *sn = y - y; *cs = y - y;  // inf/nan -> nan
        return;
    }
    const Poly p = poly_table((tsel & 2) != 0);
    const double x2 = mul_rn(x, x);
    const double xs = mul_rn(x, sg);
    const float s_poly = eval_poly(xs, x2, p, 0);
    const float c_poly = eval_poly(xs, x2, p, 1);
    if ((n & 1) == 0) { *sn = s_poly; *cs = c_poly; }
    else { *sn = c_poly; *cs = s_poly; }
}

}  // namespace slamrs_libm
