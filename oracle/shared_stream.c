/* TEST INFRASTRUCTURE -- CPU oracle side of the shared seeded stream. See shared_stream.h.
 * Build with -ffp-contract=off (oracle/Makefile does). */
#include "shared_stream.h"
#include <math.h>
#include <string.h>

static inline void mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

void ss_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, &hi0, &lo0);
        mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* natural log for x in (0, 1] (any positive normal double works):
 * x = m * 2^e, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); ln m = 2 s sum_k s^(2k)/(2k+1) */
double ss_dlog(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8);
    if (m > 0x1.6a09e667f3bcdp+0) {
        m = m * 0.5;
        e += 1;
    }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    static const double c[12] = {
        0x1.0000000000000p+0, 0x1.5555555555555p-2, 0x1.999999999999ap-3, 0x1.2492492492492p-3,
        0x1.c71c71c71c71cp-4, 0x1.745d1745d1746p-4, 0x1.3b13b13b13b14p-4, 0x1.1111111111111p-4,
        0x1.e1e1e1e1e1e1ep-5, 0x1.af286bca1af28p-5, 0x1.8618618618618p-5, 0x1.642c8590b2164p-5};
    double p = c[11];
    for (int k = 10; k >= 0; --k) {
        p = p * z;
        p = p + c[k];
    }
    double lnm = 2.0 * s;
    lnm = lnm * p;
    double el = (double)e * 0x1.62e42fefa39efp-1;
    return el + lnm;
}

/* sin and cos of 2*pi*u, u in [0,1): quadrant k = floor(4u), t = (4u-k)*pi/2, Taylor series */
void ss_dsincos2pi(double u, double* s_out, double* c_out) {
    double q = u * 4.0; /* exact */
    int k = (int)q;     /* truncation == floor, q >= 0 */
    double f = q - (double)k; /* exact */
    double t = f * 0x1.921fb54442d18p+0;
    double t2 = t * t;
    static const double sc[13] = {
        0x1.0000000000000p+0,   -0x1.5555555555555p-3,  0x1.1111111111111p-7,
        -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19,  -0x1.ae64567f544e4p-26,
        0x1.6124613a86d09p-33,  -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49,
        -0x1.2f49b46814157p-57, 0x1.71b8ef6dcf572p-66,  -0x1.761b41316381ap-75,
        0x1.3f3ccdd165fa9p-84};
    static const double cc[13] = {
        0x1.0000000000000p+0,   -0x1.0000000000000p-1,  0x1.5555555555555p-5,
        -0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-16,  -0x1.27e4fb7789f5cp-22,
        0x1.1eed8eff8d898p-29,  -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45,
        -0x1.6827863b97d97p-53, 0x1.e542ba4020225p-62,  -0x1.0ce396db7f853p-70,
        0x1.f2cf01972f578p-80};
    double ps = sc[12], pc = cc[12];
    for (int i = 11; i >= 0; --i) {
        ps = ps * t2;
        ps = ps + sc[i];
        pc = pc * t2;
        pc = pc + cc[i];
    }
    double st = ps * t, ct = pc;
    switch (k & 3) {
        case 0: *s_out = st;  *c_out = ct;  break;
        case 1: *s_out = ct;  *c_out = -st; break;
        case 2: *s_out = -st; *c_out = -ct; break;
        default: *s_out = -ct; *c_out = st; break;
    }
}

static inline double u53(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | (uint64_t)lo) >> 11;
    return (double)v * 0x1p-53;
}

void ss_motion_normals(uint64_t seed, uint64_t step, uint64_t particle, double* z1, double* z2) {
    uint32_t ctr[4] = {(uint32_t)particle, (uint32_t)step, (uint32_t)(step >> 32), 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t x[4];
    ss_philox4x32_10(ctr, key, x);
    uint64_t v1 = ((((uint64_t)x[0] << 32) | (uint64_t)x[1]) >> 11) + 1;
    double u1 = (double)v1 * 0x1p-53;
    double u2 = u53(x[2], x[3]);
    double r = sqrt(-2.0 * ss_dlog(u1));
    double s, c;
    ss_dsincos2pi(u2, &s, &c);
    *z1 = r * c;
    *z2 = r * s;
}

void ss_fill_motion_normals(uint64_t seed, uint64_t step, uint64_t first, uint64_t count, double* z) {
    for (uint64_t i = 0; i < count; ++i) ss_motion_normals(seed, step, first + i, &z[2 * i], &z[2 * i + 1]);
}

double ss_resample_uniform(uint64_t seed, uint64_t step) {
    uint32_t ctr[4] = {0u, (uint32_t)step, (uint32_t)(step >> 32), 1u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t x[4];
    ss_philox4x32_10(ctr, key, x);
    return u53(x[0], x[1]);
}

/* start pose of `particle` for a uniform initialisation (counter domains 2 and 3 of the stream) */
void ss_uniform_pose(uint64_t seed, uint64_t particle, double x0, double y0, double x1, double y1, float* out_xyt) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ca[4] = {(uint32_t)particle, 0u, 0u, 2u}, cb[4] = {(uint32_t)particle, 0u, 0u, 3u};
    uint32_t a[4], b[4];
    ss_philox4x32_10(ca, key, a);
    ss_philox4x32_10(cb, key, b);
    double d = x1 + -x0;
    out_xyt[0] = (float)(x0 + u53(a[0], a[1]) * d);
    d = y1 + -y0;
    out_xyt[1] = (float)(y0 + u53(a[2], a[3]) * d);
    out_xyt[2] = (float)(-0x1.921fb54442d18p+1 + u53(b[0], b[1]) * 0x1.921fb54442d18p+2);
}
void ss_fill_uniform_poses(uint64_t seed, uint64_t first, uint64_t count, double x0, double y0, double x1, double y1,
                           float* out_xyt) {
    for (uint64_t i = 0; i < count; ++i) ss_uniform_pose(seed, first + i, x0, y0, x1, y1, out_xyt + 3 * i);
}
