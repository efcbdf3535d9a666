// Exhaustive check: slamrs_libm::{sinf,cosf}_exact vs the host libm, all 2^32 bit patterns.
// Build: g++ -O2 -ffp-contract=off -fopenmp -mfma tools/check_libm_f32.cpp -o /tmp/check_libm -lm
// (-mfma only so that fma() is one instruction; contraction of a*b+c stays off.)
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include "../slamrs_b200/csrc/libm_f32.cuh"

int main(int argc, char** argv) {
    uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    uint64_t bad_s = 0, bad_c = 0, total = 0;
#pragma omp parallel for reduction(+ : bad_s, bad_c, total) schedule(static)
    for (uint64_t i = 0; i < (1ULL << 32); i += stride) {
        uint32_t u = (uint32_t)i;
        float x;
        memcpy(&x, &u, 4);
        float a = sinf(x), b = slamrs_libm::sinf_exact(x);
        float c = cosf(x), d = slamrs_libm::cosf_exact(x);
        float fs, fc;
        slamrs_libm::sincosf_exact(x, &fs, &fc);   // the fused form must reproduce the two separate calls
        if (memcmp(&fs, &b, 4) != 0 && !((fs != fs) && (b != b))) { if (bad_s < 5) printf("fused sin mismatch x=%a\n", x); bad_s++; }
        if (memcmp(&fc, &d, 4) != 0 && !((fc != fc) && (d != d))) { if (bad_c < 5) printf("fused cos mismatch x=%a\n", x); bad_c++; }
        uint32_t ua, ub, uc, ud;
        memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); memcpy(&uc, &c, 4); memcpy(&ud, &d, 4);
        bool nan_ok_s = (a != a) && (b != b);
        bool nan_ok_c = (c != c) && (d != d);
        if (ua != ub && !nan_ok_s) { if (bad_s < 5) printf("sin mismatch x=%a libm=%a mine=%a\n", x, a, b); bad_s++; }
        if (uc != ud && !nan_ok_c) { if (bad_c < 5) printf("cos mismatch x=%a libm=%a mine=%a\n", x, c, d); bad_c++; }
        total++;
    }
    printf("checked %llu inputs: sin mismatches %llu, cos mismatches %llu\n",
           (unsigned long long)total, (unsigned long long)bad_s, (unsigned long long)bad_c);
    return (bad_s || bad_c) ? 1 : 0;
}
