// pt_sincos.cuh -- sinf/cosf that round exactly like the host libm the
// reference calls.
//
// ConcentricSampleDisk (core/sampling.cpp:113-130) is the only place on the
// hot path where a transcendental decides a direction on the device; the
// reference reaches glibc's sinf/cosf through std::sin/std::cos.  glibc >= 2.28
// evaluates both in double precision (the "optimized routines" algorithm: one
// of two degree-7/8 minimax polynomials after a fast pi/2 reduction, result
// rounded once to float).  Restating that algorithm in fp64 on the device makes
// every sampled direction bit-identical, so no discrete decision of a path
// (hit/miss, Russian roulette, light choice) can flip against the reference.
// The coefficients are the published ones of that algorithm; parity with the
// host's libm is pinned exhaustively over |x| < 120 by tests/host_preflight.cpp
// (the hot path only ever passes |x| <= 3*pi/4).
#ifndef B200PT_SINCOS_CUH
#define B200PT_SINCOS_CUH

#include "pt_platform.h"

namespace B200PT_NS {

struct SinCosPoly {
    double sign[4];
    double hpi_inv;  // 2/pi * 2^24
    double hpi;      // pi/2
    double c0, c1, c2, c3, c4;
    double s1, s2, s3;
};

B200_HD const SinCosPoly sincos_table(int which) {
    // table[0] for quadrants whose sine keeps its sign, table[1] for the negated ones
    SinCosPoly p;
    const double sgn = which ? -1.0 : 1.0;
    p.sign[0] = 1.0;
    p.sign[1] = -1.0;
    p.sign[2] = -1.0;
    p.sign[3] = 1.0;
    p.hpi_inv = 0x1.45F306DC9C883p+23;
    p.hpi = 0x1.921FB54442D18p0;
    p.c0 = sgn * 0x1p0;
    p.c1 = sgn * -0x1.ffffffd0c621cp-2;
    p.c2 = sgn * 0x1.55553e1068f19p-5;
    p.c3 = sgn * -0x1.6c087e89a359dp-10;
    p.c4 = sgn * 0x1.99343027bf8c3p-16;
    p.s1 = -0x1.555545995a603p-3;
    p.s2 = 0x1.1107605230bc4p-7;
    p.s3 = -0x1.994eb3774cf24p-13;
    return p;
}

B200_HD unsigned abstop12(float x) { return (float_as_uint(x) >> 20) & 0x7ff; }

// sine (n even) or cosine (n odd) polynomial in double, rounded once to float
B200_HD float sincos_poly(double x, double x2, const SinCosPoly &p, int n) {
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = p.s2 + x2 * p.s3;
        double x7 = x3 * x2;
        double s = x + x3 * p.s1;
        return (float)(s + x7 * s1);
    } else {
        double x4 = x2 * x2;
        double c2 = p.c3 + x2 * p.c4;
        double c1 = p.c0 + x2 * p.c1;
        double x6 = x4 * x2;
        double c = c1 + x4 * p.c2;
        return (float)(c + x6 * c2);
    }
}

B200_HD double sincos_reduce_fast(double x, const SinCosPoly &p, int *np) {
    double r = x * p.hpi_inv;
    int n = ((int)r + 0x800000) >> 24;
    *np = n;
    return x - n * p.hpi;
}

// valid for |y| < 120 (larger arguments never occur on this path)
B200_HD float pt_sinf(float y) {
    double x = y;
    SinCosPoly p = sincos_table(0);
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {  // |y| < pi/4
        double s = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sincos_poly(x, s, p, 0);
    }
    int n;
    x = sincos_reduce_fast(x, p, &n);
    double s = p.sign[n & 3];
    if (n & 2) p = sincos_table(1);
    return sincos_poly(x * s, x * x, p, n);
}

B200_HD float pt_cosf(float y) {
    double x = y;
    SinCosPoly p = sincos_table(0);
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sincos_poly(x, x2, p, 1);
    }
    int n;
    x = sincos_reduce_fast(x, p, &n);
    double s = p.sign[n & 3];
    if (n & 2) p = sincos_table(1);
    return sincos_poly(x * s, x * x, p, n ^ 1);
}

}  // namespace B200PT_NS
#endif
