// pt_explog.cuh -- expf / logf that round exactly like the host libm the reference calls.
//
// The next widening of the path, VolPathIntegrator with homogeneous media (last item), samples
// free-flight distances with std::log (media/homogeneous.cpp:57-58) and evaluates transmittances with std::exp
// (Exp(Spectrum), core/spectrum.h:269-275; homogeneous.cpp:47-49, :69-70); both decide discrete events of a path
// (medium or surface interaction, Russian roulette via beta), so they must match the host bit for bit, like sinf / cosf
// in pt_sincos.cuh.  glibc >= 2.27 computes both in double precision with the "optimized routines" algorithms:
//   expf: x*N/ln2 = k + r (N = 32), 2^(k/N) from a 32-entry table, degree-3 polynomial in r, one rounding to float;
//   logf: z = x/c_i with c_i from a 16-entry table chosen by the top mantissa bits, degree-3 polynomial in r = z - 1,
//         result k*ln2 + log(c_i) + log1p(r), one rounding to float.
// The tables are those algorithms' published constants (read from the container's libm.so.6, glibc 2.39, where they are
// the objects __exp2f_data and __logf_data).  x86-64 glibc dispatches to a build of the same C code compiled with
// -mfma, so the polynomial steps are fused multiply-adds; the fma() calls below mirror that contraction.
// tests/libm_pin.cpp pins both functions against the host's std::exp / std::log for EVERY float input.
// Nothing on the current hot path calls them yet.
#ifndef B200PT_EXPLOG_CUH
#define B200PT_EXPLOG_CUH

#include "pt_platform.h"

namespace B200PT_NS {

B200_HD double pt_fma(double a, double b, double c) { return fma(a, b, c); }
B200_HD uint64_t double_as_u64(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
B200_HD double u64_as_double(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

B200_HD uint64_t exp2f_tab(int i) {
    // 2^(i/32) with the exponent bits of i/32's integer part removed (so that adding k << 47 gives 2^(k/32))
    const uint64_t T[32] = {
        0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
        0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
        0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
        0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
        0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
        0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
        0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
        0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
    return T[i];
}

// glibc sysdeps/ieee754/flt-32/e_expf.c (2.27+), non-TOINT_INTRINSICS path
B200_HD float pt_expf(float x) {
    const uint32_t ix = float_as_uint(x);
    const uint32_t abstop = (ix >> 20) & 0x7ffu;
    if (abstop >= 0x42bu) {  // |x| >= 88 or NaN / inf  (top12(88.0f) = 0x42b)
        if (ix == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8u) return x + x;
        if (x > 0x1.62e42ep6f) return uint_as_float(0x7f800000u);  // overflow: 0x1p97f * 0x1p97f
        if (x < -0x1.9fe368p6f) return 0.0f;                        // underflow: 0x1p-95f * 0x1p-95f
    }
    const double xd = (double)x;
    double z = 0x1.71547652b82fep+5 * xd;  // InvLn2 * N
    double kd = z + 0x1.8p+52;             // round to nearest integer in the low mantissa bits
    const uint64_t ki = double_as_u64(kd);
    kd -= 0x1.8p+52;
    const double r = pt_fma(0x1.71547652b82fep+5, xd, -kd);  // z - kd with z's product fused into the subtraction
    uint64_t t = exp2f_tab((int)(ki % 32));
    t += ki << (52 - 5);
    const double s = u64_as_double(t);
    z = pt_fma(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);
    const double r2 = r * r;
    double y = pt_fma(0x1.62e42ff0c52d6p-6, r, 1.0);
    y = pt_fma(z, r2, y);
    y = y * s;
    return (float)y;
}

// glibc sysdeps/ieee754/flt-32/e_logf.c (2.27+)
B200_HD float pt_logf(float x) {
    const double INVC[16] = {0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0,  0x1.3c995b0b80385p+0,
                             0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0,  0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
                             0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
                             0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
    const double LOGC[16] = {-0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
                             -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c81p-3,   -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
                             -0x1.252f438e10c1ep-5, 0x0p+0,                0x1.aa5aa5df25984p-5,  0x1.c5e53aa362eb4p-4,
                             0x1.526e57720db08p-3,  0x1.bc2860d22477p-3,   0x1.1058bc8a07ee1p-2,  0x1.4043057b6ee09p-2};
    uint32_t ix = float_as_uint(x);
    if (ix == 0x3f800000u) return 0.f;  // log(1) is +0 in every rounding mode
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        // x < 0x1p-126 or inf or nan
        if (ix * 2 == 0) return -uint_as_float(0x7f800000u);  // log(+-0) = -inf (divide-by-zero)
        if (ix == 0x7f800000u) return x;                      // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return (x - x) / (x - x);  // log(negative) and NaN: NaN
        // subnormal: normalise
        ix = float_as_uint(x * 0x1p23f);
        ix -= 23u << 23;
    }
    // x = 2^k z; where z is in range [OFF, 2*OFF] and exact; the range is split into 16 subintervals
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> (23 - 4)) % 16);
    const int k = (int32_t)tmp >> 23;  // arithmetic shift
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    const double z = (double)uint_as_float(iz);
    // log(x) = log1p(z/c - 1) + log(c) + k*Ln2
    const double r = pt_fma(z, INVC[i], -1.0);
    const double y0 = pt_fma((double)k, 0x1.62e42fefa39efp-1, LOGC[i]);
    // pipelined polynomial evaluation to approximate log1p(r)
    const double r2 = r * r;
    double y = pt_fma(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = pt_fma(-0x1.00ea348b88334p-2, r2, y);
    y = pt_fma(y, r2, y0 + r);
    return (float)y;
}

}  // namespace B200PT_NS
#endif
