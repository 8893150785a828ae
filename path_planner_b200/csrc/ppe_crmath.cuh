// ppe_crmath.cuh -- high-accuracy sin / cos / atan2 / acos for the per-edge Dubins solve.
//
// Why: the discrete outcomes of the reference (which of the six words wins, whether the end-state
// sample takes the `distance - 1e-5` retry, DubinsWrapper.cpp:39-42) hinge on the LAST BIT of the
// transcendental results on near-degenerate inputs (straight-ahead edges along a ribbon give
// alpha ~ 0 or ~ 2*pi depending on one ulp of atan2).  The reference runs on glibc's libm, whose
// sin/cos/atan2/acos are correctly rounded in all but a vanishing fraction of cases (documented
// max error 0.52-0.55 ulp).  CUDA's libdevice versions are 1-2 ulp functions and disagree with
// glibc often enough to flip ~0.3 % of edges.  These replacements evaluate in double-double
// (~100 bits) and round once, i.e. they return the correctly rounded double except when the true
// value lies within ~2^-45 ulp of a rounding boundary -- so they agree with glibc wherever glibc
// itself is correctly rounded.  They are used once per edge (solve + segment constants + end
// state); the per-sample positions use the fast libdevice sincos (1e-9 tolerance class).
//
// __host__ __device__: the host build (tests/host_helpers.cpp) is checked against glibc and mpmath.
#pragma once

#include <math.h>

#include "ppe_crmath_tables.h"
#include "ppe_math_base.cuh"

namespace ppe {

struct dd {
    double hi, lo;
};

// coefficient / lookup tables: one copy for device code, one for host code
#if defined(__CUDACC__)
#define PPE_TABLE_DECL(name, dims, init) \
    static __constant__ double d_##name dims = init; \
    static const double h_##name dims = init;
#else
#define PPE_TABLE_DECL(name, dims, init) static const double h_##name dims = init;
#endif
// The double-double Horner steps are ~40 instructions each.  Unrolled, the three polynomial kernels are tens of KB of
// straight-line code that every warp streams through the instruction cache once per call (ncu: no_instruction was the top
// stall of k2a_prepare); rolled, they are a few hundred bytes with the coefficients in constant memory.  Same operations,
// same order, same bits.
#if defined(__CUDACC__)
#define PPE_ROLLED _Pragma("unroll 1")
#else
#define PPE_ROLLED
#endif
#if defined(__CUDA_ARCH__)
#define PPE_TABLE(name) d_##name
#else
#define PPE_TABLE(name) h_##name
#endif
PPE_TABLE_DECL(sin_coeffs, [crtab::kSinCosTerms][2], PPE_SIN_COEFFS)
PPE_TABLE_DECL(cos_coeffs, [crtab::kSinCosTerms][2], PPE_COS_COEFFS)
PPE_TABLE_DECL(atan_table, [65][2], PPE_ATAN_TABLE)
PPE_TABLE_DECL(atan_coeffs, [crtab::kAtanTerms][2], PPE_ATAN_COEFFS)

PPE_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

PPE_HD dd two_sum(double a, double b) {
    const double s = a + b;
    const double bb = s - a;
    const double e = (a - (s - bb)) + (b - bb);
    return dd{s, e};
}
PPE_HD dd fast_two_sum(double a, double b) { // |a| >= |b|
    const double s = a + b;
    return dd{s, b - (s - a)};
}
PPE_HD dd two_prod(double a, double b) {
    const double p = a * b;
    return dd{p, fma_(a, b, -p)};
}
PPE_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = fast_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return fast_two_sum(s.hi, s.lo);
}
PPE_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo += a.lo;
    return fast_two_sum(s.hi, s.lo);
}
PPE_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
PPE_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
PPE_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return fast_two_sum(p.hi, p.lo);
}
PPE_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return fast_two_sum(p.hi, p.lo);
}
PPE_HD dd dd_div(dd a, dd b) {
    const double q1 = a.hi / b.hi;
    dd r = dd_sub(a, dd_mul_d(b, q1));
    const double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    const double q3 = r.hi / b.hi;
    dd q = fast_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
PPE_HD dd dd_sqrt(dd a) {
    if (a.hi <= 0) return dd{0.0, 0.0};
    const double x = sqrt(a.hi);
    // one Newton step in dd: x + (a - x^2) / (2x)
    const dd x2 = two_prod(x, x);
    const dd r = dd_sub(a, x2);
    const double c = r.hi / (2.0 * x);
    dd s = fast_two_sum(x, c);
    // second correction for full dd accuracy
    const dd s2 = dd_mul(s, s);
    const dd r2 = dd_sub(a, s2);
    return dd_add_d(s, r2.hi / (2.0 * s.hi));
}

// ---- sin / cos ---------------------------------------------------------------------------------------
// r = x - k*pi/2 as a double-double, |r| <= pi/4 (+ slack); valid for |x| < ~1e5
PPE_HD dd reduce_pio2(double x, int* quadrant) {
    using namespace crtab;
    const double kd = rint(x * kTwoOverPi);
    *quadrant = ((int)kd) & 3;
    // kd has <= 17 bits here; kd * (33-bit chunk) is exact
    dd r = two_sum(x, -kd * kPio2_1);
    r = dd_add(r, two_prod(-kd, kPio2_2));
    r = dd_add(r, two_prod(-kd, kPio2_3));
    r = dd_add(r, two_prod(-kd, kPio2_4h));
    r = dd_add_d(r, -kd * kPio2_4l);
    return r;
}

PPE_HD dd sin_kernel(dd r) {
    const double (*C)[2] = PPE_TABLE(sin_coeffs);
    const dd z = dd_mul(r, r);
    dd p = dd{C[crtab::kSinCosTerms - 1][0], C[crtab::kSinCosTerms - 1][1]};
    PPE_ROLLED
    for (int k = crtab::kSinCosTerms - 2; k >= 0; k--) p = dd_add(dd_mul(p, z), dd{C[k][0], C[k][1]});
    // sin r = r + r * z * p
    return dd_add(r, dd_mul(dd_mul(r, z), p));
}
PPE_HD dd cos_kernel(dd r) {
    const double (*C)[2] = PPE_TABLE(cos_coeffs);
    const dd z = dd_mul(r, r);
    dd p = dd{C[crtab::kSinCosTerms - 1][0], C[crtab::kSinCosTerms - 1][1]};
    PPE_ROLLED
    for (int k = crtab::kSinCosTerms - 2; k >= 0; k--) p = dd_add(dd_mul(p, z), dd{C[k][0], C[k][1]});
    // cos r = 1 + z * p
    return dd_add_d(dd_mul(z, p), 1.0);
}

PPE_HD_NOINLINE void cr_sincos(double x, double* s, double* c) {
    if (!(fabs(x) < 1e5)) { // outside the validated range: defer to the platform libm
        sincos_f64(x, s, c);
        return;
    }
    if (x == 0.0) { *s = x; *c = 1.0; return; }
    int q;
    const dd r = reduce_pio2(x, &q);
    const dd sr = sin_kernel(r);
    const dd cr = cos_kernel(r);
    double sv, cv;
    switch (q) {
        case 0: sv = sr.hi; cv = cr.hi; break;
        case 1: sv = cr.hi; cv = -sr.hi; break;
        case 2: sv = -sr.hi; cv = -cr.hi; break;
        default: sv = -cr.hi; cv = sr.hi; break;
    }
    *s = sv;
    *c = cv;
}
PPE_HD double cr_sin(double x) { double s, c; cr_sincos(x, &s, &c); return s; }
PPE_HD double cr_cos(double x) { double s, c; cr_sincos(x, &s, &c); return c; }

// ---- atan / atan2 / acos --------------------------------------------------------------------------------
// atan(q) for a double-double 0 <= q <= 1
PPE_HD dd atan_dd_unit(dd q) {
    const double (*T)[2] = PPE_TABLE(atan_table);
    const double (*C)[2] = PPE_TABLE(atan_coeffs);
    int i = (int)rint(q.hi * 64.0);
    if (i < 0) i = 0;
    if (i > 64) i = 64;
    const double c = (double)i * (1.0 / 64.0);
    // t = (q - c) / (1 + q c), |t| <= 2^-7
    const dd num = dd_add_d(q, -c);
    const dd den = dd_add_d(dd_mul_d(q, c), 1.0);
    const dd t = dd_div(num, den);
    const dd z = dd_mul(t, t);
    dd p = dd{C[crtab::kAtanTerms - 1][0], C[crtab::kAtanTerms - 1][1]};
    PPE_ROLLED
    for (int k = crtab::kAtanTerms - 2; k >= 0; k--) p = dd_add(dd_mul(p, z), dd{C[k][0], C[k][1]});
    const dd at = dd_add(t, dd_mul(dd_mul(t, z), p));
    return dd_add(dd{T[i][0], T[i][1]}, at);
}

// atan2 of double-double arguments, result rounded to double.  Zero / sign conventions of C99.
PPE_HD_NOINLINE double cr_atan2_dd(dd y, dd x) {
    using namespace crtab;
    const bool yneg = y.hi < 0 || (y.hi == 0 && signbit(y.hi));
    const bool xneg = x.hi < 0 || (x.hi == 0 && signbit(x.hi));
    if (y.hi != y.hi || x.hi != x.hi) return y.hi + x.hi;
    if (y.hi == 0) {
        const double r = xneg ? kPi_h : 0.0;
        return yneg ? -r : r;
    }
    if (x.hi == 0) return yneg ? -kPio2_h : kPio2_h;
    if (isinf(x.hi) || isinf(y.hi)) return atan2(y.hi, x.hi);
    const dd ay = yneg ? dd_neg(y) : y;
    const dd ax = xneg ? dd_neg(x) : x;
    dd a;
    const bool swap = ay.hi > ax.hi || (ay.hi == ax.hi && ay.lo > ax.lo);
    if (!swap) {
        a = atan_dd_unit(dd_div(ay, ax));
    } else {
        const dd t = atan_dd_unit(dd_div(ax, ay));
        // pi/2 - t with a three-part pi/2
        a = dd_add(dd_sub(dd{kPio2_h, kPio2_m}, t), dd{kPio2_l, 0.0});
    }
    if (xneg) a = dd_add(dd_sub(dd{kPi_h, kPi_m}, a), dd{kPi_l, 0.0});
    const double r = a.hi; // a is normalised: hi is the nearest double of hi + lo
    return yneg ? -r : r;
}
PPE_HD double cr_atan2(double y, double x) { return cr_atan2_dd(dd{y, 0.0}, dd{x, 0.0}); }

// acos(x) = atan2(sqrt((1 - x)(1 + x)), x)
PPE_HD_NOINLINE double cr_acos(double x) {
    if (!(fabs(x) <= 1.0)) return acos(x); // NaN for out-of-domain, as libm
    if (x == 1.0) return 0.0;
    const dd a = two_sum(1.0, -x);
    const dd b = two_sum(1.0, x);
    const dd s = dd_sqrt(dd_mul(a, b));
    return cr_atan2_dd(s, dd{x, 0.0});
}

} // namespace ppe
