// ppe_math_base.cuh -- macros, constants and bit helpers shared by ppe_math.cuh / ppe_crmath.cuh.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PPE_HD __host__ __device__ __forceinline__
#define PPE_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define PPE_HD inline
#define PPE_HD_NOINLINE inline
#endif

namespace ppe {

constexpr double kPi = 3.14159265358979323846;       // M_PI
constexpr double kPi2 = 1.57079632679489661923;      // M_PI_2
constexpr double kTwoPi = 2 * 3.14159265358979323846; // 2 * M_PI (exact doubling)

// ---- bit helpers -----------------------------------------------------------------------------
PPE_HD int64_t f64_bits(double x) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t b;
    memcpy(&b, &x, sizeof b);
    return b;
#endif
}
PPE_HD double bits_f64(int64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double x;
    memcpy(&x, &b, sizeof x);
    return x;
#endif
}
// unbiased exponent of a positive normal double
PPE_HD int f64_exponent(double x) { return (int)((f64_bits(x) >> 52) & 0x7ff) - 1023; }
// 2^e for -1022 <= e <= 1023
PPE_HD double f64_pow2(int e) { return bits_f64((int64_t)(e + 1023) << 52); }

PPE_HD void sincos_f64(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    sincos(x, s, c);
#else
    *s = sin(x);
    *c = cos(x);
#endif
}

// dubins.c: fmodr(theta, 2*M_PI)
PPE_HD double mod2pi(double theta) { return theta - kTwoPi * floor(theta / kTwoPi); }

// State::yaw(), State.h:51-55 (also the inverse map State::setYaw, State.h:62-65: the same formula)
PPE_HD double heading_to_yaw(double heading) {
    double h = kPi2 - heading;
    if (h < 0) h += kTwoPi;
    return h;
}

} // namespace ppe
