/* oracle/crmath_redirect.h -- TEST INFRASTRUCTURE.
 * Force-included when building the "oracle-B" variant (oracle/_ref/libppe_oracle_cr.so): the same
 * restatement (ppe_oracle.c + dubins.c) with the four transcendental functions of the Dubins
 * solver / sampler redirected to the correctly rounded implementations that the CUDA engine uses
 * (path_planner_b200/csrc/ppe_crmath.cuh, host build in oracle/crmath_host.cpp).
 *
 * oracle-A (glibc) is bit-identical to the compiled reference; oracle-B is bit-identical to the
 * GPU's per-edge arithmetic.  A and B differ only where glibc's own result is not the correctly
 * rounded one (measured: sin/cos 0.14 %, atan2 0.04 %, acos 0.07 % of calls, always by 1 ulp) AND
 * that last bit decides a branch; tests/test_oracle_variants.py measures that edge fraction. */
#ifndef PPE_ORACLE_CRMATH_REDIRECT_H
#define PPE_ORACLE_CRMATH_REDIRECT_H
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
double ppe_cr_sin(double x);
double ppe_cr_cos(double x);
double ppe_cr_atan2(double y, double x);
double ppe_cr_acos(double x);
#ifdef __cplusplus
}
#endif
#define sin(x) ppe_cr_sin(x)
#define cos(x) ppe_cr_cos(x)
#define atan2(y, x) ppe_cr_atan2(y, x)
#define acos(x) ppe_cr_acos(x)
#endif
