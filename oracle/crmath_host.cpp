// oracle/crmath_host.cpp -- TEST INFRASTRUCTURE: host build of the engine's correctly rounded
// sin / cos / atan2 / acos (path_planner_b200/csrc/ppe_crmath.cuh) behind C names, for oracle-B.
#include "ppe_crmath.cuh"

extern "C" {
double ppe_cr_sin(double x) { return ppe::cr_sin(x); }
double ppe_cr_cos(double x) { return ppe::cr_cos(x); }
double ppe_cr_atan2(double y, double x) { return ppe::cr_atan2(y, x); }
double ppe_cr_acos(double x) { return ppe::cr_acos(x); }
}
