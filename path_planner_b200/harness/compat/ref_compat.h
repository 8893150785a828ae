// compat/ref_compat.h -- toolchain fixes for compiling the reference's translation units in this image (g++ 13); used by
// the harness build (path_planner_b200/harness/Makefile) and by the oracle build (oracle/Makefile).
//
// Force-included (-include) in front of every reference translation unit when the reference is
// compiled *where it lies* under /root/reference (no source is copied).  Two fixes, both needed
// only because of the toolchain in this image (g++ 13):
//
//  1. `State::setYaw` is declared `double` but has no return statement
//     (path_planner_common/include/path_planner_common/State.h:62-65).  g++ >= 8 treats flowing
//     off the end as unreachable, so the unpatched function crashes (SIGILL at -O0, falls
//     through into the next function at -O2); it is called for every sample point
//     (path_planner_common/src/dubinsPlan/DubinsWrapper.cpp:47).  The macro below turns the
//     in-class definition
//         double setYaw(double yaw1) { ... }
//     into
//         double setYaw_never_defined_(double yaw1); void setYaw_void_(double yaw1) { ... }
//     i.e. the same body with a `void` return type, and then renames every later use of
//     `setYaw` to the void version.  The body and its arithmetic are untouched.
//
//  2. DynamicObstaclesManager1.h:38 uses uint32_t without including <cstdint>.
#ifndef PPE_ORACLE_REF_PREFIX_H
#define PPE_ORACLE_REF_PREFIX_H

#include <cstdint>
#include <stdexcept>
#include <cassert>
#include <memory>
#include <list>

#define setYaw(arg) setYaw_never_defined_(arg); void setYaw_void_(arg)
#include <path_planner_common/State.h>
#undef setYaw
#define setYaw setYaw_void_

#endif
