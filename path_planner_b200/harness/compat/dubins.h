/*
 * compat/dubins.h -- the C API of the external `dubins_curves` catkin package that afb2001/path_planner links
 * (find_package: path_planner_common/CMakeLists.txt:16; include sites DubinsWrapper.h:7-9, Edge.h:12-14,
 * RibbonManager.h:8-10) but does not vendor.  Declarations only, as the reference's call sites require them
 * (DubinsWrapper.cpp:13,21,38,114; field access NodeBase.h:206-212; enum order DubinsPath.msg:17).  The harness links
 * compat/dubins_host.cpp behind it -- the engine's own shared-source Dubins arithmetic built for the host -- so that
 * the planner's remaining host-side uses (plan tracing, DubinsPlan sampling by callers) agree with the device path.
 * A deployment that has the real package installs its header and library instead.
 */
#ifndef PPE_COMPAT_DUBINS_H
#define PPE_COMPAT_DUBINS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { LSL = 0, LSR = 1, RSL = 2, RSR = 3, RLR = 4, LRL = 5 } DubinsPathType;

typedef struct {
    double qi[3];        /* initial configuration (x, y, yaw) */
    double param[3];     /* lengths of the three segments, in units of rho */
    double rho;          /* turning radius */
    DubinsPathType type; /* which of the six words */
} DubinsPath;

#define EDUBOK        (0)
#define EDUBCOCONFIGS (1)
#define EDUBPARAM     (2)
#define EDUBBADRHO    (3)
#define EDUBNOPATH    (4)

int dubins_shortest_path(DubinsPath* path, double q0[3], double q1[3], double rho);
double dubins_path_length(const DubinsPath* path);
int dubins_path_sample(const DubinsPath* path, double t, double q[3]);
int dubins_extract_subpath(const DubinsPath* path, double t, DubinsPath* newpath);

#ifdef __cplusplus
}
#endif

#endif
