/*
 * oracle/dubins.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * CPU restatement of the public C API of the external `dubins_curves` catkin
 * package that afb2001/path_planner links against but does not vendor
 * (find_package: path_planner_common/CMakeLists.txt:16, path_planner/CMakeLists.txt:18;
 * <depend>dubins_curves</depend>: path_planner_common/package.xml:31).  No version is
 * pinned anywhere in the reference.  The package wraps the widely used MIT-licensed
 * "Dubins-Curves" C library (A. Walker), v1.0-style API (DubinsPathType enum,
 * dubins_shortest_path / dubins_path_sample / dubins_extract_subpath); this file and
 * dubins.c restate that library's published algorithm (Shkel & Lumelsky closed forms).
 *
 * What the reference requires of this header (call sites):
 *   DubinsWrapper.cpp:13   dubins_shortest_path(&path, q1, q2, rho)
 *   DubinsWrapper.cpp:21   dubins_path_length(const DubinsPath*)
 *   DubinsWrapper.cpp:38   dubins_path_sample(const DubinsPath*, t, q)  (+ EDUBPARAM retry :39-42)
 *   DubinsWrapper.cpp:114  dubins_extract_subpath(&copy, d, &path)
 *   RibbonManager.h:214-215, NodeBase.h:206-212 (field access qi/param/rho/type)
 *   path_planner_common/msg/DubinsPath.msg:17  enum order LSL=0 LSR=1 RSL=2 RSR=3 RLR=4 LRL=5
 *
 * PARITY NOTE: word choice / tie-breaking of the real upstream library is "parity
 * unpinned" by the reference's own tests (only straight-line and half-circle known
 * answers exist); this file is therefore the oracle's
 * *definition* of the dependency.
 */
#ifndef PPE_ORACLE_DUBINS_H
#define PPE_ORACLE_DUBINS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    LSL = 0,
    LSR = 1,
    RSL = 2,
    RSR = 3,
    RLR = 4,
    LRL = 5
} DubinsPathType;

typedef struct {
    double qi[3];        /* initial configuration (x, y, yaw) */
    double param[3];     /* lengths of the three segments, in units of rho */
    double rho;          /* turning radius */
    DubinsPathType type; /* which of the six words */
} DubinsPath;

#define EDUBOK        (0) /* no error */
#define EDUBCOCONFIGS (1) /* colocated configurations */
#define EDUBPARAM     (2) /* path parameterisation error */
#define EDUBBADRHO    (3) /* rho is invalid */
#define EDUBNOPATH    (4) /* no connection between configurations with this word */

typedef int (*DubinsPathSamplingCallback)(double q[3], double t, void* user_data);

int dubins_shortest_path(DubinsPath* path, double q0[3], double q1[3], double rho);
int dubins_path(DubinsPath* path, double q0[3], double q1[3], double rho, DubinsPathType pathType);
double dubins_path_length(const DubinsPath* path);
double dubins_segment_length(const DubinsPath* path, int i);
double dubins_segment_length_normalized(const DubinsPath* path, int i);
DubinsPathType dubins_path_type(const DubinsPath* path);
int dubins_path_sample(const DubinsPath* path, double t, double q[3]);
int dubins_path_sample_many(const DubinsPath* path, double stepSize, DubinsPathSamplingCallback cb, void* user_data);
int dubins_path_endpoint(const DubinsPath* path, double q[3]);
int dubins_extract_subpath(const DubinsPath* path, double t, DubinsPath* newpath);

#ifdef __cplusplus
}
#endif

#endif /* PPE_ORACLE_DUBINS_H */
