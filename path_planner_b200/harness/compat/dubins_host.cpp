// compat/dubins_host.cpp -- the four dubins.h entry points the reference's host code calls, on top of the engine's
// shared-source scalar arithmetic (path_planner_b200/csrc/ppe_math.cuh is __host__ __device__): the same six-word solver
// and segment sampler the kernels run, compiled for the host with -ffp-contract=off.  Used by the standalone harness
// for what stays on the host (Planner::tracePlan, DubinsPlan::sample, Edge::setEnd(wrapper), the Dubins TSP heuristics).
#include "dubins.h"

#include "ppe_math.cuh"

using namespace ppe;

static DubinsPathD to_d(const DubinsPath* p) {
    DubinsPathD d;
    d.qi[0] = p->qi[0]; d.qi[1] = p->qi[1]; d.qi[2] = p->qi[2];
    d.param[0] = p->param[0]; d.param[1] = p->param[1]; d.param[2] = p->param[2];
    d.rho = p->rho;
    d.type = (int)p->type;
    return d;
}

extern "C" {

int dubins_shortest_path(DubinsPath* path, double q0[3], double q1[3], double rho) {
    DubinsPathD d = to_d(path);
    const int e = ppe::dubins_shortest_path(&d, q0, q1, rho);
    if (e == kEdubBadRho) return EDUBBADRHO;
    path->qi[0] = d.qi[0]; path->qi[1] = d.qi[1]; path->qi[2] = d.qi[2];
    path->rho = d.rho;
    if (e != kEdubOk) return EDUBNOPATH;
    path->param[0] = d.param[0]; path->param[1] = d.param[1]; path->param[2] = d.param[2];
    path->type = (DubinsPathType)d.type;
    return EDUBOK;
}

double dubins_path_length(const DubinsPath* path) { return ppe::dubins_path_length(to_d(path)); }

int dubins_path_sample(const DubinsPath* path, double t, double q[3]) {
    PathSampler s;
    sampler_init(&s, to_d(path));
    double x, y, yaw;
    if (sampler_sample<true>(s, t, &x, &y, &yaw) != kEdubOk) return EDUBPARAM;
    q[0] = x; q[1] = y; q[2] = yaw;
    return EDUBOK;
}

// keeps the prefix [0, t] of the path: the three segment parameters clamped in turn
int dubins_extract_subpath(const DubinsPath* path, double t, DubinsPath* newpath) {
    const double tprime = t / path->rho;
    if (t < 0 || t > dubins_path_length(path)) return EDUBPARAM;
    newpath->qi[0] = path->qi[0]; newpath->qi[1] = path->qi[1]; newpath->qi[2] = path->qi[2];
    newpath->rho = path->rho;
    newpath->type = path->type;
    newpath->param[0] = fmin(path->param[0], tprime);
    newpath->param[1] = fmin(path->param[1], tprime - newpath->param[0]);
    newpath->param[2] = fmin(path->param[2], tprime - newpath->param[0] - newpath->param[1]);
    return EDUBOK;
}

} // extern "C"
