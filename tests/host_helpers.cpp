// tests/host_helpers.cpp -- host build (g++, glibc libm) of the engine's scalar building blocks
// in path_planner_b200/csrc/ppe_math.cuh, so the CPU test-suite can pin the formulas (operation
// order, exact time stepping, skip counter, hoisted path sampler) bit for bit against the oracle.
#include <stdint.h>

#include "ppe_math.cuh"

using namespace ppe;

extern "C" {

void hh_dubins_batch(int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type, double* param,
                     double* length, int32_t* err) {
    for (int64_t i = 0; i < n; i++) {
        DubinsPathD p;
        p.param[0] = p.param[1] = p.param[2] = 0;
        p.type = 0;
        const int e = dubins_shortest_path(&p, q0 + 3 * i, q1 + 3 * i, rho[i]);
        type[i] = p.type;
        param[3 * i] = p.param[0]; param[3 * i + 1] = p.param[1]; param[3 * i + 2] = p.param[2];
        length[i] = e == kEdubOk ? dubins_path_length(p) : 0.0;
        err[i] = e;
    }
}

// t_i for i = 0..n-1 through the binade walker
void hh_time_walk(double t0, double dt, int n, double* out) {
    TimeWalker tw;
    tw.init(t0, dt);
    for (int i = 0; i < n; i++) out[i] = tw.at(i);
}

// same, but strided like the lanes of a warp (lane l asks for l, l+32, ...)
void hh_time_walk_lane(double t0, double dt, int lane, int n, double* out) {
    TimeWalker tw;
    tw.init(t0, dt);
    for (int i = lane, k = 0; k < n; i += 32, k++) out[k] = tw.at(i);
}

int hh_skip_count(double x, double c, int kmax) { return skip_count(x, c, kmax); }

// DubinsWrapper::sample through the hoisted sampler: path = qi[3] param[3] rho type
void hh_sample(const double* qi, const double* param, double rho, int type, double w_start, double w_speed, int n,
               const double* times, double* x, double* y, double* heading, int32_t* ok) {
    DubinsPathD p;
    p.qi[0] = qi[0]; p.qi[1] = qi[1]; p.qi[2] = qi[2];
    p.param[0] = param[0]; p.param[1] = param[1]; p.param[2] = param[2];
    p.rho = rho; p.type = type;
    PathSampler s;
    sampler_init(&s, p);
    for (int i = 0; i < n; i++) {
        x[i] = y[i] = heading[i] = 0;
        ok[i] = wrapper_sample_pose<true>(s, w_start, w_speed, times[i], &x[i], &y[i], &heading[i]) ? 1 : 0;
    }
}

// correctly-rounded helpers of ppe_crmath.cuh
void hh_cr_sincos(int64_t n, const double* x, double* s, double* c) {
    for (int64_t i = 0; i < n; i++) cr_sincos(x[i], &s[i], &c[i]);
}
void hh_cr_atan2(int64_t n, const double* y, const double* x, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = cr_atan2(y[i], x[i]);
}
void hh_cr_acos(int64_t n, const double* x, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = cr_acos(x[i]);
}

// div_zero_aware (ppe_math.cuh) element-wise, and Ribbon::getProjection through it (4 doubles per ribbon: sx sy ex ey)
void hh_div_zero_aware(int64_t n, const double* num, const double* den, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = div_zero_aware(num[i], den[i]);
}
void hh_ribbon_projection(int64_t n, const double* ribbons, const double* x, const double* y, double* px, double* py) {
    for (int64_t i = 0; i < n; i++) {
        const RibbonD r = {ribbons[4 * i], ribbons[4 * i + 1], ribbons[4 * i + 2], ribbons[4 * i + 3]};
        ribbon_projection(r, x[i], y[i], &px[i], &py[i]);
    }
}

// the platform libm (glibc on the reference's x86-64 build), element-wise, for agreement statistics
void hh_libm_sincos(int64_t n, const double* x, double* s, double* c) {
    for (int64_t i = 0; i < n; i++) { s[i] = sin(x[i]); c[i] = cos(x[i]); }
}
void hh_libm_atan2(int64_t n, const double* y, const double* x, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = atan2(y[i], x[i]);
}
void hh_libm_acos(int64_t n, const double* x, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = acos(x[i]);
}

} // extern "C"

// ---- KeyedHeap.h against the real std::make_heap / std::pop_heap (ties included) ---------------------------------
#include <algorithm>
#include <vector>

#include "../path_planner_b200/harness/KeyedHeap.h"

extern "C" {
// keys: n doubles (duplicates welcome); pops: how many pop_heap calls follow the make_heap.  Returns 0 when the keyed
// routines leave exactly the arrangement the std:: ones leave on (key, id) pairs compared by key only.
int hh_keyed_heap_check(const double* keys, int n, int pops) {
    struct Item { double key; uint32_t id; };
    std::vector<Item> ref((size_t)n);
    std::vector<double> k(keys, keys + n);
    std::vector<uint32_t> id((size_t)n);
    for (int i = 0; i < n; i++) { ref[i].key = keys[i]; ref[i].id = (uint32_t)i; id[i] = (uint32_t)i; }
    auto comp = [](const Item& a, const Item& b) { return a.key > b.key; }; // SamplingBasedPlanner.cpp:36-40
    std::make_heap(ref.begin(), ref.end(), comp);
    ppe_heap::make_heap(k.data(), id.data(), n);
    for (int i = 0; i < n; i++) if (ref[i].id != id[i] || ref[i].key != k[i]) return 1 + i;
    for (int p = 0; p < pops && p < n; p++) {
        std::pop_heap(ref.begin(), ref.end() - p, comp);
        ppe_heap::pop_heap(k.data(), id.data(), n - p);
        for (int i = 0; i < n; i++) if (ref[i].id != id[i] || ref[i].key != k[i]) return 100000 + p;
    }
    return 0;
}
}
