// ppe_math.cuh -- scalar building blocks of the edge engine, shared by every kernel.
//
// Everything here is plain IEEE fp64 arithmetic in the operation order of the reference (the
// library is compiled with -fmad=false so nothing is contracted into FMAs the x86-64 reference
// build does not have).  The functions are __host__ __device__ so that the host-side unit tests
// (tests/host_helpers.cpp, built with g++) can pin the formulas against the oracle bit for bit
// with glibc's libm; on the device the transcendental calls resolve to CUDA's libdevice.
//
// Reference anchors:
//   dubins_curves (un-vendored; contract in SURVEY.md Appendix B): shortest path of the six words,
//     path sampling -- called from DubinsWrapper.cpp:13,38-43.
//   State::yaw / State::setYaw            path_planner_common/include/path_planner_common/State.h:51-65
//   Edge::computeTrueCost time stepping   path_planner/src/planner/search/Edge.cpp:114-125,173
//   toCoverDistance skip counter          Edge.cpp:153-154
#pragma once

#include "ppe_math_base.cuh"
#include "ppe_crmath.cuh"

namespace ppe {

// ---- Dubins shortest path ----------------------------------------------------------------------
struct DubinsPathD {
    double qi[3];
    double param[3];
    double rho;
    int type;
};

enum { kEdubOk = 0, kEdubCoconfigs = 1, kEdubParam = 2, kEdubBadRho = 3, kEdubNoPath = 4 };

// Transcendentals go through ppe_crmath.cuh (correctly rounded w.h.p.) so that word choice and the
// segment parameters agree with the glibc-based reference to the last bit.
// All six words are evaluated unconditionally (no data-dependent branch per word: the feasibility
// tests become selects), in enum order LSL, LSR, RSL, RSR, RLR, LRL with a strict `<` so ties go
// to the earliest word.
PPE_HD int dubins_shortest_path(DubinsPathD* path, const double q0[3], const double q1[3], double rho) {
    if (rho <= 0.0) return kEdubBadRho;
    const double dx = q1[0] - q0[0];
    const double dy = q1[1] - q0[1];
    const double D = sqrt(dx * dx + dy * dy);
    const double d = D / rho;
    double theta = 0;
    if (d > 0) theta = mod2pi(cr_atan2(dy, dx));
    const double alpha = mod2pi(q0[2] - theta);
    const double beta = mod2pi(q1[2] - theta);
    double sa, ca, sb, cb;
    cr_sincos(alpha, &sa, &ca);
    cr_sincos(beta, &sb, &cb);
    const double c_ab = cr_cos(alpha - beta);
    const double d_sq = d * d;

    path->qi[0] = q0[0];
    path->qi[1] = q0[1];
    path->qi[2] = q0[2];
    path->rho = rho;

    double best_cost = INFINITY;
    int best = -1;
    double b0 = 0, b1 = 0, b2 = 0;
    double at_lsl = 0, at_rsr = 0; // shared with the CCC words below

#define PPE_TAKE(word, ok, t, p, q)                         \
    {                                                       \
        const double cost_ = (t) + (p) + (q);               \
        if ((ok) && cost_ < best_cost) {                    \
            best_cost = cost_; best = (word);               \
            b0 = (t); b1 = (p); b2 = (q);                   \
        }                                                   \
    }

    { // LSL
        const double tmp0 = d + sa - sb;
        const double p_sq = 2 + d_sq - (2 * c_ab) + (2 * d * (sa - sb));
        const double tmp1 = cr_atan2((cb - ca), tmp0);
        at_lsl = tmp1;
        const double t = mod2pi(tmp1 - alpha);
        const double p = sqrt(p_sq);
        const double q = mod2pi(beta - tmp1);
        PPE_TAKE(0, p_sq >= 0, t, p, q)
    }
    { // LSR
        const double p_sq = -2 + (d_sq) + (2 * c_ab) + (2 * d * (sa + sb));
        const double p = sqrt(p_sq);
        const double tmp0 = cr_atan2((-ca - cb), (d + sa + sb)) - cr_atan2(-2.0, p);
        const double t = mod2pi(tmp0 - alpha);
        const double q = mod2pi(tmp0 - mod2pi(beta));
        PPE_TAKE(1, p_sq >= 0, t, p, q)
    }
    { // RSL
        const double p_sq = -2 + d_sq + (2 * c_ab) - (2 * d * (sa + sb));
        const double p = sqrt(p_sq);
        const double tmp0 = cr_atan2((ca + cb), (d - sa - sb)) - cr_atan2(2.0, p);
        const double t = mod2pi(alpha - tmp0);
        const double q = mod2pi(beta - tmp0);
        PPE_TAKE(2, p_sq >= 0, t, p, q)
    }
    { // RSR
        const double tmp0 = d - sa + sb;
        const double p_sq = 2 + d_sq - (2 * c_ab) + (2 * d * (sb - sa));
        const double tmp1 = cr_atan2((ca - cb), tmp0);
        at_rsr = tmp1;
        const double t = mod2pi(alpha - tmp1);
        const double p = sqrt(p_sq);
        const double q = mod2pi(tmp1 - beta);
        PPE_TAKE(3, p_sq >= 0, t, p, q)
    }
    { // RLR
        const double tmp0 = (6. - d_sq + 2 * c_ab + 2 * d * (sa - sb)) / 8.;
        const double phi = at_rsr; // atan2(ca - cb, d - sa + sb): the very call RSR made
        const double p = mod2pi((2 * kPi) - cr_acos(tmp0));
        const double t = mod2pi(alpha - phi + mod2pi(p / 2.));
        const double q = mod2pi(alpha - beta - t + mod2pi(p));
        PPE_TAKE(4, fabs(tmp0) <= 1, t, p, q)
    }
    { // LRL
        const double tmp0 = (6. - d_sq + 2 * c_ab + 2 * d * (sb - sa)) / 8.;
        // atan2(ca - cb, d + sa - sb) = atan2(-(cb - ca), same x) = -atan2(cb - ca, x): atan2 is odd in y and the
        // correctly rounded value of -v is -(correctly rounded v), so LSL's call is reused
        // (when cb == ca both differences are +0 and the identity does not apply to the zero's sign)
        const double phi = (cb - ca != 0.0) ? -at_lsl : cr_atan2(ca - cb, d + sa - sb);
        const double p = mod2pi(2 * kPi - cr_acos(tmp0));
        const double t = mod2pi(-alpha - phi + p / 2.);
        const double q = mod2pi(mod2pi(beta) - alpha - t + mod2pi(p));
        PPE_TAKE(5, fabs(tmp0) <= 1, t, p, q)
    }
#undef PPE_TAKE
    if (best < 0) return kEdubNoPath;
    path->param[0] = b0;
    path->param[1] = b1;
    path->param[2] = b2;
    path->type = best;
    return kEdubOk;
}

PPE_HD double dubins_path_length(const DubinsPathD& p) {
    double length = 0.;
    length += p.param[0];
    length += p.param[1];
    length += p.param[2];
    length = length * p.rho;
    return length;
}

// ---- path sampler with the per-path loop invariants hoisted ---------------------------------------
// dubins_path_sample walks all three segments on every call; the end configurations of segment 1
// and 2 and the sin/cos of the three segment base angles depend on the path only, so they are
// computed once per edge.  The per-sample remainder is one sincos (arcs) or none (straight).
enum { kSegL = 0, kSegS = 1, kSegR = 2 };

struct PathSampler {
    double x0, y0, rho, length;
    double p1, p12;           // param[0], param[0] + param[1]
    double p2;                // param[1]
    double bx[3], by[3], bth[3], bs[3], bc[3]; // base configuration of each segment (+ sin/cos of its angle)
    int seg[3];
};

// one dubins_segment step from base configuration k by normalised length t
template <bool kExact>
PPE_HD void sampler_segment(const PathSampler& s, int k, double t, double* qx, double* qy, double* qth) {
    const int type = s.seg[k];
    double x, y, th;
    if (type == kSegS) {
        x = s.bc[k] * t;
        y = s.bs[k] * t;
        th = 0.0;
    } else {
        double sn, cs;
        if (type == kSegL) {
            if (kExact) cr_sincos(s.bth[k] + t, &sn, &cs); else sincos_f64(s.bth[k] + t, &sn, &cs);
            x = +sn - s.bs[k];
            y = -cs + s.bc[k];
            th = t;
        } else {
            if (kExact) cr_sincos(s.bth[k] - t, &sn, &cs); else sincos_f64(s.bth[k] - t, &sn, &cs);
            x = -sn + s.bs[k];
            y = +cs - s.bc[k];
            th = -t;
        }
    }
    *qx = x + s.bx[k];
    *qy = y + s.by[k];
    *qth = th + s.bth[k];
}

PPE_HD void sampler_init(PathSampler* s, const DubinsPathD& p) {
    // DIRDATA of dubins.c: L S L / L S R / R S L / R S R / R L R / L R L
    const int t0 = (p.type == 0 || p.type == 1 || p.type == 5) ? kSegL : kSegR;
    const int t1 = (p.type < 4) ? kSegS : (p.type == 4 ? kSegL : kSegR);
    const int t2 = (p.type == 0 || p.type == 2 || p.type == 5) ? kSegL : kSegR;
    s->seg[0] = t0; s->seg[1] = t1; s->seg[2] = t2;
    s->x0 = p.qi[0]; s->y0 = p.qi[1]; s->rho = p.rho;
    s->length = dubins_path_length(p);
    s->p1 = p.param[0];
    s->p2 = p.param[1];
    s->p12 = p.param[0] + p.param[1];
    s->bx[0] = 0.0; s->by[0] = 0.0; s->bth[0] = p.qi[2];
    cr_sincos(s->bth[0], &s->bs[0], &s->bc[0]);
    sampler_segment<true>(*s, 0, p.param[0], &s->bx[1], &s->by[1], &s->bth[1]);
    cr_sincos(s->bth[1], &s->bs[1], &s->bc[1]);
    sampler_segment<true>(*s, 1, p.param[1], &s->bx[2], &s->by[2], &s->bth[2]);
    cr_sincos(s->bth[2], &s->bs[2], &s->bc[2]);
}

// dubins_path_sample(path, t, q): returns kEdubParam (q untouched) when t is outside [0, length]
template <bool kExact>
PPE_HD int sampler_sample(const PathSampler& s, double t, double* x, double* y, double* yaw) {
    const double tprime = t / s.rho;
    if (t < 0 || t > s.length) return kEdubParam;
    double qx, qy, qth;
    if (tprime < s.p1) {
        sampler_segment<kExact>(s, 0, tprime, &qx, &qy, &qth);
    } else if (tprime < s.p12) {
        sampler_segment<kExact>(s, 1, tprime - s.p1, &qx, &qy, &qth);
    } else {
        sampler_segment<kExact>(s, 2, tprime - s.p1 - s.p2, &qx, &qy, &qth);
    }
    *x = qx * s.rho + s.x0;
    *y = qy * s.rho + s.y0;
    *yaw = mod2pi(qth);
    return kEdubOk;
}

// DubinsWrapper::sample body, DubinsWrapper.cpp:36-48: distance, EDUBPARAM retry at distance-1e-5,
// yaw -> heading.  Returns false when both library calls fail (the reference then keeps the stale
// pose; the engine reports that as a per-edge status instead).
// kExact = true: per-edge samples (end state) with the correctly rounded sincos; false: the
// per-sample hot loop with libdevice's fast sincos.
template <bool kExact>
PPE_HD bool wrapper_sample_pose(const PathSampler& s, double w_start, double w_speed, double time, double* x,
                                double* y, double* heading) {
    const double distance = (time - w_start) * w_speed;
    double yaw;
    int err = sampler_sample<kExact>(s, distance, x, y, &yaw);
    if (err == kEdubParam) err = sampler_sample<kExact>(s, distance - 1e-5, x, y, &yaw);
    if (err != kEdubOk) return false;
    double h = kPi2 - yaw;
    if (h < 0) h += kTwoPi;
    *heading = h;
    return true;
}

// ---- exact replay of `t += dt` (Edge.cpp:173) without the serial dependency -------------------------
// The sample times are produced by repeated floating-point addition, so t_i != t_0 + i*dt and the
// sample COUNT depends on the accumulated rounding.  While t stays inside one binade [2^e, 2^(e+1))
// every t is a multiple of u = ulp = 2^(e-52) and fl(t + dt) = t + D with D = dt rounded to a
// multiple of u -- a constant, unless dt lies exactly half-way between two multiples (tie: the
// rounding then depends on the parity of t; those binades are stepped one true addition at a
// time).  Inside a binade t_i = t_s + (i - s) * D is exact.  A walker therefore hands every lane
// its own t_i in O(#binades) work; binade crossings are done with one true addition.
// floor(a / b) for integer-valued doubles 0 <= a < 2^53, 1 <= b < 2^53: one fp64 division and an exact integer
// correction instead of an emulated 64-bit integer division.
PPE_HD int64_t floor_div_53(double a, double b) {
    int64_t q = (int64_t)(a / b);
    const int64_t ai = (int64_t)a, bi = (int64_t)b;
    if (q * bi > ai) q--;
    else if ((q + 1) * bi <= ai) q++;
    return q;
}

struct TimeWalker {
    double dt;
    double base, D;     // current run: t_i = base + (i - i0) * D for i0 <= i < i0 + cnt
    int i0, cnt;
    double t_next;      // first time after the current run (true fp addition)
    int edt;

    PPE_HD void build() {
        // run starts at (i0, base)
        const double t = base;
        int n = 0;
        D = 0.0;
        const int64_t bits = f64_bits(t);
        const int ebits = (int)((bits >> 52) & 0x7ff);
        if (t > 0 && ebits != 0x7ff && ebits - 1023 > edt && ebits > 60) {
            const int e = ebits - 1023;
            const double u = f64_pow2(e - 52);
            const double iu = f64_pow2(52 - e);       // 1 / u: scaling by a power of two is exact
            const double kq = floor(dt * iu);
            const double r = dt - kq * u;             // exact remainder, 0 <= r < u
            if (r != 0.5 * u) {                       // no tie: rounding is parity independent
                const double Dd = kq * u + (r > 0.5 * u ? u : 0.0);
                const double lim = f64_pow2(e + 1) - u; // steps with t_j + D <= lim cannot leave the binade
                if (Dd > 0 && t + Dd <= lim) {
                    // floor((lim - t) / Dd) on integers < 2^53 held exactly in doubles
                    int64_t nn = floor_div_53((lim - t) * iu, Dd * iu);
                    if (nn > 1000000) nn = 1000000;
                    n = (int)nn;
                    D = Dd;
                }
            }
        }
        cnt = n + 1;
        t_next = (t + (double)n * D) + dt; // crossing / single step: one true addition
    }

    PPE_HD void init(double t0, double dt_) {
        dt = dt_;
        edt = (dt_ > 0) ? f64_exponent(dt_) : -2000;
        // empty run in front of t0: the first at() builds the first run (one build() call site)
        i0 = 0;
        cnt = 0;
        base = t0;
        D = 0.0;
        t_next = t0;
    }

    // t_i for non-decreasing i
    PPE_HD double at(int i) {
        while (i >= i0 + cnt) {
            i0 += cnt;
            base = t_next;
            build();
        }
        return base + (double)(i - i0) * D;
    }
};

// ---- exact replay of the toCoverDistance skip counter (Edge.cpp:153-154) ----------------------------
// Number of loop iterations that take the `toCover > inc ? toCover -= inc` branch after a
// check-point that set toCover = x, i.e. the count of successive fp subtractions until the value
// is <= c.  skip_count_walk replays the subtractions binade by binade (same argument as TimeWalker,
// walking downwards; capped at kmax).
PPE_HD_NOINLINE int skip_count_walk(double x, double c, int kmax) {
    int k = 0;
    if (!(c > 0)) return (x > c) ? kmax : 0;
    const int ec = f64_exponent(c);
    while (x > c && k < kmax) {
        const int64_t bits = f64_bits(x);
        const int ebits = (int)((bits >> 52) & 0x7ff);
        if (ebits == 0x7ff) return kmax; // inf / nan distance: never reaches the next check-point
        const int e = ebits - 1023;
        if (e >= ec + 1 && ebits > 60) {
            const double u = f64_pow2(e - 52);
            const double iu = f64_pow2(52 - e);
            const double kq = floor(c * iu);
            const double r = c - kq * u;
            if (r != 0.5 * u) {
                const double D = kq * u + (r > 0.5 * u ? u : 0.0);
                const double lo = f64_pow2(e) + u; // results >= lo stay in the binade (and > c)
                if (D > 0 && x - D >= lo) {
                    int64_t n = floor_div_53((x - lo) * iu, D * iu);
                    if (n > (int64_t)(kmax - k)) n = kmax - k;
                    x = x - (double)n * D;
                    k += (int)n;
                    continue;
                }
            }
        }
        x = x - c; // one true subtraction
        k++;
    }
    return k;
}

// In exact arithmetic the count is ceil(x / c) - 1.  Every rounded subtraction is off by at most half an ulp of
// its result (<= ulp(x) / 2), so after k steps the value differs from x - k c by less than k ulp(x) / 2: the
// count can only differ from the exact-arithmetic one when x / c lies within k ulp(x) / (2 c) (plus the
// rounding of the quotient itself) of an integer.  Outside that margin -- taken 8x wider here -- the closed
// form IS the replayed count; inside it (probability ~1e-9 per call) the subtractions are replayed.
PPE_HD int skip_count(double x, double c, int kmax) {
    if (!(x > c)) return 0;       // also NaN
    if (!(c > 0) || !(x < 1e300)) return skip_count_walk(x, c, kmax);
    const double q = x / c;
    if (!(q < (double)kmax)) return kmax;
    const double f = floor(q);
    const double frac = q - f;    // exact
    const double ulp_x = f64_pow2(f64_exponent(x) - 52);
    const double margin = 4.0 * q * (ulp_x / c) + 8.0 * q * 0x1p-52;
    if (frac > margin && (1.0 - frac) > margin) return (int)f; // ceil(q) - 1 = floor(q) for non-integer q
    return skip_count_walk(x, c, kmax);
}

// ---- Ribbon primitives (Ribbon.h / Ribbon.cpp) -------------------------------------------------------
struct RibbonD {
    double sx, sy, ex, ey;
};

PPE_HD double ribbon_sqlen(const RibbonD& r) { // Ribbon.h:134-136
    return (r.ex - r.sx) * (r.ex - r.sx) + (r.ey - r.sy) * (r.ey - r.sy);
}
PPE_HD bool ribbon_covered(const RibbonD& r, bool strict, double W) { // Ribbon.cpp:23-25, :52-58
    const double minLength = 2 * W;
    return ribbon_sqlen(r) < minLength * minLength / (strict ? 2.0 * 2.0 : 1.0);
}
// num / den.  A zero numerator over a positive denominator is the numerator itself (the sign of the zero included); it is
// returned without dividing because CUDA's fp64 division leaves its inline sequence for numerators that small and calls a
// ~100-instruction routine -- and an axis-aligned ribbon produces one such quotient in every projection (ncu, round 2:
// 16 % of the warp walker's instructions were that routine).
// (The compiler turns a guarded division into an unconditional one plus a select, so the guard is put on the OPERAND: the
// division that is executed in the zero case is 1 / den, which stays on the inline path.)
PPE_HD double div_zero_aware(double num, double den) {
    const bool zero_case = (num == 0.0) && (den > 0.0);
    double safe = zero_case ? 1.0 : num;
#ifdef __CUDA_ARCH__
    asm volatile("" : "+d"(safe)); // opaque: otherwise the select is dropped again (its result is unused in the zero case)
#endif
    const double q = safe / den;
    return zero_case ? num : q;
}

PPE_HD void ribbon_projection(const RibbonD& r, double x, double y, double* px, double* py) { // Ribbon.cpp:72-78
    const double squaredL = ribbon_sqlen(r);
    const double dot = (x - r.sx) * (r.ex - r.sx) + (y - r.sy) * (r.ey - r.sy);
    const double projectedX = div_zero_aware((r.ex - r.sx) * dot, squaredL);
    const double projectedY = div_zero_aware((r.ey - r.sy) * dot, squaredL);
    *px = projectedX + r.sx;
    *py = projectedY + r.sy;
}
PPE_HD bool ribbon_contains_projection(const RibbonD& r, double px, double py) { // Ribbon.cpp:90-95
    const double tol = 1e-5;
    return !(((px - r.sx < -tol && px - r.ex < -tol) || (px - r.sx > tol && px - r.ex > tol)) ||
             ((py - r.sy < -tol && py - r.ey < -tol) || (py - r.sy > tol && py - r.ey > tol)));
}
PPE_HD double ribbon_distance(const RibbonD& r, double x, double y) { // Ribbon.h:118-121
    return div_zero_aware(fabs((r.ey - r.sy) * x - (r.ex - r.sx) * y + r.ex * r.sy - r.ey * r.sx), sqrt(ribbon_sqlen(r)));
}
PPE_HD bool ribbon_contains(const RibbonD& r, double x, double y, double px, double py, bool strict, double W) { // Ribbon.cpp:39-43
    if (!ribbon_contains_projection(r, px, py)) return false;
    const double d = ribbon_distance(r, x, y);
    return d < (strict ? W / 2.0 : W);
}
PPE_HD double point_distance(double x1, double y1, double x2, double y2) { // RibbonManager.h:289-291
    return sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
}

} // namespace ppe
