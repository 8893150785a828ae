// ppe_kernels.cu -- hand-written sm_100a kernels of the batched Dubins edge-evaluation engine.
//
//   K1  k1_dubins_batch   thread per (q0, q1, rho): shortest Dubins word, six words evaluated without a
//                         data-dependent branch per word, correctly rounded fp64.  Replaces Edge::computeApproxCost ->
//                         DubinsWrapper::set -> dubins_shortest_path (Edge.cpp:11-20, DubinsWrapper.cpp:9-17).
//   K2a k2a_prepare       thread per edge: radius / re-solve / speed change (Edge.cpp:73-85), Edge::setEnd, per-path
//                         sampler constants, first sample time, the edge's sample-time table.
//   K2t k2t_thread_walk   thread per edge: the loop of Edge::computeTrueCost (Edge.cpp:86-203) for SIMPLE edges -- chunks
//                         proved clean by the probe are skipped, the ribbon list is read in place -- Map/GridWorldMap::
//                         isBlocked, Binary/Gaussian collisionExists, ribbon check-points, truncated end state,
//                         g (Vertex.cpp:102-104), h = MaxDistance (RibbonManager.cpp:234-248).  The rest -> heavy list.
//   K2b k2_true_cost      warp per edge (lanes = 32 consecutive samples / lanes over ribbons): the same loop with the
//                         mutable ribbon list (cover, coverage completion, end-time truncation, ribbons-after).
//   K3  k3_best_scan + k3_best_final   best feasible f = g + h over the result records (pushVertexQueue's prune record).
//   k_safe_rows / k_safe_cols          dilated free-space bitmap for the chunk probe;  k_fp64_peak  roofline denominator.
//
// No tensor cores: this is branchy fp64 transcendental + bit-gather work.  Compiled with
// -fmad=false: the x86-64 reference build has no FMA contraction and discrete outcomes (word
// choice, cell index, ribbon containment, sample count) must not flip.
#include <float.h>
#include <stdlib.h>

#include "ppe_kernels.cuh"
#include "ppe_math.cuh"
#include "ppe_device.cuh"

namespace ppe {

namespace {

// K2b CTA shapes: 4 warps x 4 CTAs per SM by default (128 registers per thread fill the register file either
// way); 16-warp CTAs stage the obstacle list once per SM instead of four times (-DPPE_K2_FORCE_NARROW=0).
// Measured equal within noise on C2 / C3 (profiles/), so the shape with the larger ribbon capacity is the default.
constexpr int kWarpsWide = 16, kWarpsNarrow = 4;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kChunk = 32; // samples per chunk of the K2b walker = one per lane
constexpr int kSkipCap = 1 << 28;
constexpr int kMaxSamples = 1 << 22;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __noinline__ double warp_min(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

__device__ __noinline__ double warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

__device__ __forceinline__ RibbonD load_ribbon(const double4* p) {
    const double4 v = *p;
    RibbonD r;
    r.sx = v.x; r.sy = v.y; r.ex = v.z; r.ey = v.w;
    return r;
}

// squared RibbonManager::distance (RibbonManager.h:289-291).  IEEE sqrt is monotone and correctly rounded, so
// min / max over end-point distances are taken on the squares and rooted once: sqrt(min(a, b)) == min(sqrt a, sqrt b).
__device__ __forceinline__ double point_distance_sq(double x1, double y1, double x2, double y2) {
    return (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2);
}

// Ribbon::contains (Ribbon.cpp:39-43) can only hold when the projection of the point lies in the segment's bounding
// box (+- 1e-5) and the point is closer than the ribbon width to it, i.e. when the point lies in that box grown by
// the width.  Outside the box grown by a safe margin the projection / distance arithmetic (three divisions and a
// square root) is skipped: contains is false either way.  (Margin 1e-3 m; coordinates beyond 1e7 m take no shortcut.)
__device__ __forceinline__ bool ribbon_may_contain(const RibbonD& r, double x, double y, double W, bool tame) {
    const double grow = W * (1 + 1e-9) + 1e-3;
    const bool outside = x < fmin(r.sx, r.ex) - grow || x > fmax(r.sx, r.ex) + grow || y < fmin(r.sy, r.ey) - grow ||
                         y > fmax(r.sy, r.ey) + grow;
    return !(outside && tame);
}

// `tame`: every coordinate the shortcut compares is below 1e7 m in magnitude (decided once per edge: the path stays
// within its length of its start, the ribbon list only ever shrinks inside the parent's extent)
__device__ __forceinline__ bool coords_tame(const RibbonD& r) {
    return fabs(r.sx) < 1e7 && fabs(r.sy) < 1e7 && fabs(r.ex) < 1e7 && fabs(r.ey) < 1e7;
}

__device__ __forceinline__ double4 pack_ribbon(double sx, double sy, double ex, double ey) {
    return make_double4(sx, sy, ex, ey);
}

// BinaryDynamicObstaclesManager::collisionExists (Binary...cpp:4-22) and
// GaussianDynamicObstaclesManager::collisionExists (Gaussian...cpp:3-13), strict = true.
// Obstacles are staged in shared memory once per CTA; every lane walks the chunk's candidate
// obstacles (bit i of `mask`; all obstacles when the set has more than 64) in container order, so
// the per-sample sum has the reference's summation order.  Obstacles outside the mask contribute
// exactly 0 (binary) or less than 1e-26 each (gaussian) to every sample of the chunk.
__device__ __forceinline__ double obstacle_binary(const ObstacleD& o, double x, double y, double time) {
    const double dtm = time - o.Time;
    const double dx = o.Speed * dtm * o.cosYaw;
    const double dy = o.Speed * dtm * o.sinYaw;
    const double X = o.X + dx, Y = o.Y + dy;
    const double tx = x - X, ty = y - Y;
    const double rx = tx * o.cosYaw - ty * o.sinYaw;
    const double ry = tx * o.sinYaw + ty * o.cosYaw;
    return (fabs(rx) < o.a && fabs(ry) < o.b) ? 1.0 : 0.0;
}

__device__ __forceinline__ double obstacle_quadform(const ObstacleD& o, double x, double y, double time) {
    const double dtm = time - o.Time;
    const double dx = o.Speed * dtm * o.cosYaw;
    const double dy = o.Speed * dtm * o.sinYaw;
    const double X = o.X + dx, Y = o.Y + dy;
    const double d0 = x - X, d1 = y - Y;
    const double r0 = d0 * o.a + d1 * o.b;  // (v^T Sigma^-1)_0 = v0*i00 + v1*i10
    const double r1 = d0 * o.c + d1 * o.d;  // (v^T Sigma^-1)_1 = v0*i01 + v1*i11
    return r0 * d0 + r1 * d1;
}

// exp(x) for the Gaussian pdf exponent (x = -q / 2 <= 0): Cody-Waite reduction by ln 2, degree-13 Taylor kernel on
// |r| <= ln2 / 2 with the coefficients in constant memory, scaling by exponent arithmetic.  <= 1.5 ulp -- the same
// class as libdevice's exp against glibc's, at a fifth of its instruction count under -fmad=false.  Results below
// 1e-304 are flushed to 0 (they cannot reach the 1e-5 threshold of collisionExists nor change a sum above it).
__constant__ double c_exp[12] = {0x1.0000000000000p-1, 0x1.5555555555555p-3, 0x1.5555555555555p-5, 0x1.1111111111111p-7,
                                 0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-16, 0x1.71de3a556c734p-19,
                                 0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26, 0x1.1eed8eff8d898p-29, 0x1.6124613a86d09p-33};
__constant__ double c_ln2[3] = {0x1.71547652b82fep+0, 0x1.62e42fee00000p-1, 0x1.a39ef35793c76p-33};

__device__ __forceinline__ double exp_pdf(double x) {
    if (!(x <= 0.0)) return exp(x);    // positive (non positive-definite covariance) or NaN: the library function
    if (x < -700.0) return 0.0;
    const double kd = rint(x * c_ln2[0]);
    double r = __fma_rn(-kd, c_ln2[1], x);
    r = __fma_rn(-kd, c_ln2[2], r);
    double p = c_exp[11];
#pragma unroll
    for (int j = 10; j >= 0; j--) p = __fma_rn(p, r, c_exp[j]);
    p = __fma_rn(p * r, r, r) + 1.0;
    return p * __longlong_as_double((long long)((int)kd + 1023) << 52);
}

// Upper bound of obstacle i's contribution to collisionExists over a set of samples that all lie within `reach`
// metres and `half_t` seconds of the probe (x, y, t): binary -- 1 if the inflated rectangle can be reached at all,
// else 0; gaussian -- an upper bound of the pdf (sqrt(q) is a norm of the relative position: it cannot shrink
// faster than cull * displacement).  Every rounding goes the safe way.
__device__ __forceinline__ float obstacle_bound(int kind, const ObstacleD& o, double x, double y, double t, double reach,
                                                double half_t) {
    const double reach_i = (reach + fabs(o.Speed) * half_t) * (1 + 1e-9) + 1e-6; // sample moves <= reach, obstacle <= |v| half_t
    if (kind == kObsBinary) {
        const double dtm = t - o.Time;
        const double X = o.X + o.Speed * dtm * o.cosYaw, Y = o.Y + o.Speed * dtm * o.sinYaw;
        const double tx = x - X, ty = y - Y;
        const double rx = tx * o.cosYaw - ty * o.sinYaw;
        const double ry = tx * o.sinYaw + ty * o.cosYaw;
        return (fabs(rx) < o.a + reach_i && fabs(ry) < o.b + reach_i) ? 1.f : 0.f;
    }
    const double q = obstacle_quadform(o, x, y, t);
    const float sq = __fsqrt_rd(__double2float_rd(fmax(q, 0.0)));
    const float m = fmaxf(sq - __double2float_ru(reach_i * o.cull), 0.f) * 0.9999f;
    const float arg = -0.5f * (m * m) * 0.9999f;                 // >= the true exponent
    const float e = __expf(arg) * 1.001f + 1e-37f;               // >= exp(arg)
    return __double2float_ru(o.norm) * e * 1.0001f;              // >= max pdf over the set
}

// obstacles that can matter at all for such a set: bound above the threshold below which a term cannot change a sum
// that reaches 1e-5 (gaussian), or reachable rectangle (binary)
constexpr float kCandidateBound = 1e-26f;

__device__ __noinline__ double collision_exists(int kind, int n_obs, const ObstacleD* __restrict__ obs,
                                                   unsigned long long mask, bool use_mask, double x, double y, double time) {
    double sum = 0;
    if (use_mask) {
        if (kind == kObsBinary) {
            while (mask) {
                const int i = __ffsll((long long)mask) - 1;
                mask &= mask - 1;
                sum += obstacle_binary(obs[i], x, y, time);
            }
            return sum;
        }
        while (mask) {
            const int i = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            sum += obs[i].norm * exp_pdf(-0.5 * obstacle_quadform(obs[i], x, y, time));
        }
    } else {
        if (kind == kObsBinary) {
#pragma unroll 1
            for (int i = 0; i < n_obs; i++) sum += obstacle_binary(obs[i], x, y, time);
            return sum;
        }
#pragma unroll 1
        for (int i = 0; i < n_obs; i++) sum += obs[i].norm * exp_pdf(-0.5 * obstacle_quadform(obs[i], x, y, time));
    }
    if (sum < 1e-5) return 0;
    return sum;
}

// One ribbon check-point = RibbonManager::minDistanceFrom (RibbonManager.cpp:142-152) on the list as it stands,
// then -- when `do_cover` -- RibbonManager::cover(x, y, strict = true) (RibbonManager.cpp:14-22 with Ribbon::split,
// Ribbon.cpp:9-17, and add, RibbonManager.cpp:154-158), in ONE pass over the ribbons, lanes parallel over ribbons:
// both need the projection of (x, y) on every ribbon and its distance to the line.  Every ribbon is split
// independently; list order is kept by a warp prefix sum over the 0 / 1 / 2 pieces each ribbon leaves behind.
// The new list goes to `alt`; the caller swaps the buffers when something changed.  Returns the new count and the
// pre-cover minDistanceFrom in *to_cover (0 as soon as any ribbon contains the point, else the nearest endpoint).
__device__ __noinline__ int warp_checkpoint(const double4* cur, double4* alt, int nr, int cap, double x, double y, double W,
                                            bool do_cover, bool tame, int lane, double* to_cover, bool* changed, bool* overflow) {
    double mn = DBL_MAX;
    bool inside = false;
    int out_base = 0;
    bool any_change = false;
#pragma unroll 1
    for (int base = 0; base < nr; base += 32) {
        const int r = base + lane;
        const bool active = r < nr;
        RibbonD rb = {0, 0, 0, 0};
        bool contained = false;
        double px = 0, py = 0;
        if (active) {
            rb = load_ribbon(cur + r);
            if (ribbon_may_contain(rb, x, y, W, tame)) {
                ribbon_projection(rb, x, y, &px, &py);
                if (ribbon_contains_projection(rb, px, py)) {   // Ribbon::contains, Ribbon.cpp:39-43
                    const double d = ribbon_distance(rb, x, y);
                    inside = inside || (d < W);                  // non-strict: minDistanceFrom
                    contained = d < W / 2.0;                     // strict: cover
                }
            }
            const double dStart = point_distance_sq(rb.sx, rb.sy, x, y);
            const double dEnd = point_distance_sq(rb.ex, rb.ey, x, y);
            mn = fmin(fmin(mn, dEnd), dStart); // squared
        }
        // nothing to cover in this pass: no ribbon contains the point and none is short enough to be erased
        const bool touch = do_cover && __any_sync(kFull, active && (contained || ribbon_covered(rb, true, W)));
        if (do_cover && !touch) {
            // these (up to) 32 ribbons pass through unchanged; `alt` still gets them, a later pass may change the list
            if (active && out_base + lane < cap) alt[out_base + lane] = pack_ribbon(rb.sx, rb.sy, rb.ex, rb.ey);
            out_base += (nr - base < 32) ? (nr - base) : 32;
        } else if (do_cover) {
            RibbonD piece = {0, 0, 0, 0};
            RibbonD rest = rb;
            if (contained) {
                piece.sx = rb.sx; piece.sy = rb.sy; piece.ex = px; piece.ey = py;
                rest.sx = px; rest.sy = py;
            }
            const bool keep_piece = active && contained && !ribbon_covered(piece, true, W);
            const bool keep_rest = active && !ribbon_covered(rest, true, W);
            const int cnt = (keep_piece ? 1 : 0) + (keep_rest ? 1 : 0);
            const int incl = warp_incl_scan(cnt, lane);
            int off = out_base + incl - cnt;
            if (keep_piece) {
                if (off < cap) alt[off] = pack_ribbon(piece.sx, piece.sy, piece.ex, piece.ey);
                off++;
            }
            if (keep_rest) {
                if (off < cap) alt[off] = pack_ribbon(rest.sx, rest.sy, rest.ex, rest.ey);
            }
            const bool ch = active && (contained ? (keep_piece || !keep_rest || px != rb.sx || py != rb.sy) : !keep_rest);
            any_change |= __any_sync(kFull, ch);
            out_base += __shfl_sync(kFull, incl, 31);
        }
    }
    __syncwarp();
    *to_cover = (nr == 0 || __any_sync(kFull, inside)) ? 0.0 : sqrt(warp_min(mn));
    if (out_base > cap) *overflow = true;
    *changed = any_change;
    return any_change ? (out_base > cap ? cap : out_base) : nr;
}

// RibbonManager::maxDistance (RibbonManager.cpp:234-248); `scratch` holds >= nr doubles.
__device__ __noinline__ double warp_max_distance(const double4* cur, int nr, double x, double y, double W,
                                                    int lane, double* scratch) {
    double mn = DBL_MAX, mx = 0;
#pragma unroll 1
    for (int r = lane; r < nr; r += 32) {
        const RibbonD rb = load_ribbon(cur + r);
        scratch[r] = sqrt(ribbon_sqlen(rb)) - 2 * W;
        const double dStart = point_distance_sq(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance_sq(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart); // squared
        mx = fmax(fmax(mx, dEnd), dStart);
    }
    __syncwarp();
    double sumLength = 0;
#pragma unroll 1
    for (int r = 0; r < nr; r++) sumLength += scratch[r]; // list order, as the reference sums
    __syncwarp();
    mn = sqrt(warp_min(mn));
    mx = sqrt(warp_max(mx));
    return fmax(sumLength + mn, mx);
}


// Per-edge state produced by K2a (one THREAD per edge) and consumed by K2b (one WARP per edge):
// the solved path, the sampler's per-path constants, the wrapper times and the connected end
// state.  Splitting the edge this way removes the 32-fold redundancy a warp would have on the
// expensive, purely scalar part (the correctly rounded Dubins solve).  The record is a flat array
// of kPrepDoubles doubles so that a warp stages it into shared memory with two coalesced loads;
// every lane then reads the (warp-uniform) values it needs as shared-memory broadcasts instead of
// holding ~60 registers of per-edge state.
enum PrepSlot {
    // per segment k = 0..2 (three entries each): start of the segment in normalised path length, turn sign
    // (+1 L, -1 R, 0 S), base angle
    kOff = 0, kSgn = 3, kBth = 6,
    // base x, y and sin / cos of the base angle per segment (dubins_segment constants)
    kRefBx = 9, kRefBy = 12, kRefBs = 15, kRefBc = 18,
    kInvRho = 21, kLength, kWStart, kWSpeed, kWEnd, kApprox, kRho, kX0, kY0, kYaw0,
    kParam0 = 31, kParam1, kParam2, kType, kStatus, kSampleFault,
    kT0 = 37,      // first sample time: src time nudged by fmod(t - startStateTime, dt), Edge.cpp:118-120
    kPrepDoubles = 40
};
// + the edge's sample-time table (see TimeTable below), built by K2a when it fits kPrepRuns runs
constexpr int kPrepRuns = 8;
struct PreparedEdge {
    double v[kPrepDoubles];
    double run_base[kPrepRuns];
    double run_D[kPrepRuns];
    double t_after;
    int run_i0[kPrepRuns];
    int n_runs;   // 0: needs more runs than fit here -- the warp walker builds its own table
    int i_after;
    int pad_[2];
};

// Edge.cpp:73-85, Edge::setEnd :208-215, DubinsWrapper.cpp:9-17,84-93 -- scalar, one thread.
__device__ void prepare_edge(const ppe_config& cfg, double dt, double horizon_end, const ppe_edge* __restrict__ edge,
                             PreparedEdge* __restrict__ out) {
    const double src_x = edge->src[0], src_y = edge->src[1], src_h = edge->src[2], src_speed = edge->src[3],
                 src_t = edge->src[4];
    const bool has_path = edge->has_path > 0;
    const bool skip = edge->has_path < 0; // empty slot of a frontier batch
    const bool cov = edge->coverage_allowed != 0;

    DubinsPathD path;
    path.qi[0] = path.qi[1] = path.qi[2] = 0;
    path.param[0] = path.param[1] = path.param[2] = 0;
    path.rho = 0; path.type = 0;
    PathSampler smp;
    memset(&smp, 0, sizeof smp);
    double w_speed = 0, w_start = -1, w_end = -1;
    bool w_init = false;
    double ex, ey, eh, es; // end()->state() pose
    double approx = -1;
    int status = skip ? PPE_EDGE_SKIPPED : PPE_EDGE_OK;
    bool sample_fault = false; // both dubins_path_sample attempts failed (stale-pose case)

    if (skip) {
        ex = ey = eh = es = 0;
        approx = 0;
    } else if (has_path) {
        path.qi[0] = edge->path_qi[0]; path.qi[1] = edge->path_qi[1]; path.qi[2] = edge->path_qi[2];
        path.param[0] = edge->path_param[0]; path.param[1] = edge->path_param[1]; path.param[2] = edge->path_param[2];
        path.rho = edge->path_rho; path.type = edge->path_type;
        w_speed = edge->w_speed;
        w_start = edge->w_start_time;
        w_init = w_start >= 0;
        sampler_init(&smp, path);
        w_end = w_start + smp.length / w_speed;
        if (edge->w_end_time < w_end) w_end = edge->w_end_time;
        ex = 0; ey = 0; eh = 0;
        if (!w_init || !(w_start <= w_end)) {
            status = PPE_EDGE_ERR_END_SAMPLE;
        } else if (!wrapper_sample_pose<true>(smp, w_start, w_speed, w_end, &ex, &ey, &eh)) {
            eh = heading_to_yaw(eh);
            sample_fault = true;
        }
        es = w_speed;
        approx = (w_end - src_t) * cfg.time_penalty_factor;
    } else {
        ex = edge->dst[0]; ey = edge->dst[1]; eh = edge->dst[2]; es = edge->dst[3];
    }
    const double speed = es;
    const double rho = cov ? cfg.coverage_turning_radius : cfg.turning_radius;
    if (status == PPE_EDGE_OK && (approx == -1 || path.rho != rho)) {
        if (src_x == ex && src_y == ey && src_h == eh) {
            approx = 0; // co-located (State::isCoLocated, State.cpp:87-91): wrapper untouched
        } else {
            const double q1[3] = {src_x, src_y, heading_to_yaw(src_h)};
            const double q2[3] = {ex, ey, heading_to_yaw(eh)};
            dubins_shortest_path(&path, q1, q2, rho); // return code ignored, as DubinsWrapper.cpp:13 does
            sampler_init(&smp, path);
            w_speed = src_speed;
            w_start = src_t;
            w_init = w_start >= 0;
            w_end = w_start + smp.length / w_speed;
            approx = smp.length / speed * cfg.time_penalty_factor;
        }
    }
    if (status == PPE_EDGE_OK && w_speed != speed) {
        if (!w_init) { // setSpeed -> setEndTime -> length() throws on an unset wrapper
            status = PPE_EDGE_ERR_NO_PATH;
        } else {
            w_speed = speed;
            w_end = w_start + smp.length / w_speed;
        }
    }
    if (status == PPE_EDGE_OK && approx < 0) status = PPE_EDGE_ERR_NO_PATH; // Edge.cpp:85

    double* v = out->v;
    const double x0 = path.qi[0], y0 = path.qi[1];
    v[kOff] = 0.0; v[kOff + 1] = smp.p1; v[kOff + 2] = smp.p12;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double sg = smp.seg[k] == kSegL ? 1.0 : (smp.seg[k] == kSegR ? -1.0 : 0.0);
        v[kSgn + k] = sg;
        v[kBth + k] = smp.bth[k];
        v[kRefBx + k] = smp.bx[k]; v[kRefBy + k] = smp.by[k]; v[kRefBs + k] = smp.bs[k]; v[kRefBc + k] = smp.bc[k];
    }
    v[kInvRho] = 1.0 / path.rho; v[kLength] = smp.length;
    v[kWStart] = w_start; v[kWSpeed] = w_speed; v[kWEnd] = w_end;
    v[kApprox] = approx;
    v[kRho] = path.rho; v[kX0] = x0; v[kY0] = y0; v[kYaw0] = path.qi[2];
    v[kParam0] = path.param[0]; v[kParam1] = path.param[1]; v[kParam2] = path.param[2];
    v[kType] = (double)path.type;
    v[kStatus] = (double)status;
    v[kSampleFault] = sample_fault ? 1.0 : 0.0;
    const double t0 = src_t + fmod(src_t - cfg.start_state_time, dt);
    v[kT0] = t0;
    v[kT0 + 1] = 0.0; v[kT0 + 2] = 0.0;

    // sample-time table (exact replay of `t += dt`): runs until one STARTS at or beyond the end time
    {
        const double end_time = fmin(horizon_end, w_end);
        const int edt = (dt > 0) ? f64_exponent(dt) : -2000;
        double t = t0;
        int i = 0, n = 0;
        bool fits = (status == PPE_EDGE_OK) && (dt > 0);
        while (fits) {
            TimeWalker tw;
            tw.dt = dt; tw.edt = edt; tw.base = t; tw.i0 = 0;
            tw.build();
            out->run_base[n] = t; out->run_D[n] = tw.D; out->run_i0[n] = i;
            n++;
            const bool past = !(t < end_time);
            i += tw.cnt;
            t = tw.t_next;
            if (past) break;
            if (n == kPrepRuns || i > (1 << 22)) fits = false;
        }
        out->n_runs = fits ? n : 0;
        out->i_after = i;
        out->t_after = t;
        // number of samples with t_i < end_time (the loop bound of Edge.cpp:125 before any truncation)
        int n_valid = -1;
        if (fits) {
            n_valid = 0;
            for (int r = 0; r < n; r++) {
                const double b = out->run_base[r], D = out->run_D[r];
                const int i0 = out->run_i0[r];
                const int cnt = (r + 1 < n ? out->run_i0[r + 1] : i) - i0;
                if (!(b < end_time)) break;
                if (D > 0) {
                    // smallest k with b + k D >= end_time (k D is exact inside the run)
                    int k = (int)fmin(floor((end_time - b) / D), (double)cnt);
                    while (k < cnt && b + (double)k * D < end_time) k++;
                    while (k > 0 && !(b + (double)(k - 1) * D < end_time)) k--;
                    n_valid = i0 + k;
                    if (k < cnt) break;
                } else {
                    n_valid = i0 + 1; // single-step run whose only sample is below the end time
                }
            }
        }
        out->pad_[0] = n_valid;
        out->pad_[1] = 0;
    }
}

// sin and cos for |x| up to a few thousand (path angles stay within a few turns): three-constant
// Cody-Waite reduction by pi/2 + the fdlibm minimax kernels on [-pi/4, pi/4] (|error| < 1 ulp).  Used for the
// per-sample poses only (1e-9 tolerance class).  The coefficients live in constant memory so that every DFMA
// takes its coefficient as a constant-bank operand instead of two register moves per coefficient per sample.
__constant__ double c_sin[6] = {-0x1.5555555555549p-3, 0x1.111111110f8a6p-7, -0x1.a01a019c161d5p-13,
                                0x1.71de357b1fe7dp-19, -0x1.ae5e68a2b9cebp-26, 0x1.5d93a5acfd57cp-33};
__constant__ double c_cos[6] = {0x1.555555555554cp-5, -0x1.6c16c16c15177p-10, 0x1.a01a019cb159p-16,
                                -0x1.27e4f809c52adp-22, 0x1.1ee9ebdb4b1c4p-29, -0x1.8fae9be8838d4p-37};
__constant__ double c_red[4] = {0x1.45f306dc9c883p-1, 0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54, -0x1.f1976b7ed8fbcp-110};

__device__ __forceinline__ void sincos_bounded(double x, double* sn, double* cs) {
    const double kd = rint(x * c_red[0]);
    const int q = (int)kd;
    double r = __fma_rn(-kd, c_red[1], x);
    r = __fma_rn(-kd, c_red[2], r);
    r = __fma_rn(-kd, c_red[3], r);
    const double z = r * r;
    double ps = c_sin[5];
    ps = __fma_rn(ps, z, c_sin[4]);
    ps = __fma_rn(ps, z, c_sin[3]);
    ps = __fma_rn(ps, z, c_sin[2]);
    ps = __fma_rn(ps, z, c_sin[1]);
    ps = __fma_rn(ps, z, c_sin[0]);
    const double s0 = __fma_rn(r * z, ps, r);
    double pc = c_cos[5];
    pc = __fma_rn(pc, z, c_cos[4]);
    pc = __fma_rn(pc, z, c_cos[3]);
    pc = __fma_rn(pc, z, c_cos[2]);
    pc = __fma_rn(pc, z, c_cos[1]);
    pc = __fma_rn(pc, z, c_cos[0]);
    const double c0 = __fma_rn(z * z, pc, __fma_rn(z, -0.5, 1.0));
    const double a = (q & 1) ? c0 : s0;
    const double b2 = (q & 1) ? s0 : c0;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b2 : b2;
}

// ---- exact replay of `t += dt` (Edge.cpp:173) as a per-edge table ------------------------------------------------
// Within one binade the repeated addition is an exact arithmetic progression (see TimeWalker in
// ppe_math.cuh); an edge crosses a handful of binades, so its sample times are a short table of runs
// (start index, start time, exact increment) held in shared memory per warp.  Any lane then gets any t_i
// in O(#runs) without carrying walker state in registers, which is what lets the probe pass look ahead.
constexpr int kTimeRuns = 24;
struct TimeTable {
    double base[kTimeRuns];
    double D[kTimeRuns];
    int i0[kTimeRuns];
    int n;
    int i_after;     // first index after the last run ...
    int overflow;    // ... reached with runs to spare (0) or because the table is full (1: step from there)
    int pad_;
    double t_after;  // ... and its time
};

// one run starting at time t: increment D (0 for a single step), sample count, time after the run
__device__ __forceinline__ void time_run(double t, double dt, int edt, double* D_out, int* cnt_out, double* t_next_out) {
    TimeWalker tw;
    tw.dt = dt; tw.edt = edt; tw.base = t; tw.i0 = 0;
    tw.build();
    *D_out = tw.D; *cnt_out = tw.cnt; *t_next_out = tw.t_next;
}

// Runs are built until one STARTS at or beyond `end_time`, so the table covers every executed sample and the
// first one past the end (whose time the reference keeps in `t` after the loop, Edge.cpp:185-191).
__device__ __noinline__ void time_table_build(TimeTable* tt, double t0, double dt, double end_time, int lane) {
    const int edt = (dt > 0) ? f64_exponent(dt) : -2000;
    double t = t0;
    int i = 0, n = 0, overflow = 0;
    for (;;) {
        double D, t_next;
        int cnt;
        time_run(t, dt, edt, &D, &cnt, &t_next);
        if (lane == 0) { tt->base[n] = t; tt->D[n] = D; tt->i0[n] = i; }
        n++;
        const bool past = !(t < end_time);
        i += cnt;
        t = t_next;
        if (past || i > kMaxSamples) break;
        if (n == kTimeRuns) { overflow = 1; break; }
    }
    if (lane == 0) { tt->n = n; tt->i_after = i; tt->t_after = t; tt->overflow = overflow; }
    __syncwarp();
}

__device__ __noinline__ double time_at_overflow(const TimeTable* tt, int i, double dt) {
    double t = tt->t_after; // table full (pathological time scales): true additions from its end
#pragma unroll 1
    for (int k = tt->i_after; k < i; k++) t = t + dt;
    return t;
}

__device__ __forceinline__ double time_at_from(const TimeTable* tt, int i, int r, double dt) {
    const int n = tt->n;
#pragma unroll 1
    for (int k = r + 1; k < n && i >= tt->i0[k]; k++) r = k;
    const double t = tt->base[r] + (double)(i - tt->i0[r]) * tt->D[r];
    if (tt->overflow && i >= tt->i_after) return time_at_overflow(tt, i, dt);
    return t;
}

// out of line: check-points of culled chunks, loop exit, probe pass
__device__ __noinline__ double time_at(const TimeTable* tt, int i, double dt) { return time_at_from(tt, i, 0, dt); }

// DubinsWrapper::sample (DubinsWrapper.cpp:29-49) with the per-path constants of the prepared record: pose at
// time t in the reference's operation order (dubins_segment / dubins_path_sample: q = segment(tl) + base,
// x = q.x * rho + x0).  On straight segments this is the reference's arithmetic bit for bit, which keeps exact
// f-ties (straight survey lines produce many) breaking the same way; on arcs it differs from glibc by the ~1 ulp
// of sincos_bounded.  `ang` is the un-wrapped path angle; the heading is derived from it only where it is needed
// (check-points, loop exit).  Returns false when both dubins_path_sample attempts fail.
__device__ __forceinline__ bool pose_eval(const double* pe, double t, double* x_out, double* y_out, double* ang_out,
                                          bool* in_time) {
    const double w_start = pe[kWStart];
    *in_time = (w_start <= t) && (pe[kWEnd] >= t); // DubinsWrapper::containsTime
    double dist = (t - w_start) * pe[kWSpeed];
    bool sample_ok = true;
    {
        const double length = pe[kLength];
        if (dist < 0 || dist > length) { // EDUBPARAM -> one retry at distance - 1e-5 (DubinsWrapper.cpp:39-42)
            dist = dist - 1e-5;
            sample_ok = !(dist < 0 || dist > length);
        }
    }
    const double tprime = dist * pe[kInvRho];
    const double p1 = pe[kOff + 1];
    const int k = tprime < p1 ? 0 : (tprime < pe[kOff + 2] ? 1 : 2);
    const double tl = k == 0 ? tprime : (k == 1 ? tprime - p1 : tprime - p1 - pe[kParam1]);
    const double* seg = pe + k;
    const double sg = seg[kSgn], bth = seg[kBth], bs = seg[kRefBs], bc = seg[kRefBc];
    const double ang = sg * tl + bth; // L: t + th, R: -t + th, S: 0 + th  (dubins_segment)
    double sn, cs;
    sincos_bounded(ang, &sn, &cs);
    const double qx = (sg == 0.0 ? bc * tl : sg * (sn - bs)) + seg[kRefBx];
    const double qy = (sg == 0.0 ? bs * tl : -sg * (cs - bc)) + seg[kRefBy];
    const double rho = pe[kRho];
    *x_out = qx * rho + pe[kX0];
    *y_out = qy * rho + pe[kY0];
    *ang_out = ang;
    return sample_ok;
}

// out-of-line copy for the uniform evaluations (loop exit, end state, a check-point on a chunk's first sample)
__device__ __noinline__ bool pose_eval_ref(const double* pe, double t, double* x_out, double* y_out, double* ang_out,
                                           bool* in_time) {
    return pose_eval(pe, t, x_out, y_out, ang_out, in_time);
}

// mod2pi + State::setYaw (State.h:51-65): heading of a path angle
__device__ __forceinline__ double heading_of(double ang) {
    const double yaw = ang - kTwoPi * floor(ang * 0x1.45f306dc9c883p-3);
    double hd = kPi2 - yaw;
    if (hd < 0) hd += kTwoPi;
    return hd;
}

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
    return (unsigned long long)__shfl_sync(kFull, (long long)v, src);
}

// ---- chunk culling: the probe pass ---------------------------------------------------------------------------------
// Lane m looks at chunk m of the next 32 chunks (samples [c0, c0 + 32), c0 = base + 32 m): it evaluates the exact pose
// of the chunk's middle sample and proves, where it can, that NOTHING discrete happens in the chunk:
//   * every sample time is below the end time and inside the wrapper, every distance inside the path
//     (times and distances are monotone, so the two ends of the chunk decide);
//   * every sample lies within `rad` (half a chunk of arc length) of the probe, and the probe's cell is `safe`
//     (all cells within that reach are in bounds and free) => no sample is blocked;
//   * the obstacles' contributions are bounded over the chunk: binary -- the probe is farther than the reach from
//     every inflated rectangle => every sample counts 0; gaussian -- sum of per-obstacle upper bounds < 1e-5 =>
//     collisionExists returns exactly 0 for every sample.
// A proved chunk costs nothing but its check-points.  For the others the lane reports which obstacles can matter
// at all in the chunk (candidate mask), so the exact per-sample evaluation touches a few obstacles, not all.
// the proof for ONE chunk, given the times of its first, middle and last sample; shared by the warp walker (lane m
// probes chunk m of the next 32) and the thread walker (a thread probes its edge's chunks one after the other)
__device__ __forceinline__ bool probe_one(const WorldD& w, const double* pe, double t_first, double t_mid, double t_last,
                                          const ObstacleD* s_obs, double end_time, double rad, unsigned long long edge_mask,
                                          unsigned long long* mask_out, const uint32_t* tile = nullptr) {
    const double w_start = pe[kWStart], w_speed = pe[kWSpeed];
    bool ok = (t_last < end_time) && (w_start <= t_first) && (pe[kWEnd] >= t_last);
    const double d_first = (t_first - w_start) * w_speed, d_last = (t_last - w_start) * w_speed;
    ok = ok && !(d_first < 0) && !(d_last > pe[kLength]);
    // arc length between the probe and either end of the chunk, from the actual sample times
    const double reach = fmax(t_mid - t_first, t_last - t_mid) * w_speed;
    bool near = (t_first <= t_mid) && (t_mid <= t_last) && (reach * (1 + 1e-9) + 1e-9 <= rad);
    double x = 0, y = 0, ang = 0;
    bool in_time = false;
    const bool sample_ok = pose_eval(pe, t_mid, &x, &y, &ang, &in_time);
    near = near && sample_ok && in_time; // every sample of the chunk lies within `reach` of a valid probe pose
    ok = ok && near;
    unsigned long long mask = w.n_obs >= 64 ? ~0ull : ((1ull << w.n_obs) - 1ull); // no bound: every obstacle is a candidate
    if (near) {
        ok = ok && map_safe(w, x, y, tile);
        if (w.obs_kind != kObsNone && w.n_obs > 0) {
            if (!w.obs_cull_ok) {
                ok = false;
            } else {
                const double half_t = fmax(t_mid - t_first, t_last - t_mid) * (1 + 1e-9);
                const double reach_m = reach * (1 + 1e-9) + 1e-9;
                const int kind = w.obs_kind;
                unsigned long long todo = edge_mask; // obstacles that can matter anywhere on this edge
                float bound = 0.f;
                mask = 0;
#pragma unroll 1
                while (todo) {
                    const int i = __ffsll((long long)todo) - 1;
                    todo &= todo - 1;
                    const float b_i = obstacle_bound(kind, s_obs[i], x, y, t_mid, reach_m, half_t);
                    if (b_i >= kCandidateBound) mask |= 1ull << i;
                    bound += b_i;
                }
                // binary: no reachable rectangle; gaussian: the sum of the upper bounds stays below the threshold
                ok = ok && (kind == kObsBinary ? (mask == 0) : (bound * 1.001f < 1e-5f));
            }
        }
    }
    *mask_out = mask;
    return ok;
}

__device__ __noinline__ void probe_chunks(const WorldD* wp, const double* pe, const TimeTable* tt, const ObstacleD* s_obs,
                                          int base, int lane, double end_time, double rad, unsigned long long edge_mask,
                                          bool* safe_out, unsigned long long* mask_out) {
    const int c0 = base + lane * kChunk;
    const double dt = wp->dt;
    const double t_first = time_at(tt, c0, dt);
    const double t_mid = time_at(tt, c0 + kChunk / 2, dt);
    const double t_last = time_at(tt, c0 + kChunk - 1, dt);
    *safe_out = probe_one(*wp, pe, t_first, t_mid, t_last, s_obs, end_time, rad, edge_mask, mask_out);
}

// ---- check-point fast path of the warp walker ---------------------------------------------------------------------------
// An edge that runs along a survey line executes a check-point at EVERY sample (toCoverDistance is 0 inside a ribbon),
// hundreds in a row, and each one only moves the start of the one ribbon it is on (Ribbon::split returns a piece too
// short to keep, RibbonManager.cpp:14-22).  The general check-point (warp_checkpoint) walks all R ribbons with lanes,
// ballots and a prefix sum per check-point.  The fast path does the same arithmetic for the few ribbons that can be
// touched at all in this chunk of 32 samples -- those whose bounding box, grown by the ribbon width, meets the bounding
// box of the chunk's poses -- as uniform scalar work, and updates the ribbon in place.  It applies only when it can
// prove the outcome equals the general one: some relevant ribbon contains the point (so minDistanceFrom is 0 whatever
// the other ribbons are), no ribbon of the list is short enough to be erased by cover() regardless of containment,
// and the split leaves the list's structure alone (piece dropped, remainder kept).  Anything else takes the general path.
constexpr int kRelCap = 8;

// ribbons of the list that any pose inside [x0, x1] x [y0, y1] could be contained in; -1 when more than kRelCap
__device__ __noinline__ int relevant_ribbons(const double4* cur, int nr, double x0, double x1, double y0, double y1, double W,
                                              int lane, int* rel) {
    const double grow = W * (1 + 1e-9) + 2e-3;
    int n = 0;
#pragma unroll 1
    for (int base = 0; base < nr; base += 32) {
        const int r = base + lane;
        bool near = false;
        if (r < nr) {
            const RibbonD rb = load_ribbon(cur + r);
            near = !(x1 < fmin(rb.sx, rb.ex) - grow || x0 > fmax(rb.sx, rb.ex) + grow || y1 < fmin(rb.sy, rb.ey) - grow ||
                     y0 > fmax(rb.sy, rb.ey) + grow);
        }
        unsigned m = __ballot_sync(kFull, near);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (n < kRelCap && lane == 0) rel[n] = base + b;
            n++;
        }
    }
    __syncwarp();
    return n <= kRelCap ? n : -1;
}

// nearest ribbon end point over the whole list (lanes over ribbons): minDistanceFrom when no ribbon contains the point
__device__ __noinline__ double warp_nearest_endpoint(const double4* cur, int nr, double x, double y, int lane) {
    double mn = DBL_MAX;
#pragma unroll 1
    for (int r = lane; r < nr; r += 32) {
        const RibbonD rb = load_ribbon(cur + r);
        const double dStart = point_distance_sq(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance_sq(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart);
    }
    return sqrt(warp_min(mn));
}

// One check-point on the relevant ribbons only (uniform across the warp).  Returns 1 when it handled the check-point with
// minDistanceFrom == 0 (cover applied in place); 2 when NO ribbon contains the point even non-strictly -- then cover()
// cannot touch anything either (strict containment implies it) and only the nearest end point is left to compute;
// 0 = nothing was modified, take the general path.
__device__ __noinline__ int fast_checkpoint(double4* cur, const int* rel, int n_rel, double x, double y, double W, bool do_cover,
                                             bool tame, int lane, bool* changed) {
    bool inside = false;
    int upd = -1;          // at most one in-place update per check-point on this path
    double nsx = 0, nsy = 0, nex = 0, ney = 0;
    bool ch = false;
#pragma unroll 1
    for (int q = 0; q < n_rel; q++) {
        const int r = rel[q];
        const RibbonD rb = load_ribbon(cur + r);
        if (!ribbon_may_contain(rb, x, y, W, tame)) continue;
        double px, py;
        ribbon_projection(rb, x, y, &px, &py);
        if (!ribbon_contains_projection(rb, px, py)) continue;
        const double d = ribbon_distance(rb, x, y);
        inside = inside || (d < W);
        if (do_cover && d < W / 2.0) {
            // Ribbon::split (Ribbon.cpp:9-17): piece = start -> projection, remainder = projection -> end; cover() keeps
            // whichever is not short enough to be "covered" (RibbonManager.cpp:14-22).  Exactly one of the two survives
            // when the boat runs along the ribbon: the remainder ahead (travelling start -> end) or the piece ahead
            // (travelling end -> start).  Either way the list keeps its structure and the ribbon is updated in place.
            RibbonD piece = {rb.sx, rb.sy, px, py};
            RibbonD rest = {px, py, rb.ex, rb.ey};
            const bool keep_piece = !ribbon_covered(piece, true, W), keep_rest = !ribbon_covered(rest, true, W);
            if (keep_piece == keep_rest) return 0; // both kept (insertion) or both erased: the list's structure changes
            if (upd >= 0) return 0;                // two ribbons at once: general path
            upd = r;
            if (keep_rest) { nsx = px; nsy = py; nex = rb.ex; ney = rb.ey; ch = (px != rb.sx) || (py != rb.sy); }
            else { nsx = rb.sx; nsy = rb.sy; nex = px; ney = py; ch = true; } // the remainder was erased
        }
    }
    if (!inside) return 2; // d >= W for every ribbon whose projection is contained: nothing to cover, distance from the end points
    if (upd >= 0 && lane == 0) cur[upd] = pack_ribbon(nsx, nsy, nex, ney);
    __syncwarp();
    *changed = ch;
    return 1;
}

// -DPPE_K2B_PROFILE (make prof -> libppe_prof.so, development only): where the warp walker's cycles go, summed over the
// edges of a launch: [0] edges [1] total [2] probe passes [3] lane poses + map [4] fast check-points [5] general
// check-points [6] obstacle penalties [7] tail (end state, final cover, heuristic, record) [8] fast count [9] general count
#ifdef PPE_K2B_PROFILE
__device__ unsigned long long g_k2b_prof[16];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0) prof[slot] += clock64() - (t0)
#define PROF_INC(slot) prof[slot] += 1
#else
#define PROF_T(var)
#define PROF_ADD(slot, t0)
#define PROF_INC(slot)
#endif

// One edge, one warp.  Per-edge scalars that every lane would hold identically live in the warp's
// shared-memory copy of the prepared record (`pe`) and in the time table (`tt`).  The sample loop of
// Edge.cpp:125-175 runs in chunks of 32 consecutive samples:
//   * a chunk the probe pass proved clean is not evaluated at all -- only its ribbon check-points are,
//     from directly evaluated poses;
//   * any other chunk is evaluated exactly, lane i owning sample base + i, against the chunk's candidate
//     obstacles.
// The two kinds share one check-point loop and one loop-exit block through `sample_pose`, which hands
// back the pose of a sample index either from the lanes (shuffle) or by direct evaluation.
__device__ void process_edge(const WorldD& w, const WorldD* ws, const ppe_edge* __restrict__ edge,
                             const PreparedEdge* __restrict__ prep, ppe_edge_result* __restrict__ result,
                             const ObstacleD* s_obs, double4* bufA, double4* bufB, double* pe, TimeTable* tt, int* rel, double* cp_pose,
                             int lane) {
    const ppe_config& cfg = w.cfg;
    const double W = cfg.ribbon_width;
    const double inc = cfg.collision_checking_increment;
    const int cap = w.ribbon_cap;
#ifdef PPE_K2B_PROFILE
    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    PROF_T(t_edge);

    // stage the prepared record (48 doubles) and the parent's ribbons into shared memory
    __syncwarp();
    pe[lane] = prep->v[lane];
    if (lane < kPrepDoubles - 32) pe[32 + lane] = prep->v[32 + lane];
    const int set = edge->ribbon_set;
    int status = PPE_EDGE_OK;
    int nr = 0;
    double cct = -1;
    if (set >= 0 && set < w.n_sets) {
        nr = w.set_count[set];
        cct = w.set_cct[set];
        const double4* src_ribbons = w.ribbons + w.set_offset[set];
        if (nr > cap) { status = PPE_EDGE_ERR_RIBBON_CAPACITY; nr = 0; }
        for (int r = lane; r < nr; r += 32) bufA[r] = src_ribbons[r]; // v->m_RibbonManager = start->m_RibbonManager
    } else {
        status = PPE_EDGE_ERR_RIBBON_CAPACITY;
    }
    __syncwarp();
    if (status == PPE_EDGE_OK && pe[kStatus] != 0.0) status = (int)pe[kStatus];
    bool tame;
    bool any_short; // some ribbon of the list is short enough for cover() to erase it wherever the point is
    {
        bool t_ = fabs(pe[kX0]) + fabs(pe[kLength]) < 1e7 && fabs(pe[kY0]) + fabs(pe[kLength]) < 1e7 && fabs(edge->src[0]) < 1e7 &&
                  fabs(edge->src[1]) < 1e7;
        bool s_ = false;
        for (int r = lane; r < nr; r += 32) {
            const RibbonD rb = load_ribbon(bufA + r);
            t_ = t_ && coords_tame(rb);
            s_ = s_ || ribbon_covered(rb, true, W);
        }
        tame = __all_sync(kFull, t_);
        any_short = __any_sync(kFull, s_);
    }

    const double src_t = edge->src[4];
    const bool cov = edge->coverage_allowed != 0;
    double4* cur = bufA;
    double4* alt = bufB;
    bool modified = false, overflow = false;
    bool sample_fault = pe[kSampleFault] != 0.0;

    // ---- loop state (Edge.cpp:86-120) -----------------------------------------------------------------------
    double endTime = fmin(w.horizon_end, pe[kWEnd]);
    int ribbonsDoneTime = -1;
    const bool startedDone = (nr == 0);
    double penalty = 0;
    bool infeasible = src_t >= endTime;
    const double dt = w.dt;
    const double t0 = pe[kT0];
    int next_cp = 0; // sample index of the next ribbon check-point (toCoverDistance starts at 0)
    // `intermediate` and `lastHeading` when the loop exits (Edge.cpp:96,174); before any sample: the source state
    double P_x = edge->src[0], P_y = edge->src[1], P_h = edge->src[2], lastHeading = edge->src[2], t_exit = t0;
    int n_samples = 0, n_cp = 0, n_culled = 0;
    {
        const double span = (endTime - t0) / dt;
        if (!((dt > 0) && (span < (double)kMaxSamples || !(t0 < endTime)))) status = PPE_EDGE_ERR_END_SAMPLE;
    }

    if (status == PPE_EDGE_OK) {
        if (prep->n_runs > 0) { // built by K2a
            if (lane < kPrepRuns) { tt->base[lane] = prep->run_base[lane]; tt->D[lane] = prep->run_D[lane]; tt->i0[lane] = prep->run_i0[lane]; }
            if (lane == 0) { tt->n = prep->n_runs; tt->i_after = prep->i_after; tt->t_after = prep->t_after; tt->overflow = 0; }
            __syncwarp();
        } else {
            time_table_build(tt, t0, dt, endTime, lane);
        }
        const double w_speed = pe[kWSpeed];
        const double rad_max = 0.5 * kChunk * inc * 1.001 + 1e-6; // the reach the safe map was dilated for
        const bool probing = !tt->overflow && (w_speed > 0) && (w_speed * dt <= inc * 1.0005);
        // obstacles that can matter anywhere on this edge (lanes over obstacles): all executed samples lie within
        // half the travelled arc length and half the time span of the pose at the middle time
        unsigned long long edge_mask = w.n_obs >= 64 ? ~0ull : ((1ull << w.n_obs) - 1ull);
        if (probing && w.obs_kind != kObsNone && w.n_obs > 0 && w.obs_cull_ok) {
            const double t_c = 0.5 * (t0 + endTime), half_T = (0.5 * (endTime - t0)) * (1 + 1e-9) + 1e-9;
            double cx_, cy_, ca_;
            bool it_;
            const bool okc = pose_eval_ref(pe, t_c, &cx_, &cy_, &ca_, &it_) && it_ && (t0 < endTime);
            const double reach_e = half_T * w_speed * (1 + 1e-9) + 1e-3;
            float b0 = 1.f, b1 = 0.f;
            if (okc) {
                b0 = lane < w.n_obs ? obstacle_bound(w.obs_kind, s_obs[lane], cx_, cy_, t_c, reach_e, half_T) : 0.f;
                b1 = lane + 32 < w.n_obs ? obstacle_bound(w.obs_kind, s_obs[lane + 32], cx_, cy_, t_c, reach_e, half_T) : 0.f;
            }
            const unsigned lo = __ballot_sync(kFull, b0 >= kCandidateBound), hi = __ballot_sync(kFull, b1 >= kCandidateBound);
            if (okc) edge_mask = ((unsigned long long)hi << 32) | lo;
        }
        int probe_base = -1;          // chunk 0 of the last probe pass; -1: no valid probe results
        bool p_safe = false;          // this lane's chunk of that pass: proved clean
        unsigned long long p_mask = ~0ull;
        int run = 0;                  // time-table run of sample `base`

        unsigned p_clean = 0;         // ballot of p_safe over the 32 chunks of that pass
        for (int base = 0;; base += kChunk) {
            if (base > kMaxSamples) { status = PPE_EDGE_ERR_END_SAMPLE; break; }
            // ---- what the probe pass knows about this chunk (one pass covers 32 chunks) ---------------------
            bool clean = false;
            unsigned long long omask = ~0ull;
            bool use_mask = false;
            if (probing) {
                int m = probe_base >= 0 ? (base - probe_base) / kChunk : 32;
                if (m >= 32) {
                    PROF_T(t_p);
                    probe_chunks(ws, pe, tt, s_obs, base, lane, endTime, rad_max, edge_mask, &p_safe, &p_mask);
                    PROF_ADD(2, t_p);
                    p_clean = __ballot_sync(kFull, p_safe);
                    probe_base = base;
                    m = 0;
                }
                // a run of clean chunks with no check-point inside is skipped in one step
                const unsigned dirty = ~(p_clean >> m);
                int nskip = dirty ? __ffs(dirty) - 1 : 32;          // clean chunks from m on (bit 32 - m of `dirty` is set for m > 0)
                const int to_cp = (next_cp - base) / kChunk;        // whole chunks before the next check-point's chunk
                nskip = nskip < to_cp ? nskip : to_cp;
                if (nskip > 0) {
                    n_samples += nskip * kChunk;
                    n_culled += nskip * kChunk;
                    base += (nskip - 1) * kChunk;
                    continue;
                }
                clean = (p_clean >> m) & 1u;
                omask = shfl_u64(p_mask, m);
                use_mask = w.obs_cull_ok != 0;
            }
            while (run + 1 < tt->n && base >= tt->i0[run + 1]) run++;

            // ---- lane data: every chunk that gets here is evaluated, lane i = sample base + i --------------------
            // (a clean chunk only gets here because a check-point falls into it: its lanes need times and poses,
            // not map or obstacle look-ups)
            double x, y, ang;
            bool valid, blocked = false;
            int limit = kChunk, fstop = kChunk;
            unsigned m_stop = 0;
            PROF_T(t_l);
            const double t_i = time_at_from(tt, base + lane, run, dt);
            valid = clean || (t_i < endTime);
            {
                bool in_time;
                const bool sample_ok = pose_eval(pe, t_i, &x, &y, &ang, &in_time);
                if (!clean) {
                    if (valid && in_time && sample_ok) blocked = map_blocked(w, x, y);
                    const unsigned m_valid = __ballot_sync(kFull, valid);
                    m_stop = __ballot_sync(kFull, valid && (!in_time || blocked));
                    if (__any_sync(kFull, valid && in_time && !sample_ok)) sample_fault = true;
                    const int nvalid = __popc(m_valid); // valid lanes form a prefix (times increase)
                    fstop = m_stop ? (__ffs(m_stop) - 1) : kChunk;
                    limit = nvalid < fstop ? nvalid : fstop; // samples [base, base + limit) execute the full loop body
                }
            }

            PROF_ADD(3, t_l);
            // ---- ribbon check-points of this chunk, in order (Edge.cpp:153-172) ---------------------------------
            int n_rel = -2; // relevant-ribbon list of this chunk: -2 not built yet, -1 too many, else count
            if (next_cp < base + limit) {
                // the chunk's poses go to shared memory once (x, y, heading per sample): the check-point loop below is
                // uniform scalar work and reads them as broadcasts instead of shuffling them out of the lanes one by one
                __syncwarp();
                cp_pose[lane] = x; cp_pose[32 + lane] = y; cp_pose[64 + lane] = heading_of(ang);
                __syncwarp();
            }
            while (next_cp < base + limit) {
                const int l = next_cp - base;
                if (nr == 0 && cct != -1 && !(cct + cfg.time_minimum < endTime)) {
                    // coverage is complete and the end time has settled: every remaining executed sample of the
                    // chunk is a check-point that only refreshes ribbonsDoneTime = (int) t (Edge.cpp:163-170)
                    n_cp += limit - l;
                    ribbonsDoneTime = (int)__shfl_sync(kFull, t_i, limit - 1);
                    next_cp = base + limit;
                    break;
                }
                // ---- run along ONE ribbon: consecutive check-points (toCoverDistance stays 0 inside a ribbon) that all
                // fall on the single relevant ribbon of this chunk and leave the list's structure alone.  The ribbon lives in
                // registers for the whole run; each step is the arithmetic of fast_checkpoint for that ribbon, nothing else.
                // The first step that is not of this kind leaves the run untouched and goes through the code below.
                if (n_rel == 1 && tame && !any_short && (cov || l > 0)) {
                    PROF_T(t_r);
                    const int ri = rel[0];
                    RibbonD rb = load_ribbon(cur + ri);
                    bool dirty = false;
                    int q = l;
#pragma unroll 1
                    for (; q < limit; q++) {
                        const double rx = cp_pose[q], ry = cp_pose[32 + q];
                        const bool do_cover = cov || cp_pose[64 + q - 1] == cp_pose[64 + q]; // lastHeading == heading, Edge.cpp:159
                        double px, py;
                        ribbon_projection(rb, rx, ry, &px, &py);
                        if (!ribbon_contains_projection(rb, px, py)) break;
                        const double d = ribbon_distance(rb, rx, ry);
                        if (!(d < W)) break; // not inside: minDistanceFrom needs the end points
                        if (do_cover && d < W / 2.0) {
                            const RibbonD piece = {rb.sx, rb.sy, px, py};
                            const RibbonD rest = {px, py, rb.ex, rb.ey};
                            const bool keep_piece = !ribbon_covered(piece, true, W), keep_rest = !ribbon_covered(rest, true, W);
                            if (keep_piece == keep_rest) break; // the list's structure changes
                            if (keep_rest) { if (px != rb.sx || py != rb.sy) modified = true; rb.sx = px; rb.sy = py; }
                            else { modified = true; rb.ex = px; rb.ey = py; }
                            dirty = true;
                        }
                        n_cp++;
#ifdef PPE_K2B_PROFILE
                        prof[8] += 1;
#endif
                    }
                    if (dirty) {
                        if (lane == 0) cur[ri] = pack_ribbon(rb.sx, rb.sy, rb.ex, rb.ey);
                        __syncwarp();
                    }
                    PROF_ADD(4, t_r);
                    if (q > l) { next_cp = base + q; continue; } // toCoverDistance was 0 at every step: the next sample is a check-point
                }
                const double cx = cp_pose[l], cy = cp_pose[32 + l];
                const double ct = __shfl_sync(kFull, t_i, l);
                const double ch = cp_pose[64 + l];
                n_cp++;
                bool do_cover = cov;
                if (!cov) { // lastHeading == intermediate.heading(), Edge.cpp:159: heading of the sample before
                    double ph;
                    if (next_cp == 0) {
                        ph = edge->src[2];
                    } else if (l > 0) {
                        ph = cp_pose[64 + l - 1];
                    } else { // the last sample of the previous chunk
                        double px_, py_, pa_;
                        bool it_;
                        pose_eval_ref(pe, time_at(tt, next_cp - 1, dt), &px_, &py_, &pa_, &it_);
                        ph = heading_of(pa_);
                    }
                    do_cover = (ph == ch);
                }
                double toCover;
                bool changed = false;
                bool handled = false;
                PROF_T(t_c);
                if (tame && !any_short && nr > 0) {
                    if (n_rel == -2) { // relevant ribbons of this chunk: bounding box of the chunk's executed poses
                        const bool in = lane < limit;
                        const double bx0 = warp_min(in ? x : DBL_MAX), bx1 = warp_max(in ? x : -DBL_MAX);
                        const double by0 = warp_min(in ? y : DBL_MAX), by1 = warp_max(in ? y : -DBL_MAX);
                        n_rel = relevant_ribbons(cur, nr, bx0, bx1, by0, by1, W, lane, rel);
                    }
                    if (n_rel >= 0) {
                        const int how = n_rel > 0 ? fast_checkpoint(cur, rel, n_rel, cx, cy, W, do_cover, tame, lane, &changed) : 2;
                        if (how == 1) { handled = true; toCover = 0.0; }
                        else if (how == 2) { handled = true; toCover = warp_nearest_endpoint(cur, nr, cx, cy, lane); }
                    }
                }
                if (handled) {
                    if (changed) modified = true;
                    PROF_ADD(4, t_c);
                    PROF_INC(8);
                } else {
                    const int nn = warp_checkpoint(cur, alt, nr, cap, cx, cy, W, do_cover, tame, lane, &toCover, &changed, &overflow);
                    if (do_cover) any_short = false; // cover() erased every ribbon short enough, contained or not
                    if (changed) {
                        double4* tmp = cur; cur = alt; alt = tmp;
                        nr = nn;
                        modified = true;
                        n_rel = -2; // indices moved
                    }
                    PROF_ADD(5, t_c);
                    PROF_INC(9);
                }
                if (nr == 0) {
                    if (cct == -1) cct = ct;
                    ribbonsDoneTime = (int)ct;
                    const double newEnd = fmin(endTime, cct + cfg.time_minimum);
                    if (newEnd < endTime) {
                        endTime = newEnd;
                        probe_base = -1; // the end moved: probe results are stale
                        valid = t_i < endTime;
                        const int nvalid = __popc(__ballot_sync(kFull, valid));
                        limit = nvalid < fstop ? nvalid : fstop;
                        if (limit < l + 1) limit = l + 1; // the check-point's own iteration has already run
                    }
                }
                next_cp = next_cp + 1 + skip_count(toCover, inc, kSkipCap);
            }

            // ---- dynamic-obstacle penalty of the executed iterations (Edge.cpp:150-151), in sample order
            PROF_T(t_o);
            if (!clean && w.obs_kind != kObsNone && w.n_obs > 0 && (!use_mask || omask != 0)) {
                double p = 0;
                if (lane < limit)
                    p = collision_exists(w.obs_kind, w.n_obs, s_obs, omask, use_mask, x, y, t_i) * cfg.collision_penalty_factor;
                if (__any_sync(kFull, p != 0)) {
                    for (int q = 0; q < limit; q++) penalty += __shfl_sync(kFull, p, q);
                }
            }

            PROF_ADD(6, t_o);
            n_samples += limit;
            n_culled += clean ? limit : 0;
            if (limit < kChunk) {
                // the loop ends inside this chunk.  The iteration at sample `limit` was entered and broke
                // out only if that sample is still inside the (possibly truncated) end time.
                bool stopped = false;
                t_exit = __shfl_sync(kFull, t_i, limit);
                if (!clean) {
                    const bool stop_lane_valid = __shfl_sync(kFull, (int)valid, limit) != 0;
                    stopped = (fstop == limit) && ((m_stop >> fstop) & 1u) && stop_lane_valid;
                }
                // pose of the last executed sample (`intermediate`), or the source state when none ran
                if (limit > 0) {
                    P_x = __shfl_sync(kFull, x, limit - 1); P_y = __shfl_sync(kFull, y, limit - 1);
                    P_h = heading_of(__shfl_sync(kFull, ang, limit - 1));
                    lastHeading = P_h;
                } else if (base > 0) {
                    double la;
                    bool it_;
                    pose_eval_ref(pe, time_at(tt, base - 1, dt), &P_x, &P_y, &la, &it_);
                    P_h = heading_of(la);
                    lastHeading = P_h;
                }
                if (stopped) {
                    infeasible = true;
                    n_samples += 1; // the breaking iteration was entered
                    if (__shfl_sync(kFull, (int)blocked, limit) != 0) { // `intermediate` holds the blocked sample
                        P_x = __shfl_sync(kFull, x, limit); P_y = __shfl_sync(kFull, y, limit);
                        P_h = heading_of(__shfl_sync(kFull, ang, limit));
                    }
                }
                break;
            }
        }
    }

    // ---- truncated end state (Edge.cpp:177-179) and the final cover (:182-191) ----------------------------------------
    PROF_T(t_tail);
    double ex = 0, ey = 0, eh = 0;
    if (status == PPE_EDGE_OK) {
        double ea;
        bool in_time;
        const bool sample_ok = pose_eval_ref(pe, endTime, &ex, &ey, &ea, &in_time);
        eh = heading_of(ea);
        if (!in_time) status = PPE_EDGE_ERR_END_SAMPLE; // sample(end state) throws, DubinsWrapper.cpp:30-35
        if (!sample_ok) sample_fault = true;
    }
    if (status == PPE_EDGE_OK) {
        if (cov || lastHeading == P_h) {
            bool changed = false;
            double unused;
            const int nn = warp_checkpoint(cur, alt, nr, cap, P_x, P_y, W, true, tame, lane, &unused, &changed, &overflow);
            if (changed) {
                double4* tmp = cur; cur = alt; alt = tmp;
                nr = nn;
                modified = true;
            }
        }
        if (nr == 0) {
            if (cct == -1) cct = t_exit;
            ribbonsDoneTime = (int)t_exit;
        }
    }

    // ---- cost, g, h and the result record (Edge.cpp:193-203) ------------------------------------------------------
    double true_cost = 0, g = 0, h = 0;
    if (status == PPE_EDGE_OK) {
        const double netTime = endTime - src_t;
        double T = fmax(netTime - (nr == 0 ? (endTime - (double)ribbonsDoneTime) : 0.0), 0.0);
        if (startedDone) T = 0;
        true_cost = T * cfg.time_penalty_factor + penalty;
        g = edge->src_g + true_cost;                                  // Vertex::setCurrentCost
        if (cfg.heuristic == PPE_H_MAX_DISTANCE) {                    // Vertex::computeApproxToGo
            const double d = nr == 0 ? 0.0 : warp_max_distance(cur, nr, ex, ey, W, lane, (double*)alt);
            h = d / cfg.max_speed * cfg.time_penalty_factor;
        } else {
            h = tsp_heuristic_or_unset(cfg, cur, nr, ex, ey); // point-robot TSP variants; -1: left to the host
        }
        if (overflow) status = PPE_EDGE_ERR_RIBBON_CAPACITY;
        else if (sample_fault) status = PPE_EDGE_ERR_END_SAMPLE;
    }

    // ribbons-after: only edges that changed their parent's set materialise a list
    long long ribbons_offset = -1;
    if (status == PPE_EDGE_OK && modified) {
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd(w.out_count, (unsigned long long)nr);
        off = __shfl_sync(kFull, off, 0);
        if (off + (unsigned long long)nr <= w.out_cap) {
            for (int r = lane; r < nr; r += 32) w.out_ribbons[off + r] = cur[r];
            ribbons_offset = (long long)off;
        } else {
            status = PPE_EDGE_ERR_RIBBON_CAPACITY;
        }
    }
    __syncwarp();
    if (lane == 0) {
        ppe_edge_result* r = result;
        const bool ok = status == PPE_EDGE_OK;
        r->true_cost = ok ? true_cost : 0.0;
        r->collision_penalty = ok ? penalty : 0.0;
        r->approx_cost = ok ? pe[kApprox] : 0.0;
        r->end[0] = ok ? ex : 0.0; r->end[1] = ok ? ey : 0.0; r->end[2] = ok ? eh : 0.0;
        r->end[3] = ok ? pe[kWSpeed] : 0.0; r->end[4] = ok ? endTime : 0.0;
        r->g = ok ? g : 0.0;
        r->h = ok ? h : 0.0;
        r->coverage_completed_time = ok ? cct : 0.0;
        r->path_qi[0] = ok ? pe[kX0] : 0.0; r->path_qi[1] = ok ? pe[kY0] : 0.0; r->path_qi[2] = ok ? pe[kYaw0] : 0.0;
        r->path_param[0] = ok ? pe[kParam0] : 0.0; r->path_param[1] = ok ? pe[kParam1] : 0.0;
        r->path_param[2] = ok ? pe[kParam2] : 0.0;
        r->path_rho = ok ? pe[kRho] : 0.0;
        r->w_speed = ok ? pe[kWSpeed] : 0.0;
        r->w_start_time = ok ? pe[kWStart] : 0.0;
        r->w_end_time = ok ? endTime : 0.0; // updateEndTime (Edge.cpp:179)
        r->ribbons_offset = ribbons_offset;
        r->path_type = ok ? (int)pe[kType] : 0;
        r->infeasible = infeasible ? 1 : 0;
        r->status = status;
        r->n_samples = ok ? n_samples : 0;
        r->n_checkpoints = ok ? n_cp : 0;
        r->n_ribbons_after = ok ? nr : 0;
        r->ribbons_changed = (ok && modified) ? 1 : 0;
        r->reserved = n_culled & 0xffffff; // instrumentation: executed samples the probe pass proved clean (never evaluated)
    }
#ifdef PPE_K2B_PROFILE
    PROF_ADD(7, t_tail);
    PROF_ADD(1, t_edge);
    prof[0] = 1;
    if (lane == 0)
        for (int q = 0; q < 10; q++) atomicAdd(&g_k2b_prof[q], (unsigned long long)prof[q]);
#endif
}

// ---- K2t: the thread walker -------------------------------------------------------------------------------------------
// Once the probe pass removes ~95 % of the per-sample work, what is left of a typical edge is SEQUENTIAL: a handful of
// ribbon check-points, the loop exit, the end state, cost and heuristic.  A warp spends 32 lanes on that uniform work;
// one thread per edge spends one, and 32 edges share every instruction fetch.  K2t therefore walks each edge in one
// thread -- the same loop, the same device functions (time table, pose_eval, probe_one, collision_exists, skip_count,
// Ribbon primitives), chunks proved clean skipped, the others evaluated sample by sample -- as long as the edge stays
// SIMPLE: the parent's ribbon list is never changed (so it is read in place from the interned pool and `cover` reduces
// to "would it change anything?") and the check-point budget is not exceeded.  Anything else (a cover that splits or
// erases a ribbon, coverage already complete, a non-OK status, unusual time scales) puts the edge on the heavy list and
// the warp walker K2b evaluates it from scratch; both paths produce the same bits.
constexpr int kThreadDirtyCap = 64;   // upper limit of K2Tuning::dirty_budget (size of the per-thread mask array): every chunk of an edge
constexpr int kDeepDirtyBudget = 8;   // K2c evaluates its dirty chunks sample by sample in the edge's own thread

struct SeqTime { // cursor over the prepared run table; the current run [lo, hi) and its coefficients stay in registers
    const PreparedEdge* p;
    int r, n_runs, lo, hi;
    double base, D;
    __device__ __forceinline__ explicit SeqTime(const PreparedEdge* prep) : p(prep), r(0), n_runs(prep->n_runs) { load(); }
    __device__ __forceinline__ void load() {
        lo = p->run_i0[r];
        hi = r + 1 < n_runs ? p->run_i0[r + 1] : INT_MAX;
        base = p->run_base[r];
        D = p->run_D[r];
    }
    __device__ __forceinline__ double at(int i) {
        if (i >= hi || (i < lo && r > 0)) { // the run with run_i0[r] <= i < run_i0[r + 1], clamped to the table
            while (r + 1 < n_runs && i >= p->run_i0[r + 1]) r++;
            while (r > 0 && i < p->run_i0[r]) r--;
            load();
        }
        return base + (double)(i - lo) * D;
    }
};

// minDistanceFrom + "would cover(x, y, strict) change the list?" over the parent's ribbons, in place
// (`any_short`: the per-set invariant "some ribbon is short enough for cover() to erase it wherever the point is")
// `box`: the grown bounding boxes of the same ribbons (WorldD::boxes), what ribbon_may_contain would compute
__device__ __forceinline__ double seq_checkpoint(const double4* __restrict__ rib, const double4* __restrict__ box, int nr, double x, double y,
                                                 double W, bool tame, bool any_short, bool* would_change) {
    double mn = DBL_MAX;
    bool inside = false, change = any_short;
#pragma unroll 1
    for (int r = 0; r < nr; r++) {
        const RibbonD rb = load_ribbon(rib + r);
        const double4 bx = box[r];
        bool contained = false;
        const bool outside = x < bx.x || x > bx.y || y < bx.z || y > bx.w;
        if (!(outside && tame)) {
            double px, py;
            ribbon_projection(rb, x, y, &px, &py);
            if (ribbon_contains_projection(rb, px, py)) {
                const double d = ribbon_distance(rb, x, y);
                inside = inside || (d < W);
                contained = d < W / 2.0;
            }
        }
        // cover leaves a ribbon as it is iff it is not (strictly) contained and not already short enough to erase
        change = change || contained;
        const double dStart = point_distance_sq(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance_sq(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart); // squared
    }
    *would_change = change;
    return inside ? 0.0 : sqrt(mn);
}

// the same with the list's length sum taken from the per-set invariant (the unchanged parent list)
__device__ __forceinline__ double seq_max_distance_presummed(const double4* __restrict__ rib, int nr, double x, double y, double sumLength) {
    double mn = DBL_MAX, mx = 0;
#pragma unroll 1
    for (int r = 0; r < nr; r++) {
        const RibbonD rb = load_ribbon(rib + r);
        const double dStart = point_distance_sq(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance_sq(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart); // squared
        mx = fmax(fmax(mx, dEnd), dStart);
    }
    return fmax(sumLength + sqrt(mn), sqrt(mx));
}

// ---- K2c: the deep thread walker ----------------------------------------------------------------------------------------
// Edges that K2t hands over because they COVER a ribbon (a check-point on every sample while the boat is on it, up to
// 1500 in a row) are strictly sequential scalar work; the warp walker executes that chain redundantly in all 32 lanes.
// K2c walks them one thread per edge -- 32 chains per warp instruction -- as long as the ribbon list keeps its STRUCTURE:
// cover() only moves the start (or the end) of the ribbon the boat is on, which is kept in a two-entry per-thread override
// table over the parent's interned list; per 32-sample chunk only the ribbons whose bounding box can be reached are
// looked at.  A split that inserts or erases a ribbon, a third modified ribbon, a list with a ribbon short enough to be
// erased wherever the point is, or more per-sample work than the dirty-chunk budget sends the edge on to the warp walker.
// The ribbon list of a K2c edge = the parent's interned list + a small delta: up to two ribbons whose start / end moved
// (or that were erased), up to two pieces a split inserted in front of a parent ribbon.  List order: for parent index r,
// the pieces inserted before r (in insertion order), then r itself unless erased.
struct RibbonDelta {
    double4 val[2];   // overrides
    double4 ins[2];   // inserted pieces
    int idx[2];       // parent index of override k
    int ins_pos[2];   // parent index piece k stands in front of
    int n, n_ins;
    unsigned erased;  // bit k: override k means "erased"
};
constexpr int kDeepRelCap = 4;

// parent ribbon r as the delta sees it; *gone = erased
__device__ __forceinline__ RibbonD load_ribbon_ov(const double4* __restrict__ rib, int r, const RibbonDelta& dl, bool* gone) {
    *gone = false;
    if (dl.n > 0 && r == dl.idx[0]) { *gone = (dl.erased & 1u) != 0; return RibbonD{dl.val[0].x, dl.val[0].y, dl.val[0].z, dl.val[0].w}; }
    if (dl.n > 1 && r == dl.idx[1]) { *gone = (dl.erased & 2u) != 0; return RibbonD{dl.val[1].x, dl.val[1].y, dl.val[1].z, dl.val[1].w}; }
    return load_ribbon(rib + r);
}

// ribbons (of the parent's list: the delta only shrinks them or cuts pieces out of them) that a pose within `reach` of
// (x, y) could be contained in
__device__ __forceinline__ int deep_relevant(const double4* __restrict__ rib, int nr, double x, double y, double reach, double W, int* rel) {
    const double grow = W * (1 + 1e-9) + 2e-3 + reach;
    int n = 0;
#pragma unroll 1
    for (int r = 0; r < nr; r++) {
        const RibbonD rb = load_ribbon(rib + r);
        const bool near = !(x < fmin(rb.sx, rb.ex) - grow || x > fmax(rb.sx, rb.ex) + grow || y < fmin(rb.sy, rb.ey) - grow ||
                            y > fmax(rb.sy, rb.ey) + grow);
        if (near) {
            if (n < kDeepRelCap) rel[n] = r;
            n++;
        }
    }
    return n <= kDeepRelCap ? n : -1;
}

// Ribbon::contains on one ribbon: non-strict (d < W) -> *inside, strict (d < W / 2) -> returns true with the projection
__device__ __forceinline__ bool deep_contains(const RibbonD& rb, double x, double y, double W, bool* inside, double* px, double* py) {
    ribbon_projection(rb, x, y, px, py);
    if (!ribbon_contains_projection(rb, *px, *py)) return false;
    const double d = ribbon_distance(rb, x, y);
    *inside = *inside || (d < W);
    return d < W / 2.0;
}

// One check-point (minDistanceFrom, then cover when `do_cover`) on the relevant ribbons.  Returns false when the delta
// cannot hold the outcome (nothing has been modified then): a third moved ribbon, a third inserted piece, a split of an
// inserted piece, two ribbons hit at once, or the list running empty (coverage completion is the warp walker's).
__device__ __forceinline__ bool deep_checkpoint(const double4* __restrict__ rib, int nr, RibbonDelta& dl, const int* rel, int n_rel,
                                                double x, double y, double W, bool do_cover, double* to_cover, bool* modified) {
    bool inside = false;
    int upd = -1, kind = 0; // kind 1: move start (rest survives), 2: move end (piece survives), 3: insert piece + move start, 4: erase
    double4 nv = make_double4(0, 0, 0, 0), piece_v = make_double4(0, 0, 0, 0);
    bool ch = false;
#pragma unroll 1
    for (int q = 0; q < n_rel; q++) {
        const int r = rel[q];
        double px, py;
        for (int k = 0; k < dl.n_ins; k++) { // pieces a split left in front of r
            if (dl.ins_pos[k] != r) continue;
            const RibbonD pc = {dl.ins[k].x, dl.ins[k].y, dl.ins[k].z, dl.ins[k].w};
            if (deep_contains(pc, x, y, W, &inside, &px, &py) && do_cover) return false; // a piece would be split again
        }
        bool gone;
        const RibbonD rb = load_ribbon_ov(rib, r, dl, &gone);
        if (gone) continue;
        if (deep_contains(rb, x, y, W, &inside, &px, &py) && do_cover) {
            const RibbonD piece = {rb.sx, rb.sy, px, py};
            const RibbonD rest = {px, py, rb.ex, rb.ey};
            const bool keep_piece = !ribbon_covered(piece, true, W), keep_rest = !ribbon_covered(rest, true, W);
            if (upd >= 0) return false; // two ribbons at once
            upd = r;
            if (keep_rest) {
                nv = make_double4(px, py, rb.ex, rb.ey);
                if (keep_piece) { kind = 3; piece_v = make_double4(rb.sx, rb.sy, px, py); ch = true; }
                else { kind = 1; ch = (px != rb.sx) || (py != rb.sy); }
            } else if (keep_piece) {
                kind = 2; nv = make_double4(rb.sx, rb.sy, px, py); ch = true;
            } else {
                kind = 4; ch = true;
            }
        }
    }
    if (upd >= 0) { // cover() changes ribbon `upd` (only reached with `inside`: strict containment implies it)
        int slot = -1;
        if (dl.n > 0 && dl.idx[0] == upd) slot = 0;
        else if (dl.n > 1 && dl.idx[1] == upd) slot = 1;
        const bool need_slot = slot < 0;
        if (need_slot && dl.n >= 2) return false;              // a third modified ribbon
        if (kind == 3 && dl.n_ins >= 2) return false;          // a third inserted piece
        if (kind == 4) {                                       // would the list run empty?
            int left = nr + dl.n_ins - 1;
            for (int k = 0; k < dl.n; k++) if ((dl.erased >> k) & 1u) left--;
            if (left <= 0) return false;
        }
        if (need_slot) { slot = dl.n; dl.n++; dl.idx[slot] = upd; }
        if (kind == 4) dl.erased |= 1u << slot;
        else dl.val[slot] = nv;
        if (kind == 3) { dl.ins[dl.n_ins] = piece_v; dl.ins_pos[dl.n_ins] = upd; dl.n_ins++; }
        *modified = *modified || ch;
    }
    if (inside) { *to_cover = 0.0; return true; }
    // no ribbon contains the point even non-strictly: nearest end point of the current list
    double mn = DBL_MAX;
#pragma unroll 1
    for (int r = 0; r < nr; r++) {
        for (int k = 0; k < dl.n_ins; k++) {
            if (dl.ins_pos[k] != r) continue;
            mn = fmin(fmin(mn, point_distance_sq(dl.ins[k].z, dl.ins[k].w, x, y)), point_distance_sq(dl.ins[k].x, dl.ins[k].y, x, y));
        }
        bool gone;
        const RibbonD rb = load_ribbon_ov(rib, r, dl, &gone);
        if (gone) continue;
        const double dStart = point_distance_sq(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance_sq(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart);
    }
    *to_cover = sqrt(mn);
    return true;
}

// RibbonManager::maxDistance (RibbonManager.cpp:234-248) over a materialised list
__device__ __forceinline__ double seq_max_distance(const double4* __restrict__ rib, int nr, double x, double y, double W) {
    double sumLength = 0, mn = DBL_MAX, mx = 0;
#pragma unroll 1
    for (int r = 0; r < nr; r++) { // list order, as the reference sums
        const RibbonD rb = load_ribbon(rib + r);
        sumLength += sqrt(ribbon_sqlen(rb)) - 2 * W;
        const double dStart = point_distance_sq(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance_sq(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart); // squared
        mx = fmax(fmax(mx, dEnd), dStart);
    }
    return fmax(sumLength + sqrt(mn), sqrt(mx));
}

// The walk of ONE edge by ONE thread: kDeep = false is K2t (simple edges, the ribbon list is only read), kDeep = true is K2c.
template <bool kDeep>
__device__ __forceinline__ void thread_walk(const WorldD& w, const long long n, const long long ei, const ppe_edge* __restrict__ edges,
                                            const PreparedEdge* __restrict__ prepared, ppe_edge_result* __restrict__ results,
                                            unsigned int* __restrict__ heavy_list, unsigned int* __restrict__ heavy_count,
                                            const int dirty_budget, const int cp_budget, const ObstacleD* s_obs, const uint32_t* tile,
                                            const bool live) {
    // `live` = false (K2t only): a lane of the batch's last warp without an edge of its own -- `ei` then names some valid
    // edge that is only read; the lane takes part in the warp-cooperative phase B and writes nothing.
    const ppe_edge* edge = edges + ei;
    const PreparedEdge* prep = prepared + ei;
    const double* pe = prep->v;
    const ppe_config& cfg = w.cfg;
    const double W = cfg.ribbon_width, inc = cfg.collision_checking_increment, dt = w.dt;

    bool skipped = false;
    if (live && pe[kStatus] == (double)PPE_EDGE_SKIPPED) { // empty slot of a frontier batch: nothing to evaluate
        ppe_edge_result* r = results + ei;
        memset(r, 0, sizeof *r);
        r->ribbons_offset = -1;
        r->status = PPE_EDGE_SKIPPED;
        if (kDeep) return;
        skipped = true; // K2t: the lane stays for phase B
    }
    // ---- is this edge simple at all? -------------------------------------------------------------------------------
    const int set = edge->ribbon_set;
    bool heavy = !live || skipped || !(set >= 0 && set < w.n_sets) || pe[kStatus] != 0.0 || pe[kSampleFault] != 0.0 || prep->n_runs <= 0;
    int nr = 0;
    double cct = -1;
    const double4* rib = nullptr;
    const double4* box = nullptr;
    if (!heavy) {
        nr = w.set_count[set];
        cct = w.set_cct[set];
        rib = w.ribbons + w.set_offset[set];
        box = w.boxes + w.set_offset[set];
        heavy = nr <= 0 || nr > w.ribbon_cap; // coverage already complete: every sample is a check-point (warp walker)
    }
    bool tame = fabs(pe[kX0]) + fabs(pe[kLength]) < 1e7 && fabs(pe[kY0]) + fabs(pe[kLength]) < 1e7 && fabs(edge->src[0]) < 1e7 &&
                fabs(edge->src[1]) < 1e7;
    bool any_short = true;
    if (!heavy) {
        any_short = (w.set_tame[set] & 2) != 0;
        tame = tame && (w.set_tame[set] & 1) != 0;
        // K2c keeps the ribbons' structure: a list holding a ribbon short enough for cover() to erase it wherever the
        // point is, or coordinates beyond the bounding-box shortcut's range, is the warp walker's
        if (kDeep) heavy = !tame || (w.set_tame[set] & 2) != 0;
    }
    const double src_t = edge->src[4];
    const bool cov = edge->coverage_allowed != 0;
    const double endTime = fmin(w.horizon_end, pe[kWEnd]);
    const double t0 = pe[kT0];
    const double w_speed = pe[kWSpeed];
    if (!heavy) {
        const double span = (endTime - t0) / dt;
        heavy = !((dt > 0) && (span < (double)kMaxSamples || !(t0 < endTime))) || !((w_speed > 0) && (w_speed * dt <= inc * 1.0005)) ||
                (w.obs_kind != kObsNone && w.n_obs > 0 && !w.obs_cull_ok);
    }

    double penalty = 0;
    bool infeasible = src_t >= endTime;
    int n_samples = 0, n_cp = 0, n_culled = 0;
    double P_x = edge->src[0], P_y = edge->src[1], P_h = edge->src[2], lastHeading = edge->src[2];
    double ex = 0, ey = 0, eh = 0;
    const int status = PPE_EDGE_OK;
    bool long_run = false; // bailed out while covering a ribbon
    RibbonDelta ov;        // K2c: what cover() has done to the parent's list so far
    ov.n = 0; ov.n_ins = 0; ov.erased = 0;
    bool modified = false;
    const int n_valid = prep->pad_[0]; // samples with t_i < endTime
    heavy = heavy || n_valid < 0 || n_valid > 64 * kChunk;

    // The walk is organised in phases that keep the 32 edges of a warp on the same instructions: every thread probes
    // ALL its chunks first (uniform, cheap); then the few chunks that were not proved clean are evaluated -- K2t: by the
    // whole warp, lane i = sample i of the chunk, one (edge, chunk) item after the other; K2c: sample by sample in the
    // edge's own thread -- then the check-points; then the tail.
    SeqTime tm(prep);
    unsigned long long dirty = 0;      // bit c: chunk c must be evaluated
    unsigned long long dmask[kThreadDirtyCap]; // its candidate obstacles, in order of appearance
    int n_dirty = 0;
    bool more_dirty = false;
    int n_exec = n_valid;      // samples that run the full loop body
    bool blocked_exit = false;
    if (!heavy) {
        const double rad_max = 0.5 * kChunk * inc * 1.001 + 1e-6;
        // obstacles that can matter anywhere on this edge
        unsigned long long edge_mask = w.n_obs >= 64 ? ~0ull : ((1ull << w.n_obs) - 1ull);
        if (w.obs_kind != kObsNone && w.n_obs > 0) {
            const double t_c = 0.5 * (t0 + endTime), half_T = (0.5 * (endTime - t0)) * (1 + 1e-9) + 1e-9;
            double cx_, cy_, ca_;
            bool it_;
            if (pose_eval(pe, t_c, &cx_, &cy_, &ca_, &it_) && it_ && (t0 < endTime)) {
                const double reach_e = half_T * w_speed * (1 + 1e-9) + 1e-3;
                edge_mask = 0;
#pragma unroll 1
                for (int i = 0; i < w.n_obs; i++)
                    if (obstacle_bound(w.obs_kind, s_obs[i], cx_, cy_, t_c, reach_e, half_T) >= kCandidateBound) edge_mask |= 1ull << i;
            }
        }

        // ---- phase A: probe every chunk (the last one over its valid samples only) -------------------------------------
        const int n_chunks = (n_valid + kChunk - 1) / kChunk;
#pragma unroll 1
        for (int c = 0; c < n_chunks; c++) {
            const int c0 = c * kChunk;
            const int last = (c0 + kChunk - 1 < n_valid - 1) ? c0 + kChunk - 1 : n_valid - 1;
            const double t_first = tm.at(c0), t_mid = tm.at(c0 + (last - c0 + 1) / 2), t_last = tm.at(last);
            unsigned long long om;
            if (!probe_one(w, pe, t_first, t_mid, t_last, s_obs, endTime, rad_max, edge_mask, &om, tile)) {
                // more dirty chunks than the budget only matter if the loop gets that far (an edge that runs into a
                // blocked area is dirty from there on, but stops at its first blocked sample)
                if (n_dirty == dirty_budget) { more_dirty = true; break; }
                dirty |= 1ull << c;
                dmask[n_dirty++] = om;
            }
        }
    }

    // ---- phase B: the dirty chunks in order (Edge.cpp:126-151) until the loop would stop -------------------------------------
    if (!kDeep) {
        // K2t: the warp takes the (edge, chunk) items of its 32 threads one after the other, lane i evaluating sample i of
        // the chunk -- 32 samples in the time one thread needs for one.  Every lane of the warp gets here (live or not,
        // simple or not); the owner of an item keeps its outcome, in the order the sequential loop would have met it.
        const int lane = threadIdx.x & 31;
        unsigned long long todo = dirty;
#pragma unroll 1
        for (int j = 0; j < dirty_budget; j++) {
            const bool want = !heavy && n_exec == n_valid && j < n_dirty;
            unsigned pending = __ballot_sync(kFull, want);
            const int c_own = want ? __ffsll((long long)todo) - 1 : 0;
            const unsigned long long om_own = want ? dmask[j] : 0ull;
            while (pending) {
                const int b = __ffs(pending) - 1;
                pending &= pending - 1;
                const PreparedEdge* pp = reinterpret_cast<const PreparedEdge*>(shfl_u64((unsigned long long)prep, b));
                const unsigned long long omask = shfl_u64(om_own, b);
                const int c0 = __shfl_sync(kFull, c_own, b) * kChunk;
                const int nv = __shfl_sync(kFull, n_valid, b);
                const int last = (c0 + kChunk - 1 < nv - 1) ? c0 + kChunk - 1 : nv - 1;
                const int i = c0 + lane;
                const bool valid = i <= last;
                double x = 0, y = 0, ang = 0, pen = 0;
                bool in_time = true, sample_ok = true, blocked = false;
                if (valid) {
                    SeqTime tb(pp);
                    const double t_i = tb.at(i);
                    sample_ok = pose_eval(pp->v, t_i, &x, &y, &ang, &in_time);
                    if (in_time && sample_ok) {
                        blocked = map_blocked(w, x, y, tile);                                     // Edge.cpp:144-147
                        if (!blocked && w.obs_kind != kObsNone && w.n_obs > 0 && omask != 0)
                            pen = collision_exists(w.obs_kind, w.n_obs, s_obs, omask, true, x, y, t_i) * cfg.collision_penalty_factor;
                    }
                }
                const unsigned m_nit = __ballot_sync(kFull, valid && !in_time);                   // sample() throws, Edge.cpp:126-133
                const unsigned m_bad = __ballot_sync(kFull, valid && in_time && !sample_ok);      // stale-pose corner: warp walker
                const unsigned m_blk = __ballot_sync(kFull, valid && in_time && sample_ok && blocked);
                const unsigned m_stop = m_nit | m_bad | m_blk;
                const int s_stop = m_stop ? __ffs(m_stop) - 1 : 0;
                const int limit = m_stop ? s_stop : last - c0 + 1; // samples [c0, c0 + limit) run the full loop body
                double acc = __shfl_sync(kFull, penalty, b);      // the penalty sum, in sample order
                if (__any_sync(kFull, lane < limit && pen != 0.0)) {
                    for (int q = 0; q < limit; q++) acc += __shfl_sync(kFull, pen, q);
                }
                const double sx_ = __shfl_sync(kFull, x, s_stop), sy_ = __shfl_sync(kFull, y, s_stop), sa_ = __shfl_sync(kFull, ang, s_stop);
                if (lane == b) {
                    penalty = acc;
                    todo &= todo - 1;
                    if (m_stop) {
                        const unsigned bit = 1u << s_stop;
                        if (m_bad & bit) {
                            heavy = true;
                        } else {
                            infeasible = true;
                            n_exec = c0 + s_stop;
                            if (m_blk & bit) { blocked_exit = true; P_x = sx_; P_y = sy_; P_h = heading_of(sa_); } // `intermediate` holds the blocked sample
                        }
                    }
                }
            }
        }
    }
    if (!heavy) {
        if (kDeep) {
            int j = 0;
            unsigned long long todo = dirty;
#pragma unroll 1
            while (todo && n_exec == n_valid && !heavy) {
                const int c = __ffsll((long long)todo) - 1;
                todo &= todo - 1;
                const unsigned long long omask = dmask[j++];
                const int c0 = c * kChunk;
                const int last = (c0 + kChunk - 1 < n_valid - 1) ? c0 + kChunk - 1 : n_valid - 1;
#pragma unroll 1
                for (int i = c0; i <= last; i++) {
                    const double t_i = tm.at(i);
                    double x, y, ang;
                    bool in_time;
                    const bool sample_ok = pose_eval(pe, t_i, &x, &y, &ang, &in_time);
                    if (!in_time) { infeasible = true; n_exec = i; break; }                        // sample() throws, Edge.cpp:126-133
                    if (!sample_ok) { heavy = true; break; }                                      // stale-pose corner: warp walker
                    if (map_blocked(w, x, y, tile)) {                                             // Edge.cpp:144-147
                        infeasible = true;
                        n_exec = i;
                        blocked_exit = true;
                        P_x = x; P_y = y; P_h = heading_of(ang); // `intermediate` holds the blocked sample
                        break;
                    }
                    if (w.obs_kind != kObsNone && w.n_obs > 0 && omask != 0)
                        penalty += collision_exists(w.obs_kind, w.n_obs, s_obs, omask, true, x, y, t_i) * cfg.collision_penalty_factor;
                }
            }
        }
        {
            // per-sample work beyond the budget belongs to the warp walker (lanes = samples)
            if (more_dirty && n_exec == n_valid) heavy = true;
            n_samples = n_exec + (infeasible && n_exec < n_valid ? 1 : 0); // the breaking iteration was entered
            // instrumentation: executed samples that lie in chunks proved clean (the last chunk may be partial)
            const int chunks_run = (n_exec + kChunk - 1) / kChunk;
            const unsigned long long clean_run = ~dirty & (chunks_run >= 64 ? ~0ull : ((1ull << chunks_run) - 1ull));
            n_culled = __popcll(clean_run) * kChunk;
            if (chunks_run > 0 && ((clean_run >> (chunks_run - 1)) & 1ull)) n_culled -= chunks_run * kChunk - n_exec;
        }

        // ---- phase C: the ribbon check-points of the executed samples (Edge.cpp:153-172) -----------------------------------
        if (!heavy) {
            int next_cp = 0;
            double prev_ang = 0;
            int prev_idx = -2;
            int rel[kDeepRelCap];
            int n_rel = 0, rel_chunk = -1;
#pragma unroll 1
            while (next_cp < n_exec) {
                if (!kDeep && ++n_cp > cp_budget) { long_run = true; heavy = true; break; }
                if (kDeep) n_cp++;
                const int idx = next_cp;
                double x, y, ang;
                bool it_;
                pose_eval(pe, tm.at(idx), &x, &y, &ang, &it_);
                bool do_cover = cov;
                if (!cov) { // lastHeading == intermediate.heading(), Edge.cpp:159
                    double ph;
                    if (idx == 0) ph = edge->src[2];
                    else if (prev_idx == idx - 1) ph = heading_of(prev_ang);
                    else {
                        double px_, py_, pa_;
                        pose_eval(pe, tm.at(idx - 1), &px_, &py_, &pa_, &it_);
                        ph = heading_of(pa_);
                    }
                    do_cover = (ph == heading_of(ang));
                }
                double toCover;
                if (!kDeep) {
                    bool would_change;
                    toCover = seq_checkpoint(rib, box, nr, x, y, W, tame, any_short, &would_change);
                    if (do_cover && would_change) { long_run = true; heavy = true; break; }
                } else {
                    const int c = idx / kChunk;
                    if (c != rel_chunk) { // ribbons any sample of this chunk can be contained in: around the chunk's middle pose
                        const int c0 = c * kChunk;
                        const int last = (c0 + kChunk - 1 < n_exec - 1) ? c0 + kChunk - 1 : n_exec - 1;
                        const double t_first = tm.at(c0), t_mid = tm.at(c0 + (last - c0 + 1) / 2), t_last = tm.at(last);
                        double mx_, my_, ma_;
                        const bool ok_mid = pose_eval(pe, t_mid, &mx_, &my_, &ma_, &it_) && it_;
                        const double reach = fmax(t_mid - t_first, t_last - t_mid) * w_speed * (1 + 1e-9) + 1e-6;
                        n_rel = ok_mid && (t_first <= t_mid) && (t_mid <= t_last) ? deep_relevant(rib, nr, mx_, my_, reach, W, rel) : -1;
                        rel_chunk = c;
                        tm.at(idx); // the cursor moves forward again
                    }
                    if (n_rel < 0 || !deep_checkpoint(rib, nr, ov, rel, n_rel, x, y, W, do_cover, &toCover, &modified)) { heavy = true; break; }
                }
                prev_ang = ang;
                prev_idx = idx;
                next_cp = idx + 1 + skip_count(toCover, inc, kSkipCap);
            }
        }

        // ---- phase D: `intermediate` / lastHeading at the loop exit, truncated end state (Edge.cpp:177-179), final cover -----
        if (!heavy) {
            if (n_exec > 0) { // the last sample that ran the full body
                double lx, ly, la;
                bool it_;
                pose_eval(pe, tm.at(n_exec - 1), &lx, &ly, &la, &it_);
                lastHeading = heading_of(la);
                if (!blocked_exit) { P_x = lx; P_y = ly; P_h = lastHeading; }
            }
            double ea;
            bool in_time;
            const bool sample_ok = pose_eval(pe, endTime, &ex, &ey, &ea, &in_time);
            eh = heading_of(ea);
            if (!in_time || !sample_ok) heavy = true; // the reference throws / stale pose: let the warp walker report it
            if (!heavy && (cov || lastHeading == P_h)) {
                if (!kDeep) {
                    bool would_change;
                    seq_checkpoint(rib, box, nr, P_x, P_y, W, tame, any_short, &would_change);
                    if (would_change) heavy = true;
                } else { // the final cover at `intermediate` (Edge.cpp:182-184)
                    int rel[kDeepRelCap];
                    const int n_rel = deep_relevant(rib, nr, P_x, P_y, 0.0, W, rel);
                    double unused;
                    if (n_rel < 0 || !deep_checkpoint(rib, nr, ov, rel, n_rel, P_x, P_y, W, true, &unused, &modified)) heavy = true;
                }
            }
        }
    }

    if (heavy) {
        if (!live || skipped) return;
        // The heavy list is filled from both ends: edges that were caught covering a ribbon (they tend to run along
        // it for hundreds of check-points, milliseconds of strictly sequential work) from the front, the rest from
        // the back.  K2b takes the front first, so the longest items start first and the batch does not end on one.
        // (the list has 2 n slots: an edge K2c passes on occupies a front and a back slot)
        if (long_run && !kDeep) {
            const unsigned int k = atomicAdd(heavy_count, 1u);
            heavy_list[k] = (unsigned int)ei;
        } else {
            const unsigned int k = atomicAdd(heavy_count + 1, 1u);
            heavy_list[2 * n - 1 - k] = (unsigned int)ei;
        }
        return;
    }

    // ---- cost, g, h and the result record (Edge.cpp:193-203); ribbons unchanged, coverage not complete -----------------
    const double netTime = endTime - src_t;
    const double T = fmax(netTime - 0.0, 0.0);
    const double true_cost = T * cfg.time_penalty_factor + penalty;
    const double g = edge->src_g + true_cost;
    double h;
    long long ribbons_offset = -1;
    int out_status = status;
    int nr_after = nr;
    const double4* list = rib;
    if (kDeep && modified) { // ribbons-after = the parent's list with the overrides applied, materialised in the pool
        int erased_n = 0;
        for (int k = 0; k < ov.n; k++) erased_n += (ov.erased >> k) & 1u;
        nr_after = nr + ov.n_ins - erased_n;
        const unsigned long long off = atomicAdd(w.out_count, (unsigned long long)nr_after);
        if (off + (unsigned long long)nr_after <= w.out_cap) {
            int o = 0;
#pragma unroll 1
            for (int q = 0; q < nr; q++) {
                for (int k = 0; k < ov.n_ins; k++)
                    if (ov.ins_pos[k] == q) w.out_ribbons[off + o++] = ov.ins[k];
                bool gone;
                const RibbonD rb = load_ribbon_ov(rib, q, ov, &gone);
                if (!gone) w.out_ribbons[off + o++] = pack_ribbon(rb.sx, rb.sy, rb.ex, rb.ey);
            }
            ribbons_offset = (long long)off;
            list = w.out_ribbons + off;
        } else {
            out_status = PPE_EDGE_ERR_RIBBON_CAPACITY; // the host grows the pool and runs the batch again
        }
    }
    if (cfg.heuristic == PPE_H_MAX_DISTANCE) {
        const double dist = (kDeep && modified) ? seq_max_distance(list, nr_after, ex, ey, W) : seq_max_distance_presummed(rib, nr, ex, ey, w.set_sumlen[set]);
        h = dist / cfg.max_speed * cfg.time_penalty_factor;
    } else {
        h = tsp_heuristic_or_unset(cfg, list, nr_after, ex, ey);
    }
    ppe_edge_result* r = results + ei;
    r->true_cost = true_cost;
    r->collision_penalty = penalty;
    r->approx_cost = pe[kApprox];
    r->end[0] = ex; r->end[1] = ey; r->end[2] = eh; r->end[3] = w_speed; r->end[4] = endTime;
    r->g = g;
    r->h = h;
    r->coverage_completed_time = cct;
    r->path_qi[0] = pe[kX0]; r->path_qi[1] = pe[kY0]; r->path_qi[2] = pe[kYaw0];
    r->path_param[0] = pe[kParam0]; r->path_param[1] = pe[kParam1]; r->path_param[2] = pe[kParam2];
    r->path_rho = pe[kRho];
    r->w_speed = w_speed;
    r->w_start_time = pe[kWStart];
    r->w_end_time = endTime;
    r->ribbons_offset = ribbons_offset;
    r->path_type = (int)pe[kType];
    r->infeasible = infeasible ? 1 : 0;
    r->status = out_status;
    r->n_samples = n_samples;
    r->n_checkpoints = n_cp;
    r->n_ribbons_after = nr_after;
    r->ribbons_changed = (kDeep && modified) ? 1 : 0;
    // instrumentation: culled samples; bit 24 = walked by a thread (K2t or K2c), bit 25 = by the deep walker K2c
    r->reserved = (n_culled & 0xffffff) | (1 << 24) | (kDeep ? (1 << 25) : 0);
}

#ifndef PPE_K2T_BLOCKS_PER_SM
#define PPE_K2T_BLOCKS_PER_SM 5 // 5 x 128 threads: at most 102 registers per thread
#endif
__global__ void __launch_bounds__(128, PPE_K2T_BLOCKS_PER_SM)
k2t_thread_walk(const __grid_constant__ WorldD w, const long long first, const long long n, const long long list_n,
                const ppe_edge* __restrict__ edges, const PreparedEdge* __restrict__ prepared,
                ppe_edge_result* __restrict__ results, unsigned int* __restrict__ heavy_list,
                unsigned int* __restrict__ heavy_count, const int dirty_budget, const int cp_budget) {
    // edges [first, n) of the batch; the heavy list belongs to the whole batch (2 * list_n slots)
    extern __shared__ double4 smem4[];
    ObstacleD* s_obs = reinterpret_cast<ObstacleD*>(smem4);
    {
        const int nd = w.n_obs * (int)(sizeof(ObstacleD) / sizeof(double));
        double* dst = reinterpret_cast<double*>(s_obs);
        const double* src = reinterpret_cast<const double*>(w.obstacles);
        for (int i = threadIdx.x; i < nd; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    // N2: shared-memory tile of the occupancy and safe bitmaps around the batch's bounding box (TMA bulk copies)
    const uint32_t* tile = nullptr;
    if (w.tile_on) {
        __shared__ unsigned long long s_tile_bar;
        uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem4 + (size_t)w.n_obs * (sizeof(ObstacleD) / sizeof(double4)));
        tile_stage(w, s_tile, &s_tile_bar);
        tile = s_tile;
    }
    const long long ei = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ei - (threadIdx.x & 31) >= n) return; // the whole warp lies beyond the range
    // phase B of the walk is warp-cooperative: the lanes beyond the last edge stay, reading (never writing) edge n - 1
    const bool live = ei < n;
    thread_walk<false>(w, list_n, live ? ei : n - 1, edges, prepared, results, heavy_list, heavy_count, dirty_budget, cp_budget, s_obs, tile, live);
}

// K2c: one thread per edge of the FRONT heavy list (the edges K2t caught covering a ribbon)
__global__ void __launch_bounds__(128)
k2c_deep_walk(const __grid_constant__ WorldD w, const long long n, const ppe_edge* __restrict__ edges,
              const PreparedEdge* __restrict__ prepared, ppe_edge_result* __restrict__ results,
              unsigned int* __restrict__ heavy_list, unsigned int* __restrict__ heavy_count, const int dirty_budget) {
    extern __shared__ double4 smem4[];
    ObstacleD* s_obs = reinterpret_cast<ObstacleD*>(smem4);
    {
        const int nd = w.n_obs * (int)(sizeof(ObstacleD) / sizeof(double));
        double* dst = reinterpret_cast<double*>(s_obs);
        const double* src = reinterpret_cast<const double*>(w.obstacles);
        for (int i = threadIdx.x; i < nd; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const unsigned int n_front = heavy_count[0];
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_front; k += gridDim.x * blockDim.x)
        thread_walk<true>(w, n, (long long)heavy_list[k], edges, prepared, results, heavy_list, heavy_count, dirty_budget, 0, s_obs, nullptr, true);
}

// K2a: one thread per edge
__global__ void __launch_bounds__(128)
k2a_prepare(const ppe_config cfg, const double dt, const double horizon_end, const long long n,
            const ppe_edge* __restrict__ edges, PreparedEdge* __restrict__ prepared) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    prepare_edge(cfg, dt, horizon_end, edges + i, prepared + i);
}

// K2b: one warp per edge, persistent CTAs pulling work from a global counter: the edges on the heavy list that
// K2t left behind, or (heavy_list == nullptr) every edge of the batch
#ifndef PPE_K2B_WARPS_PER_SM
#define PPE_K2B_WARPS_PER_SM 16 // 16 -> 128 registers per thread; 32 -> 64 (experiment: make variant K2B_WARPS=32)
#endif
template <int kWarpsPerBlock>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, PPE_K2B_WARPS_PER_SM / kWarpsPerBlock)
k2_true_cost(const __grid_constant__ WorldD w, const long long n, const ppe_edge* __restrict__ edges,
             const PreparedEdge* __restrict__ prepared, ppe_edge_result* __restrict__ results,
             unsigned long long* work_counter, const unsigned int* __restrict__ heavy_list,
             const unsigned int* __restrict__ heavy_count, const int front_too) {
    extern __shared__ double4 smem4[];
    ObstacleD* s_obs = reinterpret_cast<ObstacleD*>(smem4);
    double4* s_rib = smem4 + (size_t)w.n_obs * (sizeof(ObstacleD) / sizeof(double4));
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double4* bufA = s_rib + (size_t)warp * 2 * w.ribbon_cap;
    double4* bufB = bufA + w.ribbon_cap;
    __shared__ double s_pe[kWarpsPerBlock][kPrepDoubles];
    __shared__ TimeTable s_tt[kWarpsPerBlock];
    __shared__ int s_rel[kWarpsPerBlock][kRelCap];
    __shared__ double s_cp_pose[kWarpsPerBlock][96];
    double* pe = s_pe[warp];

    // heavy list: [0, count0) the edges K2t caught covering a ribbon (K2c's work; the warp walker's too when K2c is off),
    // [2n - count1, 2n) everything else incl. what K2c passed on
    const unsigned long long n_front = (heavy_list && front_too) ? (unsigned long long)heavy_count[0] : 0ull;
    const unsigned long long todo = heavy_list ? n_front + (unsigned long long)heavy_count[1] : (unsigned long long)n;
    if ((unsigned long long)blockIdx.x * kWarpsPerBlock >= todo) return; // nothing for this CTA: skip the staging
    {
        const int nd = w.n_obs * (int)(sizeof(ObstacleD) / sizeof(double));
        double* dst = reinterpret_cast<double*>(s_obs);
        const double* src = reinterpret_cast<const double*>(w.obstacles);
        for (int i = threadIdx.x; i < nd; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(work_counter, 1ULL);
        k = __shfl_sync(kFull, k, 0);
        if (k >= todo) break;
        const unsigned long long ei = !heavy_list ? k : (unsigned long long)(k < n_front ? heavy_list[k] : heavy_list[2 * n - 1 - (k - n_front)]);
        process_edge(w, &w, edges + ei, prepared + ei, results + ei, s_obs, bufA, bufB, pe, &s_tt[warp], s_rel[warp], s_cp_pose[warp], lane);
        __syncwarp();
    }
}

// K3a: best feasible f = g + h over the result records of one range (the prune record of pushVertexQueue,
// SamplingBasedPlanner.cpp:11-13); ties go to the smaller edge index.  One BestD per CTA.
__global__ void __launch_bounds__(256) k3_best_scan(const ppe_edge_result* __restrict__ results, const long long n, BestD* block_best) {
    __shared__ double s_f[256];
    __shared__ long long s_i[256];
    double bf = INFINITY;
    long long bi = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const ppe_edge_result* r = results + i;
        if (r->infeasible == 0 && r->status == PPE_EDGE_OK && r->h >= 0) {
            const double f = r->g + r->h;
            if (bi < 0 || f < bf) { bf = f; bi = i; } // i increases per thread: the first minimum is the smaller index
        }
    }
    s_f[threadIdx.x] = bf;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int s2 = blockDim.x / 2; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) {
            const double of = s_f[threadIdx.x + s2];
            const long long oi = s_i[threadIdx.x + s2];
            if (oi >= 0 && (s_i[threadIdx.x] < 0 || of < s_f[threadIdx.x] || (of == s_f[threadIdx.x] && oi < s_i[threadIdx.x]))) {
                s_f[threadIdx.x] = of;
                s_i[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { block_best[blockIdx.x].f = s_f[0]; block_best[blockIdx.x].idx = s_i[0]; }
}

// `index_base` turns the slice-local edge indices of one launch into batch indices; with `accumulate` the
// record of the earlier slices of the same batch (already in *out) takes part in the reduction.
__global__ void k3_best_final(const BestD* block_best, int nblocks, BestD* out, long long index_base, int accumulate) {
    __shared__ double s_f[256];
    __shared__ long long s_i[256];
    double bf = INFINITY;
    long long bi = -1;
    if (threadIdx.x == 0 && accumulate) { bf = out->f; bi = out->idx; }
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) {
        BestD b = block_best[k];
        if (b.idx >= 0) b.idx += index_base;
        if (b.idx >= 0 && (bi < 0 || b.f < bf || (b.f == bf && b.idx < bi))) { bf = b.f; bi = b.idx; }
    }
    s_f[threadIdx.x] = bf;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double of = s_f[threadIdx.x + s];
            const long long oi = s_i[threadIdx.x + s];
            if (oi >= 0 && (s_i[threadIdx.x] < 0 || of < s_f[threadIdx.x] || (of == s_f[threadIdx.x] && oi < s_i[threadIdx.x]))) {
                s_f[threadIdx.x] = of;
                s_i[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out->f = s_f[0]; out->idx = s_i[0]; }
}

// K1: one thread per Dubins solve
__global__ void __launch_bounds__(128)
k1_dubins_batch(const long long n, const double* __restrict__ q0, const double* __restrict__ q1,
                const double* __restrict__ rho, int32_t* __restrict__ type, double* __restrict__ param,
                double* __restrict__ length, int32_t* __restrict__ err) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a[3] = {q0[3 * i], q0[3 * i + 1], q0[3 * i + 2]};
    const double b[3] = {q1[3 * i], q1[3 * i + 1], q1[3 * i + 2]};
    DubinsPathD p;
    p.param[0] = p.param[1] = p.param[2] = 0;
    p.type = 0;
    const int e = dubins_shortest_path(&p, a, b, rho[i]);
    type[i] = p.type;
    param[3 * i] = p.param[0];
    param[3 * i + 1] = p.param[1];
    param[3 * i + 2] = p.param[2];
    length[i] = e == kEdubOk ? dubins_path_length(p) : 0.0;
    err[i] = e;
}

// ---- chunk-culling support: dilated free-space bitmap -------------------------------------------------------------------
// pass 1: rowfree[r][c] = cells (r, c - R .. c + R) all in bounds and free; pass 2: safe[r][c] = rowfree[r - R .. r + R][c]
// all set (rows out of bounds count as blocked).  One thread per cell, one ballot per 32 cells; the maps are a few MiB.
__global__ void k_safe_rows(const uint32_t* __restrict__ bits, uint32_t* __restrict__ out, int rows, int cols, int stride, int R) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    bool ok = false;
    if (c < cols && c - R >= 0 && c + R < cols) {
        ok = true;
        const uint32_t* row = bits + (size_t)r * stride;
        for (int k = c - R; k <= c + R; k++)
            if ((row[k >> 5] >> (k & 31)) & 1u) { ok = false; break; }
    }
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && (c >> 5) < stride) out[(size_t)r * stride + (c >> 5)] = b;
}

__global__ void k_safe_cols(const uint32_t* __restrict__ rowfree, uint32_t* __restrict__ out, int rows, int cols, int stride, int R) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    bool ok = false;
    if (c < cols && r - R >= 0 && r + R < rows) {
        ok = true;
        for (int k = r - R; k <= r + R; k++)
            if (!((rowfree[(size_t)k * stride + (c >> 5)] >> (c & 31)) & 1u)) { ok = false; break; }
    }
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && (c >> 5) < stride) out[(size_t)r * stride + (c >> 5)] = b;
}

// FP64 FMA throughput probe: the roofline denominator for K2 (MEASURED_PEAKS.json has no fp64 entry)
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;
}

} // namespace

int true_cost_block_threads() { return kWarpsWide * 32; }
int true_cost_chunk_samples() { return kChunk; }

cudaError_t launch_safe_map(const uint32_t* map_bits, uint32_t* scratch_rows, uint32_t* safe_bits, int rows, int cols,
                            int stride_words, int radius, cudaStream_t stream) {
    const dim3 block(128, 1, 1);
    const dim3 grid((unsigned)((stride_words * 32 + 127) / 128), (unsigned)rows, 1);
    k_safe_rows<<<grid, block, 0, stream>>>(map_bits, scratch_rows, rows, cols, stride_words, radius);
    k_safe_cols<<<grid, block, 0, stream>>>(scratch_rows, safe_bits, rows, cols, stride_words, radius);
    return cudaGetLastError();
}

static size_t k2_smem_bytes(int warps, int ribbon_cap, int n_obs) {
    return (size_t)n_obs * sizeof(ObstacleD) + (size_t)warps * 2 * (size_t)ribbon_cap * sizeof(double4);
}

// dynamic shared memory of the configuration that will be launched (the narrow one is the limit that matters)
size_t true_cost_smem_bytes(int ribbon_cap, int n_obs) { return k2_smem_bytes(kWarpsNarrow, ribbon_cap, n_obs); }

cudaError_t launch_dubins_batch(int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type,
                                double* param, double* length, int32_t* err, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int threads = 128;
    const long long blocks = (n + threads - 1) / threads;
    k1_dubins_batch<<<(unsigned)blocks, threads, 0, stream>>>(n, q0, q1, rho, type, param, length, err);
    return cudaGetLastError();
}

size_t prepared_edge_bytes() { return sizeof(PreparedEdge); }

template <int kW>
static cudaError_t launch_k2(const WorldD& world, int64_t n, const ppe_edge* edges, const PreparedEdge* prepared,
                             ppe_edge_result* results, unsigned long long* work_counter, const unsigned int* heavy_list,
                             const unsigned int* heavy_count, int front_too, int max_blocks, int sm_count, cudaStream_t stream,
                             int ctas_per_sm) {
    const size_t smem = k2_smem_bytes(kW, world.ribbon_cap, world.n_obs);
    cudaError_t e;
    if (smem > 32 * 1024) { // per device and per function (static + dynamic must fit): set it whenever it may be needed
        e = cudaFuncSetAttribute(k2_true_cost<kW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_true_cost<kW>, kW * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    if (ctas_per_sm > 0 && per_sm > ctas_per_sm) per_sm = ctas_per_sm;
    // persistent grid: a multiple of the SM count, never more warps than edges
    long long blocks = (long long)sm_count * per_sm;
    const long long needed = (n + kW - 1) / kW;
    if (blocks > needed) blocks = needed;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    k2_true_cost<kW><<<(unsigned)blocks, kW * 32, smem, stream>>>(world, (long long)n, edges, prepared, results, work_counter,
                                                                 heavy_list, heavy_count, front_too);
    return cudaGetLastError();
}

// K2a + (K2t) + K2b + K3a over one range of edges on `stream`; leaves one BestD per K3a CTA in block_best
// (*blocks_out of them).  `counters`: [0] 64-bit work counter of K2b, [1] two 32-bit heavy-list lengths (front, back).
// heavy_list == nullptr selects the warp walker for every edge.
cudaError_t launch_true_cost_kernels(const WorldD& world, int64_t n, const ppe_edge* edges, void* prepared_scratch,
                                     ppe_edge_result* results, unsigned long long* counters, unsigned int* heavy_list,
                                     BestD* block_best, int max_blocks, int sm_count, cudaStream_t stream, bool reset_pool,
                                     K2Tuning tuning, int* blocks_out, int* launches_out) {
    PreparedEdge* prepared = reinterpret_cast<PreparedEdge*>(prepared_scratch);
    cudaError_t e = cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    if (reset_pool) { // the ribbons-after pool runs over all slices of a batch
        e = cudaMemsetAsync(world.out_count, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
    }
    int launches = 0;
    // thread-per-edge kernels: a frontier batch (a few thousand edges) in 128-thread CTAs would occupy a fraction of the SMs
    const int bt = n < (int64_t)sm_count * 128 ? 32 : 128;
    k2a_prepare<<<(unsigned)((n + bt - 1) / bt), bt, 0, stream>>>(world.cfg, world.dt, world.horizon_end, (long long)n, edges, prepared);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    launches++;
    unsigned int* heavy_count = reinterpret_cast<unsigned int*>(counters + 1);
    if (heavy_list) {
        const size_t smem_t = (size_t)world.n_obs * sizeof(ObstacleD) +
                              (world.tile_on ? (size_t)2 * world.tile_rows * world.tile_words * sizeof(uint32_t) : 0);
        if (smem_t > 32 * 1024) {
            e = cudaFuncSetAttribute(k2t_thread_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
            if (e != cudaSuccess) return e;
        }
        const int dirty_budget = tuning.dirty_budget, cp_budget = tuning.cp_budget;
        k2t_thread_walk<<<(unsigned)((n + bt - 1) / bt), bt, smem_t, stream>>>(world, 0ll, (long long)n, (long long)n, edges, prepared,
                                                                             results, heavy_list, heavy_count, dirty_budget, cp_budget);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        launches++;
        if (tuning.deep_walker) { // K2c over the front list; what it cannot keep goes to the back list for K2b
            const size_t smem_c = (size_t)world.n_obs * sizeof(ObstacleD);
            if (smem_c > 32 * 1024) {
                e = cudaFuncSetAttribute(k2c_deep_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
                if (e != cudaSuccess) return e;
            }
            long long cblocks = (n + 127) / 128;
            if (cblocks > (long long)sm_count * 8) cblocks = (long long)sm_count * 8;
            k2c_deep_walk<<<(unsigned)cblocks, 128, smem_c, stream>>>(world, (long long)n, edges, prepared, results, heavy_list, heavy_count,
                                                                     kDeepDirtyBudget);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            launches++;
        }
    }
    const int front_too = (heavy_list && tuning.deep_walker) ? 0 : 1;
#ifndef PPE_K2_FORCE_NARROW
#define PPE_K2_FORCE_NARROW 1
#endif
    if (!PPE_K2_FORCE_NARROW && k2_smem_bytes(kWarpsWide, world.ribbon_cap, world.n_obs) <= 190 * 1024)
        e = launch_k2<kWarpsWide>(world, n, edges, prepared, results, counters, heavy_list, heavy_count, front_too, max_blocks, sm_count, stream, tuning.k2b_ctas_per_sm);
    else
        e = launch_k2<kWarpsNarrow>(world, n, edges, prepared, results, counters, heavy_list, heavy_count, front_too, max_blocks, sm_count, stream, tuning.k2b_ctas_per_sm);
    if (e != cudaSuccess) return e;
    launches++;
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    if (blocks > max_blocks) blocks = max_blocks;
    k3_best_scan<<<(unsigned)blocks, 256, 0, stream>>>(results, (long long)n, block_best);
    launches++;
    *blocks_out = (int)blocks;
    if (launches_out) *launches_out = launches;
    return cudaGetLastError();
}

// ---- host-buffer pipeline (ppe_true_cost_batch): K2a + K2t slice by slice while the copies run, K2b once ------------------------
// K2a + K2t over the edges [first, first + cnt) of a batch of n_total edges (all pointers are the batch's arrays); what K2t
// leaves behind is appended to the batch's heavy list.  `counters` must have been zeroed ahead of the first slice.
cudaError_t launch_prepare_and_walk(const WorldD& world, int64_t n_total, int64_t first, int64_t cnt, const ppe_edge* edges,
                                    void* prepared_scratch, ppe_edge_result* results, unsigned long long* counters,
                                    unsigned int* heavy_list, cudaStream_t stream, K2Tuning tuning, int* launches_out) {
    PreparedEdge* prepared = reinterpret_cast<PreparedEdge*>(prepared_scratch);
    k2a_prepare<<<(unsigned)((cnt + 127) / 128), 128, 0, stream>>>(world.cfg, world.dt, world.horizon_end, (long long)cnt, edges + first,
                                                                  prepared + first);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    unsigned int* heavy_count = reinterpret_cast<unsigned int*>(counters + 1);
    const size_t smem_t = (size_t)world.n_obs * sizeof(ObstacleD) +
                          (world.tile_on ? (size_t)2 * world.tile_rows * world.tile_words * sizeof(uint32_t) : 0);
    if (smem_t > 32 * 1024) {
        e = cudaFuncSetAttribute(k2t_thread_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
        if (e != cudaSuccess) return e;
    }
    k2t_thread_walk<<<(unsigned)((cnt + 127) / 128), 128, smem_t, stream>>>(world, (long long)first, (long long)(first + cnt),
                                                                           (long long)n_total, edges, prepared, results, heavy_list,
                                                                           heavy_count, tuning.dirty_budget, tuning.cp_budget);
    if (launches_out) *launches_out = 2;
    return cudaGetLastError();
}

// K2b over the batch's heavy list, then K3a over all results
cudaError_t launch_heavy_and_best(const WorldD& world, int64_t n_total, const ppe_edge* edges, void* prepared_scratch,
                                  ppe_edge_result* results, unsigned long long* counters, unsigned int* heavy_list,
                                  BestD* block_best, int max_blocks, int sm_count, cudaStream_t stream, K2Tuning tuning,
                                  int* blocks_out, int* launches_out) {
    PreparedEdge* prepared = reinterpret_cast<PreparedEdge*>(prepared_scratch);
    unsigned int* heavy_count = reinterpret_cast<unsigned int*>(counters + 1);
    cudaError_t e = launch_k2<kWarpsNarrow>(world, n_total, edges, prepared, results, counters, heavy_list, heavy_count, 1, max_blocks,
                                            sm_count, stream, tuning.k2b_ctas_per_sm);
    if (e != cudaSuccess) return e;
    long long blocks = (n_total + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    if (blocks > max_blocks) blocks = max_blocks;
    k3_best_scan<<<(unsigned)blocks, 256, 0, stream>>>(results, (long long)n_total, block_best);
    *blocks_out = (int)blocks;
    if (launches_out) *launches_out = 2;
    return cudaGetLastError();
}

// The records K2b wrote, one warp per record: to their places in `dst` (the caller's result array, mapped pinned host
// memory: a scatter over PCIe), or (compact_idx != nullptr) packed into dst[0 .. count) with their edge indices beside them.
static_assert(sizeof(ppe_edge_result) % 16 == 0 && sizeof(ppe_edge_result) / 16 <= 32, "k_patch_results copies 16 B per lane");
__global__ void __launch_bounds__(256)
k_patch_results(const ppe_edge_result* __restrict__ results, const unsigned int* __restrict__ heavy_list,
                const unsigned int* __restrict__ heavy_count, const long long n, ppe_edge_result* __restrict__ dst,
                unsigned int* __restrict__ compact_idx) {
    constexpr int kQuads = (int)(sizeof(ppe_edge_result) / 16);
    const unsigned int n_front = heavy_count[0], total = n_front + heavy_count[1];
    const int lane = threadIdx.x & 31;
    const unsigned int warps = gridDim.x * (blockDim.x >> 5);
    for (unsigned int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < total; k += warps) {
        const unsigned int ei = k < n_front ? heavy_list[k] : heavy_list[2 * n - 1 - (k - n_front)];
        const uint4* src = reinterpret_cast<const uint4*>(results + ei);
        uint4* out = reinterpret_cast<uint4*>(dst + (compact_idx ? k : ei));
        if (lane < kQuads) out[lane] = src[lane];
        if (compact_idx && lane == 0) compact_idx[k] = ei;
    }
}

cudaError_t launch_patch_results(const ppe_edge_result* results, const unsigned int* heavy_list, const unsigned long long* counters,
                                 int64_t n_total, ppe_edge_result* dst, unsigned int* compact_idx, int sm_count, cudaStream_t stream) {
    const unsigned int* heavy_count = reinterpret_cast<const unsigned int*>(counters + 1);
    k_patch_results<<<(unsigned)(sm_count * 4), 256, 0, stream>>>(results, heavy_list, heavy_count, (long long)n_total, dst, compact_idx);
    return cudaGetLastError();
}

// K3: reduce the per-CTA records of one range into *best; `accumulate` keeps what earlier ranges of the batch left there
cudaError_t launch_best_final(const BestD* block_best, int blocks, BestD* best, int64_t index_base, bool accumulate,
                              cudaStream_t stream) {
    k3_best_final<<<1, 256, 0, stream>>>(block_best, blocks, best, (long long)index_base, accumulate ? 1 : 0);
    return cudaGetLastError();
}

__global__ void k3_best_export(const BestD* src, BestD* dst, long long index_base) {
    BestD b = *src;
    if (b.idx >= 0) b.idx += index_base;
    *dst = b;
}

cudaError_t launch_best_export(const BestD* src, BestD* dst, int64_t index_base, cudaStream_t stream) {
    k3_best_export<<<1, 1, 0, stream>>>(src, dst, (long long)index_base);
    return cudaGetLastError();
}

K2Tuning clamp_tuning(K2Tuning t) {
    if (t.dirty_budget < 0) t.dirty_budget = 0;
    if (t.dirty_budget > kThreadDirtyCap) t.dirty_budget = kThreadDirtyCap;
    if (t.cp_budget < 1) t.cp_budget = 1;
    return t;
}

// development builds only (-DPPE_K2B_PROFILE): read and clear the warp walker's cycle counters; 0 entries otherwise
int k2b_profile_read(unsigned long long* out16) {
#ifdef PPE_K2B_PROFILE
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out16, g_k2b_prof, sizeof z) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(g_k2b_prof, z, sizeof z) != cudaSuccess) return -1;
    return 10;
#else
    (void)out16;
    return 0;
#endif
}

cudaError_t launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t stream) {
    k_fp64_peak<<<blocks, 256, 0, stream>>>(out, iters);
    return cudaGetLastError();
}

} // namespace ppe
