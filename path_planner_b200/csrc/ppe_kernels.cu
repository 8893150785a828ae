// ppe_kernels.cu -- hand-written sm_100a kernels of the batched Dubins edge-evaluation engine.
//
//   K1 k1_dubins_batch   one thread = one (q0, q1, rho) triple: shortest Dubins word, branch-free
//                        over the six words, fp64.       Replaces Edge::computeApproxCost ->
//                        DubinsWrapper::set -> dubins_shortest_path (Edge.cpp:11-20,
//                        DubinsWrapper.cpp:9-17).
//   K2 k2_true_cost      one warp = one edge; lanes = consecutive sample points of the path.
//                        Replaces Edge::computeTrueCost (Edge.cpp:68-206) including the Dubins
//                        solve for path-less edges, Map/GridWorldMap::isBlocked, Binary/Gaussian
//                        collisionExists, the RibbonManager cover state machine, the truncated end
//                        state, g (Vertex.cpp:102-104) and h = MaxDistance (RibbonManager.cpp:234-248).
//   K3 best-f epilogue   fused into K2 (per-warp running best, block reduce) + k3_best_final.
//
// No tensor cores: this is branchy fp64 transcendental + bit-gather work.  Compiled with
// -fmad=false: the x86-64 reference build has no FMA contraction and discrete outcomes (word
// choice, cell index, ribbon containment, sample count) must not flip.
#include <float.h>

#include "ppe_kernels.cuh"
#include "ppe_math.cuh"

namespace ppe {

namespace {

#ifndef PPE_K2_MIN_BLOCKS
#define PPE_K2_MIN_BLOCKS 2
#endif
constexpr int kWarpsPerBlock = 4;
constexpr int kBlockThreads = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xffffffffu;
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

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
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

__device__ __forceinline__ double4 pack_ribbon(double sx, double sy, double ex, double ey) {
    return make_double4(sx, sy, ex, ey);
}

// Map::isBlocked (Map.cpp:4-6) / GridWorldMap::isBlocked (GridWorldMap.cpp:84-93).  The bitmap
// (<= 2 MiB at 4096^2) stays L2/L1 resident; consecutive lanes are consecutive 0.05 m samples, so
// a warp's 32 lookups fall into one or two 32-byte sectors.
__device__ __forceinline__ bool map_blocked(const WorldD& w, double x, double y) {
    if (w.map_kind == kMapNone) return false;
    if (x < 0 || x / w.resolution >= (double)w.cols) return true;
    if (y < 0 || y / w.resolution >= (double)w.rows) return true;
    const unsigned long long r = (unsigned long long)(y / w.resolution);
    const unsigned long long c = (unsigned long long)(x / w.resolution);
    const uint32_t word = __ldg(&w.map_bits[r * (unsigned long long)w.stride_words + (c >> 5)]);
    return (word >> (c & 31)) & 1u;
}

// BinaryDynamicObstaclesManager::collisionExists (Binary...cpp:4-22) and
// GaussianDynamicObstaclesManager::collisionExists (Gaussian...cpp:3-13), strict = true.
// Obstacles are staged in shared memory once per CTA; every lane walks them in container order so
// the per-sample sum has the reference's summation order.
__device__ __forceinline__ double collision_exists(int kind, int n_obs, const ObstacleD* __restrict__ obs, double x,
                                                   double y, double time) {
    double sum = 0;
    if (kind == kObsBinary) {
        for (int i = 0; i < n_obs; i++) {
            const ObstacleD& o = obs[i];
            const double dtm = time - o.Time;
            const double dx = o.Speed * dtm * o.cosYaw;
            const double dy = o.Speed * dtm * o.sinYaw;
            const double X = o.X + dx, Y = o.Y + dy;
            const double tx = x - X, ty = y - Y;
            const double rx = tx * o.cosYaw - ty * o.sinYaw;
            const double ry = tx * o.sinYaw + ty * o.cosYaw;
            if (fabs(rx) < o.a && fabs(ry) < o.b) sum += 1.0;
        }
        return sum;
    }
    for (int i = 0; i < n_obs; i++) {
        const ObstacleD& o = obs[i];
        const double dtm = time - o.Time;
        const double dx = o.Speed * dtm * o.cosYaw;
        const double dy = o.Speed * dtm * o.sinYaw;
        const double X = o.X + dx, Y = o.Y + dy;
        const double d0 = x - X, d1 = y - Y;
        const double r0 = d0 * o.a + d1 * o.b;  // (v^T Sigma^-1)_0 = v0*i00 + v1*i10
        const double r1 = d0 * o.c + d1 * o.d;  // (v^T Sigma^-1)_1 = v0*i01 + v1*i11
        const double quadform = r0 * d0 + r1 * d1;
        sum += o.norm * exp(-0.5 * quadform);
    }
    if (sum < 1e-5) return 0;
    return sum;
}

// RibbonManager::minDistanceFrom (RibbonManager.cpp:142-152): lanes over ribbons.
__device__ __forceinline__ double warp_min_distance_from(const double4* cur, int nr, double x, double y, double W,
                                                         int lane) {
    if (nr == 0) return 0;
    double mn = DBL_MAX;
    bool inside = false;
    for (int r = lane; r < nr; r += 32) {
        const RibbonD rb = load_ribbon(cur + r);
        double px, py;
        ribbon_projection(rb, x, y, &px, &py);
        if (ribbon_contains(rb, x, y, px, py, false, W)) inside = true;
        const double dStart = point_distance(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart);
    }
    if (__any_sync(kFull, inside)) return 0;
    return warp_min(mn);
}

// RibbonManager::cover(x, y, strict = true) (RibbonManager.cpp:14-22 with Ribbon::split,
// Ribbon.cpp:9-17 and add, RibbonManager.cpp:154-158): every ribbon is split independently; list
// order is kept by a warp prefix sum over the 0/1/2 pieces each ribbon leaves behind.  Writes the
// new list into `alt`; the caller swaps the buffers when something changed.
__device__ __forceinline__ int warp_cover(const double4* cur, double4* alt, int nr, int cap, double x, double y,
                                          double W, int lane, bool* changed, bool* overflow) {
    int out_base = 0;
    bool any_change = false;
    for (int base = 0; base < nr; base += 32) {
        const int r = base + lane;
        const bool active = r < nr;
        RibbonD rb = {0, 0, 0, 0};
        bool contained = false;
        double px = 0, py = 0;
        if (active) {
            rb = load_ribbon(cur + r);
            ribbon_projection(rb, x, y, &px, &py);
            contained = ribbon_contains(rb, x, y, px, py, true, W);
        }
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
    __syncwarp();
    if (out_base > cap) *overflow = true;
    *changed = any_change;
    return any_change ? (out_base > cap ? cap : out_base) : nr;
}

// RibbonManager::maxDistance (RibbonManager.cpp:234-248); `scratch` holds >= nr doubles.
__device__ __forceinline__ double warp_max_distance(const double4* cur, int nr, double x, double y, double W,
                                                    int lane, double* scratch) {
    double mn = DBL_MAX, mx = 0;
    for (int r = lane; r < nr; r += 32) {
        const RibbonD rb = load_ribbon(cur + r);
        scratch[r] = sqrt(ribbon_sqlen(rb)) - 2 * W;
        const double dStart = point_distance(rb.sx, rb.sy, x, y);
        const double dEnd = point_distance(rb.ex, rb.ey, x, y);
        mn = fmin(fmin(mn, dEnd), dStart);
        mx = fmax(fmax(mx, dEnd), dStart);
    }
    __syncwarp();
    double sumLength = 0;
    for (int r = 0; r < nr; r++) sumLength += scratch[r]; // list order, as the reference sums
    __syncwarp();
    mn = warp_min(mn);
    mx = warp_max(mx);
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
    kX0 = 0, kY0, kRho, kInvRho, kLength, kP1, kP2, kP12,
    kBx = 8, kBy = 11, kBth = 14, kBs = 17, kBc = 20, kSgn = 23,   // 3 entries each, one per segment
    kWStart = 26, kWSpeed, kWEnd, kApprox, kEx, kEy, kEh, kEs,
    kYaw0 = 34, kParam0, kParam1, kParam2, kType, kStatus, kSampleFault,
    kPrepDoubles = 48
};
struct PreparedEdge {
    double v[kPrepDoubles];
};

// Edge.cpp:73-85, Edge::setEnd :208-215, DubinsWrapper.cpp:9-17,84-93 -- scalar, one thread.
__device__ void prepare_edge(const ppe_config& cfg, const ppe_edge* __restrict__ edge, PreparedEdge* __restrict__ out) {
    const double src_x = edge->src[0], src_y = edge->src[1], src_h = edge->src[2], src_speed = edge->src[3],
                 src_t = edge->src[4];
    const bool has_path = edge->has_path != 0;
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
    int status = PPE_EDGE_OK;
    bool sample_fault = false; // both dubins_path_sample attempts failed (stale-pose case)

    if (has_path) {
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
    v[kX0] = path.qi[0]; v[kY0] = path.qi[1]; v[kRho] = path.rho; v[kInvRho] = 1.0 / path.rho;
    v[kLength] = smp.length; v[kP1] = smp.p1; v[kP2] = smp.p2; v[kP12] = smp.p12;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        v[kBx + k] = smp.bx[k]; v[kBy + k] = smp.by[k]; v[kBth + k] = smp.bth[k];
        v[kBs + k] = smp.bs[k]; v[kBc + k] = smp.bc[k];
        v[kSgn + k] = smp.seg[k] == kSegL ? 1.0 : (smp.seg[k] == kSegR ? -1.0 : 0.0);
    }
    v[kWStart] = w_start; v[kWSpeed] = w_speed; v[kWEnd] = w_end;
    v[kApprox] = approx;
    v[kEx] = ex; v[kEy] = ey; v[kEh] = eh; v[kEs] = es;
    v[kYaw0] = path.qi[2]; v[kParam0] = path.param[0]; v[kParam1] = path.param[1]; v[kParam2] = path.param[2];
    v[kType] = (double)path.type;
    v[kStatus] = (double)status;
    v[kSampleFault] = sample_fault ? 1.0 : 0.0;
#pragma unroll
    for (int k = kSampleFault + 1; k < kPrepDoubles; k++) v[k] = 0.0;
}

// sin and cos for |x| up to a few thousand (path angles stay within a few turns): three-constant
// Cody-Waite reduction by pi/2 + Taylor polynomials on [-pi/4, pi/4].  ~1 ulp; used for the
// per-sample poses only (1e-9 tolerance class) -- and small enough to keep the sample loop inside
// the instruction cache, unlike libdevice's sincos with its Payne-Hanek slow path.
__device__ __forceinline__ void sincos_bounded(double x, double* sn, double* cs) {
    const double kd = rint(x * 0x1.45f306dc9c883p-1);
    const int q = (int)kd;
    double r = __fma_rn(-kd, 0x1.921fb54442d18p+0, x);
    r = __fma_rn(-kd, 0x1.1a62633145c07p-54, r);
    r = __fma_rn(-kd, -0x1.f1976b7ed8fbcp-110, r);
    const double z = r * r;
    double ps = -0x1.2f49b46814157p-57;
    ps = __fma_rn(ps, z, 0x1.952c77030ad4ap-49);
    ps = __fma_rn(ps, z, -0x1.ae7f3e733b81fp-41);
    ps = __fma_rn(ps, z, 0x1.6124613a86d09p-33);
    ps = __fma_rn(ps, z, -0x1.ae64567f544e4p-26);
    ps = __fma_rn(ps, z, 0x1.71de3a556c734p-19);
    ps = __fma_rn(ps, z, -0x1.a01a01a01a01ap-13);
    ps = __fma_rn(ps, z, 0x1.1111111111111p-7);
    ps = __fma_rn(ps, z, -0x1.5555555555555p-3);
    const double s0 = __fma_rn(r * z, ps, r);
    double pc = 0x1.e542ba4020225p-62;
    pc = __fma_rn(pc, z, -0x1.6827863b97d97p-53);
    pc = __fma_rn(pc, z, 0x1.ae7f3e733b81fp-45);
    pc = __fma_rn(pc, z, -0x1.93974a8c07c9dp-37);
    pc = __fma_rn(pc, z, 0x1.1eed8eff8d898p-29);
    pc = __fma_rn(pc, z, -0x1.27e4fb7789f5cp-22);
    pc = __fma_rn(pc, z, 0x1.a01a01a01a01ap-16);
    pc = __fma_rn(pc, z, -0x1.6c16c16c16c17p-10);
    pc = __fma_rn(pc, z, 0x1.5555555555555p-5);
    pc = __fma_rn(pc, z, -0x1.0000000000000p-1);
    const double c0 = __fma_rn(z, pc, 1.0);
    const double a = (q & 1) ? c0 : s0;
    const double b2 = (q & 1) ? s0 : c0;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b2 : b2;
}

// One edge, one warp.  Per-edge scalars that every lane would hold identically live in the warp's
// shared-memory copy of the prepared record (`pe`); lane i of a chunk owns sample index base + i.
// The pose evaluation, the ribbon cover and the end-state sample each exist exactly once in the
// instruction stream: the truncated end state (Edge.cpp:177-178) and the final cover (:182-191)
// run as one extra "tail" pass through the same loop body.
__device__ void process_edge(const WorldD& w, const ppe_edge* __restrict__ edge, const PreparedEdge* __restrict__ prep,
                             ppe_edge_result* __restrict__ result, const ObstacleD* s_obs, double4* bufA, double4* bufB,
                             double* pe, int lane, double* out_f) {
    const ppe_config& cfg = w.cfg;
    const double W = cfg.ribbon_width;
    const double inc = cfg.collision_checking_increment;
    const int cap = w.ribbon_cap;
    *out_f = INFINITY;

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
    const double t0 = src_t + fmod(src_t - cfg.start_state_time, dt);
    int next_cp = 0; // sample index of the next ribbon check-point (toCoverDistance starts at 0)
    double carry_x = edge->src[0], carry_y = edge->src[1], carry_h = edge->src[2]; // `intermediate` before the chunk
    double P_x = carry_x, P_y = carry_y, P_h = carry_h, lastHeading = carry_h, t_exit = t0;
    double ex = 0, ey = 0, eh = 0;
    int n_samples = 0, n_cp = 0;
    {
        const double span = (endTime - t0) / dt;
        if (!((dt > 0) && (span < (double)kMaxSamples || !(t0 < endTime)))) status = PPE_EDGE_ERR_END_SAMPLE;
    }

    if (status == PPE_EDGE_OK) {
        TimeWalker tw;
        tw.init(t0, dt);
        bool tail = false;
        for (int base = 0;;) {
            // ---- pose at this lane's time (tail pass: every lane at the truncated end time) ------------------
            const double t_i = tail ? endTime : tw.at(base + lane);
            bool valid = tail || (t_i < endTime);
            const double w_start = pe[kWStart];
            const bool in_time = (w_start <= t_i) && (pe[kWEnd] >= t_i); // DubinsWrapper::containsTime
            double dist = (t_i - w_start) * pe[kWSpeed];
            bool sample_ok = true;
            {
                const double length = pe[kLength];
                if (dist < 0 || dist > length) { // EDUBPARAM -> one retry at distance - 1e-5 (DubinsWrapper.cpp:39-42)
                    dist = dist - 1e-5;
                    sample_ok = !(dist < 0 || dist > length);
                }
            }
            const double tprime = dist * pe[kInvRho];
            const double p1 = pe[kP1];
            const int k = tprime < p1 ? 0 : (tprime < pe[kP12] ? 1 : 2);
            const double tl = k == 0 ? tprime : (k == 1 ? tprime - p1 : tprime - p1 - pe[kP2]);
            const double sg = pe[kSgn + k], bth = pe[kBth + k], bs = pe[kBs + k], bc = pe[kBc + k];
            const double ang = sg * tl + bth; // L: t + th, R: -t + th, S: 0 + th  (dubins_segment)
            double sn, cs;
            sincos_bounded(ang, &sn, &cs);
            const double qx = (sg == 0.0 ? bc * tl : sg * (sn - bs)) + pe[kBx + k];
            const double qy = (sg == 0.0 ? bs * tl : -sg * (cs - bc)) + pe[kBy + k];
            const double rho = pe[kRho];
            const double x = qx * rho + pe[kX0];
            const double y = qy * rho + pe[kY0];
            const double yaw = ang - kTwoPi * floor(ang * 0x1.45f306dc9c883p-3);
            double hd = kPi2 - yaw; // State::setYaw
            if (hd < 0) hd += kTwoPi;

            int limit = 0, fstop = 32;
            unsigned m_stop = 0;
            bool blocked = false;
            if (!tail) {
                if (valid && in_time && sample_ok) blocked = map_blocked(w, x, y);
                const unsigned m_valid = __ballot_sync(kFull, valid);
                m_stop = __ballot_sync(kFull, valid && (!in_time || blocked));
                if (__any_sync(kFull, valid && in_time && !sample_ok)) sample_fault = true;
                const int nvalid = __popc(m_valid); // valid lanes form a prefix (times increase)
                fstop = m_stop ? (__ffs(m_stop) - 1) : 32;
                limit = nvalid < fstop ? nvalid : fstop; // lanes [0, limit) execute the full loop body
            } else {
                if (!in_time) status = PPE_EDGE_ERR_END_SAMPLE; // sample(end state) throws, DubinsWrapper.cpp:30-35
                if (!sample_ok) sample_fault = true;
                ex = x; ey = y; eh = hd;
            }

            // ---- ribbon check-points of this chunk, in order (Edge.cpp:153-172); tail: the final cover (:182-191)
            for (;;) {
                if (!tail && !(next_cp < base + limit)) break;
                const int l = tail ? 0 : next_cp - base;
                double cx = __shfl_sync(kFull, x, l);
                double cy = __shfl_sync(kFull, y, l);
                double ch = __shfl_sync(kFull, hd, l);
                double ct = __shfl_sync(kFull, t_i, l);
                const double ph_prev = __shfl_sync(kFull, hd, l > 0 ? l - 1 : 0);
                double ph = l > 0 ? ph_prev : carry_h;
                double toCover = 0;
                if (tail) {
                    cx = P_x; cy = P_y; ch = P_h; ct = t_exit; ph = lastHeading;
                } else {
                    n_cp++;
                    toCover = warp_min_distance_from(cur, nr, cx, cy, W, lane);
                }
                if (cov || ph == ch) {
                    bool changed = false;
                    const int nn = warp_cover(cur, alt, nr, cap, cx, cy, W, lane, &changed, &overflow);
                    if (changed) {
                        double4* tmp = cur; cur = alt; alt = tmp;
                        nr = nn;
                        modified = true;
                    }
                }
                if (nr == 0) {
                    if (cct == -1) cct = ct;
                    ribbonsDoneTime = (int)ct;
                    if (!tail) {
                        endTime = fmin(endTime, cct + cfg.time_minimum);
                        valid = t_i < endTime;
                        const int nvalid = __popc(__ballot_sync(kFull, valid));
                        limit = nvalid < fstop ? nvalid : fstop;
                        if (limit < l + 1) limit = l + 1; // the check-point's own iteration has already run
                    }
                }
                if (tail) break;
                next_cp = next_cp + 1 + skip_count(toCover, inc, kSkipCap);
            }
            if (tail) break;

            // ---- dynamic-obstacle penalty of the executed iterations (Edge.cpp:150-151), in sample order
            if (w.obs_kind != kObsNone && w.n_obs > 0) {
                double p = 0;
                if (lane < limit) p = collision_exists(w.obs_kind, w.n_obs, s_obs, x, y, t_i) * cfg.collision_penalty_factor;
                if (__any_sync(kFull, p != 0)) {
                    for (int q = 0; q < limit; q++) penalty += __shfl_sync(kFull, p, q);
                }
            }

            n_samples += limit;
            if (limit < 32) {
                // the loop ends inside this chunk.  The iteration at lane `limit` was entered and broke
                // out only if that lane is still inside the (possibly truncated) end time.
                const bool stop_lane_valid = __shfl_sync(kFull, (int)valid, limit) != 0;
                const bool stopped = (fstop == limit) && ((m_stop >> fstop) & 1u) && stop_lane_valid;
                const int lp = limit > 0 ? limit - 1 : 0;
                const double px = __shfl_sync(kFull, x, lp), py = __shfl_sync(kFull, y, lp), ph = __shfl_sync(kFull, hd, lp);
                const double prev_x = limit > 0 ? px : carry_x, prev_y = limit > 0 ? py : carry_y,
                             prev_h = limit > 0 ? ph : carry_h;
                const double sx_ = __shfl_sync(kFull, x, limit), sy_ = __shfl_sync(kFull, y, limit),
                             sh_ = __shfl_sync(kFull, hd, limit);
                const bool stop_blocked = __shfl_sync(kFull, (int)blocked, limit) != 0;
                t_exit = __shfl_sync(kFull, t_i, limit);
                lastHeading = prev_h;
                P_x = prev_x; P_y = prev_y; P_h = prev_h;
                if (stopped) {
                    infeasible = true;
                    n_samples += 1; // the breaking iteration was entered
                    if (stop_blocked) { P_x = sx_; P_y = sy_; P_h = sh_; } // `intermediate` holds the blocked sample
                }
                tail = true;
                continue;
            }
            carry_x = __shfl_sync(kFull, x, 31);
            carry_y = __shfl_sync(kFull, y, 31);
            carry_h = __shfl_sync(kFull, hd, 31);
            base += 32;
            if (base > kMaxSamples) { status = PPE_EDGE_ERR_END_SAMPLE; break; }
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
            h = -1;
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
        r->reserved = 0;
    }
    if (!infeasible && status == PPE_EDGE_OK && h >= 0) *out_f = g + h;
}

// K2a: one thread per edge
__global__ void __launch_bounds__(128)
k2a_prepare(const ppe_config cfg, const long long n, const ppe_edge* __restrict__ edges, PreparedEdge* __restrict__ prepared) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    prepare_edge(cfg, edges + i, prepared + i);
}

// K2b: one warp per edge, persistent CTAs pulling edges from a global counter
__global__ void __launch_bounds__(kBlockThreads, PPE_K2_MIN_BLOCKS)
k2_true_cost(const WorldD w, const long long n, const ppe_edge* __restrict__ edges, const PreparedEdge* __restrict__ prepared,
             ppe_edge_result* __restrict__ results, unsigned long long* work_counter, BestD* block_best) {
    extern __shared__ double4 smem4[];
    ObstacleD* s_obs = reinterpret_cast<ObstacleD*>(smem4);
    double4* s_rib = smem4 + (size_t)w.n_obs * (sizeof(ObstacleD) / sizeof(double4));
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double4* bufA = s_rib + (size_t)warp * 2 * w.ribbon_cap;
    double4* bufB = bufA + w.ribbon_cap;
    __shared__ double s_pe[kWarpsPerBlock][kPrepDoubles];
    double* pe = s_pe[warp];

    {
        const int nd = w.n_obs * (int)(sizeof(ObstacleD) / sizeof(double));
        double* dst = reinterpret_cast<double*>(s_obs);
        const double* src = reinterpret_cast<const double*>(w.obstacles);
        for (int i = threadIdx.x; i < nd; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    double best_f = INFINITY;
    long long best_idx = -1;
    for (;;) {
        unsigned long long ei = 0;
        if (lane == 0) ei = atomicAdd(work_counter, 1ULL);
        ei = __shfl_sync(kFull, ei, 0);
        if (ei >= (unsigned long long)n) break;
        double f;
        process_edge(w, edges + ei, prepared + ei, results + ei, s_obs, bufA, bufB, pe, lane, &f);
        if (f < best_f || (f == best_f && (long long)ei < best_idx)) { best_f = f; best_idx = (long long)ei; }
        __syncwarp();
    }

    // K3 epilogue: block-level best (f, edge index); ties go to the smaller index
    __shared__ double s_f[kWarpsPerBlock];
    __shared__ long long s_i[kWarpsPerBlock];
    if (lane == 0) { s_f[warp] = best_f; s_i[warp] = best_idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bf = INFINITY;
        long long bi = -1;
        for (int k = 0; k < kWarpsPerBlock; k++) {
            if (s_i[k] >= 0 && (s_f[k] < bf || (s_f[k] == bf && s_i[k] < bi) || bi < 0)) { bf = s_f[k]; bi = s_i[k]; }
        }
        block_best[blockIdx.x].f = bf;
        block_best[blockIdx.x].idx = bi;
    }
}

__global__ void k3_best_final(const BestD* block_best, int nblocks, BestD* out) {
    __shared__ double s_f[256];
    __shared__ long long s_i[256];
    double bf = INFINITY;
    long long bi = -1;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) {
        const BestD b = block_best[k];
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

int true_cost_block_threads() { return kBlockThreads; }

size_t true_cost_smem_bytes(int ribbon_cap, int n_obs) {
    return (size_t)n_obs * sizeof(ObstacleD) + (size_t)kWarpsPerBlock * 2 * (size_t)ribbon_cap * sizeof(double4);
}

cudaError_t launch_dubins_batch(int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type,
                                double* param, double* length, int32_t* err, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int threads = 128;
    const long long blocks = (n + threads - 1) / threads;
    k1_dubins_batch<<<(unsigned)blocks, threads, 0, stream>>>(n, q0, q1, rho, type, param, length, err);
    return cudaGetLastError();
}

size_t prepared_edge_bytes() { return sizeof(PreparedEdge); }

cudaError_t launch_true_cost_batch(const WorldD& world, int64_t n, const ppe_edge* edges, void* prepared_scratch,
                                   ppe_edge_result* results, unsigned long long* work_counter, BestD* block_best,
                                   int max_blocks, BestD* best, int sm_count, cudaStream_t stream, int* launches) {
    PreparedEdge* prepared = reinterpret_cast<PreparedEdge*>(prepared_scratch);
    const size_t smem = true_cost_smem_bytes(world.ribbon_cap, world.n_obs);
    static size_t configured = 0;
    cudaError_t e;
    if (smem > configured) {
        e = cudaFuncSetAttribute(k2_true_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2_true_cost, kBlockThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    // persistent grid: a multiple of the SM count, never more warps than edges
    long long blocks = (long long)sm_count * per_sm;
    const long long needed = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > needed) blocks = needed;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(world.out_count, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    k2a_prepare<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(world.cfg, (long long)n, edges, prepared);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k2_true_cost<<<(unsigned)blocks, kBlockThreads, smem, stream>>>(world, (long long)n, edges, prepared, results,
                                                                   work_counter, block_best);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k3_best_final<<<1, 256, 0, stream>>>(block_best, (int)blocks, best);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t stream) {
    k_fp64_peak<<<blocks, 256, 0, stream>>>(out, iters);
    return cudaGetLastError();
}

} // namespace ppe
