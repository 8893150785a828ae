// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE (oracle/_ref), not product code.
//
// C-ABI wrapper around the UNMODIFIED reference classes, compiled from the sources where they
// lie under /root/reference (recipe: oracle/Makefile, outputs only in oracle/_ref/).  It drives
//   Vertex::makeRoot / Vertex::connect      (path_planner/src/planner/search/Vertex.cpp:21-42,125-130)
//   Edge::computeTrueCost                   (path_planner/src/planner/search/Edge.cpp:68-206)
//   Edge::computeApproxCost / DubinsWrapper (Edge.cpp:11-20, DubinsWrapper.cpp:9-17)
//   AStarPlanner::plan                      (path_planner/src/planner/AStarPlanner.cpp:12-132)
// with the same batch structs as include/ppe.h so that tests can diff engine vs reference
// field by field.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
// load this library.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <chrono>
#include <unistd.h>
#include <cstdlib>
#include <thread>
#include <atomic>

#include "planner/search/Vertex.h"
#include "planner/search/Edge.h"
#include "planner/AStarPlanner.h"
#include "planner/utilities/RibbonManager.h"
#include "common/map/GridWorldMap.h"
#include "common/dynamic_obstacles/BinaryDynamicObstaclesManager.h"
#include "common/dynamic_obstacles/GaussianDynamicObstaclesManager.h"

#include "ppe.h"

// ---- access to two private members without touching the reference sources --------------------
// (explicit template instantiation may name private members; [temp.explicit]/12)
namespace {
template <typename Tag, typename Tag::type M>
struct Rob {
    friend typename Tag::type get(Tag) { return M; }
};
struct VertexCurrentCost { typedef double Vertex::*type; friend type get(VertexCurrentCost); };
struct EdgeApproxCost { typedef double Edge::*type; friend type get(EdgeApproxCost); };
struct ManagerRibbons { typedef std::list<Ribbon> RibbonManager::*type; friend type get(ManagerRibbons); };
}
template struct Rob<VertexCurrentCost, &Vertex::m_CurrentCost>;
template struct Rob<EdgeApproxCost, &Edge::m_ApproxCost>;
template struct Rob<ManagerRibbons, &RibbonManager::m_Ribbons>;

#include "ref_shim.h"

static RibbonManager::Heuristic heuristicOf(int h) {
    switch (h) {
        case PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL: return RibbonManager::TspPointRobotNoSplitAllRibbons;
        case PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K: return RibbonManager::TspPointRobotNoSplitKRibbons;
        case PPE_H_TSP_DUBINS_NO_SPLIT_ALL: return RibbonManager::TspDubinsNoSplitAllRibbons;
        case PPE_H_TSP_DUBINS_NO_SPLIT_K: return RibbonManager::TspDubinsNoSplitKRibbons;
        default: return RibbonManager::MaxDistance;
    }
}

// the reference's planner with read access to its sample counter (for the virtual clock of ref_run_plan)
struct RefAStarPlanner : AStarPlanner {
    unsigned long attemptedSamples() const { return m_AttemptedSamples; }
};

extern "C" {

int ref_create(ref_ctx** out) {
    *out = new ref_ctx();
    // the reference chats on std::cerr (ribbon-count warnings, "Collision possible ..."); keep
    // test and bench logs readable unless asked otherwise.  The ctx is never freed before exit in
    // practice; ref_destroy restores the buffer.
    if (!getenv("PPE_REF_VERBOSE")) {
        static NullBuf quiet;
        std::cerr.rdbuf(&quiet);
    }
    (*out)->config.setMap(std::make_shared<Map>());
    (*out)->config.setNowFunction([]() -> double {
        struct timespec t; clock_gettime(CLOCK_REALTIME, &t);
        return t.tv_sec + t.tv_nsec * 1e-9;
    });
    return 0;
}

void ref_destroy(ref_ctx* ctx) { delete ctx; }

const char* ref_last_error(const ref_ctx* ctx) { return ctx->lastError.c_str(); }

int ref_set_config(ref_ctx* ctx, const ppe_config* c) {
    ctx->cfg = *c;
    ctx->config.setMaxSpeed(c->max_speed);
    ctx->config.setSlowSpeed(c->slow_speed);
    ctx->config.setTurningRadius(c->turning_radius);
    ctx->config.setCoverageTurningRadius(c->coverage_turning_radius);
    ctx->config.setTimeHorizon(c->time_horizon);
    ctx->config.setTimeMinimum(c->time_minimum);
    ctx->config.setCollisionCheckingIncrement(c->collision_checking_increment);
    ctx->config.setStartStateTime(c->start_state_time);
    ctx->config.setBranchingFactor(c->branching_factor);
    RibbonManager::setRibbonWidth(c->ribbon_width); // process-global static (Ribbon.cpp:4)
    if (c->collision_penalty_factor != Edge::collisionPenaltyFactor() ||
        c->time_penalty_factor != Edge::timePenaltyFactor()) {
        ctx->lastError = "penalty factors are compile-time constants in the reference (Edge.h:151-152)";
        return -3;
    }
    return 0;
}

int ref_set_map_none(ref_ctx* ctx) {
    ctx->config.setMap(std::make_shared<Map>());
    return 0;
}

// Writes the bitmap in the GridWorld text format (GridWorldMap.cpp:10-64: line 1 resolution,
// then rows of characters, '#' blocked, LAST line = y 0) and loads it with the reference parser.
int ref_set_map_bitmap(ref_ctx* ctx, const uint8_t* bits, int rows, int cols, int stride, double resolution) {
    char path[] = "/tmp/ppe_ref_map_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) { ctx->lastError = "mkstemp failed"; return -3; }
    close(fd);
    {
        std::ofstream f(path);
        char buf[64];
        snprintf(buf, sizeof buf, "%.17g", resolution);
        f << buf << "\n";
        std::string line((size_t)cols, '_');
        for (int r = rows - 1; r >= 0; r--) {
            for (int c = 0; c < cols; c++)
                line[c] = ((bits[(size_t)r * stride + (c >> 3)] >> (c & 7)) & 1) ? '#' : '_';
            f << line << "\n";
        }
    }
    ctx->config.setMap(std::make_shared<GridWorldMap>(path));
    unlink(path);
    return 0;
}

// The reference's own parser on a file as it lies on disk (tests of the engine's `.map` ingest).
int ref_set_map_file(ref_ctx* ctx, const char* path) {
    try {
        ctx->config.setMap(std::make_shared<GridWorldMap>(path));
    } catch (std::exception& ex) {
        ctx->lastError = ex.what();
        return -3;
    }
    return 0;
}

// Map::isBlocked of the installed map (virtual dispatch, as Edge.cpp:144 calls it)
int ref_is_blocked(ref_ctx* ctx, double x, double y) { return ctx->config.map()->isBlocked(x, y) ? 1 : 0; }

double ref_map_resolution(ref_ctx* ctx) { return ctx->config.map()->resolution(); }

int ref_set_obstacles_none(ref_ctx* ctx) {
    ctx->binary.reset(); ctx->gaussian.reset();
    ctx->config.setObstaclesManager(std::make_shared<DynamicObstaclesManager>());
    return 0;
}

// NB: takes HEADINGS (the managers' update() API); use ref_get_obstacles to read back the
// container iteration order and the stored yaws for the engine.
int ref_set_obstacles_binary(ref_ctx* ctx, int n, const double* x, const double* y, const double* heading,
                             const double* speed, const double* time, const double* width, const double* length) {
    ctx->gaussian.reset();
    ctx->binary = std::make_shared<BinaryDynamicObstaclesManager>();
    for (int i = 0; i < n; i++)
        ctx->binary->update((uint32_t)(i + 1), x[i], y[i], heading[i], speed[i], time[i], width[i], length[i]);
    ctx->config.setObstaclesManager(ctx->binary);
    return 0;
}

int ref_set_obstacles_gaussian(ref_ctx* ctx, int n, const double* x, const double* y, const double* heading,
                               const double* speed, const double* time, const double* cov) {
    ctx->binary.reset();
    ctx->gaussian = std::make_shared<GaussianDynamicObstaclesManager>();
    for (int i = 0; i < n; i++) {
        if (cov) {
            Eigen::Matrix<double, 2, 2> c;
            c << cov[4 * i], cov[4 * i + 1], cov[4 * i + 2], cov[4 * i + 3];
            ctx->gaussian->update((uint32_t)(i + 1), x[i], y[i], heading[i], speed[i], time[i], c);
        } else {
            ctx->gaussian->update((uint32_t)(i + 1), x[i], y[i], heading[i], speed[i], time[i]);
        }
    }
    ctx->config.setObstaclesManager(ctx->gaussian);
    return 0;
}

// Obstacles in the container's iteration order (what `for (auto o : m_Obstacles)` visits,
// Binary...cpp:6 / Gaussian...cpp:5).  out: n x 9 doubles: X, Y, Yaw, Speed, Time, then
// (Width, Length, 0, 0) for binary or the 2x2 covariance for gaussian.  Returns n.
int ref_get_obstacles(ref_ctx* ctx, double* out, int cap) {
    int n = 0;
    if (ctx->binary) {
        for (const auto& o : ctx->binary->get()) {
            if (n < cap) {
                double* p = out + 9 * n;
                p[0] = o.second.X; p[1] = o.second.Y; p[2] = o.second.Yaw; p[3] = o.second.Speed; p[4] = o.second.Time;
                p[5] = o.second.Width; p[6] = o.second.Length; p[7] = 0; p[8] = 0;
            }
            n++;
        }
    } else if (ctx->gaussian) {
        for (const auto& o : ctx->gaussian->get()) {
            if (n < cap) {
                double* p = out + 9 * n;
                p[0] = o.second.X; p[1] = o.second.Y; p[2] = o.second.Yaw; p[3] = o.second.Speed; p[4] = o.second.Time;
                p[5] = o.second.covariance(0, 0); p[6] = o.second.covariance(0, 1);
                p[7] = o.second.covariance(1, 0); p[8] = o.second.covariance(1, 1);
            }
            n++;
        }
    }
    return n;
}

int ref_put_ribbon_set(ref_ctx* ctx, int n, const double* xyxy, double coverageCompletedTime, int32_t* id) {
    RibbonManager m(heuristicOf(ctx->cfg.heuristic), ctx->cfg.turning_radius, ctx->cfg.tsp_k > 0 ? ctx->cfg.tsp_k : 2);
    // verbatim list, as a child vertex copies its parent's (Vertex.cpp:24,32); RibbonManager::add would
    // drop ribbons shorter than 2 * RibbonWidth (RibbonManager.cpp:154-158) that a strict cover keeps
    std::list<Ribbon>& list = m.*get(ManagerRibbons());
    for (int i = 0; i < n; i++) list.emplace_back(xyxy[4 * i], xyxy[4 * i + 1], xyxy[4 * i + 2], xyxy[4 * i + 3]);
    if (coverageCompletedTime != -1) m.setCoverageCompletedTime(coverageCompletedTime);
    ctx->sets.push_back(m);
    *id = (int32_t)ctx->sets.size() - 1;
    return 0;
}

int ref_clear_ribbon_sets(ref_ctx* ctx) { ctx->sets.clear(); return 0; }

// Ribbons of a stored set, in list order, as the reference holds them.  Returns the count.
int ref_get_ribbon_set(ref_ctx* ctx, int id, double* xyxy, int cap) {
    int n = 0;
    for (const auto& r : ctx->sets[id].get()) {
        if (n < cap) {
            xyxy[4 * n] = r.start().first; xyxy[4 * n + 1] = r.start().second;
            xyxy[4 * n + 2] = r.end().first; xyxy[4 * n + 3] = r.end().second;
        }
        n++;
    }
    return n;
}

// dubins_shortest_path through the reference's own call site (DubinsWrapper::set).
int ref_dubins_batch(ref_ctx*, int64_t n, const double* q0, const double* q1, const double* rho,
                     int32_t* type, double* param, double* length, int32_t* err) {
    for (int64_t i = 0; i < n; i++) {
        DubinsPath p;
        memset(&p, 0, sizeof p);
        double a[3] = {q0[3 * i], q0[3 * i + 1], q0[3 * i + 2]};
        double b[3] = {q1[3 * i], q1[3 * i + 1], q1[3 * i + 2]};
        err[i] = dubins_shortest_path(&p, a, b, rho[i]);
        type[i] = (int32_t)p.type;
        param[3 * i] = p.param[0]; param[3 * i + 1] = p.param[1]; param[3 * i + 2] = p.param[2];
        length[i] = err[i] == EDUBOK ? dubins_path_length(&p) : 0;
    }
    return 0;
}

static void evalOne(ref_ctx* ctx, const ppe_edge& e, ppe_edge_result& r, std::vector<double>& ribbonsOut,
                    PlannerConfig& config) {
    memset(&r, 0, sizeof r);
    r.n_samples = -1; r.n_checkpoints = -1; r.ribbons_offset = -1;
    State src(e.src[0], e.src[1], e.src[2], e.src[3], e.src[4]);
    auto root = Vertex::makeRoot(src, ctx->sets[e.ribbon_set]);
    (*root).*get(VertexCurrentCost()) = e.src_g;
    root->computeApproxToGo(config); // computeTrueCost reads start()->approxToGo() (Edge.cpp:99)
    Vertex::SharedPtr v;
    double rho = e.coverage_allowed ? config.coverageTurningRadius() : config.turningRadius();
    try {
        if (e.has_path) {
            DubinsPath p;
            p.qi[0] = e.path_qi[0]; p.qi[1] = e.path_qi[1]; p.qi[2] = e.path_qi[2];
            p.param[0] = e.path_param[0]; p.param[1] = e.path_param[1]; p.param[2] = e.path_param[2];
            p.rho = e.path_rho; p.type = (DubinsPathType)e.path_type;
            DubinsWrapper w;
            w.fill(p, e.w_speed, e.w_start_time);
            if (e.w_end_time < w.getEndTime()) w.updateEndTime(e.w_end_time);
            // the edge speed is the wrapper's speed: setEnd(wrapper) samples it into the end state (Edge.cpp:208-215)
            v = Vertex::connect(root, w, e.coverage_allowed != 0);
        } else {
            State dst(e.dst[0], e.dst[1], e.dst[2], e.dst[3], 0);
            v = Vertex::connect(root, dst, rho, e.coverage_allowed != 0);
        }
        v->parentEdge()->computeTrueCost(config);
    } catch (std::exception& ex) {
        r.status = PPE_EDGE_ERR_END_SAMPLE;
        if (v) r.infeasible = v->parentEdge()->infeasible();
        return;
    }
    const auto& edge = v->parentEdge();
    r.true_cost = edge->trueCost();
    r.collision_penalty = edge->getSavedCollisionPenalty();
    r.approx_cost = (*edge).*get(EdgeApproxCost());
    r.infeasible = edge->infeasible();
    r.end[0] = v->state().x(); r.end[1] = v->state().y(); r.end[2] = v->state().heading();
    r.end[3] = v->state().speed(); r.end[4] = v->state().time();
    r.g = v->currentCost();
    r.h = v->approxToGo();
    r.coverage_completed_time = v->ribbonManager().coverageCompletedTime();
    const DubinsWrapper& w = static_cast<const Edge&>(*edge).getPlan(config);
    const DubinsPath& p = w.unwrap();
    r.path_qi[0] = p.qi[0]; r.path_qi[1] = p.qi[1]; r.path_qi[2] = p.qi[2];
    r.path_param[0] = p.param[0]; r.path_param[1] = p.param[1]; r.path_param[2] = p.param[2];
    r.path_rho = p.rho; r.path_type = (int32_t)p.type;
    r.w_speed = w.getSpeed(); r.w_start_time = w.getStartTime(); r.w_end_time = w.getEndTime();
    ribbonsOut.clear();
    for (const auto& rb : v->ribbonManager().get()) {
        ribbonsOut.push_back(rb.start().first); ribbonsOut.push_back(rb.start().second);
        ribbonsOut.push_back(rb.end().first); ribbonsOut.push_back(rb.end().second);
    }
    r.n_ribbons_after = (int32_t)(ribbonsOut.size() / 4);
    // changed?
    const auto& parent = ctx->sets[e.ribbon_set].get();
    bool changed = parent.size() != (size_t)r.n_ribbons_after;
    if (!changed) {
        size_t k = 0;
        for (const auto& rb : parent) {
            if (rb.start().first != ribbonsOut[4 * k] || rb.start().second != ribbonsOut[4 * k + 1] ||
                rb.end().first != ribbonsOut[4 * k + 2] || rb.end().second != ribbonsOut[4 * k + 3]) { changed = true; break; }
            k++;
        }
    }
    r.ribbons_changed = changed;
}

// threads <= 0: all host threads (OpenMP); 1: the reference's own single-threaded behaviour.
// keep_ribbons: store ribbons-after for ref_get_ribbons_after (off for timing runs).
int ref_true_cost_batch_mt(ref_ctx* ctx, int64_t n, const ppe_edge* edges, ppe_edge_result* results,
                           int threads, int keep_ribbons) {
    if (keep_ribbons) ctx->ribbonsAfter.assign((size_t)n, std::vector<double>());
    else ctx->ribbonsAfter.clear();
    int nt = threads <= 0 ? (int)std::thread::hardware_concurrency() : threads;
    if (nt < 1) nt = 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        PlannerConfig config = ctx->config; // by-value copy, as Planner::plan takes it (Planner.h:50)
        std::vector<double> scratch;
        const int64_t chunk = 16;
        for (;;) {
            int64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            int64_t e = b + chunk < n ? b + chunk : n;
            for (int64_t i = b; i < e; i++)
                evalOne(ctx, edges[i], results[i], keep_ribbons ? ctx->ribbonsAfter[i] : scratch, config);
        }
    };
    if (nt == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    return 0;
}

int ref_true_cost_batch(ref_ctx* ctx, int64_t n, const ppe_edge* edges, ppe_edge_result* results) {
    return ref_true_cost_batch_mt(ctx, n, edges, results, 1, 1);
}

int ref_get_ribbons_after(ref_ctx* ctx, int64_t i, double* xyxy, int cap) {
    if (i < 0 || (size_t)i >= ctx->ribbonsAfter.size()) return -3;
    const auto& v = ctx->ribbonsAfter[i];
    int n = (int)(v.size() / 4);
    for (int k = 0; k < n && k < cap; k++) memcpy(xyxy + 4 * k, v.data() + 4 * k, 4 * sizeof(double));
    return n;
}

int ref_max_threads(void) {
    int n = (int)std::thread::hardware_concurrency();
    return n < 1 ? 1 : n;
}

// ---- planner level ---------------------------------------------------------------------------
// Plan record: per Dubins path 12 doubles: qi[3], param[3], rho, type, speed, start, end, 0.
// stats: Samples, Generated, Expanded, Iterations, PlanFValue, PlanCollisionPenalty,
//        PlanTimePenalty, PlanHValue, PlanDepth, nowCalls.
// Virtual clock (parity mode): tick > 0 => now() = clock0 + calls * tick (PlannerConfig.h:110);
// tick == 0 => wall clock.
int ref_plan(ref_ctx* ctx, int ribbon_set, const double* start5, double timeRemaining, double clock0,
             double tick, int initialSamples, int useBrownPaths, double* plan_out, int plan_cap,
             double* stats10) {
    RefAStarPlanner planner;
    return ref_run_plan(planner, ctx, ribbon_set, start5, timeRemaining, clock0, tick, initialSamples, useBrownPaths,
                        plan_out, plan_cap, stats10);
}

int ref_plan2(ref_ctx* ctx, int ribbon_set, const double* start5, double timeRemaining, double clock0, double tick,
              int initialSamples, int useBrownPaths, const double* prev_plan, int n_prev, double* plan_out, int plan_cap,
              double* stats10) {
    RefAStarPlanner planner;
    return ref_run_plan(planner, ctx, ribbon_set, start5, timeRemaining, clock0, tick, initialSamples, useBrownPaths,
                        plan_out, plan_cap, stats10, prev_plan, n_prev);
}

// + virtual time per generated sample (see ref_run_plan)
int ref_plan3(ref_ctx* ctx, int ribbon_set, const double* start5, double timeRemaining, double clock0, double tick,
              double sampleTick, int initialSamples, int useBrownPaths, const double* prev_plan, int n_prev, double* plan_out,
              int plan_cap, double* stats10) {
    RefAStarPlanner planner;
    return ref_run_plan(planner, ctx, ribbon_set, start5, timeRemaining, clock0, tick, initialSamples, useBrownPaths,
                        plan_out, plan_cap, stats10, prev_plan, n_prev, sampleTick);
}

int ref_expand_once(ref_ctx* ctx, int ribbon_set, int nSamples, int seed, double* f_out, int cap) {
    AStarPlanner planner;
    return ref_run_expand_once(planner, ctx, ribbon_set, nSamples, seed, f_out, cap);
}

} // extern "C"
