// harness_capi.cpp -- include/ppe_harness.h: the standalone, ROS-free harness around BatchedAStarPlanner.
#include "ppe_harness.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "BatchedAStarPlanner.h"
#include "common/dynamic_obstacles/BinaryDynamicObstaclesManager.h"
#include "common/dynamic_obstacles/GaussianDynamicObstaclesManager.h"
#include "common/map/GridWorldMap.h"
#include "common/map/Map.h"
#include "planner/utilities/RibbonManager.h"
#include "planner/utilities/Visualizer.h"

namespace {

// Occupancy bits behind the reference's Map interface, GridWorldMap::isBlocked semantics (GridWorldMap.cpp:84-93):
// out of bounds = blocked, cell (row, col) = (size_t)(y / res), (size_t)(x / res).
class BitmapMap : public Map {
public:
    BitmapMap(const uint8_t* bits, int rows, int cols, int stride, double res)
        : m_Bits(bits, bits + (size_t)rows * stride), m_Rows(rows), m_Cols(cols), m_Stride(stride), m_Res(res) {
        m_Extremes[0] = 0; m_Extremes[1] = cols * res; m_Extremes[2] = 0; m_Extremes[3] = rows * res;
    }
    bool isBlocked(double x, double y) const override {
        if (x < 0 || x / m_Res >= m_Cols) return true;
        if (y < 0 || y / m_Res >= m_Rows) return true;
        const size_t r = (size_t)(y / m_Res), c = (size_t)(x / m_Res);
        return (m_Bits[r * m_Stride + (c >> 3)] >> (c & 7)) & 1;
    }
    double resolution() const override { return m_Res; }
    const double* extremes() const override { return m_Extremes; }

private:
    std::vector<uint8_t> m_Bits;
    int m_Rows, m_Cols, m_Stride;
    double m_Res;
    double m_Extremes[4];
};

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };

RibbonManager::Heuristic heuristicOf(int h) {
    switch (h) {
        case PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL: return RibbonManager::TspPointRobotNoSplitAllRibbons;
        case PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K: return RibbonManager::TspPointRobotNoSplitKRibbons;
        case PPE_H_TSP_DUBINS_NO_SPLIT_ALL: return RibbonManager::TspDubinsNoSplitAllRibbons;
        case PPE_H_TSP_DUBINS_NO_SPLIT_K: return RibbonManager::TspDubinsNoSplitKRibbons;
        default: return RibbonManager::MaxDistance;
    }
}

DubinsWrapper wrapperOf(const pph_dubins_path& r) {
    DubinsPath p;
    p.qi[0] = r.initial_x; p.qi[1] = r.initial_y; p.qi[2] = r.initial_yaw;
    p.param[0] = r.length0; p.param[1] = r.length1; p.param[2] = r.length2;
    p.rho = r.rho; p.type = (DubinsPathType)r.type;
    DubinsWrapper w;
    w.fill(p, r.speed, r.start_time);
    if (r.end_time < w.getEndTime()) w.updateEndTime(r.end_time);
    return w;
}

} // namespace

struct pph_ctx {
    ppe_ctx* engine = nullptr;
    PpeWorldCache cache;
    NullBuf nullBuf;
    std::ostream nullStream;
    PlannerConfig config;
    ppe_config cfg{};
    RibbonManager ribbons;
    std::vector<double> ribbonXY;
    DubinsPlan lastPlan;
    State lastStart;
    std::string err;
    double clockNow = 0, clockTick = 0;
    uint64_t clockCalls = 0;
    pph_ctx() : nullStream(&nullBuf), config(&nullStream) {}
};

extern "C" {

int pph_create(int device, pph_ctx** out) {
    if (!out) return PPE_ERR_INVALID;
    *out = nullptr;
    ppe_ctx* engine = nullptr;
    const int rc = ppe_create(device, &engine);
    if (rc != PPE_OK) return rc; // no GPU: there is no CPU path
    pph_ctx* ctx = new pph_ctx();
    ctx->engine = engine;
    ctx->config.setMap(std::make_shared<Map>());
    *out = ctx;
    return PPE_OK;
}

void pph_destroy(pph_ctx* ctx) {
    if (!ctx) return;
    ppe_destroy(ctx->engine);
    delete ctx;
}

const char* pph_last_error(const pph_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pph_set_config(pph_ctx* ctx, const ppe_config* c) {
    if (!ctx || !c) return PPE_ERR_INVALID;
    if (c->collision_penalty_factor != Edge::collisionPenaltyFactor() || c->time_penalty_factor != Edge::timePenaltyFactor()) {
        ctx->err = "penalty factors are compile-time constants of the planner (Edge.h:151-152)";
        return PPE_ERR_INVALID;
    }
    ctx->cfg = *c;
    ctx->config.setMaxSpeed(c->max_speed);
    ctx->config.setSlowSpeed(c->slow_speed);
    ctx->config.setTurningRadius(c->turning_radius);
    ctx->config.setCoverageTurningRadius(c->coverage_turning_radius);
    ctx->config.setTimeHorizon(c->time_horizon);
    ctx->config.setTimeMinimum(c->time_minimum);
    ctx->config.setCollisionCheckingIncrement(c->collision_checking_increment);
    ctx->config.setBranchingFactor(c->branching_factor);
    RibbonManager::setRibbonWidth(c->ribbon_width); // process-global static of the reference (Ribbon.cpp:4)
    return PPE_OK;
}

int pph_set_map_none(pph_ctx* ctx) {
    if (!ctx) return PPE_ERR_INVALID;
    ctx->config.setMap(std::make_shared<Map>());
    return PPE_OK;
}

int pph_set_map_bitmap(pph_ctx* ctx, const uint8_t* bits, int rows, int cols, int stride, double resolution) {
    if (!ctx || !bits || rows <= 0 || cols <= 0 || stride * 8 < cols || !(resolution > 0)) return PPE_ERR_INVALID;
    ctx->config.setMap(std::make_shared<BitmapMap>(bits, rows, cols, stride, resolution));
    return PPE_OK;
}

int pph_load_gridworld_map(pph_ctx* ctx, const char* path) {
    if (!ctx || !path) return PPE_ERR_INVALID;
    try {
        ctx->config.setMap(std::make_shared<GridWorldMap>(path));
    } catch (std::exception& ex) { ctx->err = ex.what(); return PPE_ERR_INVALID; }
    return PPE_OK;
}

int pph_set_obstacles_none(pph_ctx* ctx) {
    if (!ctx) return PPE_ERR_INVALID;
    ctx->config.setObstaclesManager(std::make_shared<DynamicObstaclesManager>());
    return PPE_OK;
}

int pph_set_obstacles_binary(pph_ctx* ctx, int n, const double* x, const double* y, const double* heading, const double* speed,
                             const double* time, const double* width, const double* length) {
    if (!ctx || n < 0) return PPE_ERR_INVALID;
    auto m = std::make_shared<BinaryDynamicObstaclesManager>();
    for (int i = 0; i < n; i++) m->update((uint32_t)(i + 1), x[i], y[i], heading[i], speed[i], time[i], width[i], length[i]);
    ctx->config.setObstaclesManager(m);
    return PPE_OK;
}

int pph_set_obstacles_gaussian(pph_ctx* ctx, int n, const double* x, const double* y, const double* heading, const double* speed,
                               const double* time, const double* cov) {
    if (!ctx || n < 0) return PPE_ERR_INVALID;
    auto m = std::make_shared<GaussianDynamicObstaclesManager>();
    for (int i = 0; i < n; i++) {
        if (cov) {
            Eigen::Matrix<double, 2, 2> c;
            c << cov[4 * i], cov[4 * i + 1], cov[4 * i + 2], cov[4 * i + 3];
            m->update((uint32_t)(i + 1), x[i], y[i], heading[i], speed[i], time[i], c);
        } else {
            m->update((uint32_t)(i + 1), x[i], y[i], heading[i], speed[i], time[i]);
        }
    }
    ctx->config.setObstaclesManager(m);
    return PPE_OK;
}

int pph_set_ribbons(pph_ctx* ctx, int n, const double* xyxy) {
    if (!ctx || n < 0 || (n > 0 && !xyxy)) return PPE_ERR_INVALID;
    ctx->ribbons = RibbonManager(heuristicOf(ctx->cfg.heuristic), ctx->cfg.turning_radius, ctx->cfg.tsp_k > 0 ? ctx->cfg.tsp_k : 2); // executive.cpp:391
    for (int i = 0; i < n; i++) ctx->ribbons.add(xyxy[4 * i], xyxy[4 * i + 1], xyxy[4 * i + 2], xyxy[4 * i + 3]);
    return PPE_OK;
}

int pph_plan(pph_ctx* ctx, const double start5[5], const pph_dubins_path* previous, int n_previous, const pph_plan_options* opt,
             pph_dubins_path* plan_out, int cap, pph_stats* stats) {
    if (!ctx || !start5 || !opt || !stats || (cap > 0 && !plan_out)) return PPE_ERR_INVALID;
    PlannerConfig config = ctx->config;
    config.setInitialSamples(opt->initial_samples > 0 ? opt->initial_samples : 100);
    config.setUseBrownPaths(opt->use_brown_paths != 0);
    ctx->clockNow = opt->clock0; ctx->clockTick = opt->tick; ctx->clockCalls = 0;
    BatchedAStarPlanner planner(ctx->engine, opt->knn_chunk > 0 ? opt->knn_chunk : 128, &ctx->cache);
    if (opt->frontier >= 0) planner.setFrontierWidth(opt->frontier);
    if (opt->tick > 0) {
        // deterministic clock (PlannerConfig::setNowFunction, PlannerConfig.h:110): deadline tests and the sampler's
        // seed (AStarPlanner.cpp:33) then depend on call counts only
        const BatchedAStarPlanner* pl = &planner;
        const double sampleTick = opt->sample_tick;
        config.setNowFunction([ctx, pl, sampleTick]() -> double {
            return ctx->clockNow + (double)(ctx->clockCalls++) * ctx->clockTick + sampleTick * (double)pl->attemptedSamples();
        });
    } else if (opt->clock0 > 0) {
        const auto t_start = std::chrono::steady_clock::now();
        const double c0 = opt->clock0;
        config.setNowFunction([ctx, t_start, c0]() -> double {
            ctx->clockCalls++;
            return c0 + std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        });
    } else {
        config.setNowFunction([ctx]() -> double {
            ctx->clockCalls++;
            return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
        });
    }
    // visualization stream as Executive sets it up (executive.cpp:446-447): a Visualizer owning the file
    Visualizer::UniquePtr visualizer;
    if (opt->visualize && opt->visualization_path) {
        visualizer = Visualizer::UniquePtr(new Visualizer(opt->visualization_path));
        if (!visualizer->stream()) { ctx->err = "cannot open the visualization file"; return PPE_ERR_INVALID; }
        config.setVisualizer(&visualizer);
        config.setVisualizations(true);
    }
    DubinsPlan prev;
    for (int i = 0; i < n_previous; i++) prev.append(wrapperOf(previous[i]));
    const State start(start5[0], start5[1], start5[2], start5[3], start5[4]);
    Planner::Stats st;
    const auto t0 = std::chrono::steady_clock::now();
    try {
        st = planner.plan(ctx->ribbons, start, config, prev, opt->time_remaining);
    } catch (std::exception& ex) {
        ctx->err = ex.what();
        return PPE_ERR_STATE;
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int n = 0;
    for (const auto& w : st.Plan.get()) {
        if (n < cap) {
            pph_dubins_path& o = plan_out[n];
            const DubinsPath& p = w.unwrap();
            o.initial_x = p.qi[0]; o.initial_y = p.qi[1]; o.initial_yaw = p.qi[2];
            o.length0 = p.param[0]; o.length1 = p.param[1]; o.length2 = p.param[2];
            o.rho = w.getRho(); o.type = (int32_t)p.type; o.pad = 0;
            o.speed = w.getSpeed(); o.start_time = w.getStartTime(); o.end_time = w.getEndTime();
        }
        n++;
    }
    std::memset(stats, 0, sizeof *stats);
    stats->samples = st.Samples; stats->generated = st.Generated; stats->expanded = st.Expanded; stats->iterations = st.Iterations;
    stats->plan_depth = n ? st.PlanDepth : 0;
    stats->plan_f = n ? st.PlanFValue : -1; stats->plan_collision_penalty = st.PlanCollisionPenalty;
    stats->plan_time_penalty = n ? st.PlanTimePenalty : -1; stats->plan_h = n ? st.PlanHValue : -1;
    stats->plan_endtime = n ? st.Plan.getEndTime() : -1;
    stats->now_calls = ctx->clockCalls;
    stats->true_cost_edges = (uint64_t)planner.trueCostEdges(); stats->dubins_solves = (uint64_t)planner.dubinsSolves();
    stats->engine_batches = (uint64_t)planner.batches(); stats->frontier_vertices = (uint64_t)planner.frontierVertices();
    stats->frontier_hits = (uint64_t)planner.frontierHits(); stats->exact_expansions = (uint64_t)planner.exactExpansions();
    stats->wall_seconds = wall;
    stats->seconds_engine_expand = planner.secondsInEngineExpand(); stats->seconds_replay = planner.secondsInReplay();
    stats->seconds_add_samples = planner.secondsInAddSamples(); stats->seconds_exact = planner.secondsInExact();
    stats->exact_for_ties = (uint64_t)planner.exactForTies(); stats->exact_for_overflow = (uint64_t)planner.exactForOverflow();
    ctx->lastPlan = st.Plan;
    ctx->lastStart = start;
    return n;
}

int pph_advance(pph_ctx* ctx, double time, double state_out[5]) {
    if (!ctx || !state_out || ctx->lastPlan.empty()) return PPE_ERR_STATE;
    State s;
    s.time() = time;
    try {
        ctx->lastPlan.sample(s);
    } catch (std::exception& ex) { ctx->err = ex.what(); return PPE_ERR_INVALID; }
    ctx->ribbons.coverBetween(ctx->lastStart.x(), ctx->lastStart.y(), s.x(), s.y(), false); // executive.cpp:188
    state_out[0] = s.x(); state_out[1] = s.y(); state_out[2] = s.heading(); state_out[3] = s.speed(); state_out[4] = s.time();
    return PPE_OK;
}

int pph_write_plan_msg(const pph_dubins_path* plan, int n, const char* path) {
    if ((n > 0 && !plan) || !path) return PPE_ERR_INVALID;
    FILE* f = fopen(path, "w");
    if (!f) return PPE_ERR_INVALID;
    // path_planner_common/msg/Plan.msg: DubinsPath[] paths, float64 endtime -- field order of DubinsPath.msg
    fprintf(f, "paths:\n");
    for (int i = 0; i < n; i++) {
        const pph_dubins_path& p = plan[i];
        fprintf(f, "  - initial_x: %.17g\n    initial_y: %.17g\n    initial_yaw: %.17g\n    length0: %.17g\n    length1: %.17g\n"
                   "    length2: %.17g\n    rho: %.17g\n    type: %d\n    speed: %.17g\n    start_time: %.17g\n",
                p.initial_x, p.initial_y, p.initial_yaw, p.length0, p.length1, p.length2, p.rho, p.type, p.speed, p.start_time);
    }
    fprintf(f, "endtime: %.17g\n", n ? plan[n - 1].end_time : 0.0);
    fclose(f);
    return PPE_OK;
}

} // extern "C"
