// BatchedAStarPlanner.cpp -- see the header.  Host side of the restructured planner:
//   * expandFrontier / speculate: SamplingBasedPlanner::expand (SamplingBasedPlanner.cpp:52-151) for the popped vertex
//     AND the best vertices of the open list in one ppe_expand_batch call; children cached per vertex and handed to
//     pushVertexQueue in the reference's order when the reference's aStar loop pops that vertex.
//   * expandExact: the same expansion with the k-nearest heaps replayed on the host (K1 per chunk of candidates, one K2
//     launch per vertex) -- the path taken when two samples lie at exactly equal distance, where the pop order depends
//     on the arrangement std::make_heap / std::pop_heap leave in m_Samples.
//   * plan: AStarPlanner::plan (AStarPlanner.cpp:12-132) with the same control flow and now() sequence; previous-plan
//     re-validation (:46-59), Brown-path expansion (:150-162) and addSamples (SamplingBasedPlanner.cpp:157-164) call the
//     engine instead of Edge::computeTrueCost / Map::isBlocked.
#include "BatchedAStarPlanner.h"
#include "KeyedHeap.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeinfo>

#include "common/dynamic_obstacles/BinaryDynamicObstaclesManager.h"
#include "common/dynamic_obstacles/GaussianDynamicObstaclesManager.h"
#include "common/map/Map.h"
#include "planner/search/Edge.h"
#include "planner/search/Vertex.h"
#include "planner/utilities/Ribbon.h"
#include "planner/utilities/RibbonManager.h"

// ---- write access to the members Edge::computeTrueCost fills in (Edge.h:132-149, Vertex.h:180-188,
// RibbonManager.h:184-200) without modifying the reference: explicit template instantiation may
// name private members ([temp.explicit]/12).
namespace {
template <typename Tag, typename Tag::type M>
struct Access {
    friend typename Tag::type get(Tag) { return M; }
};
#define PPE_ACCESS(Tag, Class, Type, Member)                 \
    struct Tag { typedef Type Class::*type; friend type get(Tag); }; \
    template struct Access<Tag, &Class::Member>;
PPE_ACCESS(EdgeWrapper, Edge, DubinsWrapper, m_DubinsWrapper)
PPE_ACCESS(EdgeInfeasible, Edge, bool, m_Infeasible)
PPE_ACCESS(EdgeApprox, Edge, double, m_ApproxCost)
PPE_ACCESS(EdgeTrue, Edge, double, m_TrueCost)
PPE_ACCESS(EdgePenalty, Edge, double, m_CollisionPenalty)
PPE_ACCESS(VertexG, Vertex, double, m_CurrentCost)
PPE_ACCESS(VertexH, Vertex, double, m_ApproxToGo)
PPE_ACCESS(RibbonList, RibbonManager, std::list<Ribbon>, m_Ribbons)
PPE_ACCESS(RibbonHeuristic, RibbonManager, RibbonManager::Heuristic, m_Heuristic)
PPE_ACCESS(RibbonCct, RibbonManager, double, m_CoverageCompletedTime)
PPE_ACCESS(RibbonK, RibbonManager, int, m_K)
PPE_ACCESS(OpenList, SamplingBasedPlanner, std::vector<std::shared_ptr<Vertex>>, m_VertexQueue)
#undef PPE_ACCESS

struct Stopwatch { // adds the scope's wall time to a counter
    double& acc;
    std::chrono::steady_clock::time_point t0;
    explicit Stopwatch(double& a) : acc(a), t0(std::chrono::steady_clock::now()) {}
    ~Stopwatch() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

int heuristicId(RibbonManager::Heuristic h) {
    switch (h) {
        case RibbonManager::MaxDistance: return PPE_H_MAX_DISTANCE;
        case RibbonManager::TspPointRobotNoSplitAllRibbons: return PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL;
        case RibbonManager::TspPointRobotNoSplitKRibbons: return PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K;
        case RibbonManager::TspDubinsNoSplitAllRibbons: return PPE_H_TSP_DUBINS_NO_SPLIT_ALL;
        default: return PPE_H_TSP_DUBINS_NO_SPLIT_K;
    }
}
} // namespace

BatchedAStarPlanner::BatchedAStarPlanner(ppe_ctx* ctx, int knnChunk, PpeWorldCache* cache)
    : m_Ctx(ctx), m_KnnChunk(knnChunk), m_Cache(cache) {}

void BatchedAStarPlanner::check(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + ppe_last_error(m_Ctx));
}

// Read-only world state, once per plan (PlannerConfig is passed by value per plan, Planner.h:50).
void BatchedAStarPlanner::uploadWorld(const RibbonManager& ribbonManager, const State& start, const PlannerConfig& config) {
    ppe_config c;
    c.max_speed = config.maxSpeed();
    c.slow_speed = config.slowSpeed();
    c.turning_radius = config.turningRadius();
    c.coverage_turning_radius = config.coverageTurningRadius();
    c.time_horizon = config.timeHorizon();
    c.time_minimum = config.timeMinimum();
    c.collision_checking_increment = config.collisionCheckingIncrement();
    c.start_state_time = start.time();                      // AStarPlanner.cpp:16
    c.ribbon_width = Ribbon::RibbonWidth;
    c.collision_penalty_factor = Edge::collisionPenaltyFactor();
    c.time_penalty_factor = Edge::timePenaltyFactor();
    c.branching_factor = config.branchingFactor();
    // changeHeuristicIfTooManyRibbons (AStarPlanner.cpp:18, RibbonManager.cpp:381-385)
    RibbonManager::Heuristic h = ribbonManager.*get(RibbonHeuristic());
    if (ribbonManager.get().size() > 5) h = RibbonManager::MaxDistance;
    m_Heuristic = heuristicId(h);
    // MaxDistance and the point-robot TSP variants are evaluated by the kernels (h >= 0 in the results; -1 = not evaluated)
    m_HOnDevice = m_Heuristic == PPE_H_MAX_DISTANCE || m_Heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL ||
                  m_Heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K;
    c.heuristic = m_Heuristic;
    c.tsp_k = (m_Heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K || m_Heuristic == PPE_H_TSP_DUBINS_NO_SPLIT_K) ? ribbonManager.*get(RibbonK()) : 0;
    c.reserved0 = 0;
    check(ppe_set_config(m_Ctx, &c), "ppe_set_config");

    // static map: rasterise Map::isBlocked at cell centres (GridWorldMap cells are res x res squares).  The Executive
    // hands the same immutable Map object to every planning cycle (executive.cpp:182-190), so the bitmap already on the
    // device is kept while the caller's PpeWorldCache (it lives next to the ppe_ctx) says this very object went up last
    // AND the engine's map generation shows nobody else has uploaded a map to this context since.
    const Map::SharedPtr& map = config.map();
    const bool sameMap = m_Cache && m_Cache->valid && m_Cache->generation == ppe_map_generation(m_Ctx) &&
                         !m_Cache->map.expired() && m_Cache->map.lock() == map;
    if (!sameMap) {
        const double res = map->resolution();
        const double* ext = map->extremes();
        if (typeid(*map) == typeid(Map)) {
            check(ppe_set_map_none(m_Ctx), "ppe_set_map_none"); // the base class: never blocked (Map.cpp:4-6)
        } else {
            // any other Map must describe itself as a grid (positive resolution, finite extremes anchored at the origin,
            // as GridWorldMap does); GeoTiffMap overrides neither -> refuse instead of silently dropping its isBlocked
            if (!(res > 0) || !(ext[1] < 1e300) || !(ext[3] < 1e300) || ext[0] != 0 || ext[2] != 0)
                throw std::runtime_error("BatchedAStarPlanner: this Map subclass does not expose a grid (resolution() > 0, finite "
                                         "extremes() with origin 0,0); convert it to an occupancy bitmap first");
            const int cols = (int)std::llround((ext[1] - ext[0]) / res), rows = (int)std::llround((ext[3] - ext[2]) / res);
            const int stride = (cols + 7) / 8;
            std::vector<uint8_t> bits((size_t)rows * stride, 0);
            for (int r = 0; r < rows; r++)
                for (int cc = 0; cc < cols; cc++)
                    if (map->isBlocked((cc + 0.5) * res, (r + 0.5) * res)) bits[(size_t)r * stride + (cc >> 3)] |= (uint8_t)(1u << (cc & 7));
            check(ppe_set_map_bitmap(m_Ctx, bits.data(), rows, cols, stride, res), "ppe_set_map_bitmap");
        }
        if (m_Cache) {
            m_Cache->map = map;
            m_Cache->generation = ppe_map_generation(m_Ctx);
            m_Cache->valid = true;
        }
    }

    // dynamic obstacles in container iteration order (the summation order of collisionExists)
    const DynamicObstaclesManager* mgr = &config.obstaclesManager();
    if (auto* b = dynamic_cast<const BinaryDynamicObstaclesManager*>(mgr)) {
        std::vector<double> x, y, yaw, sp, tm, wd, ln;
        for (const auto& o : b->get()) {
            x.push_back(o.second.X); y.push_back(o.second.Y); yaw.push_back(o.second.Yaw); sp.push_back(o.second.Speed);
            tm.push_back(o.second.Time); wd.push_back(o.second.Width); ln.push_back(o.second.Length);
        }
        check(ppe_set_obstacles_binary(m_Ctx, (int)x.size(), x.data(), y.data(), yaw.data(), sp.data(), tm.data(), wd.data(), ln.data()),
              "ppe_set_obstacles_binary");
    } else if (auto* g = dynamic_cast<const GaussianDynamicObstaclesManager*>(mgr)) {
        std::vector<double> x, y, yaw, sp, tm, cov;
        for (const auto& o : g->get()) {
            x.push_back(o.second.X); y.push_back(o.second.Y); yaw.push_back(o.second.Yaw); sp.push_back(o.second.Speed);
            tm.push_back(o.second.Time);
            cov.push_back(o.second.covariance(0, 0)); cov.push_back(o.second.covariance(0, 1));
            cov.push_back(o.second.covariance(1, 0)); cov.push_back(o.second.covariance(1, 1));
        }
        check(ppe_set_obstacles_gaussian(m_Ctx, (int)x.size(), x.data(), y.data(), yaw.data(), sp.data(), tm.data(), cov.data()),
              "ppe_set_obstacles_gaussian");
    } else {
        check(ppe_set_obstacles_none(m_Ctx), "ppe_set_obstacles_none");
    }
    check(ppe_clear_ribbon_sets(m_Ctx), "ppe_clear_ribbon_sets");
}

void BatchedAStarPlanner::prepareWorld(const RibbonManager& ribbonManager, const State& start, const PlannerConfig& config) {
    uploadWorld(ribbonManager, start, config);
    m_Perm.clear();
    m_SampleXY.clear();
    m_Log.clear();
    m_LogApplied = 0;
    m_Expansions.clear();
    check(ppe_clear_samples(m_Ctx), "ppe_clear_samples");
}

// SamplingBasedPlanner::addSamples (SamplingBasedPlanner.cpp:157-164): the generator is the reference's own (the
// sample sequence must be the reference's); Map::isBlocked runs on the device for the whole batch and the free
// states are appended to the resident sample set in the same order as to m_Samples.
void BatchedAStarPlanner::addSamplesResident(StateGenerator& generator, int n) {
    Stopwatch sw(m_TSamples);
    m_AttemptedSamples += n;
    if (n <= 0) return;
    std::vector<State>& gen = m_Scratch;
    gen.clear();
    m_GenX.resize(n); m_GenY.resize(n); m_GenH.resize(n); m_Keep.resize(n);
    for (int i = 0; i < n; i++) {
        gen.push_back(generator.generate());
        m_GenX[i] = gen.back().x(); m_GenY[i] = gen.back().y(); m_GenH[i] = gen.back().heading();
    }
    // the resident set must mirror m_Samples before anything is appended to either
    if ((size_t)ppe_sample_count(m_Ctx) != m_Samples.size()) {
        check(ppe_clear_samples(m_Ctx), "ppe_clear_samples");
        m_Samples.clear();
    }
    const int64_t kept = ppe_add_samples(m_Ctx, n, m_GenX.data(), m_GenY.data(), m_GenH.data(), m_Keep.data());
    check((int)std::min<int64_t>(kept, 0), "ppe_add_samples");
    for (int i = 0; i < n; i++)
        if (m_Keep[i]) m_Samples.push_back(gen[i]);
    m_Batches++;
}

int32_t BatchedAStarPlanner::internRibbons(const RibbonManager& ribbons) {
    m_RibbonBuf.clear();
    for (const auto& r : ribbons.get()) {
        m_RibbonBuf.push_back(r.start().first); m_RibbonBuf.push_back(r.start().second);
        m_RibbonBuf.push_back(r.end().first); m_RibbonBuf.push_back(r.end().second);
    }
    int32_t setId = -1;
    check(ppe_put_ribbon_set(m_Ctx, (int)(m_RibbonBuf.size() / 4), m_RibbonBuf.data(), ribbons.coverageCompletedTime(), &setId),
          "ppe_put_ribbon_set");
    return setId;
}

// Edge members written by computeTrueCost (Edge.cpp:177-199), Vertex members (Vertex.cpp:102-104, :49-64) and the
// child's ribbon manager, from one engine result.
void BatchedAStarPlanner::fillChild(const std::shared_ptr<Vertex>& v, const double qi[3], const double param[3], double rho, int type,
                                    double wSpeed, double wStart, double wEnd, bool infeasible, double approx, double trueCost,
                                    double penalty, double g, double h, double cct, bool ribbonsChanged, const double* ribbons,
                                    int nRibbons) {
    Edge& edge = *v->parentEdge();
    DubinsPath p;
    p.qi[0] = qi[0]; p.qi[1] = qi[1]; p.qi[2] = qi[2];
    p.param[0] = param[0]; p.param[1] = param[1]; p.param[2] = param[2];
    p.rho = rho; p.type = (DubinsPathType)type;
    DubinsWrapper w;
    w.fill(p, wSpeed, wStart);
    if (wEnd < w.getEndTime()) w.updateEndTime(wEnd);
    edge.*get(EdgeWrapper()) = w;
    edge.*get(EdgeInfeasible()) = infeasible;
    edge.*get(EdgeApprox()) = approx;
    edge.*get(EdgeTrue()) = trueCost;
    edge.*get(EdgePenalty()) = penalty;
    (*v).*get(VertexG()) = g;
    RibbonManager& rm = v->ribbonManager();
    if (ribbonsChanged) {
        std::list<Ribbon>& list = rm.*get(RibbonList());
        list.clear();
        for (int k2 = 0; k2 < nRibbons; k2++) list.emplace_back(ribbons[4 * k2], ribbons[4 * k2 + 1], ribbons[4 * k2 + 2], ribbons[4 * k2 + 3]);
    }
    rm.*get(RibbonCct()) = cct;
    if (m_HOnDevice && h >= 0) (*v).*get(VertexH()) = h;
    else v->computeApproxToGo(m_Config); // what the engine does not evaluate stays on the host (Dubins TSP variants, long lists)
    if (m_Config.visualizations()) dumpTrajectory(v);
}

// The "Trajectory:" block Edge::computeTrueCost writes for every evaluated edge when visualizations are on
// (Edge.cpp:122-143): one State line per int(1 / increment) + 1 sample points, f / g / h as the reference prints them
// (g so far = parent g + time so far; the running collision penalty is only known for the whole edge and is left out).
void BatchedAStarPlanner::dumpTrajectory(const std::shared_ptr<Vertex>& v) {
    std::ostream& os = m_Config.visualizationStream();
    os << "Trajectory:" << std::endl;
    Edge& edge = *v->parentEdge();
    const DubinsWrapper& w = edge.*get(EdgeWrapper());
    const auto parent = v->parent();
    const double startG = parent->currentCost(), startH = parent->approxToGo();
    const double dt = m_Config.collisionCheckingIncrement() / m_Config.maxSpeed();
    State s = parent->state();
    s.time() += fmod(s.time() - m_Config.startStateTime(), dt);
    int visCount = 0;
    const double endTime = v->state().time();
    while (s.time() < endTime) {
        try { w.sample(s); } catch (std::runtime_error&) { break; }
        if (visCount-- <= 0) {
            visCount = int(1.0 / m_Config.collisionCheckingIncrement());
            const double gSoFar = startG + (s.time() - parent->state().time());
            os << "State: (" << s.toStringRad() << "), f: " << gSoFar + startH << ", g: " << gSoFar << ", h: " << startH << " trajectory" << std::endl;
        }
        s.time() += dt;
    }
}

// AStarPlanner.cpp:46-59: the previous plan, wrapper by wrapper, each one an edge from the end of the one before
// (its ribbons-after and g feed the next), evaluated by the engine as has_path edges.
Vertex::SharedPtr BatchedAStarPlanner::revalidatePreviousPlan(const Vertex::SharedPtr& startV, const DubinsPlan& previousPlan, bool visualize) {
    Vertex::SharedPtr lastPlanEnd = startV;
    if (previousPlan.empty()) return lastPlanEnd;
    for (const auto& p : previousPlan.get()) {
        if (p.getEndTime() <= startV->state().time()) continue;
        if (p.getNetTime() == 0) continue;
        const bool cov = p.getRho() == m_Config.coverageTurningRadius();
        const State& src = lastPlanEnd->state();
        ppe_edge e;
        std::memset(&e, 0, sizeof e);
        e.src[0] = src.x(); e.src[1] = src.y(); e.src[2] = src.heading(); e.src[3] = src.speed(); e.src[4] = src.time();
        e.src_g = lastPlanEnd->currentCost();
        e.ribbon_set = internRibbons(lastPlanEnd->ribbonManager());
        const DubinsPath& path = p.unwrap();
        e.has_path = 1;
        e.path_qi[0] = path.qi[0]; e.path_qi[1] = path.qi[1]; e.path_qi[2] = path.qi[2];
        e.path_param[0] = path.param[0]; e.path_param[1] = path.param[1]; e.path_param[2] = path.param[2];
        e.path_rho = path.rho;
        e.path_type = (int32_t)path.type;
        e.w_speed = p.getSpeed();
        e.w_start_time = p.getStartTime();
        e.w_end_time = p.getEndTime();
        e.dst[3] = p.getSpeed();
        e.coverage_allowed = cov;
        ppe_edge_result r;
        check(ppe_true_cost_batch(m_Ctx, 1, &e, &r), "ppe_true_cost_batch");
        m_TrueCostEdges++;
        m_Batches++;
        if (r.status != PPE_EDGE_OK)
            throw std::runtime_error("previous-plan edge failed where the reference throws (status " + std::to_string(r.status) + ")");
        lastPlanEnd = Vertex::connect(lastPlanEnd, p, cov); // Edge::setEnd(wrapper), Edge.cpp:208-215
        lastPlanEnd->state() = State(r.end[0], r.end[1], r.end[2], r.end[3], r.end[4]);
        if (r.ribbons_changed) {
            m_RibbonBuf.resize((size_t)std::max(1, r.n_ribbons_after) * 4);
            check(ppe_get_ribbons_after(m_Ctx, 0, m_RibbonBuf.data(), r.n_ribbons_after), "ppe_get_ribbons_after");
        }
        fillChild(lastPlanEnd, r.path_qi, r.path_param, r.path_rho, r.path_type, r.w_speed, r.w_start_time, r.w_end_time,
                  r.infeasible != 0, r.approx_cost, r.true_cost, r.collision_penalty, r.g, r.h, r.coverage_completed_time,
                  r.ribbons_changed != 0, m_RibbonBuf.data(), r.n_ribbons_after);
        if (visualize) { // AStarPlanner.cpp:78-79
            lastPlanEnd->computeApproxToGo(m_Config);
            visualizeVertex(lastPlanEnd, "lastPlanEnd", false);
        }
        if (lastPlanEnd->parentEdge()->infeasible()) {
            lastPlanEnd = startV;
            break;
        }
        if (goalCondition(lastPlanEnd)) break;
    }
    return lastPlanEnd;
}

// AStarPlanner::expandToCoverSpecificSamples (AStarPlanner.cpp:150-162): the Brown-path states, both speeds, coverage
// radius -- one engine batch per call.
void BatchedAStarPlanner::expandSpecific(const Vertex::SharedPtr& root, const std::vector<State>& samples, bool coverageAllowed) {
    if (!(m_Config.coverageTurningRadius() > 0) || samples.empty()) return;
    const State& src = root->state();
    const int32_t setId = internRibbons(root->ribbonManager());
    m_Edges.clear();
    for (auto s : samples) {
        for (const auto& speed : {m_Config.maxSpeed(), m_Config.slowSpeed()}) {
            ppe_edge e;
            std::memset(&e, 0, sizeof e);
            e.src[0] = src.x(); e.src[1] = src.y(); e.src[2] = src.heading(); e.src[3] = src.speed(); e.src[4] = src.time();
            e.src_g = root->currentCost();
            e.ribbon_set = setId;
            e.dst[0] = s.x(); e.dst[1] = s.y(); e.dst[2] = s.heading(); e.dst[3] = speed;
            e.has_path = 0;
            e.coverage_allowed = coverageAllowed; // radius of the edge follows from it (Edge.cpp:73-77)
            m_Edges.push_back(e);
        }
    }
    m_Results.resize(m_Edges.size());
    check(ppe_true_cost_batch(m_Ctx, (int64_t)m_Edges.size(), m_Edges.data(), m_Results.data()), "ppe_true_cost_batch");
    m_TrueCostEdges += (long)m_Edges.size();
    m_Batches++;
    for (size_t i = 0; i < m_Edges.size(); i++) {
        const ppe_edge& e = m_Edges[i];
        const ppe_edge_result& r = m_Results[i];
        if (r.status != PPE_EDGE_OK)
            throw std::runtime_error("edge evaluation failed where the reference throws (status " + std::to_string(r.status) + ")");
        State end(r.end[0], r.end[1], r.end[2], r.end[3], r.end[4]);
        auto v = Vertex::connect(root, end, m_Config.coverageTurningRadius(), coverageAllowed);
        if (r.ribbons_changed) {
            m_RibbonBuf.resize((size_t)std::max(1, r.n_ribbons_after) * 4);
            check(ppe_get_ribbons_after(m_Ctx, (int64_t)i, m_RibbonBuf.data(), r.n_ribbons_after), "ppe_get_ribbons_after");
        }
        fillChild(v, r.path_qi, r.path_param, r.path_rho, r.path_type, r.w_speed, r.w_start_time, r.w_end_time, r.infeasible != 0,
                  r.approx_cost, r.true_cost, r.collision_penalty, r.g, r.h, r.coverage_completed_time, r.ribbons_changed != 0,
                  m_RibbonBuf.data(), r.n_ribbons_after);
        pushVertexQueue(v);
        (void)e;
    }
}

// AStarPlanner::plan (AStarPlanner.cpp:12-132), statement for statement where the order is observable (every now()
// call, every push into the open list, the sample generator's draw sequence), with the engine behind the three places
// that evaluate edges or test samples.
Planner::Stats BatchedAStarPlanner::plan(const RibbonManager& ribbonManager, const State& start, PlannerConfig config,
                                         const DubinsPlan& previousPlan, double timeRemaining) {
    m_TrueCostEdges = m_DubinsSolves = m_Batches = m_FrontierVertices = m_FrontierHits = m_ExactExpansions = 0;
    m_ExactTies = m_ExactOverflow = 0;
    m_TEngine = m_TReplay = m_TSamples = m_TExact = 0;
    m_Config = std::move(config); // before the first now(), :14
    const double endTime = timeRemaining + now();
    m_Config.setStartStateTime(start.time());
    prepareWorld(ribbonManager, start, m_Config);
    m_RibbonManager = ribbonManager;
    m_RibbonManager.changeHeuristicIfTooManyRibbons();
    if (m_RibbonManager.done()) m_RibbonManager.setCoverageCompletedTime(start.time());
    m_Stats = Stats();
    m_IterationCount = 0;
    m_StartStateTime = start.time();
    m_Samples.clear();
    m_AttemptedSamples = 0;
    const double minSpeed = m_Config.maxSpeed(), maxSpeed = m_Config.maxSpeed();
    const double magnitude = m_Config.maxSpeed() * m_Config.timeHorizon();
    const double* mapExtremes = m_Config.map()->extremes();
    const double minX = fmax(start.x() - magnitude, mapExtremes[0]);
    const double maxX = fmin(start.x() + magnitude, mapExtremes[1]);
    const double minY = fmax(start.y() - magnitude, mapExtremes[2]);
    const double maxY = fmin(start.y() + magnitude, mapExtremes[3]);
    const auto seed = (unsigned long)endTime; // :33
    StateGenerator generator(minX, maxX, minY, maxY, minSpeed, maxSpeed, seed, m_RibbonManager);
    auto startV = Vertex::makeRoot(start, m_RibbonManager);
    startV->state().speed() = m_Config.maxSpeed();
    startV->computeApproxToGo(m_Config);
    m_BestVertex = nullptr;
    std::vector<State> brownPathSamples;
    if (m_Config.useBrownPaths()) brownPathSamples = m_RibbonManager.findNearStatesOnRibbons(start, m_Config.coverageTurningRadius());

    Vertex::SharedPtr lastPlanEnd = revalidatePreviousPlan(startV, previousPlan, false); // :46-59

    while (now() < endTime) { // :61
        clearVertexQueue();
        // cached frontier results refer to vertices of the search tree that is being thrown away; their ribbon sets go too
        m_Expansions.clear();
        check(ppe_clear_ribbon_sets(m_Ctx), "ppe_clear_ribbon_sets");
        if (m_BestVertex && m_BestVertex->f() <= startV->f()) {
            *m_Config.output() << "Found best possible plan, assuming heuristic admissibility" << std::endl;
            break;
        }
        visualizeVertex(startV, "start", false);
        if (m_Config.visualizations()) {
            lastPlanEnd = revalidatePreviousPlan(startV, previousPlan, true); // :69-86
            m_Config.visualizationStream() << "Incumbent f-value: " << (m_BestVertex ? m_BestVertex->f() : 0) << std::endl;
            m_Config.visualizationStream() << m_RibbonManager.dumpRibbons() << "End Ribbons" << std::endl;
        }
        pushVertexQueue(startV);
        if (lastPlanEnd != startV) pushVertexQueue(lastPlanEnd);
        expandSpecific(startV, brownPathSamples, true); // :99
        if (m_Samples.size() < (size_t)m_Config.initialSamples()) addSamplesResident(generator, m_Config.initialSamples());
        else addSamplesResident(generator, (int)m_Samples.size()); // double the samples, :101-102
        if (m_Config.visualizations()) {
            for (const auto& s : m_Samples)
                m_Config.visualizationStream() << "State: (" << s.toStringRad() << "), f: " << 0 << ", g: " << 0 << ", h: " << 0
                                               << " sample" << std::endl;
        }
        auto v = aStar(m_Config.obstaclesManager(), endTime); // the reference's own loop; it calls our expand()
        if (!m_BestVertex || (v && v->f() + 0.0 < m_BestVertex->f())) {
            m_BestVertex = v;
            if (v && m_Config.visualizations()) {
                visualizePlan(tracePlan(v, false, m_Config.obstaclesManager()));
                visualizeVertex(v, "goal", false);
            }
        }
        m_Stats.Iterations++;
    }
    m_Expansions.clear();
    m_Stats.Samples = m_Samples.size();
    if (!m_BestVertex) {
        *m_Config.output() << "Failed to find a plan" << std::endl;
    } else {
        m_Stats.PlanFValue = m_BestVertex->f();
        m_Stats.PlanDepth = m_BestVertex->getDepth();
        m_Stats.PlanTimePenalty = (m_BestVertex->state().time() - m_StartStateTime) * Edge::timePenaltyFactor();
        m_Stats.PlanHValue = m_BestVertex->approxToGo();
        m_Stats.Plan = std::move(tracePlan(m_BestVertex, false, m_Config.obstaclesManager()));
    }
    return m_Stats;
}

void BatchedAStarPlanner::expand(const std::shared_ptr<Vertex>& sourceVertex, const DynamicObstaclesManager& obstacles) {
    (void)obstacles;
    if (m_Frontier <= 0 || m_Config.branchingFactor() > 16 || m_Config.branchingFactor() < 1) expandExact(sourceVertex);
    else expandFrontier(sourceVertex);
}

// Applies the logged device expansions to the (m_Keys, m_Perm) arrangement: what std::make_heap and the pop_heap calls
// of each of those expansions would have left in the reference's m_Samples.
void BatchedAStarPlanner::syncSampleHeap() {
    if (m_Perm.size() > m_Samples.size()) { m_Perm.clear(); m_SampleXY.clear(); }
    for (; m_LogApplied < m_Log.size(); m_LogApplied++) {
        const ExpansionLog& e = m_Log[m_LogApplied];
        const size_t nAll = e.nSamples;
        for (size_t i = m_Perm.size(); i < nAll; i++) {
            m_Perm.push_back((uint32_t)i);
            m_SampleXY.push_back(m_Samples[i].x());
            m_SampleXY.push_back(m_Samples[i].y());
        }
        m_Dist.resize(nAll);
        const double* xy = m_SampleXY.data();
        for (size_t i = 0; i < nAll; i++) {
            const double x = xy[2 * i], y = xy[2 * i + 1];
            m_Dist[i] = sqrt((x - e.x) * (x - e.x) + (y - e.y) * (y - e.y));
        }
        m_Keys.resize(nAll);
        for (size_t i = 0; i < nAll; i++) m_Keys[i] = m_Dist[m_Perm[i]];
        ppe_heap::make_heap(m_Keys.data(), m_Perm.data(), (std::ptrdiff_t)nAll);
        for (uint32_t q = 0; q < e.pops; q++) ppe_heap::pop_heap(m_Keys.data(), m_Perm.data(), (std::ptrdiff_t)(nAll - q));
    }
}

// The vertex the reference's loop just popped, plus the best vertices still on the open list, in one engine call.
void BatchedAStarPlanner::speculate(const std::shared_ptr<Vertex>& first) {
    Stopwatch sw(m_TEngine);
    std::vector<std::shared_ptr<Vertex>> batch;
    batch.push_back(first);
    if (m_Frontier > 1) {
        // best-first walk over the open list's heap (front = smallest f, AStarPlanner.cpp:6-10) without modifying it
        const std::vector<std::shared_ptr<Vertex>>& open = this->*get(OpenList());
        typedef std::pair<double, size_t> Node;
        auto worse = [](const Node& a, const Node& b) { return a.first > b.first || (a.first == b.first && a.second > b.second); };
        std::vector<Node> walk;
        if (!open.empty()) walk.push_back(Node(open[0]->f(), 0));
        size_t visited = 0;
        while ((int)batch.size() < m_Frontier && !walk.empty() && visited < 4 * (size_t)m_Frontier) {
            std::pop_heap(walk.begin(), walk.end(), worse);
            const size_t idx = walk.back().second;
            walk.pop_back();
            visited++;
            const std::shared_ptr<Vertex>& u = open[idx];
            if (u != first && !goalCondition(u) && !m_Expansions.count(u.get())) batch.push_back(u);
            for (size_t child = 2 * idx + 1; child <= 2 * idx + 2 && child < open.size(); child++) {
                walk.push_back(Node(open[child]->f(), child));
                std::push_heap(walk.begin(), walk.end(), worse);
            }
        }
    }
    const int n = (int)batch.size();
    const double inc = m_Config.collisionCheckingIncrement();
    m_Verts.resize(n);
    for (int i = 0; i < n; i++) {
        const Vertex& u = *batch[i];
        ppe_vertex& pv = m_Verts[i];
        std::memset(&pv, 0, sizeof pv);
        const State& st = u.state();
        pv.state[0] = st.x(); pv.state[1] = st.y(); pv.state[2] = st.heading(); pv.state[3] = st.speed(); pv.state[4] = st.time();
        pv.g = u.currentCost();
        pv.ribbon_set = internRibbons(u.ribbonManager());
        if (!u.done()) { // SamplingBasedPlanner.cpp:65-68
            const State s = u.getNearestPointAsState();
            if (st.distanceTo(s) > inc) {
                pv.has_endpoint = 1;
                pv.endpoint[0] = s.x(); pv.endpoint[1] = s.y(); pv.endpoint[2] = s.heading();
            }
        }
    }
    const int stride = ppe_expand_stride(m_Ctx);
    m_Children.resize((size_t)n * stride);
    m_NChildren.resize(n); m_Flags.resize(n); m_Popped.resize(n);
    const int64_t solves0 = ppe_expand_solve_count(m_Ctx);
    check(ppe_expand_batch(m_Ctx, n, m_Verts.data(), m_NChildren.data(), m_Children.data(), m_Flags.data(), m_Popped.data()),
          "ppe_expand_batch");
    m_DubinsSolves += (long)(ppe_expand_solve_count(m_Ctx) - solves0);
    m_Batches++;
    m_FrontierVertices += n;
    int64_t nPool = 0;
    const double* pool = ppe_ribbon_pool(m_Ctx, &nPool);
    for (int i = 0; i < n; i++) {
        Expansion& ex = m_Expansions[batch[i].get()];
        ex.vertex = batch[i];
        ex.flags = m_Flags[i];
        ex.popped = m_Popped[i];
        const int nc = m_NChildren[i];
        ex.children.assign(m_Children.begin() + (size_t)i * stride, m_Children.begin() + (size_t)i * stride + nc);
        ex.ribbonStart.assign(nc, -1);
        ex.ribbons.clear();
        for (int c = 0; c < nc; c++) {
            const ppe_child& ch = ex.children[c];
            if (ch.status == PPE_EDGE_OK && ch.ribbons_changed) {
                if (ch.ribbons_offset < 0 || ch.ribbons_offset + ch.n_ribbons_after > nPool)
                    throw std::runtime_error("ppe_expand_batch: ribbons-after of a child were not materialised");
                ex.ribbonStart[c] = (int)(ex.ribbons.size() / 4);
                ex.ribbons.insert(ex.ribbons.end(), pool + 4 * ch.ribbons_offset, pool + 4 * (ch.ribbons_offset + ch.n_ribbons_after));
            }
        }
        m_TrueCostEdges += nc;
    }
}

void BatchedAStarPlanner::expandFrontier(const std::shared_ptr<Vertex>& sourceVertex) {
    // callers that filled m_Samples through the base class (SamplingBasedPlanner::addSamples is not virtual): mirror it
    if ((size_t)ppe_sample_count(m_Ctx) != m_Samples.size()) {
        check(ppe_clear_samples(m_Ctx), "ppe_clear_samples");
        const size_t n = m_Samples.size();
        m_GenX.resize(n); m_GenY.resize(n); m_GenH.resize(n); m_Keep.resize(n);
        for (size_t i = 0; i < n; i++) { m_GenX[i] = m_Samples[i].x(); m_GenY[i] = m_Samples[i].y(); m_GenH[i] = m_Samples[i].heading(); }
        const int64_t kept = n ? ppe_add_samples(m_Ctx, (int64_t)n, m_GenX.data(), m_GenY.data(), m_GenH.data(), m_Keep.data()) : 0;
        if (kept != (int64_t)n) throw std::runtime_error("BatchedAStarPlanner: m_Samples holds states the map blocks");
    }
    auto it = m_Expansions.find(sourceVertex.get());
    if (it == m_Expansions.end()) {
        speculate(sourceVertex);
        it = m_Expansions.find(sourceVertex.get());
    } else {
        m_FrontierHits++;
    }
    static const bool forceExact = getenv("PPE_HARNESS_TEST_FORCE_EXACT") != nullptr; // tests: every 3rd expansion replays on the host
    if ((it->second.flags & (PPE_EXPAND_TIE | PPE_EXPAND_OVERFLOW)) || (forceExact && m_Stats.Expanded % 3 == 2)) {
        if (it->second.flags & PPE_EXPAND_TIE) m_ExactTies++;
        if (it->second.flags & PPE_EXPAND_OVERFLOW) m_ExactOverflow++;
        m_Expansions.erase(it);
        expandExact(sourceVertex);
        return;
    }
    Stopwatch sw(m_TReplay);
    visualizeVertex(sourceVertex, "vertex", true);
    const Expansion& ex = it->second;
    const State& src = sourceVertex->state();
    const double qi[3] = {src.x(), src.y(), src.yaw()};
    for (size_t c = 0; c < ex.children.size(); c++) {
        const ppe_child& ch = ex.children[c];
        if (ch.status != PPE_EDGE_OK)
            throw std::runtime_error("edge evaluation failed where the reference throws (status " + std::to_string(ch.status) + ")");
        State end(ch.end[0], ch.end[1], ch.end[2], ch.end[3], ch.end[4]);
        const double rho = ch.coverage_allowed ? m_Config.coverageTurningRadius() : m_Config.turningRadius();
        auto v = Vertex::connect(sourceVertex, end, rho, ch.coverage_allowed != 0);
        const double* rib = ex.ribbonStart[c] >= 0 ? ex.ribbons.data() + 4 * (size_t)ex.ribbonStart[c] : nullptr;
        fillChild(v, qi, ch.path_param, rho, ch.path_type, ch.end[3], src.time(), ch.w_end_time, ch.infeasible != 0, ch.approx_cost,
                  ch.true_cost, ch.collision_penalty, ch.g, ch.h, ch.coverage_completed_time, ch.ribbons_changed != 0, rib,
                  ch.n_ribbons_after);
        pushVertexQueue(v);
    }
    ExpansionLog lg;
    lg.x = src.x(); lg.y = src.y(); lg.pops = (uint32_t)ex.popped; lg.nSamples = (uint32_t)m_Samples.size();
    m_Log.push_back(lg);
    m_Expansions.erase(it);
    m_Stats.Expanded++;
}

void BatchedAStarPlanner::expandExact(const std::shared_ptr<Vertex>& sourceVertex) {
    Stopwatch sw(m_TExact);
    m_ExactExpansions++;
    syncSampleHeap(); // the arrangement of m_Samples as the reference would have it right now
    visualizeVertex(sourceVertex, "vertex", true);
    const State& src = sourceVertex->state();
    const double inc = m_Config.collisionCheckingIncrement();

    // configurations, SamplingBasedPlanner.cpp:58-63
    const double speeds[2] = {m_Config.maxSpeed(), m_Config.maxSpeed() == m_Config.slowSpeed() ? -1 : m_Config.slowSpeed()};
    const int nTurningRadii = 2;
    const double turningRadii[nTurningRadii] = {m_Config.turningRadius(),
                                                m_Config.coverageTurningRadius() == m_Config.turningRadius() ? -1 : m_Config.coverageTurningRadius()};

    // intern the parent's ribbon set once; every edge of this expansion refers to it
    const int32_t setId = internRibbons(sourceVertex->ribbonManager());

    m_Edges.clear();
    auto baseEdge = [&]() {
        ppe_edge e;
        std::memset(&e, 0, sizeof e);
        e.src[0] = src.x(); e.src[1] = src.y(); e.src[2] = src.heading(); e.src[3] = src.speed(); e.src[4] = src.time();
        e.src_g = sourceVertex->currentCost();
        e.ribbon_set = setId;
        return e;
    };

    // (a) nearest ribbon endpoint, :65-81 -- path-less edges, the engine solves them (Edge.cpp:78-80)
    if (!sourceVertex->done()) {
        auto s = sourceVertex->getNearestPointAsState();
        if (src.distanceTo(s) > inc) {
            for (double speed : speeds) {
                if (speed <= 0) continue;
                for (double turningRadius : turningRadii) {
                    if (turningRadius <= 0) continue;
                    ppe_edge e = baseEdge();
                    e.dst[0] = s.x(); e.dst[1] = s.y(); e.dst[2] = s.heading(); e.dst[3] = speed;
                    e.has_path = 0;
                    e.coverage_allowed = turningRadius == m_Config.coverageTurningRadius();
                    m_Edges.push_back(e);
                }
            }
        }
    }
    const size_t nEndpointEdges = m_Edges.size();

    // (b) k nearest samples by Dubins distance, :83-133.  Same heap operations on m_Samples as the reference, in the
    // same order.  The Dubins solves of the samples the loop is about to pop ride in one K1 launch per chunk: the next
    // `m_KnnChunk` samples in pop order are read off the heap WITHOUT touching it (best-first walk over the heap's
    // tree with a small auxiliary queue), solved on the device, and then the reference's loop body is replayed with
    // real pops.  A popped sample that is not where the walk predicted (only possible among samples at exactly equal
    // distance) is looked up in the chunk, and solved on its own if it is not there.
    auto dubinsComp = [](const Candidate& a, const Candidate& b) { return a.approxCost < b.approxCost; };
    // std::make_heap(m_Samples.begin(), m_Samples.end(), comp) with comp = "farther from src is lower priority", on
    // the (distance, index) representation: every distance is computed once instead of twice per comparison
    const size_t nAll = m_Samples.size();
    if (m_Perm.size() > nAll) { m_Perm.clear(); m_SampleXY.clear(); }
    for (size_t i = m_Perm.size(); i < nAll; i++) { // addSamples appended at the end
        m_Perm.push_back((uint32_t)i);
        m_SampleXY.push_back(m_Samples[i].x());
        m_SampleXY.push_back(m_Samples[i].y());
    }
    // State::distanceTo (State.cpp:91-93), same expression, streamed over the stored samples; then gathered into
    // the heap's arrangement
    m_Dist.resize(nAll);
    {
        const double sx = src.x(), sy = src.y();
        const double* xy = m_SampleXY.data();
        for (size_t i = 0; i < nAll; i++) {
            const double x = xy[2 * i], y = xy[2 * i + 1];
            m_Dist[i] = sqrt((x - sx) * (x - sx) + (y - sy) * (y - sy));
        }
    }
    m_Keys.resize(nAll);
    for (size_t i = 0; i < nAll; i++) m_Keys[i] = m_Dist[m_Perm[i]];
    ppe_heap::make_heap(m_Keys.data(), m_Perm.data(), (std::ptrdiff_t)nAll);
    std::vector<Candidate> bestSamplesHeaps[nTurningRadii];
    bool doneChecks[nTurningRadii] = {false, false};
    const size_t nSamples = m_Samples.size();
    const size_t kBranch = (size_t)k();
    size_t pops = 0; // how many samples the reference's loop has consumed = how far the heap has shrunk
    std::vector<State>& chunk = m_Scratch;
    typedef std::pair<double, size_t> Node; // (distance to src, index in the heap array)
    auto nodeGreater = [](const Node& a, const Node& b) { return a.first > b.first || (a.first == b.first && a.second > b.second); };
    std::vector<Node> walk;
    std::vector<char> used;
    struct Solve { double q0[3], param[3], length; int type; bool valid; };
    static const bool mispredict = getenv("PPE_HARNESS_TEST_MISPREDICT") != nullptr; // tests/test_harness_host_logic.py
    auto sameState = [](const State& a, const State& b) {
        return a.x() == b.x() && a.y() == b.y() && a.heading() == b.heading() && a.speed() == b.speed() && a.time() == b.time();
    };
    // solves src -> every sample of `chunk` at every radius still in play; slot[2 c + j] indexes the K1 outputs
    std::vector<long> slot;
    auto solveChunk = [&]() {
        const size_t m = chunk.size();
        m_Q0.resize(6 * m); m_Q1.resize(6 * m); m_Rho.resize(2 * m);
        m_Param.resize(6 * m); m_Length.resize(2 * m); m_Type.resize(2 * m); m_Err.resize(2 * m);
        slot.assign(2 * m, -1);
        size_t nq = 0;
        for (size_t c = 0; c < m; c++)
            for (int j = 0; j < nTurningRadii; j++) {
                if (turningRadii[j] <= 0 || doneChecks[j]) continue;
                m_Q0[3 * nq] = src.x(); m_Q0[3 * nq + 1] = src.y(); m_Q0[3 * nq + 2] = src.yaw();
                m_Q1[3 * nq] = chunk[c].x(); m_Q1[3 * nq + 1] = chunk[c].y(); m_Q1[3 * nq + 2] = chunk[c].yaw();
                m_Rho[nq] = turningRadii[j];
                slot[2 * c + j] = (long)nq++;
            }
        if (nq) {
            check(ppe_dubins_batch(m_Ctx, (int64_t)nq, m_Q0.data(), m_Q1.data(), m_Rho.data(), m_Type.data(), m_Param.data(),
                                   m_Length.data(), m_Err.data()), "ppe_dubins_batch");
            m_DubinsSolves += (long)nq;
            m_Batches++;
        }
    };
    while (pops < nSamples && (!doneChecks[0] || !doneChecks[1])) {
        // the next chunk of samples in pop order, read off the heap [0, heapLen) without modifying it
        const size_t heapLen = nSamples - pops;
        const size_t want = std::min<size_t>((size_t)m_KnnChunk, heapLen);
        chunk.clear();
        walk.clear();
        walk.push_back(Node(m_Keys[0], 0));
        while (chunk.size() < want && !walk.empty()) {
            std::pop_heap(walk.begin(), walk.end(), nodeGreater);
            const size_t idx = walk.back().second;
            walk.pop_back();
            chunk.push_back(m_Samples[m_Perm[idx]]);
            for (size_t child = 2 * idx + 1; child <= 2 * idx + 2 && child < heapLen; child++) {
                walk.push_back(Node(m_Keys[child], child));
                std::push_heap(walk.begin(), walk.end(), nodeGreater);
            }
        }
        if (mispredict && chunk.size() > 2) { // test hook: scramble the prediction so that the look-up and single-solve paths run
            std::reverse(chunk.begin(), chunk.end());
            chunk.resize(chunk.size() - chunk.size() / 3);
        }
        solveChunk();
        const size_t m = chunk.size();
        used.assign(m, 0);
        // replay of the reference's loop body: real pops
        for (size_t c = 0; c < m && (!doneChecks[0] || !doneChecks[1]); c++) {
            State sample = m_Samples[m_Perm[0]]; // m_Samples.front() of the reference's arrangement
            ppe_heap::pop_heap(m_Keys.data(), m_Perm.data(), (std::ptrdiff_t)(nSamples - pops));
            pops++;
            // where are this sample's solves?
            size_t at = m;
            if (!used[c] && sameState(chunk[c], sample)) at = c;
            else for (size_t o = 0; o < m; o++) if (!used[o] && sameState(chunk[o], sample)) { at = o; break; }
            Solve solves[nTurningRadii];
            if (at < m) {
                used[at] = 1;
                for (int j = 0; j < nTurningRadii; j++) {
                    const long q = slot[2 * at + j];
                    solves[j].valid = q >= 0;
                    if (q < 0) continue;
                    for (int d = 0; d < 3; d++) { solves[j].q0[d] = m_Q0[3 * q + d]; solves[j].param[d] = m_Param[3 * q + d]; }
                    solves[j].length = m_Length[q];
                    solves[j].type = m_Type[q];
                }
            } else {
                // not predicted (samples at exactly equal distance popped in another order): solve this one on its own
                for (int j = 0; j < nTurningRadii; j++) {
                    solves[j].valid = false;
                    if (turningRadii[j] <= 0 || doneChecks[j]) continue;
                    const double q0[3] = {src.x(), src.y(), src.yaw()}, q1[3] = {sample.x(), sample.y(), sample.yaw()};
                    const double rho = turningRadii[j];
                    int32_t type = 0, err = 0;
                    check(ppe_dubins_batch(m_Ctx, 1, q0, q1, &rho, &type, solves[j].param, &solves[j].length, &err), "ppe_dubins_batch");
                    m_DubinsSolves++;
                    m_Batches++;
                    for (int d = 0; d < 3; d++) solves[j].q0[d] = q0[d];
                    solves[j].type = type;
                    solves[j].valid = true;
                }
            }
            for (int j = 0; j < nTurningRadii; j++) {
                if (doneChecks[j]) continue;
                const double turningRadius = turningRadii[j];
                if (turningRadius <= 0) { doneChecks[j] = true; continue; }
                auto& bestSamples = bestSamplesHeaps[j];
                if (bestSamples.size() < kBranch || bestSamples.front().length > sample.distanceTo(src)) {
                    if (src.distanceTo(sample) > inc) {
                        sample.speed() = m_Config.maxSpeed();
                        const Solve& sv = solves[j];
                        if (!sv.valid) throw std::logic_error("BatchedAStarPlanner: missing Dubins solve for a popped sample");
                        Candidate cand;
                        cand.sample = sample;
                        cand.coverageAllowed = turningRadius == m_Config.coverageTurningRadius();
                        cand.path[0] = sv.q0[0]; cand.path[1] = sv.q0[1]; cand.path[2] = sv.q0[2];
                        cand.path[3] = sv.param[0]; cand.path[4] = sv.param[1]; cand.path[5] = sv.param[2];
                        cand.path[6] = turningRadius;
                        cand.type = sv.type;
                        cand.length = sv.length;
                        // Edge::computeApproxCost(): length / end speed (= max speed) * timePenaltyFactor, Edge.cpp:11-20,64-66
                        cand.approxCost = cand.length / sample.speed() * Edge::timePenaltyFactor();
                        bestSamples.push_back(cand);
                        std::push_heap(bestSamples.begin(), bestSamples.end(), dubinsComp);
                        if (bestSamples.size() > kBranch) {
                            std::pop_heap(bestSamples.begin(), bestSamples.end(), dubinsComp);
                            bestSamples.pop_back();
                        }
                    }
                } else {
                    doneChecks[j] = true;
                }
            }
        }
    }

    // (c) winners x speeds, :134-149 -- wrapper reused, speed set per edge
    struct Winner { const Candidate* cand; double speed; };
    std::vector<Winner> winners;
    for (auto& bestSamples : bestSamplesHeaps) {
        if (bestSamples.size() > kBranch) throw std::runtime_error("Somehow got too many samples in the heap");
        for (auto& cand : bestSamples) {
            for (double speed : speeds) {
                if (speed <= 0) continue;
                ppe_edge e = baseEdge();
                e.has_path = 1;
                e.path_qi[0] = cand.path[0]; e.path_qi[1] = cand.path[1]; e.path_qi[2] = cand.path[2];
                e.path_param[0] = cand.path[3]; e.path_param[1] = cand.path[4]; e.path_param[2] = cand.path[5];
                e.path_rho = cand.path[6];
                e.path_type = cand.type;
                e.w_speed = speed;                          // wrapper.setSpeed(speed)
                e.w_start_time = src.time();                // DubinsWrapper::set, DubinsWrapper.cpp:15
                e.w_end_time = src.time() + cand.length / speed; // setEndTime, DubinsWrapper.cpp:96-98
                e.dst[0] = cand.sample.x(); e.dst[1] = cand.sample.y(); e.dst[2] = cand.sample.heading(); e.dst[3] = speed;
                e.coverage_allowed = cand.coverageAllowed;
                m_Edges.push_back(e);
                winners.push_back(Winner{&cand, speed});
            }
        }
    }

    // ---- one K2 launch for the whole expansion ---------------------------------------------------------
    m_Results.resize(m_Edges.size());
    if (!m_Edges.empty()) {
        check(ppe_true_cost_batch(m_Ctx, (int64_t)m_Edges.size(), m_Edges.data(), m_Results.data()), "ppe_true_cost_batch");
        m_TrueCostEdges += (long)m_Edges.size();
        m_Batches++;
    }

    // ---- hand the results back through the reference's own objects, in the reference's push order ------
    for (size_t i = 0; i < m_Edges.size(); i++) {
        const ppe_edge& e = m_Edges[i];
        const ppe_edge_result& r = m_Results[i];
        if (r.status != PPE_EDGE_OK)
            throw std::runtime_error("edge evaluation failed where the reference throws (status " + std::to_string(r.status) + ")");
        State end(r.end[0], r.end[1], r.end[2], r.end[3], r.end[4]);
        const double rho = e.coverage_allowed ? m_Config.coverageTurningRadius() : m_Config.turningRadius();
        auto v = Vertex::connect(sourceVertex, end, rho, e.coverage_allowed != 0);
        if (r.ribbons_changed) {
            m_RibbonBuf.resize((size_t)std::max(1, r.n_ribbons_after) * 4);
            check(ppe_get_ribbons_after(m_Ctx, (int64_t)i, m_RibbonBuf.data(), r.n_ribbons_after), "ppe_get_ribbons_after");
        }
        fillChild(v, r.path_qi, r.path_param, r.path_rho, r.path_type, r.w_speed, r.w_start_time, r.w_end_time, r.infeasible != 0,
                  r.approx_cost, r.true_cost, r.collision_penalty, r.g, r.h, r.coverage_completed_time, r.ribbons_changed != 0,
                  m_RibbonBuf.data(), r.n_ribbons_after);
        pushVertexQueue(v);
    }
    (void)nEndpointEdges;
    m_Stats.Expanded++;
}
