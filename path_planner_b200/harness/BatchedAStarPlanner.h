// BatchedAStarPlanner -- the reference's anytime A* planner with its expansion loop restructured
// to emit whole edge batches into the B200 edge engine (include/ppe.h).
//
// Drop-in: it IS an AStarPlanner (path_planner/src/planner/AStarPlanner.h); only the virtual
//   SamplingBasedPlanner::expand (SamplingBasedPlanner.h:43, SamplingBasedPlanner.cpp:52-151)
// is overridden, plus a thin plan() wrapper that uploads the read-only world state once per plan.
// Planner::Stats, tracePlan, pushVertexQueue, goalCondition, the open list and the anytime loop
// of AStarPlanner::plan are the reference's own code, untouched.
//
// Compiles against the reference headers where they lie (-I<ref>/path_planner/src
// -I<ref>/path_planner_common/include) and any dubins.h; links libppe.so.
#ifndef PPE_BATCHED_ASTAR_PLANNER_H
#define PPE_BATCHED_ASTAR_PLANNER_H

#include <cstdint>
#include <memory>
#include <vector>

#include "planner/AStarPlanner.h"
#include "ppe.h"

// "Which Map object did this engine context receive last": owned by whoever owns the ppe_ctx (one per planning
// thread), handed to every planner instance created for it (the Executive makes a new planner per cycle,
// executive.cpp:85-90).  Never shared between contexts or threads.
struct PpeWorldCache {
    std::weak_ptr<Map> map;
    uint64_t generation = 0; // ppe_map_generation(ctx) right after the upload
    bool valid = false;
};

class BatchedAStarPlanner : public AStarPlanner {
public:
    // `ctx` is borrowed (one ppe_ctx per planning thread); `knnChunk` = how many nearest samples
    // (Euclidean order) get their Dubins paths solved per K1 launch while replaying the
    // reference's k-nearest selection.
    explicit BatchedAStarPlanner(ppe_ctx* ctx, int knnChunk = 128, PpeWorldCache* cache = nullptr);
    ~BatchedAStarPlanner() override = default;

    Stats plan(const RibbonManager& ribbonManager, const State& start, PlannerConfig config,
               const DubinsPlan& previousPlan, double timeRemaining) override;

    void expand(const std::shared_ptr<Vertex>& sourceVertex, const DynamicObstaclesManager& obstacles) override;

    // What plan() does before it hands over to AStarPlanner::plan: config, map, obstacles to the engine.  Public for
    // callers that drive expand() themselves (the reference's ExpandTest1Ribbons does).
    void prepareWorld(const RibbonManager& ribbonManager, const State& start, const PlannerConfig& config) {
        uploadWorld(ribbonManager, start, config);
        m_Perm.clear();
        m_SampleXY.clear();
    }

    // instrumentation
    long trueCostEdges() const { return m_TrueCostEdges; }
    long dubinsSolves() const { return m_DubinsSolves; }
    long batches() const { return m_Batches; }

private:
    struct Candidate {
        State sample;          // destination state (speed = max speed)
        double path[7];        // qi[3], param[3], rho
        int type;
        double length;         // dubins_path_length
        double approxCost;     // Edge::approxCost() of the candidate edge
        bool coverageAllowed;
    };

    void uploadWorld(const RibbonManager& ribbonManager, const State& start, const PlannerConfig& config);
    void check(int rc, const char* what);

    ppe_ctx* m_Ctx;
    int m_KnnChunk;
    PpeWorldCache* m_Cache; // borrowed, may be null (then the map is uploaded for every plan)
    int m_Heuristic = PPE_H_MAX_DISTANCE;
    long m_TrueCostEdges = 0, m_DubinsSolves = 0, m_Batches = 0;

    // The reference keeps m_Samples itself heap-ordered (SamplingBasedPlanner.cpp:85-93).  Here the States stay where
    // addSamples put them; the arrangement the reference's vector would have is m_Samples[m_Perm[i]], and the heap
    // operations run on (m_Keys, m_Perm) -- see KeyedHeap.h.
    std::vector<uint32_t> m_Perm;
    std::vector<double> m_Keys;
    std::vector<double> m_SampleXY; // x, y of m_Samples in storage order (the distance pass streams 16 B per sample)
    std::vector<double> m_Dist;     // distance of every stored sample to the vertex being expanded

    // scratch reused across expansions
    std::vector<State> m_Scratch;
    std::vector<double> m_Q0, m_Q1, m_Rho, m_Param, m_Length;
    std::vector<int32_t> m_Type, m_Err;
    std::vector<ppe_edge> m_Edges;
    std::vector<ppe_edge_result> m_Results;
    std::vector<double> m_RibbonBuf;
};

#endif
