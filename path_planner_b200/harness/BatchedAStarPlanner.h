// BatchedAStarPlanner -- the reference's anytime A* planner with every call site of the edge
// evaluation path routed through the B200 edge engine (include/ppe.h).
//
// Drop-in: it IS an AStarPlanner (path_planner/src/planner/AStarPlanner.h).  Overridden virtuals:
//   SamplingBasedPlanner::expand (SamplingBasedPlanner.h:43, SamplingBasedPlanner.cpp:52-151)
//       -- whole FRONTIER batches: the vertex being expanded plus the best vertices of the open list go
//          to the device in one ppe_expand_batch call (k-nearest selection over the resident sample set,
//          Dubins solves, true cost); the children are cached per vertex and handed to the open list in
//          the reference's push order when the reference's own aStar loop asks for them.
//   Planner::plan (Planner.h:50, AStarPlanner.cpp:12-132)
//       -- the same control flow and the same now() call sequence, with the previous-plan re-validation
//          (AStarPlanner.cpp:46-59), the Brown-path expansion (:150-162) and addSamples
//          (SamplingBasedPlanner.cpp:157-164) going through the engine as well.
// Planner::Stats, tracePlan, pushVertexQueue, popVertexQueue, goalCondition, aStar and the open list
// are the reference's own code, untouched.
//
// Compiles against the reference headers where they lie (-I<ref>/path_planner/src
// -I<ref>/path_planner_common/include) and any dubins.h; links libppe.so.
#ifndef PPE_BATCHED_ASTAR_PLANNER_H
#define PPE_BATCHED_ASTAR_PLANNER_H

#include <cstdint>
#include <memory>
#include <unordered_map>
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
    // `ctx` is borrowed (one ppe_ctx per planning thread).  `knnChunk`: how many nearest samples get their Dubins
    // paths solved per K1 launch on the exact host-replay path.  `cache`: see PpeWorldCache (may be null).
    explicit BatchedAStarPlanner(ppe_ctx* ctx, int knnChunk = 128, PpeWorldCache* cache = nullptr);
    ~BatchedAStarPlanner() override = default;

    // Vertices per ppe_expand_batch call.  1 = one vertex per call (no speculation); 0 = the exact host-replay
    // expansion for every vertex (round-1 behaviour: host k-nearest heaps, K1 per chunk, K2 per vertex).
    void setFrontierWidth(int m) { m_Frontier = m; }

    Stats plan(const RibbonManager& ribbonManager, const State& start, PlannerConfig config,
               const DubinsPlan& previousPlan, double timeRemaining) override;

    void expand(const std::shared_ptr<Vertex>& sourceVertex, const DynamicObstaclesManager& obstacles) override;

    // What plan() does before it starts searching: config, map, obstacles to the engine, empty sample set.  Public for
    // callers that drive addSamples() / expand() themselves (the reference's ExpandTest1Ribbons does); such callers
    // add their samples with addSamplesResident.
    void prepareWorld(const RibbonManager& ribbonManager, const State& start, const PlannerConfig& config);
    // SamplingBasedPlanner::addSamples (SamplingBasedPlanner.cpp:157-164) with the blocked test and the resident
    // copy of the sample set on the device
    void addSamplesResident(StateGenerator& generator, int n);

    // states drawn from the generator so far (SamplingBasedPlanner::m_AttemptedSamples)
    unsigned long attemptedSamples() const { return m_AttemptedSamples; }

    // instrumentation
    long trueCostEdges() const { return m_TrueCostEdges; }
    long dubinsSolves() const { return m_DubinsSolves; }
    long batches() const { return m_Batches; }
    long frontierVertices() const { return m_FrontierVertices; }  // vertices sent to ppe_expand_batch
    long frontierHits() const { return m_FrontierHits; }          // expansions served from a cached frontier result
    long exactExpansions() const { return m_ExactExpansions; }    // expansions replayed on the host (ties, width 0)
    long exactForTies() const { return m_ExactTies; }             // ... because the device flagged an exact distance tie
    long exactForOverflow() const { return m_ExactOverflow; }     // ... because the device ran out of candidate capacity
    // wall seconds spent in: ppe_expand_batch calls (incl. building the vertex records), handing cached children to the open
    // list, addSamplesResident, exact host expansions
    double secondsInEngineExpand() const { return m_TEngine; }
    double secondsInReplay() const { return m_TReplay; }
    double secondsInAddSamples() const { return m_TSamples; }
    double secondsInExact() const { return m_TExact; }

private:
    struct Candidate {
        State sample;          // destination state (speed = max speed)
        double path[7];        // qi[3], param[3], rho
        int type;
        double length;         // dubins_path_length
        double approxCost;     // Edge::approxCost() of the candidate edge
        bool coverageAllowed;
    };
    // children of one vertex as ppe_expand_batch returned them, waiting for the reference's loop to pop that vertex
    struct Expansion {
        std::shared_ptr<Vertex> vertex; // keeps the address (the key) from being reused while the record exists
        std::vector<ppe_child> children;
        std::vector<double> ribbons;    // ribbons-after of the changed children, 4 doubles each, child order
        std::vector<int> ribbonStart;   // per child: first ribbon in `ribbons` (-1: unchanged)
        int flags = 0;
        int popped = 0;
    };
    struct ExpansionLog { double x, y; uint32_t pops; uint32_t nSamples; };

    void uploadWorld(const RibbonManager& ribbonManager, const State& start, const PlannerConfig& config);
    void check(int rc, const char* what);
    void expandExact(const std::shared_ptr<Vertex>& sourceVertex);
    void expandFrontier(const std::shared_ptr<Vertex>& sourceVertex);
    void speculate(const std::shared_ptr<Vertex>& first);
    void syncSampleHeap();
    int32_t internRibbons(const RibbonManager& ribbons);
    // result of one engine-evaluated edge -> the reference's Edge / Vertex / RibbonManager members
    void fillChild(const std::shared_ptr<Vertex>& v, const double qi[3], const double param[3], double rho, int type, double wSpeed,
                   double wStart, double wEnd, bool infeasible, double approx, double trueCost, double penalty, double g, double h,
                   double cct, bool ribbonsChanged, const double* ribbons, int nRibbons);
    Vertex::SharedPtr revalidatePreviousPlan(const Vertex::SharedPtr& startV, const DubinsPlan& previousPlan, bool visualize);
    void expandSpecific(const Vertex::SharedPtr& root, const std::vector<State>& samples, bool coverageAllowed);
    void dumpTrajectory(const std::shared_ptr<Vertex>& v);

    ppe_ctx* m_Ctx;
    int m_KnnChunk;
    PpeWorldCache* m_Cache; // borrowed, may be null (then the map is uploaded for every plan)
    int m_Frontier = 64;
    int m_Heuristic = PPE_H_MAX_DISTANCE;
    bool m_HOnDevice = true;
    long m_TrueCostEdges = 0, m_DubinsSolves = 0, m_Batches = 0, m_FrontierVertices = 0, m_FrontierHits = 0, m_ExactExpansions = 0, m_ExactTies = 0,
         m_ExactOverflow = 0;
    double m_TEngine = 0, m_TReplay = 0, m_TSamples = 0, m_TExact = 0;

    // The reference keeps m_Samples itself heap-ordered (SamplingBasedPlanner.cpp:85-93).  Here the States stay where
    // addSamples put them; the arrangement the reference's vector would have is m_Samples[m_Perm[i]], and the heap
    // operations run on (m_Keys, m_Perm) -- see KeyedHeap.h.  Expansions served by the device do not touch the
    // arrangement; they are logged (source position, pop count) and replayed into it only if an exact host expansion
    // is ever needed (two samples at exactly equal distance: then, and only then, the arrangement decides).
    std::vector<uint32_t> m_Perm;
    std::vector<double> m_Keys;
    std::vector<double> m_SampleXY; // x, y of m_Samples in storage order (the distance pass streams 16 B per sample)
    std::vector<double> m_Dist;     // distance of every stored sample to the vertex being expanded
    std::vector<ExpansionLog> m_Log;
    size_t m_LogApplied = 0;

    std::unordered_map<const Vertex*, Expansion> m_Expansions;

    // scratch reused across expansions
    std::vector<State> m_Scratch;
    std::vector<double> m_Q0, m_Q1, m_Rho, m_Param, m_Length;
    std::vector<int32_t> m_Type, m_Err;
    std::vector<ppe_edge> m_Edges;
    std::vector<ppe_edge_result> m_Results;
    std::vector<double> m_RibbonBuf;
    std::vector<ppe_vertex> m_Verts;
    std::vector<ppe_child> m_Children;
    std::vector<int32_t> m_NChildren, m_Flags, m_Popped;
    std::vector<double> m_GenX, m_GenY, m_GenH;
    std::vector<uint8_t> m_Keep;
};

#endif
