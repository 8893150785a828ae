// oracle/harness_shim.cpp -- TEST INFRASTRUCTURE: runs the product's BatchedAStarPlanner
// (path_planner_b200/harness/) on exactly the inputs ref_plan gives the reference's AStarPlanner,
// so tests can compare the final plans (plan identity in virtual-clock mode) and bench.py can
// report plan cost at a wall-clock budget.  Links the compiled reference objects + libppe.so.
#include "ref_shim.h"

#include "BatchedAStarPlanner.h"

#include <map>
#include <mutex>

// one engine context (+ its world cache) per (test context, device): nothing process-global is shared between
// planning threads
struct EngineSlot { ppe_ctx* engine = nullptr; PpeWorldCache cache; };
static EngineSlot& engineFor(ref_ctx* ctx, int device) {
    static std::mutex mu;
    static std::map<std::pair<ref_ctx*, int>, EngineSlot> slots;
    std::lock_guard<std::mutex> lock(mu);
    EngineSlot& s = slots[std::make_pair(ctx, device)];
    if (!s.engine && ppe_create(device, &s.engine) != PPE_OK) s.engine = nullptr;
    return s;
}

extern "C" {

// stats16 = the 10 values of ref_plan + true-cost edges, Dubins solves, engine batches, frontier vertices,
// frontier hits, exact (host-replayed) expansions.  frontier < 0: the adapter's default width.
int harness_plan2(ref_ctx* ctx, int device, int ribbon_set, const double* start5, double timeRemaining, double clock0,
                  double tick, int initialSamples, int useBrownPaths, int knnChunk, int frontier, const double* prev_plan,
                  int n_prev, double* plan_out, int plan_cap, double* stats16) {
    EngineSlot& slot = engineFor(ctx, device);
    if (!slot.engine) { ctx->lastError = "ppe_create failed: no CUDA device (the engine has no CPU path)"; return -2; }
    BatchedAStarPlanner planner(slot.engine, knnChunk > 0 ? knnChunk : 128, &slot.cache);
    if (frontier >= 0) planner.setFrontierWidth(frontier);
    int n = ref_run_plan(planner, ctx, ribbon_set, start5, timeRemaining, clock0, tick, initialSamples, useBrownPaths,
                         plan_out, plan_cap, stats16, prev_plan, n_prev);
    stats16[10] = (double)planner.trueCostEdges();
    stats16[11] = (double)planner.dubinsSolves();
    stats16[12] = (double)planner.batches();
    stats16[13] = (double)planner.frontierVertices();
    stats16[14] = (double)planner.frontierHits();
    stats16[15] = (double)planner.exactExpansions();
    return n;
}

int harness_plan(ref_ctx* ctx, int device, int ribbon_set, const double* start5, double timeRemaining, double clock0,
                 double tick, int initialSamples, int useBrownPaths, int knnChunk, double* plan_out, int plan_cap,
                 double* stats13) {
    double stats16[16];
    int n = harness_plan2(ctx, device, ribbon_set, start5, timeRemaining, clock0, tick, initialSamples, useBrownPaths, knnChunk,
                          -1, nullptr, 0, plan_out, plan_cap, stats16);
    for (int i = 0; i < 13; i++) stats13[i] = stats16[i];
    return n;
}

// the same single expansion through the product's adapter (its expand() needs the world on the engine: plan() would
// upload it, so this entry point does)
int harness_expand_once2(ref_ctx* ctx, int device, int ribbon_set, int nSamples, int seed, int frontier, double* f_out, int cap) {
    EngineSlot& slot = engineFor(ctx, device);
    if (!slot.engine) { ctx->lastError = "ppe_create failed"; return -2; }
    ppe_ctx* engine = slot.engine;
    BatchedAStarPlanner planner(engine, 128, &slot.cache);
    if (frontier >= 0) planner.setFrontierWidth(frontier);
    try {
        PlannerConfig config = ctx->config;
        config.setStartStateTime(1);
        planner.prepareWorld(ctx->sets[ribbon_set], State(0, 0, 0, 2.5, 1), config);
    } catch (std::exception& ex) { ctx->lastError = ex.what(); return -1; }
    return ref_run_expand_once(planner, ctx, ribbon_set, nSamples, seed, f_out, cap);
}

int harness_expand_once(ref_ctx* ctx, int device, int ribbon_set, int nSamples, int seed, double* f_out, int cap) {
    return harness_expand_once2(ctx, device, ribbon_set, nSamples, seed, -1, f_out, cap);
}

} // extern "C"
