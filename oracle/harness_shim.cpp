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

// stats13 = the 10 values of ref_plan + true-cost edges, Dubins solves, engine batches
int harness_plan(ref_ctx* ctx, int device, int ribbon_set, const double* start5, double timeRemaining, double clock0,
                 double tick, int initialSamples, int useBrownPaths, int knnChunk, double* plan_out, int plan_cap,
                 double* stats13) {
    EngineSlot& slot = engineFor(ctx, device);
    if (!slot.engine) { ctx->lastError = "ppe_create failed: no CUDA device (the engine has no CPU path)"; return -2; }
    ppe_ctx* engine = slot.engine;
    BatchedAStarPlanner planner(engine, knnChunk > 0 ? knnChunk : 128, &slot.cache);
    int n = ref_run_plan(planner, ctx, ribbon_set, start5, timeRemaining, clock0, tick, initialSamples, useBrownPaths,
                         plan_out, plan_cap, stats13);
    stats13[10] = (double)planner.trueCostEdges();
    stats13[11] = (double)planner.dubinsSolves();
    stats13[12] = (double)planner.batches();
    return n;
}

// the same single expansion through the product's adapter (its expand() needs the world on the engine: plan() would
// upload it, so this entry point does)
int harness_expand_once(ref_ctx* ctx, int device, int ribbon_set, int nSamples, int seed, double* f_out, int cap) {
    EngineSlot& slot = engineFor(ctx, device);
    if (!slot.engine) { ctx->lastError = "ppe_create failed"; return -2; }
    ppe_ctx* engine = slot.engine;
    BatchedAStarPlanner planner(engine, 128, &slot.cache);
    try {
        PlannerConfig config = ctx->config;
        config.setStartStateTime(1);
        planner.prepareWorld(ctx->sets[ribbon_set], State(0, 0, 0, 2.5, 1), config);
    } catch (std::exception& ex) { ctx->lastError = ex.what(); return -1; }
    return ref_run_expand_once(planner, ctx, ribbon_set, nSamples, seed, f_out, cap);
}

} // extern "C"
