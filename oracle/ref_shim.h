// oracle/ref_shim.h -- TEST INFRASTRUCTURE: context of the compiled-reference shim (ref_shim.cpp),
// shared with the harness shim that runs the product's BatchedAStarPlanner on the same inputs.
#ifndef PPE_ORACLE_REF_SHIM_H
#define PPE_ORACLE_REF_SHIM_H

#include <chrono>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "planner/AStarPlanner.h"
#include "planner/utilities/RibbonManager.h"
#include "planner/utilities/StateGenerator.h"
#include "planner/search/Vertex.h"
#include <stdexcept>
#include "common/dynamic_obstacles/BinaryDynamicObstaclesManager.h"
#include "common/dynamic_obstacles/GaussianDynamicObstaclesManager.h"
#include "ppe.h"

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };

struct ref_ctx {
    NullBuf nullBuf;
    std::ostream nullStream;
    PlannerConfig config;
    ppe_config cfg;
    std::vector<RibbonManager> sets;
    std::vector<std::vector<double>> ribbonsAfter;
    std::shared_ptr<BinaryDynamicObstaclesManager> binary;
    std::shared_ptr<GaussianDynamicObstaclesManager> gaussian;
    std::string lastError;
    // virtual clock for ref_plan
    double clockNow = 0, clockTick = 0;
    long clockCalls = 0;
    ref_ctx() : nullStream(&nullBuf), config(&nullStream) {}
};


// Runs `planner.plan` with the ctx's world and (optionally) a virtual clock; fills the plan record
// (12 doubles per Dubins path: qi[3], param[3], rho, type, speed, start, end, 0) and
// stats10 = Samples, Generated, Expanded, Iterations, PlanFValue, PlanCollisionPenalty,
// PlanTimePenalty, PlanHValue, PlanDepth, now() calls.  Returns the number of paths or -1.
// `prev_plan` / `n_prev`: a previous plan in the same 12-double record layout (nullptr / 0 = none): what the
// Executive passes on every cycle after the first (executive.cpp:146,189), re-validated at AStarPlanner.cpp:46-59.
template <typename PlannerT>
int ref_run_plan(PlannerT& planner, ref_ctx* ctx, int ribbon_set, const double* start5, double timeRemaining,
                 double clock0, double tick, int initialSamples, int useBrownPaths, double* plan_out, int plan_cap,
                 double* stats10, const double* prev_plan = nullptr, int n_prev = 0, double sampleTick = 0) {
    PlannerConfig config = ctx->config;
    config.setInitialSamples(initialSamples);
    config.setUseBrownPaths(useBrownPaths != 0);
    ctx->clockNow = clock0; ctx->clockTick = tick; ctx->clockCalls = 0;
    if (tick > 0) {
        // parity mode: deterministic clock (PlannerConfig::setNowFunction, PlannerConfig.h:110);
        // the RNG seed (AStarPlanner.cpp:33) and every deadline test then depend on call counts only
        // sampleTick > 0 also charges virtual time per generated sample (planner.attemptedSamples()), which bounds the
        // sample doubling of AStarPlanner.cpp:101-102 the way generation time bounds it on a real clock
        PlannerT* pl = &planner;
        config.setNowFunction([ctx, pl, sampleTick]() -> double {
            double t = ctx->clockNow + (double)ctx->clockCalls * ctx->clockTick + sampleTick * (double)pl->attemptedSamples();
            ctx->clockCalls++;
            return t;
        });
    }
    else if (clock0 > 0) {
        // real clock rebased so that the plan starts at clock0: the sampler's seed is the integer second of the deadline
        // (AStarPlanner.cpp:33), so a chosen clock0 gives both planners the same sample sequence at a real budget
        const auto t_start = std::chrono::steady_clock::now();
        config.setNowFunction([ctx, t_start, clock0]() -> double {
            ctx->clockCalls++;
            return clock0 + std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        });
    }
    State start(start5[0], start5[1], start5[2], start5[3], start5[4]);
    Planner::Stats stats;
    DubinsPlan previous;
    for (int i = 0; i < n_prev; i++) {
        const double* o = prev_plan + 12 * i;
        DubinsPath p;
        p.qi[0] = o[0]; p.qi[1] = o[1]; p.qi[2] = o[2];
        p.param[0] = o[3]; p.param[1] = o[4]; p.param[2] = o[5];
        p.rho = o[6]; p.type = (DubinsPathType)(int)o[7];
        DubinsWrapper w;
        w.fill(p, o[8], o[9]);
        if (o[10] < w.getEndTime()) w.updateEndTime(o[10]);
        previous.append(w);
    }
    try {
        stats = planner.plan(ctx->sets[ribbon_set], start, config, previous, timeRemaining);
    } catch (std::exception& ex) {
        ctx->lastError = ex.what();
        return -1;
    }
    int n = 0;
    for (const auto& w : stats.Plan.get()) {
        if (n < plan_cap) {
            double* o = plan_out + 12 * n;
            const DubinsPath& p = w.unwrap();
            o[0] = p.qi[0]; o[1] = p.qi[1]; o[2] = p.qi[2];
            o[3] = p.param[0]; o[4] = p.param[1]; o[5] = p.param[2];
            o[6] = p.rho; o[7] = (double)p.type; o[8] = w.getSpeed();
            o[9] = w.getStartTime(); o[10] = w.getEndTime(); o[11] = 0;
        }
        n++;
    }
    stats10[0] = (double)stats.Samples; stats10[1] = (double)stats.Generated;
    stats10[2] = (double)stats.Expanded; stats10[3] = (double)stats.Iterations;
    stats10[4] = n ? stats.PlanFValue : -1; stats10[5] = stats.PlanCollisionPenalty;
    stats10[6] = n ? stats.PlanTimePenalty : -1; stats10[7] = n ? stats.PlanHValue : -1;
    stats10[8] = n ? (double)stats.PlanDepth : -1; stats10[9] = (double)ctx->clockCalls;
    return n;
}

// ExpandTest1Ribbons (test_planner.cpp:1061-1082): one expansion of a root vertex over `nSamples` samples drawn by
// StateGenerator(-50, 50, -50, 50, 2.5, 2.5, seed); the start state is the generator's first draw with time 1.
// Pops the open list dry: returns the number of vertices it held, f_out receives their f-values in pop order.
template <typename PlannerT>
int ref_run_expand_once(PlannerT& planner, ref_ctx* ctx, int ribbon_set, int nSamples, int seed, double* f_out, int cap) {
    try {
        StateGenerator generator(-50, 50, -50, 50, 2.5, 2.5, seed);
        State start = generator.generate();
        start.time() = 1;
        PlannerConfig config = ctx->config;
        config.setStartStateTime(1);
        auto root = Vertex::makeRoot(start, ctx->sets[ribbon_set]);
        root->computeApproxToGo(config);
        planner.setConfig(config);
        planner.addSamples(generator, nSamples);
        planner.expand(root, config.obstaclesManager());
        int n = 0;
        for (;;) {
            Vertex::SharedPtr v;
            try { v = planner.popVertexQueue(); } catch (std::out_of_range&) { break; }
            if (n < cap) f_out[n] = v->f();
            n++;
        }
        return n;
    } catch (std::exception& ex) {
        ctx->lastError = ex.what();
        return -1;
    }
}

#endif
