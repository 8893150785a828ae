/*
 * ppe_harness.h -- C ABI of the standalone, ROS-free planning harness (path_planner_b200/libppe_harness.so):
 * the reference's Planner / Edge / Vertex / State classes (compiled from the reference's own sources, linked as
 * path_planner_common + planner) driven by the product's BatchedAStarPlanner over the B200 edge engine (ppe.h).
 * It stands where the reference's Executive::planLoop stands (path_planner/src/executive/executive.cpp:79-279):
 * world state in, one Planner::plan call per cycle, plan + Planner::Stats out.
 *
 * Plain C, int status returns (0 = ok, negative = error, pph_last_error describes it).  One pph_ctx per planning
 * thread; it owns its ppe_ctx (one GPU) and the "which map is on the device" cache.
 */
#ifndef PPE_HARNESS_H
#define PPE_HARNESS_H

#include <stdint.h>

#include "ppe.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pph_ctx pph_ctx;

int pph_create(int device, pph_ctx** out);
void pph_destroy(pph_ctx* ctx);
const char* pph_last_error(const pph_ctx* ctx);

/* PlannerConfig scalars + Ribbon::RibbonWidth (executive.cpp:394-422 fan-out); cfg->heuristic = RibbonManager heuristic */
int pph_set_config(pph_ctx* ctx, const ppe_config* cfg);
/* Map base class: never blocked (Map.cpp:4-6) */
int pph_set_map_none(pph_ctx* ctx);
/* occupancy bits (layout of ppe_set_map_bitmap) served through a Map subclass with GridWorldMap::isBlocked semantics */
int pph_set_map_bitmap(pph_ctx* ctx, const uint8_t* bits, int rows, int cols, int row_stride_bytes, double resolution);
/* GridWorldMap text file (GridWorldMap.cpp:10-82), parsed by the reference's own loader */
int pph_load_gridworld_map(pph_ctx* ctx, const char* path);
int pph_set_obstacles_none(pph_ctx* ctx);
/* Binary / GaussianDynamicObstaclesManager::update(mmsi = i + 1, x, y, heading, speed, time, ...) in this order */
int pph_set_obstacles_binary(pph_ctx* ctx, int n, const double* x, const double* y, const double* heading, const double* speed,
                             const double* time, const double* width, const double* length);
int pph_set_obstacles_gaussian(pph_ctx* ctx, int n, const double* x, const double* y, const double* heading, const double* speed,
                               const double* time, const double* cov /* n x 4 or NULL = manager default */);
/* RibbonManager(heuristic from the config).add(x1, y1, x2, y2) per ribbon (executive.cpp:371-392) */
int pph_set_ribbons(pph_ctx* ctx, int n, const double* xyxy);

typedef struct {
    double time_remaining;    /* Planner::plan's budget, seconds                                          */
    double clock0, tick;      /* tick > 0: virtual clock now() = clock0 + calls * tick + samples drawn * sample_tick;
                                 tick == 0 and clock0 > 0: the real clock rebased so that the plan starts at clock0 (the
                                 sampler's seed is the integer second of the deadline, AStarPlanner.cpp:33: a chosen clock0
                                 makes runs comparable); both 0: the system clock */
    double sample_tick;       /* virtual seconds per generated sample (bounds the anytime loop's sample doubling)  */
    int32_t initial_samples;  /* PlannerConfig::initialSamples()                                          */
    int32_t use_brown_paths;  /* PlannerConfig::useBrownPaths()                                           */
    int32_t frontier;         /* vertices per ppe_expand_batch; < 0 default, 0 = exact host replay        */
    int32_t knn_chunk;        /* K1 chunk of the exact path; <= 0 default                                 */
    int32_t visualize;        /* 1: write the reference's visualization text stream to visualization_path */
    int32_t reserved;
    const char* visualization_path;
} pph_plan_options;

/* One record per DubinsWrapper of a plan, the fields of path_planner_common/msg/DubinsPath.msg in message order
 * (NodeBase.h:201-220) + the wrapper's end time (Plan.msg carries only the last one as `endtime`). */
typedef struct {
    double initial_x, initial_y, initial_yaw;
    double length0, length1, length2;
    double rho;
    int32_t type;
    int32_t pad;
    double speed;
    double start_time;
    double end_time;
} pph_dubins_path;

/* Planner::Stats (Planner.h:24-35) + engine counters */
typedef struct {
    uint64_t samples, generated, expanded, iterations, plan_depth;
    double plan_f, plan_collision_penalty, plan_time_penalty, plan_h;
    double plan_endtime;          /* Plan.msg endtime */
    uint64_t now_calls;
    uint64_t true_cost_edges, dubins_solves, engine_batches, frontier_vertices, frontier_hits, exact_expansions;
    double wall_seconds;
    double seconds_engine_expand, seconds_replay, seconds_add_samples, seconds_exact; /* where the wall time went */
    uint64_t exact_for_ties, exact_for_overflow; /* why expansions were replayed on the host */
} pph_stats;

/* Planner::plan(ribbons, start, config, previousPlan, timeRemaining): start = x, y, heading, speed, time.
 * Returns the number of paths in the plan (<= cap written) or a negative status. */
int pph_plan(pph_ctx* ctx, const double start[5], const pph_dubins_path* previous, int n_previous, const pph_plan_options* opt,
             pph_dubins_path* plan_out, int cap, pph_stats* stats);

/* RibbonManager::coverBetween + DubinsPlan::sample of the last plan: the Executive's bookkeeping between two cycles
 * (executive.cpp:146,188): covers the ribbons along the last plan up to `time` and returns the state there. */
int pph_advance(pph_ctx* ctx, double time, double state_out[5]);

/* Plan.msg as the text `rostopic echo` prints (paths: - initial_x: ... endtime: ...) */
int pph_write_plan_msg(const pph_dubins_path* plan, int n, const char* path);

#ifdef __cplusplus
}
#endif

#endif
