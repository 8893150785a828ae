/*
 * ppe.h -- C ABI of the B200 batched Dubins edge-evaluation engine ("ppe" = path-planner edges).
 *
 * Drop-in boundary for the hot path of afb2001/path_planner: the batched equivalent of
 *   Edge::computeApproxCost  (path_planner/src/planner/search/Edge.cpp:11-20,64-66)
 *   Edge::computeTrueCost    (path_planner/src/planner/search/Edge.cpp:68-206)
 * as called from SamplingBasedPlanner::expand (path_planner/src/planner/SamplingBasedPlanner.cpp:76,119,145)
 * and AStarPlanner::plan (path_planner/src/planner/AStarPlanner.cpp:52,157).
 *
 * The reference has no FFI layer; its seams are C++ virtuals.  A maintainer binds this library
 * from a subclass that overrides `virtual SamplingBasedPlanner::expand`
 * (SamplingBasedPlanner.h:43) -- see INTEGRATION.md and path_planner_b200/harness/.
 *
 * Conventions: plain C, no exceptions cross the boundary.  Every call returns PPE_OK (0) or a
 * negative ppe_status; ppe_last_error(ctx) describes the last failure.  The caller owns all host
 * buffers; the engine owns device memory.  One ppe_ctx per planning thread and per GPU (one
 * process per GPU); a ctx is not thread-safe.  There is NO CPU fallback: without a CUDA device
 * ppe_create fails with PPE_ERR_NO_DEVICE.
 *
 * All angles are radians.  "heading" is east of north (State.h:9-12), "yaw" is counter-clockwise
 * from +x (DubinsPath.msg:6); yaw = pi/2 - heading wrapped to [0, 2pi) (State.h:51-55).
 */
#ifndef PPE_H
#define PPE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPE_ABI_VERSION 3

typedef struct ppe_ctx ppe_ctx;

typedef enum {
    PPE_OK = 0,
    PPE_ERR_NO_DEVICE = -1,   /* no CUDA device / wrong architecture: there is no CPU path */
    PPE_ERR_CUDA = -2,        /* a CUDA runtime call failed; see ppe_last_error */
    PPE_ERR_INVALID = -3,     /* bad argument */
    PPE_ERR_STATE = -4,       /* call order (e.g. batch before ppe_set_config) */
    PPE_ERR_CAPACITY = -5     /* a fixed device capacity was exceeded (ribbon pool ...) */
} ppe_status;

/* Dubins words, same numbering as DubinsPath.msg:17 / dubins.h DubinsPathType */
enum { PPE_LSL = 0, PPE_LSR = 1, PPE_RSL = 2, PPE_RSR = 3, PPE_RLR = 4, PPE_LRL = 5 };

/* error codes of the dubins library (dubins.h), reported per solve in ppe_dubins_batch */
enum { PPE_EDUBOK = 0, PPE_EDUBCOCONFIGS = 1, PPE_EDUBPARAM = 2, PPE_EDUBBADRHO = 3, PPE_EDUBNOPATH = 4 };

/* RibbonManager::Heuristic (RibbonManager.h:19-25).  MaxDistance and the two point-robot TSP
 * variants (RibbonManager.cpp:53-95; lists of up to 8 ribbons) are evaluated on the device; for the
 * Dubins TSP variants (and longer lists) the engine returns h = -1 and the host adapter calls the
 * reference's Vertex::computeApproxToGo on the returned ribbon set. */
enum {
    PPE_H_MAX_DISTANCE = 0,
    PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL = 1,
    PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K = 2,
    PPE_H_TSP_DUBINS_NO_SPLIT_ALL = 3,
    PPE_H_TSP_DUBINS_NO_SPLIT_K = 4
};

/* PlannerConfig scalars (PlannerConfig.h:179-207) + Ribbon::RibbonWidth (Ribbon.cpp:4) +
 * Edge penalty factors (Edge.h:151-152). */
typedef struct {
    double max_speed;                    /* PlannerConfig::maxSpeed()                 default 2.5 */
    double slow_speed;                   /* PlannerConfig::slowSpeed()                default 0.5 */
    double turning_radius;               /* PlannerConfig::turningRadius()            default 8   */
    double coverage_turning_radius;      /* PlannerConfig::coverageTurningRadius()    default 16  */
    double time_horizon;                 /* PlannerConfig::timeHorizon()              default 30  */
    double time_minimum;                 /* PlannerConfig::timeMinimum()              default 5   */
    double collision_checking_increment; /* PlannerConfig::collisionCheckingIncrement default 0.05 */
    double start_state_time;             /* PlannerConfig::startStateTime()                        */
    double ribbon_width;                 /* Ribbon::RibbonWidth                       default 1.5 */
    double collision_penalty_factor;     /* Edge::collisionPenaltyFactor()            600         */
    double time_penalty_factor;          /* Edge::timePenaltyFactor()                 1           */
    int32_t heuristic;                   /* PPE_H_*                                               */
    int32_t branching_factor;            /* PlannerConfig::branchingFactor()          default 9   */
    int32_t tsp_k;                       /* RibbonManager::m_K of the K-ribbon TSP heuristics (executive.cpp:391: 2) */
    int32_t reserved0;
} ppe_config;

/* One edge = the inputs Edge::computeTrueCost reads (Edge.cpp:68-96). */
typedef struct {
    double src[5];       /* start()->state(): x, y, heading, speed, time                      */
    double src_g;        /* start()->currentCost()                                            */
    double dst[4];       /* end()->state() as connected: x, y, heading, speed (the edge speed) */
    double path_qi[3];   /* DubinsWrapper::unwrap() -- read when has_path != 0                */
    double path_param[3];
    double path_rho;
    double w_speed;      /* DubinsWrapper::getSpeed()      (has_path)                         */
    double w_start_time; /* DubinsWrapper m_StartTime      (has_path)                         */
    double w_end_time;   /* DubinsWrapper::getEndTime()    (has_path; may be truncated)       */
    int32_t path_type;   /* PPE_LSL.. (has_path)                                              */
    int32_t has_path;    /* 0: solve src->dst at the edge's radius first (Edge.cpp:78-80); < 0: skip */
    int32_t coverage_allowed; /* end()->coverageAllowed()                                     */
    int32_t ribbon_set;  /* id from ppe_put_ribbon_set: start()->ribbonManager()              */
} ppe_edge;

/* per-edge status: conditions under which the reference throws out of computeTrueCost */
enum {
    PPE_EDGE_OK = 0,
    PPE_EDGE_ERR_END_SAMPLE = 1,     /* DubinsWrapper::sample(end state) would throw (DubinsWrapper.cpp:30-35) */
    PPE_EDGE_ERR_NO_PATH = 2,        /* dubins_shortest_path failed (unset wrapper, Edge.cpp:85)                */
    PPE_EDGE_ERR_RIBBON_CAPACITY = 3,/* ribbon set outgrew the per-edge device capacity                         */
    PPE_EDGE_SKIPPED = 4             /* has_path < 0: an empty slot of a frontier batch, not evaluated           */
};

/* What Edge::computeTrueCost writes into Edge/Vertex members (Edge.cpp:177-203). */
typedef struct {
    double true_cost;           /* Edge::m_TrueCost                                   */
    double collision_penalty;   /* Edge::m_CollisionPenalty                           */
    double approx_cost;         /* Edge::m_ApproxCost                                 */
    double end[5];              /* end()->state() after truncation: x,y,heading,speed,time */
    double g;                   /* Vertex::m_CurrentCost (Vertex.cpp:102-104)         */
    double h;                   /* Vertex::m_ApproxToGo  (Vertex.cpp:49-64); -1 if not on device */
    double coverage_completed_time; /* end()->ribbonManager().coverageCompletedTime() */
    double path_qi[3];          /* the wrapper's path after the call                  */
    double path_param[3];
    double path_rho;
    double w_speed;             /* wrapper speed after the call                       */
    double w_start_time;
    double w_end_time;          /* after updateEndTime (Edge.cpp:179)                 */
    int64_t ribbons_offset;     /* first ribbon of this edge's ribbons-after in the pool; -1 if unchanged */
    int32_t path_type;
    int32_t infeasible;         /* Edge::m_Infeasible                                 */
    int32_t status;             /* PPE_EDGE_*                                         */
    int32_t n_samples;          /* executed iterations of the while loop (Edge.cpp:125) */
    int32_t n_checkpoints;      /* executions of the ribbon branch (Edge.cpp:155-171) */
    int32_t n_ribbons_after;    /* size of end()->ribbonManager().get() after the call */
    int32_t ribbons_changed;    /* 0: identical to the parent's set                   */
    int32_t reserved;           /* engine instrumentation: bits 0-23 executed samples proved clean by the culling probe
                                   (never evaluated one by one), bit 24 set when a K2t thread walked the edge */
} ppe_edge_result;

/* ---- lifetime ---------------------------------------------------------------------------- */
int ppe_abi_version(void);
int ppe_create(int device, ppe_ctx** out);
void ppe_destroy(ppe_ctx* ctx);
const char* ppe_last_error(const ppe_ctx* ctx);

/* ---- world state (replicated per GPU, uploaded once per plan) ----------------------------- */
int ppe_set_config(ppe_ctx* ctx, const ppe_config* cfg);

/* Map (Map.cpp:4-6): never blocked. */
int ppe_set_map_none(ppe_ctx* ctx);
/* GridWorldMap (GridWorldMap.cpp:84-93): occupancy bits, row 0 = y 0, bit (r, c) at
 * bits[r * row_stride_bytes + c/8] >> (c%8) & 1; out of bounds = blocked. */
int ppe_set_map_bitmap(ppe_ctx* ctx, const uint8_t* bits, int rows, int cols, int row_stride_bytes,
                       double resolution);

/* DynamicObstaclesManager base (DynamicObstaclesManager.h:23): always 0. */
int ppe_set_obstacles_none(ppe_ctx* ctx);
/* BinaryDynamicObstaclesManager::get() in container iteration order (Binary...h:14-25). */
int ppe_set_obstacles_binary(ppe_ctx* ctx, int n, const double* x, const double* y, const double* yaw,
                             const double* speed, const double* time, const double* width,
                             const double* length);
/* GaussianDynamicObstaclesManager::get() in container iteration order (Gaussian...h:21-44);
 * cov is n x 4 row-major 2x2 covariances. */
int ppe_set_obstacles_gaussian(ppe_ctx* ctx, int n, const double* x, const double* y, const double* yaw,
                               const double* speed, const double* time, const double* cov);

/* RibbonManager state of a parent vertex (RibbonManager::get() in list order, 4 doubles per
 * ribbon: startX, startY, endX, endY) + coverageCompletedTime.  The list is taken verbatim, as a
 * child vertex copies its parent's (Vertex.cpp:24,32): the covered-filter of RibbonManager::add
 * (RibbonManager.cpp:154-158) is the caller's.  Sets are interned: every edge leaving that vertex
 * refers to the returned id. */
int ppe_put_ribbon_set(ppe_ctx* ctx, int n, const double* xyxy, double coverage_completed_time,
                       int32_t* set_id);
int ppe_clear_ribbon_sets(ppe_ctx* ctx);

/* ---- K1: batched Dubins solve = Edge::computeApproxCost / DubinsWrapper::set ------------- */
/* q0, q1: n x 3 (x, y, yaw); rho: n.  Outputs: type n, param n x 3, length n (= dubins_path_length),
 * err n (PPE_EDUB*). */
int ppe_dubins_batch(ppe_ctx* ctx, int64_t n, const double* q0, const double* q1, const double* rho,
                     int32_t* type, double* param, double* length, int32_t* err);

/* ---- K2: batched true cost = Edge::computeTrueCost ---------------------------------------- */
/* Host buffers; H2D / D2H copies are part of the call.  Batches of 2^16 edges and more are pipelined: the edges travel in
 * slices of an eighth of the batch (16 Ki .. 128 Ki edges; PPE_LATE_SLICE fixes the size), K2a + K2t of a slice run while the next slice arrives and the previous slice's
 * records leave, K2b runs once over the heavy list of the whole batch and its records are scattered into `results` by a
 * kernel when `results` is pinned, mapped memory (cudaHostAlloc / cudaHostRegister) or by the host when it is pageable.
 * Pinned buffers also make the slice copies asynchronous.  Smaller batches: one H2D, one launch group, one D2H.
 * Environment knobs read at ppe_create (tuning / testing only): PPE_THREAD_WALKER=0 evaluates every edge with the warp
 * walker K2b instead of K2t + K2b; PPE_K2T_DIRTY=<n> non-clean chunks of one edge a K2t warp evaluates before handing
 * the edge to K2b (default 64 = all); PPE_K2T_CPS=<n> ribbon check-points a K2t thread walks (default 6);
 * PPE_K2B_CTAS=<n> K2b CTAs per SM; PPE_LATE_K2B=0 slices the whole kernel sequence instead (the round-1 pipeline). */
int ppe_true_cost_batch(ppe_ctx* ctx, int64_t n, const ppe_edge* edges, ppe_edge_result* results);
/* Ribbons-after of edge `edge_index` of the last batch (4 doubles per ribbon, list order).
 * Returns the number of ribbons (<= cap written) or a negative status. */
int ppe_get_ribbons_after(ppe_ctx* ctx, int64_t edge_index, double* xyxy, int cap);

/* ---- K3: best feasible f = g + h of the last batch (prune / gather record) ---------------- */
/* *f = +inf and *edge_index = -1 when no feasible edge exists. */
int ppe_best(ppe_ctx* ctx, double* f, int64_t* edge_index);

/* ---- frontier expansion: SamplingBasedPlanner::expand for MANY vertices in one launch group ----
 * (SamplingBasedPlanner.cpp:52-151).  The sample set is resident on the device; per vertex the
 * engine orders the samples by Euclidean distance (:85-94), replays the per-radius k-best heaps over
 * Dubins lengths solved on the fly (:95-133, std::push_heap / std::pop_heap arrangement included),
 * emits the <= 4 nearest-endpoint edges (:65-81) and the winners x speeds (:134-149) in the
 * reference's push order and evaluates their true cost (K2) -- no host round trip in between. */

/* m_Samples.clear() (AStarPlanner.cpp:24) */
int ppe_clear_samples(ppe_ctx* ctx);
/* SamplingBasedPlanner::addSamples (SamplingBasedPlanner.cpp:157-164): n generated states in generation
 * order; the device evaluates Map::isBlocked for each, appends the free ones (order kept) to the resident
 * sample set and writes keep[i] = 1 / 0 so the caller's m_Samples mirrors it.  Returns the number kept. */
int64_t ppe_add_samples(ppe_ctx* ctx, int64_t n, const double* x, const double* y, const double* heading, uint8_t* keep);
int64_t ppe_sample_count(const ppe_ctx* ctx);

/* A vertex of the open list about to be expanded. */
typedef struct {
    double state[5];      /* Vertex::state(): x, y, heading, speed, time                          */
    double g;             /* Vertex::currentCost()                                                */
    double endpoint[3];   /* Vertex::getNearestPointAsState(): x, y, heading (has_endpoint != 0)  */
    int32_t ribbon_set;   /* id from ppe_put_ribbon_set: Vertex::ribbonManager()                  */
    int32_t has_endpoint; /* !done() && distanceTo(endpoint) > increment (:65-68)                 */
} ppe_vertex;

/* One child edge of an expanded vertex: what computeTrueCost + the k-nearest selection produced.  The
 * wrapper is (qi = vertex pose as yaw, path_param, rho by coverage_allowed, path_type, speed = end[3],
 * start time = vertex time, end time = w_end_time). */
typedef struct {
    double true_cost, collision_penalty, approx_cost;
    double end[5];                  /* end()->state() after truncation                          */
    double g, h;                    /* h = -1 when the heuristic is not evaluated on the device */
    double coverage_completed_time;
    double path_param[3];
    double w_end_time;
    int64_t ribbons_offset;         /* into the batch's ribbons-after pool; -1 if unchanged     */
    int32_t sample_index;           /* resident sample the edge leads to; -1 = endpoint edge    */
    int32_t path_type;
    int32_t infeasible;
    int32_t status;                 /* PPE_EDGE_*                                               */
    int32_t coverage_allowed;
    int32_t n_ribbons_after;
    int32_t ribbons_changed;
    int32_t reserved;
} ppe_child;

enum {
    PPE_EXPAND_TIE = 1,      /* two samples at exactly equal distance inside the consumed prefix: their pop order
                                depends on the std::heap arrangement, the caller must replay this vertex exactly */
    PPE_EXPAND_OVERFLOW = 2  /* internal candidate capacity exceeded (caller falls back as for a tie) */
};

/* children per vertex slot: 4 endpoint edges + 2 radii x branching_factor x 2 speeds */
int ppe_expand_stride(const ppe_ctx* ctx);
/* Expands n vertices.  Outputs (host): n_children[n]; children[n * stride] (vertex v's children at
 * v * stride, in push order); flags[n] (PPE_EXPAND_*); n_popped[n] = samples the k-nearest loop consumed.
 * The ribbons-after of changed children are fetched with ppe_ribbon_pool(). */
int ppe_expand_batch(ppe_ctx* ctx, int n, const ppe_vertex* vertices, int32_t* n_children, ppe_child* children,
                     int32_t* flags, int32_t* n_popped);
/* Host copy (pinned, owned by the context, valid until the next batch) of the ribbons-after pool of the last
 * ppe_expand_batch: 4 doubles per ribbon; a child's list is [ribbons_offset, ribbons_offset + n_ribbons_after). */
const double* ppe_ribbon_pool(ppe_ctx* ctx, int64_t* n_ribbons);

/* ---- device-resident variants (inputs already in HBM; `stream` is a cudaStream_t) --------- */
int ppe_dubins_batch_device(ppe_ctx* ctx, int64_t n, const double* d_q0, const double* d_q1,
                            const double* d_rho, int32_t* d_type, double* d_param, double* d_length,
                            int32_t* d_err, void* stream);
int ppe_true_cost_batch_device(ppe_ctx* ctx, int64_t n, const ppe_edge* d_edges,
                               ppe_edge_result* d_results, void* stream);
/* (f, edge_index) of the last *_device batch; synchronises `stream`. */
int ppe_best_device(ppe_ctx* ctx, double* f, int64_t* edge_index, void* stream);
/* Same record as 16 bytes {double f; int64 edge_index + index_base} written into d_dst on `stream`
 * without synchronising: the send buffer of the per-batch NCCL gather (one record per GPU) when one
 * edge batch is sharded over ranks; index_base = the rank's first edge in the whole batch, so the
 * gathered records carry GLOBAL edge indices and ties break towards the smaller global index. */
int ppe_best_copy_device(ppe_ctx* ctx, void* d_dst16, int64_t index_base, void* stream);

/* Counter bumped by every ppe_set_map_none / ppe_set_map_bitmap on this context.  A caller that keeps
 * "which Map object did I upload last" next to its context (path_planner_b200/harness WorldCache)
 * compares it to notice uploads made by anybody else. */
uint64_t ppe_map_generation(const ppe_ctx* ctx);

/* Bounding box (metres) of the edges of the batches that follow; with PPE_MAP_TILE=1 the thread walker stages the
 * occupancy and safe bitmaps of that box in shared memory (TMA bulk copies).  x1 <= x0 clears it.  (north_star's
 * "map tiles around each batch's bounding box"; measured A/B in profiles/, off by default.) */
int ppe_set_map_window(ppe_ctx* ctx, double x0, double y0, double x1, double y1);

/* ---- instrumentation ----------------------------------------------------------------------- */
/* number of engine kernels launched on this ctx since creation */
int64_t ppe_launch_count(const ppe_ctx* ctx);
/* Dubins solves of ppe_expand_batch that entered a k-best heap, since creation */
int64_t ppe_expand_solve_count(const ppe_ctx* ctx);
/* measured FP64 FMA throughput of the device in TFLOP/s (2 flop per DFMA), `ms` of work */
int ppe_measure_fp64_peak(ppe_ctx* ctx, double* tflops, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* PPE_H */
