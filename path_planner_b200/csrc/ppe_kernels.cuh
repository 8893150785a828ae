// ppe_kernels.cuh -- device-side world description and kernel launch prototypes (internal).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ppe.h"

namespace ppe {

// Dynamic obstacle with the loop invariants hoisted on the host (the reference recomputes
// cos/sin(Yaw), Width+2, the covariance inverse and the normaliser for every sample point:
// BinaryDynamicObstaclesManager.cpp:8-20, GaussianDynamicObstaclesManager.h:31-44).
struct ObstacleD {
    double X, Y, Time, Speed;
    double cosYaw, sinYaw;
    // binary: a = (Length+2)/2, b = (Width+2)/2 (strict inflation, Binary...cpp:9-12,18)
    // gaussian: a..d = inverse covariance i00, i10, i01, i11; norm = 1/(2 pi)/sqrt(det)
    double a, b, c, d;
    double norm;
    // chunk-culling bound (host-computed): gaussian: sqrt(lambda_max) of the symmetric part of the inverse
    // covariance, so that sqrt(q(v + e)) >= sqrt(q(v)) - |e| * cull; unused for binary obstacles
    double cull;
};
static_assert(sizeof(ObstacleD) == 96, "ObstacleD layout");

enum { kMapNone = 0, kMapBitmap = 1 };
enum { kObsNone = 0, kObsBinary = 1, kObsGaussian = 2 };

// Read-only world state resident in HBM for the lifetime of a plan.
struct WorldD {
    ppe_config cfg;
    double dt;            // collisionCheckingIncrement / maxSpeed        (Edge.cpp:114)
    double horizon_end;   // timeHorizon + 1e-12 + startStateTime        (Edge.cpp:90)
    // static map: rows x stride_words 32-bit words, bit c of row r = blocked[r][c]
    const uint32_t* map_bits;
    // same layout: bit = every cell within `safe_radius_cells` (Chebyshev) is in bounds and free; nullptr when
    // the radius is too large to be useful (chunk culling then never skips map look-ups)
    const uint32_t* safe_bits;
    int map_kind, rows, cols, stride_words;
    double resolution;
    double inv_resolution;    // exact when the resolution is a power of two (res_pow2): x / res == x * inv
    int res_pow2;
    // optional shared-memory tile of both bitmaps for the thread walker (N2 A/B, PPE_MAP_TILE): rows [tile_r0, tile_r0 +
    // tile_rows) x words [tile_w0, tile_w0 + tile_words) staged per CTA with cp.async.bulk; look-ups outside fall back to L2
    int tile_on, tile_r0, tile_w0, tile_rows, tile_words;
    int obs_cull_ok;          // obstacle set admits the chunk bound (SPD symmetric covariances, <= 64 obstacles)
    // dynamic obstacles
    const ObstacleD* obstacles;
    int obs_kind, n_obs;
    // interned ribbon sets (RibbonManager state of parent vertices)
    const double4* ribbons;     // pool: sx, sy, ex, ey
    const double4* boxes;       // per ribbon of the pool: x_lo, x_hi, y_lo, y_hi of its bounding box grown by the ribbon width + margin
    const int* set_offset;
    const int* set_count;
    const double* set_cct;      // coverageCompletedTime
    // per-set invariants of the UNCHANGED list (what the thread walker evaluates): sum over the list, in list order, of
    // length - 2 * RibbonWidth (RibbonManager::maxDistance, RibbonManager.cpp:238-240) and "all coordinates below 1e7"
    const double* set_sumlen;
    const int* set_tame;        // bit 0: tame; bit 1: the list holds a ribbon short enough for a strict cover() to erase it
    int n_sets;
    int ribbon_cap;             // per-warp working capacity (ribbons)
    // ribbons-after output pool
    double4* out_ribbons;
    unsigned long long* out_count;
    unsigned long long out_cap;
};

struct BestD {
    double f;
    long long idx;
};

// K2t hand-over budgets (measured best on C2 / C3 / C5; PPE_K2T_DIRTY / PPE_K2T_CPS override them per context)
struct K2Tuning {
    int dirty_budget = 64; // non-clean chunks of an edge that K2t evaluates (warp-cooperatively) before handing the edge to K2b
    int cp_budget = 6;    // ribbon check-points a K2t thread walks
    int k2b_ctas_per_sm = 0; // K2b CTAs (4 warps each) resident per SM; 0 = as many as fit (4).  PPE_K2B_CTAS
    int deep_walker = 0;  // K2c: thread-per-edge walk of the edges K2t caught covering a ribbon (PPE_DEEP_WALKER=1; measured
                          // slower than handing them to K2b -- DESIGN.md section 4.5 -- so off by default)
};
K2Tuning clamp_tuning(K2Tuning t);
// heuristics other than MaxDistance that the kernels evaluate themselves (h >= 0 in the result records)
inline bool tsp_on_device(int heuristic) {
    return heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL || heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K;
}

// ---- frontier expansion (ppe_expand.cu) ------------------------------------------------------------------------------
struct ExpandParamsD {
    const double* sx;      // resident samples, SoA
    const double* sy;
    const double* sh;
    int n_samples;
    int k;                 // PlannerConfig::branchingFactor()
    int stride;            // child slots per vertex: 4 + 4 k
    double inc;            // collisionCheckingIncrement
    double max_speed;
    double time_factor;    // Edge::timePenaltyFactor()
    double speed[2];       // {max, slow or -1}                       (SamplingBasedPlanner.cpp:58-59)
    double rho[2];         // {turning, coverage or -1}               (:61-63)
    double coverage_rho;
    double r2_init;        // first search radius (squared) of the candidate collection
};
int expand_max_branch();
cudaError_t launch_sample_filter(const WorldD& w, int64_t n, const double* x, const double* y, const double* h, uint8_t* keep,
                                 unsigned int* block_count, unsigned int* total, double* sx, double* sy, double* sh, int64_t base,
                                 cudaStream_t stream, int phase);
cudaError_t launch_expand_select(const ExpandParamsD& p, int n_vertices, const ppe_vertex* verts, ppe_edge* edges,
                                 int32_t* edge_sample, int32_t* n_children, int32_t* flags, int32_t* n_popped, int32_t* n_solved,
                                 cudaStream_t stream);
cudaError_t launch_expand_pack(int64_t n, const ppe_edge* edges, const ppe_edge_result* results, const int32_t* edge_sample,
                               ppe_child* out, cudaStream_t stream);

// launchers (ppe_kernels.cu)
cudaError_t launch_dubins_batch(int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type,
                                double* param, double* length, int32_t* err, cudaStream_t stream);
cudaError_t launch_true_cost_kernels(const WorldD& world, int64_t n, const ppe_edge* edges, void* prepared_scratch,
                                     ppe_edge_result* results, unsigned long long* counters, unsigned int* heavy_list,
                                     BestD* block_best, int max_blocks, int sm_count, cudaStream_t stream, bool reset_pool,
                                     K2Tuning tuning, int* blocks_out, int* launches_out);
// host-buffer pipeline of ppe_true_cost_batch: K2a + K2t per slice of the batch, K2b + K3a once, then the K2b records
// to the caller's array (see ppe_kernels.cu)
cudaError_t launch_prepare_and_walk(const WorldD& world, int64_t n_total, int64_t first, int64_t cnt, const ppe_edge* edges,
                                    void* prepared_scratch, ppe_edge_result* results, unsigned long long* counters,
                                    unsigned int* heavy_list, cudaStream_t stream, K2Tuning tuning, int* launches_out);
cudaError_t launch_heavy_and_best(const WorldD& world, int64_t n_total, const ppe_edge* edges, void* prepared_scratch,
                                  ppe_edge_result* results, unsigned long long* counters, unsigned int* heavy_list,
                                  BestD* block_best, int max_blocks, int sm_count, cudaStream_t stream, K2Tuning tuning,
                                  int* blocks_out, int* launches_out);
cudaError_t launch_patch_results(const ppe_edge_result* results, const unsigned int* heavy_list, const unsigned long long* counters,
                                 int64_t n_total, ppe_edge_result* dst, unsigned int* compact_idx, int sm_count, cudaStream_t stream);
cudaError_t launch_best_final(const BestD* block_best, int blocks, BestD* best, int64_t index_base, bool accumulate,
                              cudaStream_t stream);
// *dst = {src->f, src->idx + index_base}: the NCCL send record of a rank that holds edges [index_base, ...)
cudaError_t launch_best_export(const BestD* src, BestD* dst, int64_t index_base, cudaStream_t stream);
size_t prepared_edge_bytes();
cudaError_t launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t stream);
int k2b_profile_read(unsigned long long* out16);
// safe[r][c] = all cells within Chebyshev distance `radius` of (r, c) are in bounds and free
cudaError_t launch_safe_map(const uint32_t* map_bits, uint32_t* scratch_rows, uint32_t* safe_bits, int rows, int cols,
                            int stride_words, int radius, cudaStream_t stream);
// samples per chunk of the K2b walker (the unit of culling); the safe-map radius is derived from it
int true_cost_chunk_samples();
size_t true_cost_smem_bytes(int ribbon_cap, int n_obs);
int true_cost_block_threads();

} // namespace ppe
