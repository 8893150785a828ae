/* oracle/ppe_on_oracle.c -- TEST INFRASTRUCTURE (a test double, never shipped, never measured).
 *
 * Implements the subset of include/ppe.h that the C++ host adapter
 * (path_planner_b200/harness/BatchedAStarPlanner.cpp) calls by forwarding to the CPU oracle
 * (oracle/ppe_oracle.c, glibc variant == the compiled reference bit for bit).  Linked ONLY into
 * oracle/_ref/libplan_compare_cpu.so so that `-m "not gpu"` tests can check the adapter's host
 * logic -- batch assembly, the replay of the k-nearest heaps, the push order into the open list --
 * in a container without a GPU: with a bit-identical evaluator behind it, BatchedAStarPlanner must
 * return the reference's plan bit for bit.  The product library libppe.so has no such path.
 */
#include <stdint.h>
#include <stdlib.h>

#include "ppe.h"

typedef struct oracle_ctx oracle_ctx;
int oracle_create(oracle_ctx** out);
void oracle_destroy(oracle_ctx* c);
const char* oracle_last_error(const oracle_ctx* c);
int oracle_set_config(oracle_ctx* c, const ppe_config* cfg);
int oracle_set_map_none(oracle_ctx* c);
int oracle_set_map_bitmap(oracle_ctx* c, const uint8_t* bits, int rows, int cols, int stride, double res);
int oracle_set_obstacles_none(oracle_ctx* c);
int oracle_set_obstacles_binary(oracle_ctx* c, int n, const double* x, const double* y, const double* yaw,
                                const double* speed, const double* time, const double* width, const double* length);
int oracle_set_obstacles_gaussian(oracle_ctx* c, int n, const double* x, const double* y, const double* yaw,
                                  const double* speed, const double* time, const double* cov);
int oracle_put_ribbon_set(oracle_ctx* c, int n, const double* xyxy, double cct, int32_t* id);
int oracle_clear_ribbon_sets(oracle_ctx* c);
int oracle_dubins_batch(oracle_ctx* c, int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type,
                        double* param, double* length, int32_t* err);
int oracle_true_cost_batch(oracle_ctx* c, int64_t n, const ppe_edge* edges, ppe_edge_result* results);
int oracle_get_ribbons_after(oracle_ctx* c, int64_t i, double* xyxy, int cap);

int oracle_clear_samples(oracle_ctx* c);
int64_t oracle_sample_count(oracle_ctx* c);
int64_t oracle_expand_solve_count(oracle_ctx* c);
int64_t oracle_add_samples(oracle_ctx* c, int64_t n, const double* x, const double* y, const double* heading, uint8_t* keep);
int oracle_expand_stride(oracle_ctx* c);
int oracle_expand_batch(oracle_ctx* c, int n, const ppe_vertex* verts, int32_t* n_children, ppe_child* children, int32_t* flags,
                        int32_t* n_popped);
const double* oracle_ribbon_pool(oracle_ctx* c, int64_t* n_ribbons);

#define O(ctx) ((oracle_ctx*)(ctx))

int ppe_abi_version(void) { return PPE_ABI_VERSION; }
static uint64_t g_map_generation = 0;
uint64_t ppe_map_generation(const ppe_ctx* ctx) { (void)ctx; return g_map_generation; }
int ppe_create(int device, ppe_ctx** out) { (void)device; return oracle_create((oracle_ctx**)out); }
void ppe_destroy(ppe_ctx* ctx) { oracle_destroy(O(ctx)); }
const char* ppe_last_error(const ppe_ctx* ctx) { return oracle_last_error((const oracle_ctx*)ctx); }
int ppe_set_config(ppe_ctx* ctx, const ppe_config* cfg) { return oracle_set_config(O(ctx), cfg); }
int ppe_set_map_none(ppe_ctx* ctx) { g_map_generation++; return oracle_set_map_none(O(ctx)); }
int ppe_set_map_bitmap(ppe_ctx* ctx, const uint8_t* bits, int rows, int cols, int stride, double res) {
    g_map_generation++;
    return oracle_set_map_bitmap(O(ctx), bits, rows, cols, stride, res);
}
int ppe_set_obstacles_none(ppe_ctx* ctx) { return oracle_set_obstacles_none(O(ctx)); }
int ppe_set_obstacles_binary(ppe_ctx* ctx, int n, const double* x, const double* y, const double* yaw, const double* speed,
                             const double* time, const double* width, const double* length) {
    return oracle_set_obstacles_binary(O(ctx), n, x, y, yaw, speed, time, width, length);
}
int ppe_set_obstacles_gaussian(ppe_ctx* ctx, int n, const double* x, const double* y, const double* yaw, const double* speed,
                               const double* time, const double* cov) {
    return oracle_set_obstacles_gaussian(O(ctx), n, x, y, yaw, speed, time, cov);
}
int ppe_put_ribbon_set(ppe_ctx* ctx, int n, const double* xyxy, double cct, int32_t* id) {
    return oracle_put_ribbon_set(O(ctx), n, xyxy, cct, id);
}
int ppe_clear_ribbon_sets(ppe_ctx* ctx) { return oracle_clear_ribbon_sets(O(ctx)); }
int ppe_dubins_batch(ppe_ctx* ctx, int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type,
                     double* param, double* length, int32_t* err) {
    return oracle_dubins_batch(O(ctx), n, q0, q1, rho, type, param, length, err);
}
int ppe_true_cost_batch(ppe_ctx* ctx, int64_t n, const ppe_edge* edges, ppe_edge_result* results) {
    return oracle_true_cost_batch(O(ctx), n, edges, results);
}
int ppe_get_ribbons_after(ppe_ctx* ctx, int64_t i, double* xyxy, int cap) { return oracle_get_ribbons_after(O(ctx), i, xyxy, cap); }

/* frontier expansion (oracle/ppe_oracle_expand.c) */
int ppe_clear_samples(ppe_ctx* ctx) { return oracle_clear_samples(O(ctx)); }
int64_t ppe_sample_count(const ppe_ctx* ctx) { return oracle_sample_count((oracle_ctx*)ctx); }
int64_t ppe_expand_solve_count(const ppe_ctx* ctx) { return oracle_expand_solve_count((oracle_ctx*)ctx); }
int64_t ppe_add_samples(ppe_ctx* ctx, int64_t n, const double* x, const double* y, const double* heading, uint8_t* keep) {
    return oracle_add_samples(O(ctx), n, x, y, heading, keep);
}
int ppe_expand_stride(const ppe_ctx* ctx) { return oracle_expand_stride((oracle_ctx*)ctx); }
int ppe_expand_batch(ppe_ctx* ctx, int n, const ppe_vertex* v, int32_t* n_children, ppe_child* children, int32_t* flags, int32_t* n_popped) {
    return oracle_expand_batch(O(ctx), n, v, n_children, children, flags, n_popped);
}
const double* ppe_ribbon_pool(ppe_ctx* ctx, int64_t* n) { return oracle_ribbon_pool(O(ctx), n); }
