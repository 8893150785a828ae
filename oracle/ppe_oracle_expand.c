/* oracle/ppe_oracle_expand.c -- TEST INFRASTRUCTURE (CPU checker; never shipped, never measured).
 *
 * Plain-C restatement of the expansion step of the reference's sampling planner, one vertex at a
 * time, with the batch structs of include/ppe.h so the CUDA path (ppe_add_samples /
 * ppe_expand_batch, path_planner_b200/csrc/ppe_expand.cu) can be diffed against it field by field:
 *
 *   oracle_add_samples    SamplingBasedPlanner::addSamples        SamplingBasedPlanner.cpp:157-164
 *   oracle_expand_batch   SamplingBasedPlanner::expand            SamplingBasedPlanner.cpp:52-151
 *       nearest-endpoint edges :65-81, Euclidean pop order :85-94, per-radius k-best heaps over
 *       Edge::computeApproxCost (Edge.cpp:11-20,64-66) with std::push_heap / std::pop_heap
 *       (libstdc++ bits/stl_heap.h) :95-133, winners x speeds :134-149, then Edge::computeTrueCost for
 *       every emitted edge through oracle_true_cost_batch (oracle/ppe_oracle.c).
 *
 * The Euclidean order is a full sort by (distance, sample index): it equals the reference's heap pop
 * order whenever no two popped samples are at exactly the same distance; when they are, the vertex is
 * flagged PPE_EXPAND_TIE exactly as the device flags it and the caller replays the reference's heap.
 * Pinned end to end by the whole-plan identity tests (tests/test_harness_host_logic.py): behind this
 * evaluator the product's adapter must return the compiled reference's plan and search counters bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ppe.h"

typedef struct oracle_ctx oracle_ctx;
int oracle_is_blocked(oracle_ctx* c, double x, double y);
int oracle_dubins_batch(oracle_ctx* c, int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type,
                        double* param, double* length, int32_t* err);
int oracle_true_cost_batch(oracle_ctx* c, int64_t n, const ppe_edge* edges, ppe_edge_result* results);
int oracle_get_ribbons_after(oracle_ctx* c, int64_t i, double* xyxy, int cap);
const ppe_config* oracle_config(const oracle_ctx* c);

#define MAX_BRANCH 16

typedef struct {
    double* x; double* y; double* h;
    int64_t n, cap;
    double* pool; int64_t n_pool, cap_pool;
    int64_t solves;
} expand_state;

/* one expand_state per oracle context, kept in a small table (test infrastructure: a handful of contexts) */
static struct { oracle_ctx* c; expand_state s; } g_tab[64];
static expand_state* state_of(oracle_ctx* c) {
    for (int i = 0; i < 64; i++) if (g_tab[i].c == c) return &g_tab[i].s;
    for (int i = 0; i < 64; i++) if (!g_tab[i].c) { g_tab[i].c = c; memset(&g_tab[i].s, 0, sizeof(expand_state)); return &g_tab[i].s; }
    return NULL;
}

int oracle_clear_samples(oracle_ctx* c) { expand_state* s = state_of(c); if (!s) return PPE_ERR_CAPACITY; s->n = 0; return PPE_OK; }
int64_t oracle_sample_count(oracle_ctx* c) { expand_state* s = state_of(c); return s ? s->n : 0; }
int64_t oracle_expand_solve_count(oracle_ctx* c) { expand_state* s = state_of(c); return s ? s->solves : 0; }

int64_t oracle_add_samples(oracle_ctx* c, int64_t n, const double* x, const double* y, const double* heading, uint8_t* keep) {
    expand_state* s = state_of(c);
    if (!s) return PPE_ERR_CAPACITY;
    if (s->n + n > s->cap) {
        int64_t cap = s->cap ? s->cap : 1024;
        while (cap < s->n + n) cap *= 2;
        s->x = (double*)realloc(s->x, cap * sizeof(double));
        s->y = (double*)realloc(s->y, cap * sizeof(double));
        s->h = (double*)realloc(s->h, cap * sizeof(double));
        s->cap = cap;
    }
    int64_t kept = 0;
    for (int64_t i = 0; i < n; i++) {
        keep[i] = oracle_is_blocked(c, x[i], y[i]) ? 0 : 1; /* SamplingBasedPlanner.cpp:160 */
        if (keep[i]) { s->x[s->n] = x[i]; s->y[s->n] = y[i]; s->h[s->n] = heading[i]; s->n++; kept++; }
    }
    return kept;
}

static double yaw_of(double heading) { /* State::yaw(), State.h:51-55 */
    double h = M_PI_2 - heading;
    if (h < 0) h += 2 * M_PI;
    return h;
}

/* std::push_heap / std::pop_heap with comp(a, b) = a.cost < b.cost (SamplingBasedPlanner.cpp:171-176) */
typedef struct { double cost, len, param[3]; int type, sample; } cand_t;
static void push_up(cand_t* a, int hole, int top, cand_t v) {
    int parent = (hole - 1) / 2;
    while (hole > top && a[parent].cost < v.cost) { a[hole] = a[parent]; hole = parent; parent = (hole - 1) / 2; }
    a[hole] = v;
}
static void heap_push(cand_t* a, int* n, cand_t v) { (*n)++; push_up(a, *n - 1, 0, v); }
static void heap_pop(cand_t* a, int* n) {
    const int len = *n - 1;
    const cand_t v = a[len];
    a[len] = a[0];
    int hole = 0, second = 0;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (a[second].cost < a[second - 1].cost) second--;
        a[hole] = a[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        a[hole] = a[second - 1];
        hole = second - 1;
    }
    push_up(a, hole, 0, v);
    *n = len;
}

typedef struct { double d; int idx; } key_t_;
static int key_cmp(const void* a, const void* b) {
    const key_t_* p = (const key_t_*)a; const key_t_* q = (const key_t_*)b;
    if (p->d < q->d) return -1;
    if (p->d > q->d) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

int oracle_expand_stride(oracle_ctx* c) { return 4 + 4 * oracle_config(c)->branching_factor; }

int oracle_expand_batch(oracle_ctx* c, int n, const ppe_vertex* verts, int32_t* n_children, ppe_child* children, int32_t* flags,
                        int32_t* n_popped) {
    expand_state* s = state_of(c);
    if (!s) return PPE_ERR_CAPACITY;
    const ppe_config* cfg = oracle_config(c);
    const int k = cfg->branching_factor;
    if (k < 1 || k > MAX_BRANCH) return PPE_ERR_CAPACITY;
    const int stride = 4 + 4 * k;
    const double speeds[2] = {cfg->max_speed, cfg->max_speed == cfg->slow_speed ? -1 : cfg->slow_speed};
    const double radii[2] = {cfg->turning_radius, cfg->coverage_turning_radius == cfg->turning_radius ? -1 : cfg->coverage_turning_radius};
    const double inc = cfg->collision_checking_increment;
    key_t_* keys = (key_t_*)malloc((size_t)(s->n > 0 ? s->n : 1) * sizeof(key_t_));
    ppe_edge* edges = (ppe_edge*)calloc((size_t)stride, sizeof(ppe_edge));
    ppe_edge_result* res = (ppe_edge_result*)calloc((size_t)stride, sizeof(ppe_edge_result));
    int32_t* esample = (int32_t*)malloc((size_t)stride * sizeof(int32_t));
    s->n_pool = 0;
    for (int v = 0; v < n; v++) {
        const ppe_vertex* vx = &verts[v];
        const double sx = vx->state[0], sy = vx->state[1];
        for (int64_t i = 0; i < s->n; i++) {
            keys[i].d = sqrt((sx - s->x[i]) * (sx - s->x[i]) + (sy - s->y[i]) * (sy - s->y[i])); /* State::distanceTo */
            keys[i].idx = (int)i;
        }
        qsort(keys, (size_t)s->n, sizeof(key_t_), key_cmp);
        cand_t heap[2][MAX_BRANCH + 1];
        int hn[2] = {0, 0};
        int done[2] = {0, 0};
        int pops = 0;
        const double q0[3] = {sx, sy, yaw_of(vx->state[2])};
        for (int64_t i = 0; i < s->n && (!done[0] || !done[1]); i++) { /* :91 */
            const int si = keys[i].idx;
            const double dist = keys[i].d;
            pops++;
            for (int j = 0; j < 2; j++) {
                if (done[j]) continue;
                if (radii[j] <= 0) { done[j] = 1; continue; }
                if (hn[j] < k || heap[j][0].len > dist) {
                    if (dist > inc) {
                        const double q1[3] = {s->x[si], s->y[si], yaw_of(s->h[si])};
                        cand_t cd;
                        int32_t err = 0;
                        oracle_dubins_batch(c, 1, q0, q1, &radii[j], &cd.type, cd.param, &cd.len, &err);
                        cd.cost = cd.len / cfg->max_speed * cfg->time_penalty_factor;
                        cd.sample = si;
                        heap_push(heap[j], &hn[j], cd);
                        s->solves++;
                        if (hn[j] > k) heap_pop(heap[j], &hn[j]);
                    }
                } else {
                    done[j] = 1;
                }
            }
        }
        int flag = 0;
        {
            const int64_t upto = pops < s->n ? pops : s->n - 1;
            for (int64_t q = 0; q < upto; q++) if (keys[q].d == keys[q + 1].d) flag |= PPE_EXPAND_TIE;
        }
        /* edges in push order */
        int ne = 0;
        if (vx->has_endpoint) {
            for (int si = 0; si < 2; si++) {
                if (speeds[si] <= 0) continue;
                for (int j = 0; j < 2; j++) {
                    if (radii[j] <= 0) continue;
                    ppe_edge* e = &edges[ne];
                    memset(e, 0, sizeof *e);
                    memcpy(e->src, vx->state, sizeof e->src);
                    e->src_g = vx->g; e->ribbon_set = vx->ribbon_set;
                    e->dst[0] = vx->endpoint[0]; e->dst[1] = vx->endpoint[1]; e->dst[2] = vx->endpoint[2]; e->dst[3] = speeds[si];
                    e->has_path = 0;
                    e->coverage_allowed = radii[j] == cfg->coverage_turning_radius;
                    esample[ne++] = -1;
                }
            }
        }
        for (int j = 0; j < 2; j++) {
            for (int w = 0; w < hn[j]; w++) {
                for (int si = 0; si < 2; si++) {
                    if (speeds[si] <= 0) continue;
                    const cand_t* cd = &heap[j][w];
                    ppe_edge* e = &edges[ne];
                    memset(e, 0, sizeof *e);
                    memcpy(e->src, vx->state, sizeof e->src);
                    e->src_g = vx->g; e->ribbon_set = vx->ribbon_set;
                    e->has_path = 1;
                    e->path_qi[0] = q0[0]; e->path_qi[1] = q0[1]; e->path_qi[2] = q0[2];
                    memcpy(e->path_param, cd->param, sizeof e->path_param);
                    e->path_rho = radii[j];
                    e->path_type = cd->type;
                    e->w_speed = speeds[si];
                    e->w_start_time = vx->state[4];
                    e->w_end_time = vx->state[4] + cd->len / speeds[si];
                    e->dst[0] = s->x[cd->sample]; e->dst[1] = s->y[cd->sample]; e->dst[2] = s->h[cd->sample]; e->dst[3] = speeds[si];
                    e->coverage_allowed = radii[j] == cfg->coverage_turning_radius;
                    esample[ne++] = cd->sample;
                }
            }
        }
        if (ne) oracle_true_cost_batch(c, ne, edges, res);
        for (int e = 0; e < stride; e++) {
            ppe_child* ch = &children[(size_t)v * stride + e];
            memset(ch, 0, sizeof *ch);
            ch->ribbons_offset = -1;
            ch->sample_index = -1;
            if (e >= ne) { ch->status = PPE_EDGE_SKIPPED; continue; }
            const ppe_edge_result* r = &res[e];
            ch->true_cost = r->true_cost; ch->collision_penalty = r->collision_penalty; ch->approx_cost = r->approx_cost;
            memcpy(ch->end, r->end, sizeof ch->end);
            ch->g = r->g; ch->h = r->h; ch->coverage_completed_time = r->coverage_completed_time;
            memcpy(ch->path_param, r->path_param, sizeof ch->path_param);
            ch->w_end_time = r->w_end_time;
            ch->sample_index = esample[e];
            ch->path_type = r->path_type; ch->infeasible = r->infeasible; ch->status = r->status;
            ch->coverage_allowed = edges[e].coverage_allowed;
            ch->n_ribbons_after = r->n_ribbons_after; ch->ribbons_changed = r->ribbons_changed;
            if (r->status == PPE_EDGE_OK && r->ribbons_changed) {
                if (s->n_pool + r->n_ribbons_after > s->cap_pool) {
                    int64_t cap = s->cap_pool ? s->cap_pool : 4096;
                    while (cap < s->n_pool + r->n_ribbons_after) cap *= 2;
                    s->pool = (double*)realloc(s->pool, (size_t)cap * 4 * sizeof(double));
                    s->cap_pool = cap;
                }
                ch->ribbons_offset = s->n_pool;
                oracle_get_ribbons_after(c, e, s->pool + 4 * s->n_pool, r->n_ribbons_after);
                s->n_pool += r->n_ribbons_after;
            }
        }
        n_children[v] = ne;
        flags[v] = flag;
        n_popped[v] = pops;
    }
    free(keys); free(edges); free(res); free(esample);
    return PPE_OK;
}

const double* oracle_ribbon_pool(oracle_ctx* c, int64_t* n_ribbons) {
    expand_state* s = state_of(c);
    if (n_ribbons) *n_ribbons = s ? s->n_pool : 0;
    return s ? s->pool : NULL;
}
