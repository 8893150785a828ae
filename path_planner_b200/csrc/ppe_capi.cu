// ppe_capi.cu -- the C ABI of include/ppe.h: context, HBM-resident world state, batch calls.
//
// Host side of the engine.  Loop invariants that the reference recomputes per sample point with
// the host libm (cos/sin of obstacle yaws, covariance inverse, normaliser) are computed here, once,
// with the same host libm; everything per edge and per sample runs in the kernels of
// ppe_kernels.cu.  There is no CPU evaluation path: without a usable CUDA device ppe_create fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ppe.h"
#include "ppe_kernels.cuh"

using namespace ppe;

struct ppe_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // host-buffer batches are pipelined in slices (batch_pipelined): H2D, K2a + K2t and D2H of different slices overlap,
    // consecutive slices alternating between two kernel lanes; K2b runs once per batch on `stream`.  (PPE_LATE_K2B=0: the
    // round-1 pipeline, the whole kernel sequence per slice, each lane with its own scratch.)
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    struct Lane {
        cudaStream_t stream = nullptr;
        unsigned char* d_prepared = nullptr;
        size_t cap_prepared = 0;
        unsigned int* d_heavy = nullptr;   // heavy list of the lane's slice (K2t -> K2b)
        size_t cap_heavy = 0;
        unsigned long long* d_work = nullptr; // [0] K2b work counter, [1] heavy-list length
        BestD* d_block_best = nullptr;
        cudaEvent_t ev_k3 = nullptr; // K3 of the lane's previous slice has consumed d_block_best
        bool used = false;
    } lanes[2];
    std::vector<cudaEvent_t> ev_in, ev_k2, ev_done;
    std::string err;

    ppe_config cfg{};
    bool have_cfg = false;

    // static map
    uint32_t* d_map = nullptr;
    uint32_t* d_safe = nullptr;     // dilated free-space bitmap for chunk culling (built lazily per config)
    uint32_t* d_safe_tmp = nullptr;
    int safe_radius = -1;           // radius d_safe was built for; -1 = stale
    int map_kind = kMapNone, rows = 0, cols = 0, stride_words = 0;
    double resolution = 1;
    uint64_t map_generation = 0;
    // N2 A/B (PPE_MAP_TILE=1): bounding box in metres of the batch about to be evaluated; empty = no tile
    double win_x0 = 0, win_y0 = 0, win_x1 = -1, win_y1 = -1;
    int tile_mode = 0;              // PPE_MAP_TILE
    int l2_persist = 0;             // PPE_MAP_L2_PERSIST: persisting L2 access-policy window over the two bitmaps    // bumped by every ppe_set_map_*: lets a caller-side cache notice foreign uploads
    int obs_cull_ok = 1;

    // dynamic obstacles
    ObstacleD* d_obs = nullptr;
    int obs_kind = kObsNone, n_obs = 0;

    // interned ribbon sets: host mirror + device pool (uploaded lazily before a batch)
    std::vector<double> h_ribbons; // 4 doubles per ribbon
    std::vector<int> h_off, h_cnt;
    std::vector<double> h_cct;
    std::vector<double> h_sumlen;  // per set, see WorldD::set_sumlen
    std::vector<int> h_tame;
    double sets_width = -1;        // ribbon width the per-set sums were computed with
    double* d_sumlen = nullptr;
    int* d_tame = nullptr;
    bool sets_dirty = true;
    size_t uploaded_ribbons = 0, uploaded_sets = 0; // the pool is append-only: only the tail is copied
    int max_set = 0;
    double4* d_ribbons = nullptr;
    double4* d_boxes = nullptr;    // per ribbon of the pool: bounding box grown by the ribbon width (WorldD::boxes)
    size_t cap_boxes = 0, uploaded_boxes = 0;
    std::vector<double> h_boxes;
    int* d_off = nullptr;
    int* d_cnt = nullptr;
    double* d_cct = nullptr;
    size_t cap_ribbons = 0, cap_sets = 0;

    // grow-only batch buffers for the host-pointer entry points
    ppe_edge* d_edges = nullptr;
    ppe_edge_result* d_results = nullptr;
    size_t cap_edges = 0, cap_results = 0, cap_dubi = 0;
    unsigned char* d_prepared = nullptr; // K2a -> K2b per-edge scratch
    size_t cap_prepared = 0;
    double* d_dub = nullptr; // q0 q1 rho param length
    int32_t* d_dubi = nullptr; // type err
    size_t cap_dub = 0;

    // ribbons-after pool
    double4* d_out_ribbons = nullptr;
    unsigned long long* d_out_count = nullptr;
    size_t out_cap = 0;

    unsigned long long* d_work = nullptr; // [0] K2b work counter, [1] heavy-list length
    unsigned int* d_heavy = nullptr;
    size_t cap_heavy = 0;
    unsigned char* d_patch = nullptr;     // K2b's records packed, then their edge indices (pageable result arrays only)
    size_t cap_patch = 0;
    bool thread_walker = true;            // PPE_THREAD_WALKER=0: the warp walker evaluates every edge
    bool late_k2b = true;                 // PPE_LATE_K2B=0: slice the whole kernel sequence of host-buffer batches (A/B)
    int64_t late_slice = 131072;          // PPE_LATE_SLICE: edges per slice of the late-K2b pipeline (default: n / 8, 16 Ki .. 128 Ki)
    bool late_slice_fixed = false;
    K2Tuning tuning;                      // PPE_K2T_DIRTY / PPE_K2T_CPS, read at ppe_create
    BestD* d_block_best = nullptr;
    BestD* d_best = nullptr;
    int max_blocks = 0;

    // bookkeeping of the last true-cost batch (for ppe_get_ribbons_after / ppe_best): the device copies of
    // its edges and results stay valid until the next batch, records are fetched on demand
    int64_t last_count = 0;
    std::vector<double> h_out_ribbons;
    bool out_downloaded = false;
    unsigned long long last_out_count = 0;
    bool have_batch = false;

    // ---- frontier expansion: resident samples + per-batch buffers (ppe_expand.cu) ----
    double *d_sx = nullptr, *d_sy = nullptr, *d_sh = nullptr; // m_Samples, SoA
    size_t cap_samples = 0;
    int64_t n_samples = 0;
    double* d_stage = nullptr;            // x, y, heading of the states being added
    uint8_t* d_keep = nullptr;
    unsigned int* d_blockcnt = nullptr;   // per-block keep counts / offsets + [last] total
    size_t cap_stage = 0;
    ppe_vertex* d_verts = nullptr;
    ppe_edge* d_xedges = nullptr;
    ppe_edge_result* d_xresults = nullptr;
    ppe_child* d_children = nullptr;
    int32_t* d_xints = nullptr;           // [edge_sample: n * stride][n_children n][flags n][n_popped n][n_solved n]
    size_t cap_verts = 0;
    void* h_pinned = nullptr;             // pinned staging of one expand batch (inputs and outputs)
    size_t cap_pinned = 0;
    double* h_pool = nullptr;             // pinned host copy of the ribbons-after pool of the last expand batch
    size_t cap_pool = 0;
    int64_t n_pool = 0;
    int64_t expand_solves = 0;            // instrumentation: Dubins solves that ended up in a k-best heap
    bool last_was_expand = false;

    int64_t launches = 0;
};

namespace {

int fail(ppe_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    if (ctx) ctx->err = buf;
    return code;
}

#define PPE_CUDA(ctx, call)                                                   \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return fail((ctx), PPE_ERR_CUDA, #call, e_);   \
    } while (0)

template <typename T>
int grow(ppe_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (need <= *cap) return PPE_OK;
    size_t ncap = *cap ? *cap : 1024;
    while (ncap < need) ncap *= 2;
    if (*p) PPE_CUDA(ctx, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    PPE_CUDA(ctx, cudaMalloc((void**)p, ncap * sizeof(T)));
    *cap = ncap;
    return PPE_OK;
}

int ribbon_cap_for(int max_set) {
    int cap = ((max_set + 32 + 31) / 32) * 32; // head-room for splits (a split adds one ribbon)
    if (cap < 32) cap = 32;
    return cap;
}

// grows a device array to hold `need` elements, keeping its first `keep` (device-to-device copy)
template <typename T>
int grow_keep(ppe_ctx* ctx, T** p, size_t* cap, size_t need, size_t keep, size_t first_cap) {
    if (need <= *cap) return PPE_OK;
    size_t c = *cap ? *cap : first_cap;
    while (c < need) c *= 2;
    T* fresh = nullptr;
    PPE_CUDA(ctx, cudaMalloc((void**)&fresh, c * sizeof(T)));
    if (*p && keep) PPE_CUDA(ctx, cudaMemcpyAsync(fresh, *p, keep * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    if (*p) {
        PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PPE_CUDA(ctx, cudaFree(*p));
    }
    *p = fresh;
    *cap = c;
    return PPE_OK;
}

// The interned pool is append-only between ppe_clear_ribbon_sets calls: only the sets added since the last
// upload travel (one expansion interns one set, so a plan of E expansions copies O(E) ribbons, not O(E^2)).
int upload_sets(ppe_ctx* ctx) {
    if (!ctx->sets_dirty) return PPE_OK;
    const size_t nr = ctx->h_ribbons.size() / 4, ns = ctx->h_cnt.size();
    if (ctx->uploaded_ribbons > nr || ctx->uploaded_sets > ns) ctx->uploaded_ribbons = ctx->uploaded_sets = 0;
    if (ctx->sets_width != ctx->cfg.ribbon_width) { // the per-set sums depend on the ribbon width: recompute and upload all
        ctx->sets_width = ctx->cfg.ribbon_width;
        for (size_t k = 0; k < ns; k++) {
            double sum = 0;
            int any_short = 0;
            const double ml = 2 * ctx->sets_width; // Ribbon::minLength(); covered(strict): |e|^2 < ml^2 / (2 * 2), Ribbon.cpp:23-25
            const double* r = ctx->h_ribbons.data() + 4 * (size_t)ctx->h_off[k];
            for (int q = 0; q < ctx->h_cnt[k]; q++, r += 4) {
                const double sq = (r[2] - r[0]) * (r[2] - r[0]) + (r[3] - r[1]) * (r[3] - r[1]);
                sum += sqrt(sq) - 2 * ctx->sets_width;
                if (sq < ml * ml / (2.0 * 2.0)) any_short = 1;
            }
            ctx->h_sumlen[k] = sum;
            ctx->h_tame[k] = (ctx->h_tame[k] & 1) | (any_short ? 2 : 0);
        }
        ctx->uploaded_sets = 0;
        ctx->uploaded_boxes = 0; // the boxes are grown by the width
    }
    const size_t r0 = ctx->uploaded_ribbons, s0 = ctx->uploaded_sets;
    int rc = grow_keep(ctx, &ctx->d_ribbons, &ctx->cap_ribbons, nr + 1, r0, 1024);
    if (rc != PPE_OK) return rc;
    {
        // the box ribbon_may_contain (ppe_kernels.cu) compares a point with: the segment's bounding box grown by the ribbon
        // width and the shortcut's margin -- the same IEEE operations the kernel used to repeat per ribbon and check-point
        if (ctx->uploaded_boxes > nr || ctx->uploaded_boxes > r0) ctx->uploaded_boxes = 0;
        const size_t b0 = ctx->uploaded_boxes;
        ctx->h_boxes.resize(nr * 4);
        const double grow = ctx->cfg.ribbon_width * (1 + 1e-9) + 1e-3;
        for (size_t i = b0; i < nr; i++) {
            const double* r = ctx->h_ribbons.data() + 4 * i;
            double* b = ctx->h_boxes.data() + 4 * i;
            b[0] = fmin(r[0], r[2]) - grow; b[1] = fmax(r[0], r[2]) + grow;
            b[2] = fmin(r[1], r[3]) - grow; b[3] = fmax(r[1], r[3]) + grow;
        }
        rc = grow_keep(ctx, &ctx->d_boxes, &ctx->cap_boxes, nr + 1, b0, 1024);
        if (rc != PPE_OK) return rc;
        if (nr > b0)
            PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_boxes + b0, ctx->h_boxes.data() + 4 * b0, (nr - b0) * 4 * sizeof(double),
                                          cudaMemcpyHostToDevice, ctx->stream));
        ctx->uploaded_boxes = nr;
    }
    {
        size_t c1 = ctx->cap_sets, c2 = ctx->cap_sets, c3 = ctx->cap_sets, c4 = ctx->cap_sets, c5 = ctx->cap_sets;
        rc = grow_keep(ctx, &ctx->d_off, &c1, ns + 1, s0, 64);
        if (rc == PPE_OK) rc = grow_keep(ctx, &ctx->d_cnt, &c2, ns + 1, s0, 64);
        if (rc == PPE_OK) rc = grow_keep(ctx, &ctx->d_cct, &c3, ns + 1, s0, 64);
        if (rc == PPE_OK) rc = grow_keep(ctx, &ctx->d_sumlen, &c4, ns + 1, s0, 64);
        if (rc == PPE_OK) rc = grow_keep(ctx, &ctx->d_tame, &c5, ns + 1, s0, 64);
        if (rc != PPE_OK) return rc;
        ctx->cap_sets = c1;
    }
    if (nr > r0)
        PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_ribbons + r0, ctx->h_ribbons.data() + 4 * r0, (nr - r0) * 4 * sizeof(double),
                                      cudaMemcpyHostToDevice, ctx->stream));
    if (ns > s0) {
        PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_off + s0, ctx->h_off.data() + s0, (ns - s0) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_cnt + s0, ctx->h_cnt.data() + s0, (ns - s0) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_cct + s0, ctx->h_cct.data() + s0, (ns - s0) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_sumlen + s0, ctx->h_sumlen.data() + s0, (ns - s0) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_tame + s0, ctx->h_tame.data() + s0, (ns - s0) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    // batches run on other streams (caller's / the lanes'): the tail must have landed before they start
    PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->uploaded_ribbons = nr;
    ctx->uploaded_sets = ns;
    ctx->sets_dirty = false;
    return PPE_OK;
}

int ensure_pool(ppe_ctx* ctx, size_t want) {
    if (ctx->d_out_ribbons && ctx->out_cap >= want) return PPE_OK;
    if (ctx->d_out_ribbons) PPE_CUDA(ctx, cudaFree(ctx->d_out_ribbons));
    ctx->d_out_ribbons = nullptr;
    ctx->out_cap = 0;
    PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_out_ribbons, want * sizeof(double4)));
    ctx->out_cap = want;
    return PPE_OK;
}

int make_world(ppe_ctx* ctx, WorldD* w) {
    if (!ctx->have_cfg) return fail(ctx, PPE_ERR_STATE, "ppe_set_config must be called before a batch");
    int rc = upload_sets(ctx);
    if (rc != PPE_OK) return rc;
    if (!ctx->d_out_ribbons) {
        size_t want = (size_t)1 << 22; // 4 Mi ribbons = 128 MiB; grows on demand
        const char* env = getenv("PPE_RIBBON_POOL");
        if (env && atoll(env) > 0) want = (size_t)atoll(env);
        rc = ensure_pool(ctx, want);
        if (rc != PPE_OK) return rc;
    }
    // chunk culling: a probe sample whose cell has every neighbour within `radius` cells free proves that the
    // other samples of its chunk (at most half a chunk of arc length away) are free and in bounds too
    const uint32_t* safe = nullptr;
    if (ctx->map_kind == kMapBitmap) {
        const double reach = 0.5 * true_cost_chunk_samples() * ctx->cfg.collision_checking_increment * 1.001 + 1e-6;
        const double cells = floor(reach / ctx->resolution);
        if (cells <= 62) {
            const int radius = (int)cells + 2; // +1: position inside the cell, +1: rounding of x / res at a cell edge
            if (radius != ctx->safe_radius) {
                const size_t words = (size_t)ctx->rows * ctx->stride_words;
                if (!ctx->d_safe) {
                    PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_safe, words * sizeof(uint32_t)));
                    PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_safe_tmp, words * sizeof(uint32_t)));
                }
                PPE_CUDA(ctx, launch_safe_map(ctx->d_map, ctx->d_safe_tmp, ctx->d_safe, ctx->rows, ctx->cols, ctx->stride_words,
                                              radius, ctx->stream));
                PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                ctx->launches += 2;
                ctx->safe_radius = radius;
            }
            safe = ctx->d_safe;
        }
    }
    memset(w, 0, sizeof *w);
    w->cfg = ctx->cfg;
    w->dt = ctx->cfg.collision_checking_increment / ctx->cfg.max_speed;
    w->horizon_end = ctx->cfg.time_horizon + 1e-12 + ctx->cfg.start_state_time;
    w->map_bits = ctx->d_map;
    w->map_kind = ctx->map_kind;
    w->rows = ctx->rows; w->cols = ctx->cols; w->stride_words = ctx->stride_words;
    w->resolution = ctx->resolution;
    w->safe_bits = safe;
    {
        int e = 0;
        const double m = frexp(ctx->resolution, &e);
        w->res_pow2 = (m == 0.5 && e > -1000 && e < 1000) ? 1 : 0;
        w->inv_resolution = 1.0 / ctx->resolution;
    }
    // N2: shared-memory tile for the thread walker when a window is set and the tile fits
    w->tile_on = 0;
    if (ctx->tile_mode && safe && ctx->win_x1 > ctx->win_x0 && ctx->win_y1 > ctx->win_y0 && ctx->stride_words % 4 == 0) {
        const double res = ctx->resolution;
        long long r0 = (long long)floor(ctx->win_y0 / res) - 1, r1 = (long long)floor(ctx->win_y1 / res) + 1;
        long long c0 = (long long)floor(ctx->win_x0 / res) - 1, c1 = (long long)floor(ctx->win_x1 / res) + 1;
        if (r0 < 0) r0 = 0;
        if (c0 < 0) c0 = 0;
        if (r1 > ctx->rows - 1) r1 = ctx->rows - 1;
        if (c1 > ctx->cols - 1) c1 = ctx->cols - 1;
        if (r1 >= r0 && c1 >= c0) {
            const long long w0 = (c0 >> 5) & ~3ll;                                   // 16-byte aligned rows
            long long w1 = (((c1 >> 5) + 4) & ~3ll);                                 // exclusive, multiple of 4 words
            if (w1 > ctx->stride_words) w1 = ctx->stride_words;
            const long long rows = r1 - r0 + 1, words = w1 - w0;
            if (words > 0 && words % 4 == 0 && 2 * rows * words * 4 <= 96 * 1024) {
                w->tile_on = 1;
                w->tile_r0 = (int)r0; w->tile_w0 = (int)w0; w->tile_rows = (int)rows; w->tile_words = (int)words;
            }
        }
    }
    w->obs_cull_ok = ctx->obs_cull_ok;
    w->obstacles = ctx->d_obs;
    w->obs_kind = ctx->n_obs > 0 ? ctx->obs_kind : kObsNone;
    w->n_obs = ctx->obs_kind == kObsNone ? 0 : ctx->n_obs;
    w->ribbons = ctx->d_ribbons;
    w->boxes = ctx->d_boxes;
    w->set_offset = ctx->d_off;
    w->set_count = ctx->d_cnt;
    w->set_cct = ctx->d_cct;
    w->set_sumlen = ctx->d_sumlen;
    w->set_tame = ctx->d_tame;
    w->n_sets = (int)ctx->h_cnt.size();
    w->ribbon_cap = ribbon_cap_for(ctx->max_set);
    w->out_ribbons = ctx->d_out_ribbons;
    w->out_count = ctx->d_out_count;
    w->out_cap = ctx->out_cap;
    if (true_cost_smem_bytes(w->ribbon_cap, w->n_obs) > 200 * 1024)
        return fail(ctx, PPE_ERR_CAPACITY, "ribbon sets / obstacle list too large for the per-CTA shared-memory working set");
    return PPE_OK;
}

} // namespace

extern "C" {

int ppe_abi_version(void) { return PPE_ABI_VERSION; }
int ppe_abi_sizeof_config(void) { return (int)sizeof(ppe_config); }
int ppe_abi_sizeof_edge(void) { return (int)sizeof(ppe_edge); }
int ppe_abi_sizeof_edge_result(void) { return (int)sizeof(ppe_edge_result); }

int ppe_create(int device, ppe_ctx** out) {
    if (!out) return PPE_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return PPE_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PPE_ERR_NO_DEVICE;
    if (prop.major < 10) return PPE_ERR_NO_DEVICE; // sm_100a SASS only; no other code path is shipped
    if (cudaSetDevice(device) != cudaSuccess) return PPE_ERR_NO_DEVICE;
    ppe_ctx* ctx = new ppe_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PPE_ERR_CUDA; }
    ctx->max_blocks = ctx->sm_count * 32;
    {
        const char* env = getenv("PPE_THREAD_WALKER");
        if (env && env[0] == '0') ctx->thread_walker = false;
        const char* env_tile = getenv("PPE_MAP_TILE");
        ctx->tile_mode = env_tile && env_tile[0] == '1';
        const char* env_l2 = getenv("PPE_MAP_L2_PERSIST");
        ctx->l2_persist = env_l2 && env_l2[0] == '1';
        if (ctx->l2_persist) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 8u << 20);
        const char* env_cp = getenv("PPE_K2T_CPS");   // tuning knob: check-points a K2t thread may walk
        if (env_cp && atoi(env_cp) > 0) ctx->tuning.cp_budget = atoi(env_cp);
        const char* env_d = getenv("PPE_K2T_DIRTY");  // tuning knob: non-clean chunks a K2t thread may evaluate
        if (env_d) ctx->tuning.dirty_budget = atoi(env_d);
        const char* env_k2b = getenv("PPE_K2B_CTAS");  // tuning knob: K2b CTAs per SM
        if (env_k2b) ctx->tuning.k2b_ctas_per_sm = atoi(env_k2b);
        const char* env_late = getenv("PPE_LATE_K2B");
        if (env_late) ctx->late_k2b = atoi(env_late) != 0;
        const char* env_ls = getenv("PPE_LATE_SLICE");
        if (env_ls && atoll(env_ls) >= 1024) { ctx->late_slice = atoll(env_ls); ctx->late_slice_fixed = true; }
        const char* env_k2c = getenv("PPE_DEEP_WALKER");
        if (env_k2c) ctx->tuning.deep_walker = env_k2c[0] == '1';
        ctx->tuning = clamp_tuning(ctx->tuning);
    }
    bool ok = cudaMalloc((void**)&ctx->d_work, 2 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_out_count, sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_block_best, (size_t)ctx->max_blocks * sizeof(BestD)) == cudaSuccess &&
              cudaMalloc((void**)&ctx->d_best, sizeof(BestD)) == cudaSuccess;
    for (auto& ln : ctx->lanes) {
        ok = ok && cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaMalloc((void**)&ln.d_work, 2 * sizeof(unsigned long long)) == cudaSuccess &&
             cudaMalloc((void**)&ln.d_block_best, (size_t)ctx->max_blocks * sizeof(BestD)) == cudaSuccess &&
             cudaEventCreateWithFlags(&ln.ev_k3, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) { ppe_destroy(ctx); return PPE_ERR_CUDA; }
    *out = ctx;
    return PPE_OK;
}

void ppe_destroy(ppe_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_map); cudaFree(ctx->d_safe); cudaFree(ctx->d_safe_tmp); cudaFree(ctx->d_obs);
    cudaFree(ctx->d_ribbons); cudaFree(ctx->d_boxes); cudaFree(ctx->d_off); cudaFree(ctx->d_cnt); cudaFree(ctx->d_cct); cudaFree(ctx->d_sumlen); cudaFree(ctx->d_tame);
    cudaFree(ctx->d_edges); cudaFree(ctx->d_results); cudaFree(ctx->d_prepared); cudaFree(ctx->d_dub); cudaFree(ctx->d_dubi);
    cudaFree(ctx->d_out_ribbons); cudaFree(ctx->d_out_count);
    cudaFree(ctx->d_work); cudaFree(ctx->d_heavy); cudaFree(ctx->d_block_best); cudaFree(ctx->d_best);
    if (ctx->d_patch) cudaFree(ctx->d_patch);
    cudaFree(ctx->d_sx); cudaFree(ctx->d_sy); cudaFree(ctx->d_sh); cudaFree(ctx->d_stage); cudaFree(ctx->d_keep);
    cudaFree(ctx->d_blockcnt); cudaFree(ctx->d_verts); cudaFree(ctx->d_xedges); cudaFree(ctx->d_xresults);
    cudaFree(ctx->d_children); cudaFree(ctx->d_xints);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_pool) cudaFreeHost(ctx->h_pool);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream_in) cudaStreamDestroy(ctx->stream_in);
    if (ctx->stream_out) cudaStreamDestroy(ctx->stream_out);
    for (auto& ln : ctx->lanes) {
        if (ln.stream) { cudaStreamSynchronize(ln.stream); cudaStreamDestroy(ln.stream); }
        cudaFree(ln.d_prepared); cudaFree(ln.d_heavy); cudaFree(ln.d_work); cudaFree(ln.d_block_best);
        if (ln.ev_k3) cudaEventDestroy(ln.ev_k3);
    }
    for (cudaEvent_t e : ctx->ev_k2) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_done) cudaEventDestroy(e);
    delete ctx;
}

const char* ppe_last_error(const ppe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ppe_set_config(ppe_ctx* ctx, const ppe_config* cfg) {
    if (!ctx || !cfg) return PPE_ERR_INVALID;
    if (!(cfg->max_speed > 0) || !(cfg->collision_checking_increment > 0) || !(cfg->ribbon_width > 0) ||
        !(cfg->turning_radius > 0) || !(cfg->coverage_turning_radius > 0))
        return fail(ctx, PPE_ERR_INVALID, "config: speeds, radii, increment and ribbon width must be positive");
    ctx->cfg = *cfg;
    ctx->have_cfg = true;
    return PPE_OK;
}

int ppe_set_map_none(ppe_ctx* ctx) {
    if (!ctx) return PPE_ERR_INVALID;
    ctx->map_kind = kMapNone;
    ctx->map_generation++;
    return PPE_OK;
}

int ppe_set_map_bitmap(ppe_ctx* ctx, const uint8_t* bits, int rows, int cols, int row_stride_bytes, double resolution) {
    if (!ctx || !bits || rows <= 0 || cols <= 0 || row_stride_bytes * 8 < cols || !(resolution > 0))
        return fail(ctx, PPE_ERR_INVALID, "ppe_set_map_bitmap: bad arguments");
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    // repack rows to whole 32-bit words so one lane-load serves 32 cells
    const int words = (cols + 31) / 32;
    std::vector<uint32_t> packed((size_t)rows * words, 0u);
    for (int r = 0; r < rows; r++) {
        const uint8_t* src = bits + (size_t)r * row_stride_bytes;
        uint32_t* dst = packed.data() + (size_t)r * words;
        const int nbytes = (cols + 7) / 8;
        for (int b = 0; b < nbytes; b++) dst[b >> 2] |= (uint32_t)src[b] << (8 * (b & 3));
        // clear padding bits past `cols`
        if (cols & 31) dst[words - 1] &= (1u << (cols & 31)) - 1u;
    }
    if (ctx->d_map) PPE_CUDA(ctx, cudaFree(ctx->d_map));
    ctx->d_map = nullptr;
    if (ctx->d_safe) PPE_CUDA(ctx, cudaFree(ctx->d_safe));
    if (ctx->d_safe_tmp) PPE_CUDA(ctx, cudaFree(ctx->d_safe_tmp));
    ctx->d_safe = ctx->d_safe_tmp = nullptr;
    ctx->safe_radius = -1;
    PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_map, packed.size() * sizeof(uint32_t)));
    PPE_CUDA(ctx, cudaMemcpy(ctx->d_map, packed.data(), packed.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->map_kind = kMapBitmap;
    ctx->rows = rows; ctx->cols = cols; ctx->stride_words = words; ctx->resolution = resolution;
    ctx->map_generation++;
    return PPE_OK;
}

int ppe_set_obstacles_none(ppe_ctx* ctx) {
    if (!ctx) return PPE_ERR_INVALID;
    ctx->obs_kind = kObsNone;
    ctx->n_obs = 0;
    return PPE_OK;
}

static int upload_obstacles(ppe_ctx* ctx, const std::vector<ObstacleD>& h, int kind) {
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->d_obs) PPE_CUDA(ctx, cudaFree(ctx->d_obs));
    ctx->d_obs = nullptr;
    if (!h.empty()) {
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_obs, h.size() * sizeof(ObstacleD)));
        PPE_CUDA(ctx, cudaMemcpy(ctx->d_obs, h.data(), h.size() * sizeof(ObstacleD), cudaMemcpyHostToDevice));
    }
    ctx->obs_kind = kind;
    ctx->n_obs = (int)h.size();
    // the chunk bound needs a 64-bit candidate mask and, for Gaussians, a norm induced by the quadratic form
    ctx->obs_cull_ok = h.size() <= 64;
    for (const ObstacleD& o : h)
        if (!(o.cull >= 0) || !isfinite(o.cull)) ctx->obs_cull_ok = 0;
    return PPE_OK;
}

int ppe_set_obstacles_binary(ppe_ctx* ctx, int n, const double* x, const double* y, const double* yaw,
                             const double* speed, const double* time, const double* width, const double* length) {
    if (!ctx || n < 0 || (n > 0 && (!x || !y || !yaw || !speed || !time || !width || !length)))
        return fail(ctx, PPE_ERR_INVALID, "ppe_set_obstacles_binary: bad arguments");
    std::vector<ObstacleD> h((size_t)n);
    for (int i = 0; i < n; i++) {
        ObstacleD& o = h[i];
        memset(&o, 0, sizeof o);
        o.X = x[i]; o.Y = y[i]; o.Time = time[i]; o.Speed = speed[i];
        o.cosYaw = cos(yaw[i]); o.sinYaw = sin(yaw[i]);
        o.a = (length[i] + 2) / 2; // strict: Length += 2, then `fabs(rotatedX) < Length / 2`
        o.b = (width[i] + 2) / 2;
        o.cull = 0;
    }
    return upload_obstacles(ctx, h, kObsBinary);
}

int ppe_set_obstacles_gaussian(ppe_ctx* ctx, int n, const double* x, const double* y, const double* yaw,
                               const double* speed, const double* time, const double* cov) {
    if (!ctx || n < 0 || (n > 0 && (!x || !y || !yaw || !speed || !time)))
        return fail(ctx, PPE_ERR_INVALID, "ppe_set_obstacles_gaussian: bad arguments");
    std::vector<ObstacleD> h((size_t)n);
    for (int i = 0; i < n; i++) {
        ObstacleD& o = h[i];
        memset(&o, 0, sizeof o);
        o.X = x[i]; o.Y = y[i]; o.Time = time[i]; o.Speed = speed[i];
        o.cosYaw = cos(yaw[i]); o.sinYaw = sin(yaw[i]);
        // default covariance of GaussianDynamicObstaclesManager::Obstacle (Gaussian...h:24-25)
        const double c00 = cov ? cov[4 * i] : 30, c01 = cov ? cov[4 * i + 1] : 10, c10 = cov ? cov[4 * i + 2] : 10,
                     c11 = cov ? cov[4 * i + 3] : 30;
        // Obstacle::pdf (Gaussian...h:39-44): 2x2 inverse and determinant in Eigen's closed form
        const double det = c00 * c11 - c10 * c01;
        const double invdet = 1.0 / det;
        o.a = c11 * invdet;   // i00
        o.b = -c10 * invdet;  // i10
        o.c = -c01 * invdet;  // i01
        o.d = c00 * invdet;   // i11
        const double twoPi = 2 * M_PI;
        o.norm = 1.0 / twoPi / sqrt(det);
        // q(v) = a v0^2 + (b + c) v0 v1 + d v1^2; with S = its symmetric matrix positive definite, sqrt(q) is a norm
        // and |sqrt(q(v + e)) - sqrt(q(v))| <= sqrt(lambda_max(S)) |e|.  Otherwise no bound (cull = -1).
        const double s01 = 0.5 * (o.b + o.c);
        const double tr = o.a + o.d, dt_ = o.a * o.d - s01 * s01;
        if (o.a > 0 && dt_ > 0 && isfinite(o.norm) && o.norm > 0) {
            const double lmax = 0.5 * tr + sqrt(fmax(0.25 * tr * tr - dt_, 0.0));
            o.cull = sqrt(lmax) * (1 + 1e-9);
        } else {
            o.cull = -1;
        }
    }
    return upload_obstacles(ctx, h, kObsGaussian);
}

int ppe_put_ribbon_set(ppe_ctx* ctx, int n, const double* xyxy, double coverage_completed_time, int32_t* set_id) {
    if (!ctx || n < 0 || (n > 0 && !xyxy) || !set_id) return fail(ctx, PPE_ERR_INVALID, "ppe_put_ribbon_set: bad arguments");
    if (!ctx->have_cfg) return fail(ctx, PPE_ERR_STATE, "ppe_set_config must precede ppe_put_ribbon_set (ribbon width)");
    const int off = (int)(ctx->h_ribbons.size() / 4);
    // verbatim, in list order: a child vertex copies its parent's list (Vertex.cpp:24,32).  The
    // covered-filter of RibbonManager::add (RibbonManager.cpp:154-158) is NOT applied here -- a strict
    // cover keeps remainders shorter than 2 * RibbonWidth that add() would drop.
    ctx->h_ribbons.insert(ctx->h_ribbons.end(), xyxy, xyxy + (size_t)4 * n);
    const int kept = n;
    if (ribbon_cap_for(kept) > 512) {
        ctx->h_ribbons.resize((size_t)off * 4);
        return fail(ctx, PPE_ERR_CAPACITY, "ribbon set larger than the per-edge device working set (480 ribbons)");
    }
    ctx->h_off.push_back(off);
    ctx->h_cnt.push_back(kept);
    ctx->h_cct.push_back(coverage_completed_time);
    {
        // invariants of the list as it stands: the same sum RibbonManager::maxDistance accumulates (list order, IEEE sqrt)
        double sum = 0;
        int tame = 1, any_short = 0;
        const double ml = 2 * ctx->cfg.ribbon_width; // Ribbon::minLength(); covered(strict): |e|^2 < ml^2 / (2 * 2), Ribbon.cpp:23-25
        for (int q = 0; q < n; q++) {
            const double* r = xyxy + 4 * (size_t)q;
            const double sq = (r[2] - r[0]) * (r[2] - r[0]) + (r[3] - r[1]) * (r[3] - r[1]);
            sum += sqrt(sq) - 2 * ctx->cfg.ribbon_width;
            if (sq < ml * ml / (2.0 * 2.0)) any_short = 1;
            if (!(fabs(r[0]) < 1e7 && fabs(r[1]) < 1e7 && fabs(r[2]) < 1e7 && fabs(r[3]) < 1e7)) tame = 0;
        }
        if (ctx->sets_width < 0 || ctx->h_sumlen.empty()) ctx->sets_width = ctx->cfg.ribbon_width;
        ctx->h_sumlen.push_back(sum);
        ctx->h_tame.push_back(tame | (any_short ? 2 : 0));
    }
    if (kept > ctx->max_set) ctx->max_set = kept;
    ctx->sets_dirty = true;
    *set_id = (int32_t)ctx->h_cnt.size() - 1;
    return PPE_OK;
}

int ppe_clear_ribbon_sets(ppe_ctx* ctx) {
    if (!ctx) return PPE_ERR_INVALID;
    ctx->h_ribbons.clear(); ctx->h_off.clear(); ctx->h_cnt.clear(); ctx->h_cct.clear(); ctx->h_sumlen.clear(); ctx->h_tame.clear();
    ctx->max_set = 0;
    ctx->uploaded_ribbons = ctx->uploaded_sets = 0;
    ctx->uploaded_boxes = 0;
    ctx->sets_dirty = true;
    ctx->have_batch = false;
    return PPE_OK;
}

int ppe_set_map_window(ppe_ctx* ctx, double x0, double y0, double x1, double y1) {
    if (!ctx) return PPE_ERR_INVALID;
    ctx->win_x0 = x0; ctx->win_y0 = y0; ctx->win_x1 = x1; ctx->win_y1 = y1;
    return PPE_OK;
}

// N2 A/B: persisting L2 window over the occupancy bitmap on `stream` (the safe map is allocated right after it is built;
// one window per stream, so the occupancy map -- read per sample in non-clean chunks -- gets it)
static void apply_l2_window(ppe_ctx* ctx, cudaStream_t stream) {
    if (!ctx->l2_persist || !ctx->d_map) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    v.accessPolicyWindow.base_ptr = (void*)ctx->d_map;
    v.accessPolicyWindow.num_bytes = (size_t)ctx->rows * ctx->stride_words * sizeof(uint32_t);
    v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v);
}

// ---- K1 ------------------------------------------------------------------------------------------
int ppe_dubins_batch_device(ppe_ctx* ctx, int64_t n, const double* d_q0, const double* d_q1, const double* d_rho,
                            int32_t* d_type, double* d_param, double* d_length, int32_t* d_err, void* stream) {
    if (!ctx || n < 0) return PPE_ERR_INVALID;
    if (n == 0) return PPE_OK;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    PPE_CUDA(ctx, launch_dubins_batch(n, d_q0, d_q1, d_rho, d_type, d_param, d_length, d_err, (cudaStream_t)stream));
    ctx->launches += 1;
    return PPE_OK;
}

int ppe_dubins_batch(ppe_ctx* ctx, int64_t n, const double* q0, const double* q1, const double* rho, int32_t* type,
                     double* param, double* length, int32_t* err) {
    if (!ctx || n < 0 || (n > 0 && (!q0 || !q1 || !rho || !type || !param || !length || !err)))
        return fail(ctx, PPE_ERR_INVALID, "ppe_dubins_batch: bad arguments");
    if (n == 0) return PPE_OK;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    // layout: q0[3n] q1[3n] rho[n] param[3n] length[n]
    int rc = grow(ctx, &ctx->d_dub, &ctx->cap_dub, (size_t)n * 11);
    if (rc != PPE_OK) return rc;
    rc = grow(ctx, &ctx->d_dubi, &ctx->cap_dubi, (size_t)n * 2);
    if (rc != PPE_OK) return rc;
    double* dq0 = ctx->d_dub;
    double* dq1 = dq0 + 3 * n;
    double* drho = dq1 + 3 * n;
    double* dpar = drho + n;
    double* dlen = dpar + 3 * n;
    int32_t* dtype = ctx->d_dubi;
    int32_t* derr = dtype + n;
    cudaStream_t s = ctx->stream;
    PPE_CUDA(ctx, cudaMemcpyAsync(dq0, q0, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    PPE_CUDA(ctx, cudaMemcpyAsync(dq1, q1, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    PPE_CUDA(ctx, cudaMemcpyAsync(drho, rho, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    rc = ppe_dubins_batch_device(ctx, n, dq0, dq1, drho, dtype, dpar, dlen, derr, s);
    if (rc != PPE_OK) return rc;
    PPE_CUDA(ctx, cudaMemcpyAsync(param, dpar, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    PPE_CUDA(ctx, cudaMemcpyAsync(length, dlen, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    PPE_CUDA(ctx, cudaMemcpyAsync(type, dtype, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    PPE_CUDA(ctx, cudaMemcpyAsync(err, derr, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    PPE_CUDA(ctx, cudaStreamSynchronize(s));
    return PPE_OK;
}

// ---- K2 / K3 ----------------------------------------------------------------------------------------
// one launch group (K2a, K2b, K3) over the edges of a device-resident batch, all on `stream`
static int run_batch_device(ppe_ctx* ctx, const WorldD& w, int64_t n, const ppe_edge* d_edges, ppe_edge_result* d_results,
                            cudaStream_t stream) {
    int rc = grow(ctx, &ctx->d_prepared, &ctx->cap_prepared, (size_t)n * prepared_edge_bytes());
    if (rc != PPE_OK) return rc;
    if (ctx->thread_walker) {
        rc = grow(ctx, &ctx->d_heavy, &ctx->cap_heavy, 2 * (size_t)n);
        if (rc != PPE_OK) return rc;
    }
    int blocks = 1, launches = 0;
    PPE_CUDA(ctx, launch_true_cost_kernels(w, n, d_edges, ctx->d_prepared, d_results, ctx->d_work,
                                           ctx->thread_walker ? ctx->d_heavy : nullptr, ctx->d_block_best, ctx->max_blocks,
                                           ctx->sm_count, stream, true, ctx->tuning, &blocks, &launches));
    PPE_CUDA(ctx, launch_best_final(ctx->d_block_best, blocks, ctx->d_best, 0, false, stream));
    ctx->launches += launches + 1;
    return PPE_OK;
}

int ppe_true_cost_batch_device(ppe_ctx* ctx, int64_t n, const ppe_edge* d_edges, ppe_edge_result* d_results, void* stream) {
    if (!ctx || n < 0) return PPE_ERR_INVALID;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    WorldD w;
    int rc = make_world(ctx, &w);
    if (rc != PPE_OK) return rc;
    ctx->have_batch = false;
    ctx->out_downloaded = false;
    if (n == 0) return PPE_OK;
    apply_l2_window(ctx, (cudaStream_t)stream);
    return run_batch_device(ctx, w, n, d_edges, d_results, (cudaStream_t)stream);
}

int ppe_best_device(ppe_ctx* ctx, double* f, int64_t* edge_index, void* stream) {
    if (!ctx || !f || !edge_index) return PPE_ERR_INVALID;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->cfg.heuristic != PPE_H_MAX_DISTANCE && !tsp_on_device(ctx->cfg.heuristic))
        return fail(ctx, PPE_ERR_STATE, "ppe_best*: h is not evaluated on the device for this heuristic (f = g + h unknown)");
    BestD b;
    PPE_CUDA(ctx, cudaMemcpyAsync(&b, ctx->d_best, sizeof b, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PPE_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    *f = b.idx >= 0 ? b.f : INFINITY;
    *edge_index = b.idx;
    return PPE_OK;
}

int ppe_best_copy_device(ppe_ctx* ctx, void* d_dst16, int64_t index_base, void* stream) {
    if (!ctx || !d_dst16) return PPE_ERR_INVALID;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->cfg.heuristic != PPE_H_MAX_DISTANCE && !tsp_on_device(ctx->cfg.heuristic))
        return fail(ctx, PPE_ERR_STATE, "ppe_best*: h is not evaluated on the device for this heuristic (f = g + h unknown)");
    PPE_CUDA(ctx, launch_best_export(ctx->d_best, reinterpret_cast<BestD*>(d_dst16), index_base, (cudaStream_t)stream));
    ctx->launches += 1;
    return PPE_OK;
}

// Large host-buffer batches when K2t is on.  The copies of a 2^20-edge batch (176 B in, 208 B out per edge) take about as
// long as its kernels, and slicing the WHOLE kernel sequence makes every slice end on its own tail of survey-line edges.
// So only the uniform part is sliced: K2a + K2t of slice k run while slice k + 1 arrives and the results of slice k - 1
// leave -- 95 % of the records are final after K2t.  K2b then runs ONCE over the heavy list of the whole batch, and the
// records it wrote (5 %) are scattered by a kernel straight into the caller's array when that is mapped pinned memory, or
// packed, copied and scattered by the host when it is pageable.
static int grow_pinned(ppe_ctx* ctx, void** p, size_t* cap, size_t need);
static int batch_pipelined(ppe_ctx* ctx, WorldD& w, int64_t n, const ppe_edge* edges, ppe_edge_result* results, int64_t slice) {
    const int64_t n_slices = (n + slice - 1) / slice;
    int rc = grow(ctx, &ctx->d_prepared, &ctx->cap_prepared, (size_t)n * prepared_edge_bytes());
    if (rc != PPE_OK) return rc;
    rc = grow(ctx, &ctx->d_heavy, &ctx->cap_heavy, 2 * (size_t)n);
    if (rc != PPE_OK) return rc;
    // can a kernel write into `results`?
    ppe_edge_result* mapped = nullptr;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, results) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            mapped = reinterpret_cast<ppe_edge_result*>(attr.devicePointer);
        else
            (void)cudaGetLastError();
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        PPE_CUDA(ctx, cudaMemsetAsync(ctx->d_out_count, 0, sizeof(unsigned long long), ctx->stream_in));
        PPE_CUDA(ctx, cudaMemsetAsync(ctx->d_work, 0, 2 * sizeof(unsigned long long), ctx->stream_in));
        for (int64_t k = 0; k < n_slices; k++) {
            const int64_t lo = k * slice, cnt = (lo + slice <= n) ? slice : n - lo;
            ppe_ctx::Lane& ln = ctx->lanes[k & 1];
            PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_edges + lo, edges + lo, (size_t)cnt * sizeof(ppe_edge), cudaMemcpyHostToDevice, ctx->stream_in));
            PPE_CUDA(ctx, cudaEventRecord(ctx->ev_in[k], ctx->stream_in));
            PPE_CUDA(ctx, cudaStreamWaitEvent(ln.stream, ctx->ev_in[k], 0));
            int launches = 0;
            PPE_CUDA(ctx, launch_prepare_and_walk(w, n, lo, cnt, ctx->d_edges, ctx->d_prepared, ctx->d_results, ctx->d_work, ctx->d_heavy,
                                                  ln.stream, ctx->tuning, &launches));
            PPE_CUDA(ctx, cudaEventRecord(ctx->ev_k2[k], ln.stream));
            ctx->launches += launches;
            PPE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_out, ctx->ev_k2[k], 0));
            PPE_CUDA(ctx, cudaMemcpyAsync(results + lo, ctx->d_results + lo, (size_t)cnt * sizeof(ppe_edge_result), cudaMemcpyDeviceToHost, ctx->stream_out));
            PPE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_k2[k], 0));
        }
        PPE_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], ctx->stream_out));
        int blocks = 1, launches = 0;
        PPE_CUDA(ctx, launch_heavy_and_best(w, n, ctx->d_edges, ctx->d_prepared, ctx->d_results, ctx->d_work, ctx->d_heavy, ctx->d_block_best,
                                            ctx->max_blocks, ctx->sm_count, ctx->stream, ctx->tuning, &blocks, &launches));
        PPE_CUDA(ctx, launch_best_final(ctx->d_block_best, blocks, ctx->d_best, 0, false, ctx->stream));
        ctx->launches += launches + 1;
        PPE_CUDA(ctx, cudaMemcpyAsync(&ctx->last_out_count, ctx->d_out_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        // K2b's records overwrite what the slice copies brought: after the last of those copies
        PPE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_done[0], 0));
        if (mapped) {
            PPE_CUDA(ctx, launch_patch_results(ctx->d_results, ctx->d_heavy, ctx->d_work, n, mapped, nullptr, ctx->sm_count, ctx->stream));
            ctx->launches += 1;
            PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            unsigned int counts[2] = {0, 0};
            PPE_CUDA(ctx, cudaMemcpyAsync(counts, ctx->d_work + 1, sizeof counts, cudaMemcpyDeviceToHost, ctx->stream));
            PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const size_t m = (size_t)counts[0] + counts[1];
            if (m > 0) {
                const size_t rec_bytes = m * sizeof(ppe_edge_result), bytes = rec_bytes + m * sizeof(unsigned int);
                rc = grow(ctx, &ctx->d_patch, &ctx->cap_patch, bytes);
                if (rc != PPE_OK) return rc;
                rc = grow_pinned(ctx, &ctx->h_pinned, &ctx->cap_pinned, bytes);
                if (rc != PPE_OK) return rc;
                unsigned int* d_idx = reinterpret_cast<unsigned int*>(ctx->d_patch + rec_bytes);
                PPE_CUDA(ctx, launch_patch_results(ctx->d_results, ctx->d_heavy, ctx->d_work, n, reinterpret_cast<ppe_edge_result*>(ctx->d_patch),
                                                   d_idx, ctx->sm_count, ctx->stream));
                ctx->launches += 1;
                PPE_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_patch, bytes, cudaMemcpyDeviceToHost, ctx->stream));
                PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                const ppe_edge_result* rec = reinterpret_cast<const ppe_edge_result*>(ctx->h_pinned);
                const unsigned int* idx = reinterpret_cast<const unsigned int*>(static_cast<const unsigned char*>(ctx->h_pinned) + rec_bytes);
                for (size_t i = 0; i < m; i++) results[idx[i]] = rec[i];
            }
        }
        PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream_out));
        if (ctx->last_out_count <= ctx->out_cap) break;
        // the ribbons-after pool was too small for this batch: grow it and run the batch again
        rc = ensure_pool(ctx, (size_t)(ctx->last_out_count + ctx->last_out_count / 4 + 1024));
        if (rc != PPE_OK) return rc;
        rc = make_world(ctx, &w);
        if (rc != PPE_OK) return rc;
    }
    return PPE_OK;
}

int ppe_true_cost_batch(ppe_ctx* ctx, int64_t n, const ppe_edge* edges, ppe_edge_result* results) {
    if (!ctx || n < 0 || (n > 0 && (!edges || !results))) return fail(ctx, PPE_ERR_INVALID, "ppe_true_cost_batch: bad arguments");
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->have_batch = false;
    if (n == 0) return PPE_OK;
    {
        int rc = grow(ctx, &ctx->d_edges, &ctx->cap_edges, (size_t)n);
        if (rc != PPE_OK) return rc;
        rc = grow(ctx, &ctx->d_results, &ctx->cap_results, (size_t)n);
        if (rc != PPE_OK) return rc;
    }
    WorldD w;
    int rc = make_world(ctx, &w);
    if (rc != PPE_OK) return rc;
    ctx->out_downloaded = false;
    // slice size of the late-K2b pipeline: an eighth of the batch (so that a shard of a batch split over several GPUs is
    // pipelined like the whole batch is on one), between 16 Ki and 128 Ki edges; PPE_LATE_SLICE fixes it
    int64_t late_slice = ctx->late_slice;
    if (!ctx->late_slice_fixed) {
        late_slice = ((n / 8 + 4095) / 4096) * 4096;
        if (late_slice < 16384) late_slice = 16384;
        if (late_slice > 131072) late_slice = 131072;
    }
    if (ctx->thread_walker && !ctx->tuning.deep_walker && ctx->late_k2b && n >= 4 * late_slice) {
        const int64_t n_slices = (n + late_slice - 1) / late_slice;
        while ((int64_t)ctx->ev_in.size() < n_slices) {
            cudaEvent_t a, b, c;
            PPE_CUDA(ctx, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            PPE_CUDA(ctx, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            PPE_CUDA(ctx, cudaEventCreateWithFlags(&c, cudaEventDisableTiming));
            ctx->ev_in.push_back(a);
            ctx->ev_k2.push_back(b);
            ctx->ev_done.push_back(c);
        }
        rc = batch_pipelined(ctx, w, n, edges, results, late_slice);
        if (rc != PPE_OK) return rc;
        ctx->last_count = n;
        ctx->have_batch = true;
        ctx->last_was_expand = false;
        ctx->out_downloaded = false;
        return PPE_OK;
    }
    // Slices of the batch flow through the streams: H2D of slice k + 1 (copy engine), the kernels of slices k and
    // k - 1 (two lanes) and D2H of finished slices (second copy engine) run concurrently; K3 accumulates the best
    // record slice by slice on the context stream.  Small batches are one slice.
    int64_t slice_edges = 262144;
    {
        const char* env = getenv("PPE_SLICE_EDGES");
        if (env && atoll(env) >= 1024) slice_edges = atoll(env);
    }
    const int64_t slice = n <= slice_edges + slice_edges / 2 ? n : slice_edges;
    const int64_t n_slices = (n + slice - 1) / slice;
    while ((int64_t)ctx->ev_in.size() < n_slices) {
        cudaEvent_t a, b, c;
        PPE_CUDA(ctx, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        PPE_CUDA(ctx, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        PPE_CUDA(ctx, cudaEventCreateWithFlags(&c, cudaEventDisableTiming));
        ctx->ev_in.push_back(a);
        ctx->ev_k2.push_back(b);
        ctx->ev_done.push_back(c);
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        for (auto& ln : ctx->lanes) ln.used = false;
        // the ribbons-after pool counter runs over all slices: reset once, ahead of slice 0's copy
        PPE_CUDA(ctx, cudaMemsetAsync(ctx->d_out_count, 0, sizeof(unsigned long long), ctx->stream_in));
        for (int64_t k = 0; k < n_slices; k++) {
            const int64_t lo = k * slice, cnt = (lo + slice <= n) ? slice : n - lo;
            ppe_ctx::Lane& ln = ctx->lanes[k & 1];
            rc = grow(ctx, &ln.d_prepared, &ln.cap_prepared, (size_t)cnt * prepared_edge_bytes());
            if (rc != PPE_OK) return rc;
            if (ctx->thread_walker) {
                rc = grow(ctx, &ln.d_heavy, &ln.cap_heavy, 2 * (size_t)cnt);
                if (rc != PPE_OK) return rc;
            }
            PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_edges + lo, edges + lo, (size_t)cnt * sizeof(ppe_edge), cudaMemcpyHostToDevice, ctx->stream_in));
            PPE_CUDA(ctx, cudaEventRecord(ctx->ev_in[k], ctx->stream_in));
            PPE_CUDA(ctx, cudaStreamWaitEvent(ln.stream, ctx->ev_in[k], 0));
            if (ln.used) PPE_CUDA(ctx, cudaStreamWaitEvent(ln.stream, ln.ev_k3, 0)); // its block_best has been consumed
            int blocks = 1, launches = 0;
            PPE_CUDA(ctx, launch_true_cost_kernels(w, cnt, ctx->d_edges + lo, ln.d_prepared, ctx->d_results + lo, ln.d_work,
                                                   ctx->thread_walker ? ln.d_heavy : nullptr, ln.d_block_best, ctx->max_blocks,
                                                   ctx->sm_count, ln.stream, false, ctx->tuning, &blocks, &launches));
            PPE_CUDA(ctx, cudaEventRecord(ctx->ev_k2[k], ln.stream));
            PPE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_k2[k], 0));
            PPE_CUDA(ctx, launch_best_final(ln.d_block_best, blocks, ctx->d_best, lo, k > 0, ctx->stream));
            PPE_CUDA(ctx, cudaEventRecord(ln.ev_k3, ctx->stream));
            ln.used = true;
            ctx->launches += launches + 1;
            PPE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_out, ctx->ev_k2[k], 0));
            PPE_CUDA(ctx, cudaMemcpyAsync(results + lo, ctx->d_results + lo, (size_t)cnt * sizeof(ppe_edge_result), cudaMemcpyDeviceToHost, ctx->stream_out));
        }
        PPE_CUDA(ctx, cudaMemcpyAsync(&ctx->last_out_count, ctx->d_out_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PPE_CUDA(ctx, cudaStreamSynchronize(ctx->stream_out));
        if (ctx->last_out_count <= ctx->out_cap) break;
        // the ribbons-after pool was too small for this batch: grow it and run the batch again
        rc = ensure_pool(ctx, (size_t)(ctx->last_out_count + ctx->last_out_count / 4 + 1024));
        if (rc != PPE_OK) return rc;
        rc = make_world(ctx, &w);
        if (rc != PPE_OK) return rc;
    }
    ctx->last_count = n;
    ctx->have_batch = true;
    ctx->last_was_expand = false;
    ctx->out_downloaded = false;
    return PPE_OK;
}

int ppe_get_ribbons_after(ppe_ctx* ctx, int64_t i, double* xyxy, int cap) {
    if (!ctx || !ctx->have_batch || i < 0 || i >= ctx->last_count || (cap > 0 && !xyxy))
        return fail(ctx, PPE_ERR_INVALID, "ppe_get_ribbons_after: no such edge in the last ppe_true_cost_batch");
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    // the batch's device copies are still in place: fetch this edge's record (and its parent set id) on demand
    ppe_edge_result r;
    PPE_CUDA(ctx, cudaMemcpy(&r, ctx->d_results + i, sizeof r, cudaMemcpyDeviceToHost));
    if (!r.ribbons_changed) {
        // identical to the parent's interned set
        int32_t set = -1;
        PPE_CUDA(ctx, cudaMemcpy(&set, (const char*)(ctx->d_edges + i) + offsetof(ppe_edge, ribbon_set), sizeof set, cudaMemcpyDeviceToHost));
        if (set < 0 || (size_t)set >= ctx->h_cnt.size()) return fail(ctx, PPE_ERR_INVALID, "unknown ribbon set");
        const int n = ctx->h_cnt[set];
        const double* src = ctx->h_ribbons.data() + (size_t)ctx->h_off[set] * 4;
        for (int k = 0; k < n && k < cap; k++) memcpy(xyxy + 4 * k, src + 4 * k, 4 * sizeof(double));
        return n;
    }
    if (r.ribbons_offset < 0) return fail(ctx, PPE_ERR_CAPACITY, "ribbons-after of this edge were not materialised");
    const int n = r.n_ribbons_after;
    const int m = n < cap ? n : cap;
    if (m > 0) PPE_CUDA(ctx, cudaMemcpy(xyxy, ctx->d_out_ribbons + r.ribbons_offset, (size_t)m * sizeof(double4), cudaMemcpyDeviceToHost));
    return n;
}

int ppe_best(ppe_ctx* ctx, double* f, int64_t* edge_index) {
    if (!ctx || !ctx->have_batch) return fail(ctx, PPE_ERR_STATE, "ppe_best: no batch has been evaluated");
    return ppe_best_device(ctx, f, edge_index, ctx->stream);
}

// ---- frontier expansion ------------------------------------------------------------------------------
int ppe_clear_samples(ppe_ctx* ctx) {
    if (!ctx) return PPE_ERR_INVALID;
    ctx->n_samples = 0;
    return PPE_OK;
}

int64_t ppe_sample_count(const ppe_ctx* ctx) { return ctx ? ctx->n_samples : 0; }

static int grow_pinned(ppe_ctx* ctx, void** p, size_t* cap, size_t need) {
    if (need <= *cap) return PPE_OK;
    size_t c = *cap ? *cap : (1 << 16);
    while (c < need) c *= 2;
    if (*p) PPE_CUDA(ctx, cudaFreeHost(*p));
    *p = nullptr;
    *cap = 0;
    PPE_CUDA(ctx, cudaMallocHost(p, c));
    *cap = c;
    return PPE_OK;
}

static int64_t add_samples_chunk(ppe_ctx* ctx, int64_t n, const double* x, const double* y, const double* heading, uint8_t* keep);

int64_t ppe_add_samples(ppe_ctx* ctx, int64_t n, const double* x, const double* y, const double* heading, uint8_t* keep) {
    if (!ctx || n < 0 || (n > 0 && (!x || !y || !heading || !keep))) return fail(ctx, PPE_ERR_INVALID, "ppe_add_samples: bad arguments");
    if ((int64_t)ctx->n_samples + n > ((int64_t)1 << 31) - 1) return fail(ctx, PPE_ERR_CAPACITY, "ppe_add_samples: more than 2^31 resident samples");
    int64_t kept = 0;
    const int64_t chunk = (int64_t)1 << 22; // staging buffers stay bounded however far the anytime loop doubles the sample set
    for (int64_t lo = 0; lo < n; lo += chunk) {
        const int64_t m = n - lo < chunk ? n - lo : chunk;
        const int64_t k = add_samples_chunk(ctx, m, x + lo, y + lo, heading + lo, keep + lo);
        if (k < 0) return k;
        kept += k;
    }
    return kept;
}

static int64_t add_samples_chunk(ppe_ctx* ctx, int64_t n, const double* x, const double* y, const double* heading, uint8_t* keep) {
    if (n == 0) return 0;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    WorldD w;
    int rc = make_world(ctx, &w);
    if (rc != PPE_OK) return rc;
    cudaStream_t st = ctx->stream;
    const int blocks = (int)((n + 255) / 256);
    if ((size_t)n > ctx->cap_stage) {
        size_t c = ctx->cap_stage ? ctx->cap_stage : 4096;
        while (c < (size_t)n) c *= 2;
        cudaFree(ctx->d_stage); cudaFree(ctx->d_keep); cudaFree(ctx->d_blockcnt);
        ctx->d_stage = nullptr; ctx->d_keep = nullptr; ctx->d_blockcnt = nullptr; ctx->cap_stage = 0;
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_stage, 3 * c * sizeof(double)));
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_keep, c));
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_blockcnt, ((c + 255) / 256 + 1) * sizeof(unsigned int)));
        ctx->cap_stage = c;
    }
    // room for all n: the resident arrays keep what they hold (device-to-device) when they grow
    {
        const size_t need = (size_t)ctx->n_samples + (size_t)n;
        size_t c1 = ctx->cap_samples, c2 = ctx->cap_samples, c3 = ctx->cap_samples;
        rc = grow_keep(ctx, &ctx->d_sx, &c1, need, (size_t)ctx->n_samples, 1 << 14);
        if (rc == PPE_OK) rc = grow_keep(ctx, &ctx->d_sy, &c2, need, (size_t)ctx->n_samples, 1 << 14);
        if (rc == PPE_OK) rc = grow_keep(ctx, &ctx->d_sh, &c3, need, (size_t)ctx->n_samples, 1 << 14);
        if (rc != PPE_OK) return rc;
        ctx->cap_samples = c1;
    }
    double* dx = ctx->d_stage;
    double* dy = dx + ctx->cap_stage;
    double* dh = dy + ctx->cap_stage;
    PPE_CUDA(ctx, cudaMemcpyAsync(dx, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    PPE_CUDA(ctx, cudaMemcpyAsync(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    PPE_CUDA(ctx, cudaMemcpyAsync(dh, heading, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    unsigned int* d_total = ctx->d_blockcnt + blocks;
    PPE_CUDA(ctx, launch_sample_filter(w, n, dx, dy, dh, ctx->d_keep, ctx->d_blockcnt, d_total, nullptr, nullptr, nullptr, 0, st, 0));
    PPE_CUDA(ctx, launch_sample_filter(w, n, dx, dy, dh, ctx->d_keep, ctx->d_blockcnt, d_total, ctx->d_sx, ctx->d_sy, ctx->d_sh,
                                       ctx->n_samples, st, 1));
    unsigned int total = 0;
    PPE_CUDA(ctx, cudaMemcpyAsync(keep, ctx->d_keep, (size_t)n, cudaMemcpyDeviceToHost, st));
    PPE_CUDA(ctx, cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, st));
    PPE_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 3;
    ctx->n_samples += (int64_t)total;
    return (int64_t)total;
}

int ppe_expand_stride(const ppe_ctx* ctx) {
    if (!ctx || !ctx->have_cfg) return 0;
    return 4 + 4 * ctx->cfg.branching_factor;
}

int ppe_expand_batch(ppe_ctx* ctx, int n, const ppe_vertex* vertices, int32_t* n_children, ppe_child* children,
                     int32_t* flags, int32_t* n_popped) {
    if (!ctx || n < 0 || (n > 0 && (!vertices || !n_children || !children || !flags || !n_popped)))
        return fail(ctx, PPE_ERR_INVALID, "ppe_expand_batch: bad arguments");
    if (n == 0) return PPE_OK;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->have_cfg) return fail(ctx, PPE_ERR_STATE, "ppe_set_config must be called before a batch");
    const int k = ctx->cfg.branching_factor;
    if (k < 1 || k > expand_max_branch()) return fail(ctx, PPE_ERR_CAPACITY, "ppe_expand_batch: branching factor outside [1, 16]");
    if (ctx->tile_mode) { // every edge of the batch stays within max_speed * horizon of its vertex
        double x0 = vertices[0].state[0], x1 = x0, y0 = vertices[0].state[1], y1 = y0;
        for (int v = 1; v < n; v++) {
            x0 = fmin(x0, vertices[v].state[0]); x1 = fmax(x1, vertices[v].state[0]);
            y0 = fmin(y0, vertices[v].state[1]); y1 = fmax(y1, vertices[v].state[1]);
        }
        const double reach = ctx->cfg.max_speed * ctx->cfg.time_horizon + 2.0;
        ppe_set_map_window(ctx, x0 - reach, y0 - reach, x1 + reach, y1 + reach);
    }
    WorldD w;
    int rc = make_world(ctx, &w);
    if (rc != PPE_OK) return rc;
    const int stride = ppe_expand_stride(ctx);
    const size_t slots = (size_t)n * stride;
    if ((size_t)n > ctx->cap_verts) {
        size_t c = ctx->cap_verts ? ctx->cap_verts : 64;
        while (c < (size_t)n) c *= 2;
        cudaFree(ctx->d_verts); cudaFree(ctx->d_xedges); cudaFree(ctx->d_xresults); cudaFree(ctx->d_children); cudaFree(ctx->d_xints);
        ctx->d_verts = nullptr; ctx->d_xedges = nullptr; ctx->d_xresults = nullptr; ctx->d_children = nullptr; ctx->d_xints = nullptr;
        ctx->cap_verts = 0;
        const size_t cs = c * (size_t)(4 + 4 * expand_max_branch());
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_verts, c * sizeof(ppe_vertex)));
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_xedges, cs * sizeof(ppe_edge)));
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_xresults, cs * sizeof(ppe_edge_result)));
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_children, cs * sizeof(ppe_child)));
        PPE_CUDA(ctx, cudaMalloc((void**)&ctx->d_xints, (cs + 4 * c) * sizeof(int32_t)));
        ctx->cap_verts = c;
    }
    rc = grow(ctx, &ctx->d_prepared, &ctx->cap_prepared, slots * prepared_edge_bytes());
    if (rc != PPE_OK) return rc;
    if (ctx->thread_walker) {
        rc = grow(ctx, &ctx->d_heavy, &ctx->cap_heavy, 2 * slots);
        if (rc != PPE_OK) return rc;
    }
    // pinned staging: [vertices][children][4 int arrays]
    const size_t bytes_v = (size_t)n * sizeof(ppe_vertex), bytes_c = slots * sizeof(ppe_child), bytes_i = 4 * (size_t)n * sizeof(int32_t);
    rc = grow_pinned(ctx, &ctx->h_pinned, &ctx->cap_pinned, bytes_v + bytes_c + bytes_i + 64);
    if (rc != PPE_OK) return rc;
    char* hp = (char*)ctx->h_pinned;
    ppe_vertex* h_v = (ppe_vertex*)hp;
    ppe_child* h_c = (ppe_child*)(hp + bytes_v);
    int32_t* h_i = (int32_t*)(hp + bytes_v + bytes_c);
    memcpy(h_v, vertices, bytes_v);

    int32_t* d_edge_sample = ctx->d_xints;
    int32_t* d_nchild = ctx->d_xints + slots;
    int32_t* d_flags = d_nchild + n;
    int32_t* d_pops = d_flags + n;
    int32_t* d_solved = d_pops + n;

    ExpandParamsD p;
    memset(&p, 0, sizeof p);
    p.sx = ctx->d_sx; p.sy = ctx->d_sy; p.sh = ctx->d_sh;
    p.n_samples = (int)ctx->n_samples;
    p.k = k;
    p.stride = stride;
    p.inc = ctx->cfg.collision_checking_increment;
    p.max_speed = ctx->cfg.max_speed;
    p.time_factor = ctx->cfg.time_penalty_factor;
    p.speed[0] = ctx->cfg.max_speed;
    p.speed[1] = ctx->cfg.max_speed == ctx->cfg.slow_speed ? -1 : ctx->cfg.slow_speed;
    p.rho[0] = ctx->cfg.turning_radius;
    p.rho[1] = ctx->cfg.coverage_turning_radius == ctx->cfg.turning_radius ? -1 : ctx->cfg.coverage_turning_radius;
    p.coverage_rho = ctx->cfg.coverage_turning_radius;
    {
        // first search disc: about 24 k candidates' worth of the sampling square (2 v T)^2 -- AStarPlanner.cpp:26-32;
        // the kernel widens / shrinks it on its own when the guess is off
        const double side = 2.0 * ctx->cfg.max_speed * ctx->cfg.time_horizon;
        const double density = (double)(ctx->n_samples > 0 ? ctx->n_samples : 1) / (side * side);
        const char* env = getenv("PPE_EXPAND_R2"); // testing only: force the retry paths
        p.r2_init = env ? atof(env) : (24.0 * k) / (3.141592653589793 * density);
    }
    cudaStream_t st = ctx->stream;
    PPE_CUDA(ctx, cudaMemcpyAsync(ctx->d_verts, h_v, bytes_v, cudaMemcpyHostToDevice, st));
    PPE_CUDA(ctx, launch_expand_select(p, n, ctx->d_verts, ctx->d_xedges, d_edge_sample, d_nchild, d_flags, d_pops, d_solved, st));
    for (int attempt = 0; attempt < 2; attempt++) {
        int blocks = 1, launches = 0;
        PPE_CUDA(ctx, launch_true_cost_kernels(w, (int64_t)slots, ctx->d_xedges, ctx->d_prepared, ctx->d_xresults, ctx->d_work,
                                               ctx->thread_walker ? ctx->d_heavy : nullptr, ctx->d_block_best, ctx->max_blocks,
                                               ctx->sm_count, st, true, ctx->tuning, &blocks, &launches));
        PPE_CUDA(ctx, launch_best_final(ctx->d_block_best, blocks, ctx->d_best, 0, false, st));
        ctx->launches += launches + 1;
        PPE_CUDA(ctx, cudaMemcpyAsync(&ctx->last_out_count, ctx->d_out_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        if (attempt == 0) {
            PPE_CUDA(ctx, launch_expand_pack((int64_t)slots, ctx->d_xedges, ctx->d_xresults, d_edge_sample, ctx->d_children, st));
            PPE_CUDA(ctx, cudaMemcpyAsync(h_c, ctx->d_children, bytes_c, cudaMemcpyDeviceToHost, st));
            PPE_CUDA(ctx, cudaMemcpyAsync(h_i, d_nchild, bytes_i, cudaMemcpyDeviceToHost, st));
        }
        PPE_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->last_out_count <= ctx->out_cap) {
            if (attempt == 1) { // the pool had to grow: pack again from the second evaluation
                PPE_CUDA(ctx, launch_expand_pack((int64_t)slots, ctx->d_xedges, ctx->d_xresults, d_edge_sample, ctx->d_children, st));
                PPE_CUDA(ctx, cudaMemcpyAsync(h_c, ctx->d_children, bytes_c, cudaMemcpyDeviceToHost, st));
                PPE_CUDA(ctx, cudaStreamSynchronize(st));
            }
            break;
        }
        rc = ensure_pool(ctx, (size_t)(ctx->last_out_count + ctx->last_out_count / 4 + 1024));
        if (rc != PPE_OK) return rc;
        rc = make_world(ctx, &w);
        if (rc != PPE_OK) return rc;
    }
    ctx->launches += 2;
    // the ribbons-after pool: one copy for the whole batch
    ctx->n_pool = (int64_t)ctx->last_out_count;
    if (ctx->n_pool > 0) {
        void* pp = ctx->h_pool;
        rc = grow_pinned(ctx, &pp, &ctx->cap_pool, (size_t)ctx->n_pool * sizeof(double4));
        ctx->h_pool = (double*)pp;
        if (rc != PPE_OK) return rc;
        PPE_CUDA(ctx, cudaMemcpyAsync(ctx->h_pool, ctx->d_out_ribbons, (size_t)ctx->n_pool * sizeof(double4), cudaMemcpyDeviceToHost, st));
        PPE_CUDA(ctx, cudaStreamSynchronize(st));
    }
    memcpy(children, h_c, bytes_c);
    memcpy(n_children, h_i, (size_t)n * sizeof(int32_t));
    memcpy(flags, h_i + n, (size_t)n * sizeof(int32_t));
    memcpy(n_popped, h_i + 2 * (size_t)n, (size_t)n * sizeof(int32_t));
    for (int v = 0; v < n; v++) ctx->expand_solves += h_i[3 * (size_t)n + v];
    ctx->have_batch = false; // ppe_get_ribbons_after refers to ppe_true_cost_batch only
    ctx->last_was_expand = true;
    return PPE_OK;
}

const double* ppe_ribbon_pool(ppe_ctx* ctx, int64_t* n_ribbons) {
    if (!ctx || !ctx->last_was_expand) { if (n_ribbons) *n_ribbons = 0; return nullptr; }
    if (n_ribbons) *n_ribbons = ctx->n_pool;
    return ctx->h_pool;
}

int64_t ppe_expand_solve_count(const ppe_ctx* ctx) { return ctx ? ctx->expand_solves : 0; }

int64_t ppe_launch_count(const ppe_ctx* ctx) { return ctx ? ctx->launches : 0; }
// development builds (make prof): cycle counters of the warp walker since the last read; returns 0 entries in the shipped library
int ppe_debug_k2b_profile(ppe_ctx* ctx, unsigned long long* out16) {
    if (!ctx || !out16) return PPE_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    return k2b_profile_read(out16);
}
uint64_t ppe_map_generation(const ppe_ctx* ctx) { return ctx ? ctx->map_generation : 0; }



int ppe_measure_fp64_peak(ppe_ctx* ctx, double* tflops, void* stream) {
    if (!ctx || !tflops) return PPE_ERR_INVALID;
    PPE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    double* d_out = nullptr;
    PPE_CUDA(ctx, cudaMalloc((void**)&d_out, sizeof(double)));
    cudaEvent_t a, b;
    PPE_CUDA(ctx, cudaEventCreate(&a));
    PPE_CUDA(ctx, cudaEventCreate(&b));
    const int blocks = ctx->sm_count * 8, iters = 1 << 15;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        PPE_CUDA(ctx, cudaEventRecord(a, s));
        PPE_CUDA(ctx, launch_fp64_peak(d_out, blocks, iters, s));
        PPE_CUDA(ctx, cudaEventRecord(b, s));
        PPE_CUDA(ctx, cudaEventSynchronize(b));
        float ms = 0;
        PPE_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
        const double flops = (double)blocks * 256.0 * 8.0 * (double)iters * 2.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    ctx->launches += 4;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d_out);
    *tflops = best;
    return PPE_OK;
}

} // extern "C"
