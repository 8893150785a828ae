// ppe_device.cuh -- device helpers shared by ppe_kernels.cu (K1-K3) and ppe_expand.cu (frontier expansion).
#pragma once

#include <float.h>

#include "ppe_kernels.cuh"

namespace ppe {

// Map::isBlocked (Map.cpp:4-6) / GridWorldMap::isBlocked (GridWorldMap.cpp:84-93).  The bitmap
// (<= 2 MiB at 4096^2) stays L2/L1 resident; consecutive lanes are consecutive 0.05 m samples, so
// a warp's 32 lookups fall into one or two 32-byte sectors.  x / res is a multiplication by the
// exact reciprocal when the resolution is a power of two (bit-identical quotient), else a division.
static __device__ __forceinline__ bool map_cell(const WorldD& w, double x, double y, unsigned long long* r, unsigned long long* c) {
    double qx, qy;
    if (w.res_pow2) { qx = x * w.inv_resolution; qy = y * w.inv_resolution; }
    else { qx = x / w.resolution; qy = y / w.resolution; }
    if (x < 0 || qx >= (double)w.cols) return false;
    if (y < 0 || qy >= (double)w.rows) return false;
    *r = (unsigned long long)qy;
    *c = (unsigned long long)qx;
    return true;
}

// word (r, c >> 5) of bitmap `which` (0 = occupancy, 1 = safe) from the CTA's shared-memory tile when it covers the
// cell, else from global memory (L2-resident)
static __device__ __forceinline__ uint32_t map_word(const WorldD& w, const uint32_t* tile, int which, unsigned long long r,
                                                    unsigned long long c) {
    const unsigned long long cw = c >> 5;
    if (tile != nullptr) {
        const long long tr = (long long)r - w.tile_r0, tw = (long long)cw - w.tile_w0;
        if (tr >= 0 && tr < w.tile_rows && tw >= 0 && tw < w.tile_words)
            return tile[((size_t)which * w.tile_rows + (size_t)tr) * w.tile_words + (size_t)tw];
    }
    const uint32_t* bits = which ? w.safe_bits : w.map_bits;
    return __ldg(&bits[r * (unsigned long long)w.stride_words + cw]);
}

static __device__ __forceinline__ bool map_blocked(const WorldD& w, double x, double y, const uint32_t* tile = nullptr) {
    if (w.map_kind == kMapNone) return false;
    unsigned long long r, c;
    if (!map_cell(w, x, y, &r, &c)) return true; // out of bounds = blocked
    const uint32_t word = map_word(w, tile, 0, r, c);
    return (word >> (c & 31)) & 1u;
}

// Chunk culling: true when every cell within the dilation radius of (x, y)'s cell is in bounds and free.
static __device__ __forceinline__ bool map_safe(const WorldD& w, double x, double y, const uint32_t* tile = nullptr) {
    if (w.map_kind == kMapNone) return true;
    if (w.safe_bits == nullptr) return false;
    unsigned long long r, c;
    if (!map_cell(w, x, y, &r, &c)) return false;
    const uint32_t word = map_word(w, tile, 1, r, c);
    return (word >> (c & 31)) & 1u;
}

// ---- TMA bulk copy of the map tile (cp.async.bulk, completion on an mbarrier) ----------------------------------------
static __device__ __forceinline__ void tile_stage(const WorldD& w, uint32_t* tile, unsigned long long* bar) {
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned row_bytes = (unsigned)w.tile_words * 4u;
        const unsigned total = 2u * (unsigned)w.tile_rows * row_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(total) : "memory");
        for (int which = 0; which < 2; which++) {
            const uint32_t* bits = which ? w.safe_bits : w.map_bits;
            for (int r = 0; r < w.tile_rows; r++) {
                const uint32_t* src = bits + (size_t)(w.tile_r0 + r) * w.stride_words + w.tile_w0;
                const unsigned dst_s = (unsigned)__cvta_generic_to_shared(tile + ((size_t)which * w.tile_rows + r) * w.tile_words);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
                             "l"(src), "r"(row_bytes), "r"(bar_s)
                             : "memory");
            }
        }
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar_s)
                     : "memory");
    }
}


// RibbonManager::tspPointRobotNoSplitAllRibbons (RibbonManager.cpp:53-67) / ...KRibbons (:69-95) for the lists of up to
// kTspMaxRibbons ribbons a vertex holds when Executive's default heuristic is in force (executive.cpp:391; more than five
// ribbons at the root force MaxDistance, RibbonManager.cpp:381-385).  The reference recurses with list copies; here the
// depth-first search runs on an explicit stack.  A level holds the ribbons left IN LIST ORDER (4 bits each) -- the K variant
// sorts the list it was handed with std::list::sort (stable) by "nearest end point farther first" (comparator :73-77) and
// tries the first K -- so ties resolve as the reference's.  Sums and fmax / fmin run in the reference's order.
constexpr int kTspMaxRibbons = 8;

static __device__ __noinline__ double tsp_point_robot(const double4* rib, int n, double x, double y, int K, double W) {
    struct Level {
        unsigned long long order; // ribbons left, list order, 4 bits each
        double soFar, px, py, mn;
        int cnt, b;               // b: next branch = 2 * position + direction
    };
    Level st[kTspMaxRibbons + 1];
    int l = 0;
    {
        unsigned long long o = 0;
        for (int i = 0; i < n; i++) o |= (unsigned long long)i << (4 * i);
        st[0].order = o; st[0].cnt = n; st[0].soFar = 0; st[0].px = x; st[0].py = y; st[0].mn = DBL_MAX; st[0].b = -1;
    }
    double ret = 0;
    for (;;) {
        Level& L = st[l];
        if (L.b < 0) { // entering the level
            if (L.cnt == 0) { // :54 / :71: nothing left
                ret = L.soFar;
                if (l == 0) return ret;
                l--;
                st[l].mn = fmin(st[l].mn, ret);
                continue;
            }
            if (K > 0) { // ribbonsLeft.sort(comp), stable
                int ord[kTspMaxRibbons];
                double key[kTspMaxRibbons];
                for (int i = 0; i < L.cnt; i++) {
                    ord[i] = (int)((L.order >> (4 * i)) & 15ull);
                    const double4 r = rib[ord[i]];
                    key[i] = fmin(sqrt((L.px - r.x) * (L.px - r.x) + (L.py - r.y) * (L.py - r.y)),
                                  sqrt((L.px - r.z) * (L.px - r.z) + (L.py - r.w) * (L.py - r.w)));
                }
                for (int i = 1; i < L.cnt; i++) {
                    const int o = ord[i];
                    const double kv = key[i];
                    int j = i;
                    while (j > 0 && kv > key[j - 1]) { ord[j] = ord[j - 1]; key[j] = key[j - 1]; j--; }
                    ord[j] = o; key[j] = kv;
                }
                unsigned long long o2 = 0;
                for (int i = 0; i < L.cnt; i++) o2 |= (unsigned long long)ord[i] << (4 * i);
                L.order = o2;
            }
            L.mn = DBL_MAX;
            L.b = 0;
        }
        const int pos = L.b >> 1, dir = L.b & 1;
        const int limit = (K > 0 && K < L.cnt) ? K : L.cnt;
        if (pos >= limit) { // all branches of this level done
            ret = L.mn;
            if (l == 0) return ret;
            l--;
            st[l].mn = fmin(st[l].mn, ret);
            continue;
        }
        L.b++;
        const int ri = (int)((L.order >> (4 * pos)) & 15ull);
        const double4 r = rib[ri];
        const double len = sqrt((r.z - r.x) * (r.z - r.x) + (r.w - r.y) * (r.w - r.y)); // Ribbon::length()
        const double tx = dir == 0 ? r.x : r.z, ty = dir == 0 ? r.y : r.w;             // the end approached first
        const double d = sqrt((L.px - tx) * (L.px - tx) + (L.py - ty) * (L.py - ty));
        Level& C2 = st[l + 1];
        const unsigned long long low = L.order & ((1ull << (4 * pos)) - 1ull);
        const unsigned long long high = pos + 1 < 16 ? (L.order >> (4 * (pos + 1))) : 0ull;
        C2.order = low | (high << (4 * pos));
        C2.cnt = L.cnt - 1;
        C2.soFar = fmax(L.soFar + len - 2 * W + d, 0.0);
        C2.px = dir == 0 ? r.z : r.x; C2.py = dir == 0 ? r.w : r.y; // leave from the other end
        C2.b = -1;
        l++;
    }
}

// Vertex::computeApproxToGo (Vertex.cpp:49-64) for the heuristics the device evaluates besides MaxDistance; -1 otherwise
static __device__ __forceinline__ double tsp_heuristic_or_unset(const ppe_config& cfg, const double4* rib, int nr, double x, double y) {
    const bool pr = cfg.heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL || cfg.heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K;
    if (!pr || nr > kTspMaxRibbons) return -1.0;
    const int K = cfg.heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K ? (cfg.tsp_k > 0 ? cfg.tsp_k : 2) : 0;
    const double d = nr == 0 ? 0.0 : tsp_point_robot(rib, nr, x, y, K, cfg.ribbon_width);
    return d / cfg.max_speed * cfg.time_penalty_factor;
}

} // namespace ppe
