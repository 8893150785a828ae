// ppe_device.cuh -- device helpers shared by ppe_kernels.cu (K1-K3) and ppe_expand.cu (frontier expansion).
#pragma once

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

static __device__ __forceinline__ bool map_blocked(const WorldD& w, double x, double y) {
    if (w.map_kind == kMapNone) return false;
    unsigned long long r, c;
    if (!map_cell(w, x, y, &r, &c)) return true; // out of bounds = blocked
    const uint32_t word = __ldg(&w.map_bits[r * (unsigned long long)w.stride_words + (c >> 5)]);
    return (word >> (c & 31)) & 1u;
}

// Chunk culling: true when every cell within the dilation radius of (x, y)'s cell is in bounds and free.
static __device__ __forceinline__ bool map_safe(const WorldD& w, double x, double y) {
    if (w.map_kind == kMapNone) return true;
    if (w.safe_bits == nullptr) return false;
    unsigned long long r, c;
    if (!map_cell(w, x, y, &r, &c)) return false;
    const uint32_t word = __ldg(&w.safe_bits[r * (unsigned long long)w.stride_words + (c >> 5)]);
    return (word >> (c & 31)) & 1u;
}


} // namespace ppe
