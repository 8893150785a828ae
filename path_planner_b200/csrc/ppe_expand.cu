// ppe_expand.cu -- frontier expansion on the device: SamplingBasedPlanner::expand (SamplingBasedPlanner.cpp:52-151)
// for many open-list vertices per launch group, against the RESIDENT sample set.
//
//   KS  k_sample_keep / k_sample_scan / k_sample_append   SamplingBasedPlanner::addSamples (:157-164): Map::isBlocked
//                         for every generated state, order-preserving append of the free ones to the resident SoA.
//   KX  k_expand_select   one CTA per vertex: (1) all threads stream the resident samples (coalesced, L2-resident
//                         16 B per sample) and collect those inside a search radius; (2) bitonic sort by (distance,
//                         index) = the pop order of the reference's Euclidean heap (:85-94) whenever no two popped
//                         distances are equal; (3) the loop of :91-133 is replayed on 256 candidates at a time -- every
//                         thread solves the Dubins words of one candidate x 2 radii (K1's branch-free solver), thread 0
//                         runs the k-best max-heaps with libstdc++'s push_heap / pop_heap arrangement -- until both radii
//                         are done; (4) the CTA emits the vertex's edges (endpoint edges :65-81, winners x
//                         speeds :134-149) in the reference's push order as ppe_edge records for K2.
//                         The F x N length matrix is never materialised: only ~#popped solves happen per vertex.
//   KP  k_expand_pack     compact child records (ppe_child, 160 B) from the K2 result records.
#include <float.h>

#include "ppe_device.cuh"
#include "ppe_math.cuh"

namespace ppe {

namespace {

constexpr int kSelThreads = 256;
constexpr int kSelCap = 8192;     // candidates per pass (a power of two: the bitonic sort pads to one): 64 KB + 32 KB of dynamic shared memory
constexpr int kMaxBranch = 16;    // PlannerConfig::branchingFactor() upper limit served on the device (default 9)
constexpr unsigned kFullMask = 0xffffffffu;

// ---- KS ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sample_keep(const __grid_constant__ WorldD w, const long long n, const double* __restrict__ x,
                                                      const double* __restrict__ y, uint8_t* __restrict__ keep,
                                                      unsigned int* __restrict__ block_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool k = false;
    if (i < n) {
        k = !map_blocked(w, x[i], y[i]); // SamplingBasedPlanner.cpp:160
        keep[i] = k ? 1 : 0;
    }
    const int c = __syncthreads_count(k);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (unsigned)c;
}

// exclusive scan of the per-block counts (<= 2^22 samples per call -> <= 16384 blocks), one CTA
__global__ void __launch_bounds__(1024) k_sample_scan(unsigned int* __restrict__ block_count, int nblocks, unsigned int* total) {
    __shared__ unsigned int s[1024];
    unsigned int carry = 0;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned int v = i < nblocks ? block_count[i] : 0u;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const unsigned int t = threadIdx.x >= d ? s[threadIdx.x - d] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_count[i] = carry + s[threadIdx.x] - v;
        carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) k_sample_append(const long long n, const double* __restrict__ x, const double* __restrict__ y,
                                                        const double* __restrict__ h, const uint8_t* __restrict__ keep,
                                                        const unsigned int* __restrict__ block_offset, double* __restrict__ sx,
                                                        double* __restrict__ sy, double* __restrict__ sh, const long long base) {
    __shared__ unsigned int s_warp[8];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool k = i < n && keep[i] != 0;
    const unsigned ballot = __ballot_sync(kFullMask, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    unsigned int off = block_offset[blockIdx.x];
    for (int q = 0; q < warp; q++) off += s_warp[q];
    off += __popc(ballot & ((1u << lane) - 1u));
    if (k) {
        sx[base + off] = x[i];
        sy[base + off] = y[i];
        sh[base + off] = h[i];
    }
}

// ---- KX ----------------------------------------------------------------------------------------------------------------
// libstdc++ std::push_heap / std::pop_heap (bits/stl_heap.h __push_heap, __adjust_heap) on parallel arrays with the
// comparator of SamplingBasedPlanner::getDubinsComparator (:171-176): comp(a, b) = a.approxCost < b.approxCost, a max-heap.
struct KBest {
    double cost[kMaxBranch + 1];  // Edge::approxCost() of the candidate edge
    double len[kMaxBranch + 1];   // wrapper length (dubins_path_length)
    int id[kMaxBranch + 1];       // entry of the winner pool below
    int size;
};

__device__ __forceinline__ void kbest_push_up(KBest& h, int hole, int top, double vcost, double vlen, int vid) {
    int parent = (hole - 1) / 2;
    while (hole > top && h.cost[parent] < vcost) {
        h.cost[hole] = h.cost[parent]; h.len[hole] = h.len[parent]; h.id[hole] = h.id[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h.cost[hole] = vcost; h.len[hole] = vlen; h.id[hole] = vid;
}

__device__ __forceinline__ void kbest_push(KBest& h, double vcost, double vlen, int vid) { // push_back + std::push_heap
    h.size++;
    kbest_push_up(h, h.size - 1, 0, vcost, vlen, vid);
}

// std::pop_heap + pop_back; returns the pool id of the removed (largest) element
__device__ __forceinline__ int kbest_pop(KBest& h) {
    const int removed = h.id[0];
    const int len = h.size - 1; // heap range after the swap-out
    const double vcost = h.cost[len], vlen = h.len[len];
    const int vid = h.id[len];
    int hole = 0, second = 0;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (h.cost[second] < h.cost[second - 1]) second--;
        h.cost[hole] = h.cost[second]; h.len[hole] = h.len[second]; h.id[hole] = h.id[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        h.cost[hole] = h.cost[second - 1]; h.len[hole] = h.len[second - 1]; h.id[hole] = h.id[second - 1];
        hole = second - 1;
    }
    kbest_push_up(h, hole, 0, vcost, vlen, vid);
    h.size = len;
    return removed;
}

struct Winner {   // what a heap entry carries besides its keys
    double param[3];
    int type;
    int sample;   // index into the resident sample set
};

struct SelShared {
    KBest heap[2];
    int free_id[2]; // pool slot the next push into a full heap uses (the slot of the last evicted entry)
    Winner pool[2][kMaxBranch + 1];
    // solves of the current kSelThreads candidates
    double c_len[2][kSelThreads];
    double c_par[2][kSelThreads][3];
    int c_type[2][kSelThreads];
    int count;
    int done[2];
    int pops;
    int solves;
    int consumed;   // candidates consumed by the replay in the current pass
    int tie;
};

__global__ void __launch_bounds__(kSelThreads)
k_expand_select(const ExpandParamsD p, const ppe_vertex* __restrict__ verts, ppe_edge* __restrict__ edges,
                int32_t* __restrict__ edge_sample, int32_t* __restrict__ n_children, int32_t* __restrict__ flags,
                int32_t* __restrict__ n_popped, int32_t* __restrict__ n_solved) {
    __shared__ SelShared s;
    extern __shared__ double sel_dyn[];
    double* const s_d = sel_dyn;                                       // [kSelCap] squared distance, then distance
    int* const s_idx = reinterpret_cast<int*>(sel_dyn + kSelCap);      // [kSelCap] resident sample index
    const int v = blockIdx.x;
    const int tid = threadIdx.x;
    const ppe_vertex vx = verts[v];
    const double sx = vx.state[0], sy = vx.state[1];
    const int N = p.n_samples;
    int flag = 0;

    // Candidates are taken in BANDS of increasing distance: (r_lo2, r_hi2] in squared distance, first band a disc sized from
    // the sample density.  A band that does not fit the shared-memory capacity is shrunk; when a band is consumed and the
    // loop of :91 is still running, the next band continues with the heaps as they stand -- the consumed prefix can be
    // arbitrarily long (deep anytime iterations: 10^6..10^7 samples) with a fixed shared-memory footprint.
    double r_lo2 = -1.0, r_hi2 = p.r2_init > 0 ? p.r2_init : 1.0;
    int n = 0;
    long long seen = 0;      // samples in the bands consumed so far
    bool complete = false;   // every resident sample lies in a band consumed so far
    bool tie_probe = false;  // both radii finished exactly at the end of a band: only the first sample beyond it is needed
    double last_d = -1.0;    // distance of the last candidate of the previous band
    if (tid == 0) {
        s.heap[0].size = s.heap[1].size = 0;
        s.done[0] = s.done[1] = 0;
        s.free_id[0] = s.free_id[1] = p.k;
        s.pops = 0; s.solves = 0; s.consumed = 0; s.tie = 0;
    }
    __syncthreads();
    for (int pass = 0; pass < 4096; pass++) {
        // ---- (1) collect the samples of the band ------------------------------------------------------------------------
        if (tid == 0) s.count = 0;
        __syncthreads();
        for (int i = tid; i < N; i += kSelThreads) {
            const double dx = sx - p.sx[i], dy = sy - p.sy[i];
            const double d2 = dx * dx + dy * dy; // State::distanceTo (State.cpp:91-93), squared
            if (d2 > r_lo2 && d2 <= r_hi2) {
                const int slot = atomicAdd(&s.count, 1);
                if (slot < kSelCap) { s_d[slot] = d2; s_idx[slot] = i; }
            }
        }
        __syncthreads();
        const int cnt = s.count;
        if (cnt > kSelCap) { // too many: shrink the band (uniform density -> count ~ area)
            const double lo = r_lo2 > 0 ? r_lo2 : 0.0;
            const double shrunk_hi = lo + (r_hi2 - lo) * (0.75 * (double)kSelCap / (double)cnt);
            if (!(shrunk_hi < r_hi2) || !(shrunk_hi > lo)) { flag |= PPE_EXPAND_OVERFLOW; break; } // > capacity samples at one distance
            r_hi2 = shrunk_hi;
            __syncthreads();
            continue;
        }
        n = cnt;
        seen += n;
        complete = (seen >= (long long)N);
        // ---- (2) sort by (distance, sample index) -------------------------------------------------------------------------
        int P = 32;
        while (P < n) P <<= 1;
        for (int i = tid; i < P; i += kSelThreads) {
            if (i < n) s_d[i] = sqrt(s_d[i]); // the comparator's key
            else { s_d[i] = DBL_MAX; s_idx[i] = 0x7fffffff; }
        }
        __syncthreads();
        for (int k2 = 2; k2 <= P; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += kSelThreads) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const double a = s_d[i], b = s_d[ixj];
                        const int ai = s_idx[i], bi = s_idx[ixj];
                        const bool a_gt_b = a > b || (a == b && ai > bi);
                        const bool up = (i & k2) == 0;
                        if (a_gt_b == up) { s_d[i] = b; s_d[ixj] = a; s_idx[i] = bi; s_idx[ixj] = ai; }
                    }
                }
                __syncthreads();
            }
        }
        // a sample of this band at exactly the distance of the previous band's last one (distinct squares can round to
        // the same square root): the reference's pop order between the two depends on its heap arrangement
        if (n > 0 && last_d >= 0 && s_d[0] == last_d) flag |= PPE_EXPAND_TIE;
        if (tie_probe) break; // the loop had already ended: this band was only collected for the check above
        if (tid == 0) s.consumed = 0;
        __syncthreads();
        // ---- (3) replay of the k-nearest loop, kSelThreads candidates per step ---------------------------------------------------------
        {
            // every thread of the CTA solves one candidate (both radii) per step; thread 0 then replays the loop body for
            // those kSelThreads candidates in order
            const double q0[3] = {sx, sy, heading_to_yaw(vx.state[2])};
            int pos = 0;
            while (pos < n && !(s.done[0] && s.done[1])) {
                const int c = pos + tid;
                if (c < n) {
                    const int si = s_idx[c];
                    const double q1[3] = {p.sx[si], p.sy[si], heading_to_yaw(p.sh[si])};
#pragma unroll 1
                    for (int j = 0; j < 2; j++) {
                        if (s.done[j] || !(p.rho[j] > 0)) continue;
                        DubinsPathD path;
                        path.param[0] = path.param[1] = path.param[2] = 0;
                        path.type = 0;
                        const int e = dubins_shortest_path(&path, q0, q1, p.rho[j]); // Edge::computeApproxCost, Edge.cpp:11-20
                        s.c_len[j][tid] = e == kEdubOk ? dubins_path_length(path) : 0.0;
                        s.c_par[j][tid][0] = path.param[0]; s.c_par[j][tid][1] = path.param[1]; s.c_par[j][tid][2] = path.param[2];
                        s.c_type[j][tid] = path.type;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    const int m = n - pos < kSelThreads ? n - pos : kSelThreads;
                    int q = 0;
                    for (; q < m && !(s.done[0] && s.done[1]); q++) { // one iteration of the loop at :91
                        const double dist = s_d[pos + q];
                        s.pops++;
#pragma unroll 1
                        for (int j = 0; j < 2; j++) {
                            if (s.done[j]) continue;
                            if (!(p.rho[j] > 0)) { s.done[j] = 1; continue; }      // :103-106
                            KBest& h = s.heap[j];
                            if (h.size < p.k || h.len[0] > dist) {                // :109-110
                                if (dist > p.inc) {                               // :111
                                    const double len = s.c_len[j][q];
                                    const double cost = len / p.max_speed * p.time_factor; // Edge.cpp:64-66 at the max speed
                                    // pool slots 0 .. k-1 fill in push order; a full heap pushes into the free slot
                                    const int id = h.size < p.k ? h.size : s.free_id[j];
                                    Winner& wn = s.pool[j][id];
                                    wn.param[0] = s.c_par[j][q][0]; wn.param[1] = s.c_par[j][q][1]; wn.param[2] = s.c_par[j][q][2];
                                    wn.type = s.c_type[j][q];
                                    wn.sample = s_idx[pos + q];
                                    kbest_push(h, cost, len, id);
                                    s.solves++;
                                    if (h.size > p.k) s.free_id[j] = kbest_pop(h); // :121-124
                                }
                            } else {
                                s.done[j] = 1;                                    // :129-131
                            }
                        }
                    }
                    s.consumed = pos + q;
                }
                __syncthreads();
                pos = s.consumed;
            }
            // exact-distance ties inside the consumed prefix (or between the last popped and the next sample): the pop
            // order then depends on the heap arrangement of the reference's m_Samples
            int tie = 0;
            const int upto = s.consumed < n ? s.consumed : n - 1;
            for (int q = tid; q < upto; q += kSelThreads) tie |= (s_d[q] == s_d[q + 1]) ? 1 : 0;
            tie = __syncthreads_or(tie);
            if (tid == 0) s.tie = tie;
        }
        __syncthreads();
        if (s.tie) flag |= PPE_EXPAND_TIE;
        const bool finished = s.done[0] && s.done[1];
        if (complete) break;                       // nothing beyond this band
        if (finished && s.consumed < n) break;     // the loop ended inside the band: the next candidate was compared above
        // the band is consumed: go on with the next one -- to continue the loop, or (finished exactly at its end) only to
        // compare its first sample with the last one popped
        if (n > 0) last_d = s_d[n - 1];
        tie_probe = finished;
        r_lo2 = r_hi2;
        r_hi2 = pass > 48 ? DBL_MAX : r_hi2 * 4.0;
        __syncthreads();
    }

    // ---- (4) emit the edges in the reference's push order ---------------------------------------------------------------------
    // slot layout: [endpoint edges: speed-major, radius-minor] [radius 0 winners in heap-array order x speeds] [radius 1 ...]
    const int stride = p.stride;
    const int n_speeds = p.speed[1] > 0 ? 2 : 1;
    const int n_radii = (p.rho[0] > 0 ? 1 : 0) + (p.rho[1] > 0 ? 1 : 0);
    const int n_ep = vx.has_endpoint ? n_speeds * n_radii : 0;
    const int n0 = s.heap[0].size * n_speeds, n1 = s.heap[1].size * n_speeds;
    const int total = n_ep + n0 + n1;
    ppe_edge* out = edges + (size_t)v * stride;
    for (int e = tid; e < stride; e += kSelThreads) {
        ppe_edge ed;
        memset(&ed, 0, sizeof ed);
        int sample = -1;
        if (e >= total) {
            ed.has_path = -1; // empty slot
        } else {
            ed.src[0] = vx.state[0]; ed.src[1] = vx.state[1]; ed.src[2] = vx.state[2]; ed.src[3] = vx.state[3]; ed.src[4] = vx.state[4];
            ed.src_g = vx.g;
            ed.ribbon_set = vx.ribbon_set;
            if (e < n_ep) {
                // for speed in speeds: for radius in radii (:69-80)
                const int si = e / n_radii, ri = e % n_radii;
                const int j = (p.rho[0] > 0) ? ri : 1;
                const double speed = p.speed[si];
                ed.dst[0] = vx.endpoint[0]; ed.dst[1] = vx.endpoint[1]; ed.dst[2] = vx.endpoint[2]; ed.dst[3] = speed;
                ed.has_path = 0;
                ed.coverage_allowed = p.rho[j] == p.coverage_rho ? 1 : 0;
            } else {
                const int e2 = e - n_ep;
                const int j = e2 < n0 ? 0 : 1;
                const int w = (e2 - (j ? n0 : 0)) / n_speeds, si = (e2 - (j ? n0 : 0)) % n_speeds;
                const KBest& h = s.heap[j];
                const Winner& wn = s.pool[j][h.id[w]];
                const double speed = p.speed[si];
                sample = wn.sample;
                ed.has_path = 1;
                ed.path_qi[0] = sx; ed.path_qi[1] = sy; ed.path_qi[2] = heading_to_yaw(vx.state[2]);
                ed.path_param[0] = wn.param[0]; ed.path_param[1] = wn.param[1]; ed.path_param[2] = wn.param[2];
                ed.path_rho = p.rho[j];
                ed.path_type = wn.type;
                ed.w_speed = speed;                               // wrapper.setSpeed(speed), :141
                ed.w_start_time = vx.state[4];                    // DubinsWrapper::set, DubinsWrapper.cpp:15
                ed.w_end_time = vx.state[4] + h.len[w] / speed;   // setEndTime, DubinsWrapper.cpp:96-98
                ed.dst[0] = p.sx[sample]; ed.dst[1] = p.sy[sample]; ed.dst[2] = p.sh[sample]; ed.dst[3] = speed;
                ed.coverage_allowed = p.rho[j] == p.coverage_rho ? 1 : 0;
            }
        }
        out[e] = ed;
        edge_sample[(size_t)v * stride + e] = sample;
    }
    if (tid == 0) {
        n_children[v] = total;
        flags[v] = flag;
        n_popped[v] = s.pops;
        n_solved[v] = s.solves;
    }
}

// ---- KP ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_expand_pack(const long long n, const ppe_edge* __restrict__ edges,
                                                      const ppe_edge_result* __restrict__ results,
                                                      const int32_t* __restrict__ edge_sample, ppe_child* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ppe_edge_result& r = results[i];
    ppe_child c;
    c.true_cost = r.true_cost; c.collision_penalty = r.collision_penalty; c.approx_cost = r.approx_cost;
    c.end[0] = r.end[0]; c.end[1] = r.end[1]; c.end[2] = r.end[2]; c.end[3] = r.end[3]; c.end[4] = r.end[4];
    c.g = r.g; c.h = r.h;
    c.coverage_completed_time = r.coverage_completed_time;
    c.path_param[0] = r.path_param[0]; c.path_param[1] = r.path_param[1]; c.path_param[2] = r.path_param[2];
    c.w_end_time = r.w_end_time;
    c.ribbons_offset = r.ribbons_offset;
    c.sample_index = edge_sample[i];
    c.path_type = r.path_type;
    c.infeasible = r.infeasible;
    c.status = r.status;
    c.coverage_allowed = edges[i].coverage_allowed;
    c.n_ribbons_after = r.n_ribbons_after;
    c.ribbons_changed = r.ribbons_changed;
    c.reserved = 0;
    out[i] = c;
}

} // namespace

int expand_max_branch() { return kMaxBranch; }

cudaError_t launch_sample_filter(const WorldD& w, int64_t n, const double* x, const double* y, const double* h, uint8_t* keep,
                                 unsigned int* block_count, unsigned int* total, double* sx, double* sy, double* sh, int64_t base,
                                 cudaStream_t stream, int phase) {
    const int blocks = (int)((n + 255) / 256);
    if (phase == 0) {
        k_sample_keep<<<blocks, 256, 0, stream>>>(w, (long long)n, x, y, keep, block_count);
        k_sample_scan<<<1, 1024, 0, stream>>>(block_count, blocks, total);
    } else {
        k_sample_append<<<blocks, 256, 0, stream>>>((long long)n, x, y, h, keep, block_count, sx, sy, sh, (long long)base);
    }
    return cudaGetLastError();
}

cudaError_t launch_expand_select(const ExpandParamsD& p, int n_vertices, const ppe_vertex* verts, ppe_edge* edges,
                                 int32_t* edge_sample, int32_t* n_children, int32_t* flags, int32_t* n_popped, int32_t* n_solved,
                                 cudaStream_t stream) {
    const size_t smem = (size_t)kSelCap * (sizeof(double) + sizeof(int));
    cudaError_t e = cudaFuncSetAttribute(k_expand_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_expand_select<<<n_vertices, kSelThreads, smem, stream>>>(p, verts, edges, edge_sample, n_children, flags, n_popped, n_solved);
    return cudaGetLastError();
}

cudaError_t launch_expand_pack(int64_t n, const ppe_edge* edges, const ppe_edge_result* results, const int32_t* edge_sample,
                               ppe_child* out, cudaStream_t stream) {
    k_expand_pack<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((long long)n, edges, results, edge_sample, out);
    return cudaGetLastError();
}

} // namespace ppe
