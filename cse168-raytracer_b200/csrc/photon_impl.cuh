// photon_impl.cuh -- device kNN gather over the reference's heap-ordered photon kd-tree.
//
// One thread per query runs Photon_map::locate_photons (reference PhotonMap.cpp:152-243) as an explicit-stack
// walk that visits nodes in the reference's order (near child, then far child if the splitting plane is within
// the current search radius, then the node's own photon -- post-order), applies the reference's direction
// filter (PhotonMap.cpp:183-186) and maintains the k-nearest set in the same array-then-max-heap structure.
// Because order and comparisons are the same, the candidate array ends in the same permutation, the power sum
// adds in the same order, and the estimate is bit-identical to the reference's (no FMA on this path either).
// The k <= 512 candidate list (distance + index) lives in per-thread local memory.
#ifndef MIROGPU_PHOTON_IMPL_CUH
#define MIROGPU_PHOTON_IMPL_CUH

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace mirogpu {

static inline int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int PhotonMapDevice::upload(const void* photons28, int n, std::string& err)
{
    release();
    struct Photon28 { float pos[3]; short plane; unsigned char theta, phi; float power[3]; };
    static_assert(sizeof(Photon28) == 28, "reference Photon is 28 bytes (PhotonMap.h:16-22)");
    std::vector<float4> packed((size_t)2 * (n + 1), make_float4(0.f, 0.f, 0.f, 0.f));
    const Photon28* src = static_cast<const Photon28*>(photons28);
    for (int i = 1; i <= n; ++i) {
        const Photon28& p = src[i];
        const uint32_t bits = ((uint32_t)(uint16_t)p.plane & 3u) | ((uint32_t)p.theta << 8) | ((uint32_t)p.phi << 16);
        float w; memcpy(&w, &bits, 4);
        packed[2 * (size_t)i] = make_float4(p.pos[0], p.pos[1], p.pos[2], w);
        packed[2 * (size_t)i + 1] = make_float4(p.power[0], p.power[1], p.power[2], 0.f);
    }
    // direction tables exactly as the reference fills them (PhotonMap.cpp:47-53): double trig, stored as float
    std::vector<float> tab(1024);
    for (int i = 0; i < 256; ++i) {
        const double angle = double(i) * (1.0 / 256.0) * M_PI;
        tab[i] = (float)cos(angle); tab[256 + i] = (float)sin(angle);
        tab[512 + i] = (float)cos(2.0 * angle); tab[768 + i] = (float)sin(2.0 * angle);
    }
    cudaError_t e = cudaMalloc(&d_photons, packed.size() * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&d_tables, tab.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_search, packed.size() * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&d_tickets, (size_t)MIRO_GW_TICKET_SLOTS * MIRO_GW_TICKET_SPAN * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemcpy(d_photons, packed.data(), packed.size() * sizeof(float4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        // the power half of each record becomes the photon's direction, multiplied out of the tables the way
        // Photon_map::photon_dir does (PhotonMap.cpp:68-74: single binary32 products, so host and device agree bit for bit)
        for (int i = 1; i <= n; ++i) {
            const Photon28& p = src[i];
            const float st = tab[256 + p.theta];
            packed[2 * (size_t)i + 1] = make_float4(xmul(st, tab[512 + p.phi]), xmul(st, tab[768 + p.phi]), tab[p.theta], 0.f);
        }
        e = cudaMemcpy(d_search, packed.data(), packed.size() * sizeof(float4), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMemcpy(d_tables, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { err = std::string("photon upload: ") + cudaGetErrorString(e); release(); return e == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA; }
    stored = n;
    half_stored = n / 2 - 1;   // PhotonMap.cpp:358
    return MIROGPU_OK;
}

__global__ void __launch_bounds__(128) k_photon_gather(const float4* __restrict__ photons, const float* __restrict__ tables, int stored,
                                                       int half_stored, const float* __restrict__ pos3, const float* __restrict__ nrm3,
                                                       const float4* __restrict__ active, size_t n, const uint32_t* __restrict__ d_n,
                                                       float max_dist, int kmax, float* __restrict__ irr3)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d_n) n = min(n, (size_t)*d_n);
    if (q >= n) return;
    if (active && active[q].w == 0.f) { irr3[3 * q] = irr3[3 * q + 1] = irr3[3 * q + 2] = 0.f; return; }
    const float px = pos3[3 * q], py = pos3[3 * q + 1], pz = pos3[3 * q + 2];
    const float nx = nrm3[3 * q], ny = nrm3[3 * q + 1], nz = nrm3[3 * q + 2];
    float d2[MIRO_PHOTON_KMAX + 1];
    uint32_t id[MIRO_PHOTON_KMAX + 1];
    int found = 0;
    bool heap = false;
    d2[0] = xmul(max_dist, max_dist);
    uint32_t stack[72];     // (index << 2) | stage; the tree has < 2^30 nodes and depth <= 30
    int sp = 0;
    if (stored >= 1) stack[sp++] = (1u << 2) | 0u;
    while (sp > 0) {
        const uint32_t e = stack[--sp];
        const uint32_t index = e >> 2, stage = e & 3u;
        const float4 ph = __ldg(photons + 2 * (size_t)index);
        const uint32_t bits = __float_as_uint(ph.w);
        if (stage < 2 && (int)index < half_stored) {
            const uint32_t plane = bits & 3u;
            const float qc = plane == 0 ? px : (plane == 1 ? py : pz);
            const float pc = plane == 0 ? ph.x : (plane == 1 ? ph.y : ph.z);
            const float dist1 = xsub(qc, pc);
            const uint32_t near_child = dist1 > 0.0f ? 2 * index + 1 : 2 * index;
            const uint32_t far_child = dist1 > 0.0f ? 2 * index : 2 * index + 1;
            if (stage == 0) {
                stack[sp++] = (index << 2) | 1u;
                stack[sp++] = (near_child << 2) | 0u;
                continue;
            }
            // stage 1: back from the near side
            if (xmul(dist1, dist1) < d2[0]) {
                stack[sp++] = (index << 2) | 2u;
                stack[sp++] = (far_child << 2) | 0u;
                continue;
            }
        }
        // ---- the node's own photon ------------------------------------------------------------------
        float t = xsub(ph.x, px);
        float dist2 = xmul(t, t);
        t = xsub(ph.y, py); dist2 = xadd(dist2, xmul(t, t));
        t = xsub(ph.z, pz); dist2 = xadd(dist2, xmul(t, t));
        const uint32_t th = (bits >> 8) & 0xffu, phi = (bits >> 16) & 0xffu;
        const float st = __ldg(tables + 256 + th);
        const float dx = xmul(st, __ldg(tables + 512 + phi)), dy = xmul(st, __ldg(tables + 768 + phi)), dz = __ldg(tables + th);
        if (!(dist2 < d2[0] && xdot(dx, dy, dz, nx, ny, nz) < 0.0f)) continue;
        if (found < kmax) {
            ++found;
            d2[found] = dist2; id[found] = index;
            continue;
        }
        if (!heap) {
            // first overflow: turn the filled array into a max-heap on distance (bottom-up sift)
            const int half_found = found >> 1;
            for (int k = half_found; k >= 1; --k) {
                int parent = k;
                const uint32_t pid = id[k];
                const float pd = d2[k];
                while (parent <= half_found) {
                    int j = parent + parent;
                    if (j < found && d2[j] < d2[j + 1]) ++j;
                    if (pd >= d2[j]) break;
                    d2[parent] = d2[j]; id[parent] = id[j];
                    parent = j;
                }
                d2[parent] = pd; id[parent] = pid;
            }
            heap = true;
        }
        // replace the farthest candidate and restore the heap
        int parent = 1, j = 2;
        while (j <= found) {
            if (j < found && d2[j] < d2[j + 1]) ++j;
            if (dist2 > d2[j]) break;
            d2[parent] = d2[j]; id[parent] = id[j];
            parent = j;
            j += j;
        }
        id[parent] = index; d2[parent] = dist2;
        d2[0] = d2[1];
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = 1; i <= found; ++i) {
        const float4 pw = __ldg(photons + 2 * (size_t)id[i] + 1);
        sx = xadd(sx, pw.x); sy = xadd(sy, pw.y); sz = xadd(sz, pw.z);
    }
    // density estimate, formed in double like the reference (PhotonMap.cpp:136)
    const float tmp = (float)(((double)1.0f / 3.14159265358979323846) / (double)d2[0]);
    irr3[3 * q] = xmul(sx, tmp); irr3[3 * q + 1] = xmul(sy, tmp); irr3[3 * q + 2] = xmul(sz, tmp);
}

// ---- one query per WARP ---------------------------------------------------------------------------------------------
// The same k-nearest set as the reference's search, found cooperatively: the warp keeps ONE stack of pending kd nodes
// (node, lower bound of its cell's squared distance) and ONE candidate buffer in shared memory.  Each iteration the 32
// lanes pop the 64 topmost nodes (two each), fetch their photons, test plane and photon distances against the current radius, and
// append surviving children (far sides first, so near sides are popped next) and accepted photons by ballot + popc
// compaction.  When the buffer is about to overflow, the k-th smallest distance is found by bisection on the float
// bits (ballot-free: per-lane counts + one warp reduction per step), the buffer is compacted to those k and the radius
// shrinks to that distance.  Per visited node this costs ~2 warp instructions instead of a divergent per-thread stack
// machine with a 4 KB heap in local memory.  The result holds the k nearest photons that pass the direction filter --
// what the reference's heap ends with -- summed in another order: estimates agree to ~1e-6 relative (1e-4).
#ifndef MIRO_GW_STACK
#define MIRO_GW_STACK 1024      /* pending nodes per warp (32 x tree depth is the pseudo-DFS worst case; see the throttle below) */
#endif
#ifndef MIRO_GW_CAND
#define MIRO_GW_CAND 768        /* candidate buffer per warp; must be >= k + 128 */
#endif
#define MIRO_GW_WARPS 4

struct GatherWarpShared {
    uint32_t stack_node[MIRO_GW_STACK];
    float stack_bound[MIRO_GW_STACK];
    float cand_d2[MIRO_GW_CAND];
    uint32_t cand_id[MIRO_GW_CAND];
};

// k-th smallest of cand_d2[0..n) (n > k), compaction of the buffer to exactly those k; returns that distance.
__device__ __forceinline__ float gather_select(GatherWarpShared& sh, int& n, int k, unsigned lane)
{
    // distances are non-negative floats: their bit patterns order like unsigned integers
    uint32_t lo = 0u, hi = 0x7f800000u;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int i = (int)lane; i < n; i += 32) c += __float_as_uint(sh.cand_d2[i]) <= mid;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= k) hi = mid; else lo = mid + 1u;
    }
    const uint32_t tau = lo;
    int below = 0;
    for (int i = (int)lane; i < n; i += 32) below += __float_as_uint(sh.cand_d2[i]) < tau;
    below = __reduce_add_sync(0xffffffffu, below);
    int quota = k - below;      // how many of the candidates AT the k-th distance are kept (> 1 only on exact ties)
    int out = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + (int)lane;
        float d = 0.f; uint32_t id = 0u; bool keep = false, tie = false;
        if (i < n) {
            d = sh.cand_d2[i]; id = sh.cand_id[i];
            const uint32_t b = __float_as_uint(d);
            keep = b < tau; tie = b == tau;
        }
        const unsigned tm = __ballot_sync(0xffffffffu, tie);
        if (tie && __popc(tm & ((1u << lane) - 1u)) < quota) keep = true;
        quota -= min(quota, __popc(tm));
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) { const int o = out + __popc(km & ((1u << lane) - 1u)); sh.cand_d2[o] = d; sh.cand_id[o] = id; }
        out += __popc(km);
        __syncwarp();
    }
    n = out;
    return __uint_as_float(tau);
}

// Queries are claimed in chunks of `chunk` consecutive indices, because consecutive queries are neighbouring pixels: the k
// photons that answered one query bound the search radius of the next one the warp answers -- if all k also face that
// query's normal, its k-th nearest facing photon is no farther than the farthest of them (true for any two queries; it
// is a tight bound for neighbours).
// Starting the walk with that radius instead of max_dist saves most of the visits on evenly lit surfaces (config 5:
// 12 k -> 3 k per query on the walls); the result is the same k nearest photons.
template <int NPL>   // nodes per lane and iteration
__global__ void __launch_bounds__(32 * MIRO_GW_WARPS) k_photon_gather_warp(const float4* __restrict__ photons, const float4* __restrict__ search, int stored,
                                                                          int half_stored, const float* __restrict__ pos3, const float* __restrict__ nrm3,
                                                                          const float4* __restrict__ active, size_t n, const uint32_t* __restrict__ d_n,
                                                                          float max_dist, int kmax, float* __restrict__ irr3, unsigned int* ticket,
                                                                          int chunk, int seed_on)
{
    extern __shared__ unsigned char gw_smem[];
    GatherWarpShared& sh = reinterpret_cast<GatherWarpShared*>(gw_smem)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    if (d_n) n = min(n, (size_t)*d_n);
    const float full_r2 = xmul(max_dist, max_dist);
    // Work distribution: the query range is cut into one contiguous segment per warp; a warp takes chunks of its home segment in
    // order (so nearly every query has its predecessor's photons to seed from, not just 7 of 8) and, when that is exhausted, takes
    // chunks from the other segments' cursors -- segments differ 100-fold in cost, the chunk stays the unit of load balance.
    const uint32_t W = gridDim.x * MIRO_GW_WARPS, wid = blockIdx.x * MIRO_GW_WARPS + (threadIdx.x >> 5);
    const size_t seg_len = (((n + W - 1) / W + (size_t)chunk - 1) / (size_t)chunk) * (size_t)chunk;
    bool prev_valid = false;    // sh.cand_id[0 .. kmax) holds the k photons of the previous query this warp answered
    uint32_t scanned = 0;       // segments (from home, cyclically) already found exhausted
    while (scanned < W) {
        bool has = false;
        if (scanned + lane < W) {
            const uint32_t v = (wid + scanned + lane) % W;
            const size_t sv = (size_t)v * seg_len;
            if (sv < n) has = *(volatile unsigned int*)(ticket + v) < (unsigned int)(min(n, sv + seg_len) - sv);
        }
        const unsigned work = __ballot_sync(0xffffffffu, has);
        if (work == 0u) { scanned += 32u; continue; }
        scanned += (uint32_t)(__ffs(work) - 1);              // stay on this segment until it is exhausted
        const uint32_t v = (wid + scanned) % W;
        const size_t sv = (size_t)v * seg_len, lv = min(n, sv + seg_len) - sv;
        unsigned int off = 0u;
        if (lane == 0) off = atomicAdd(ticket + v, (unsigned int)chunk);
        off = __shfl_sync(0xffffffffu, off, 0);
        if (off >= lv) continue;                              // another warp took the last chunk meanwhile
        const size_t qb = sv + off;
        const size_t qe = min(sv + lv, qb + (size_t)chunk);
        for (size_t q = qb; q < qe; ++q) {
            if (active && active[q].w == 0.f) {
                if (lane < 3) irr3[3 * q + lane] = 0.f;
                continue;
            }
            const float px = pos3[3 * q], py = pos3[3 * q + 1], pz = pos3[3 * q + 2];
            const float nx = nrm3[3 * q], ny = nrm3[3 * q + 1], nz = nrm3[3 * q + 2];
            float r2_seed = full_r2;
            bool seeded = false;
            if (seed_on && prev_valid) {
                bool all_face = true;
                float dmax = 0.f;
                for (int i = (int)lane; i < kmax; i += 32) {
                    const F8 ph = ld256(search + 2 * (size_t)sh.cand_id[i]);
                    float t = xsub(ph.lo.x, px);
                    float d = xmul(t, t);
                    t = xsub(ph.lo.y, py); d = xadd(d, xmul(t, t));
                    t = xsub(ph.lo.z, pz); d = xadd(d, xmul(t, t));
                    all_face = all_face && xdot(ph.hi.x, ph.hi.y, ph.hi.z, nx, ny, nz) < 0.0f;
                    dmax = fmaxf(dmax, d);
                }
                all_face = __all_sync(0xffffffffu, all_face);
                dmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(dmax)));
                if (all_face) {
                    r2_seed = fmaxf(__fmul_ru(dmax, 1.000001f), 1e-37f);    // strictly above the farthest of them
                    seeded = r2_seed < full_r2;
                }
            }
            __syncwarp();
            float r2;                   // warp-uniform: current squared search radius
            bool overflowed;            // more than k photons have been accepted (the reference's heap was built)
            int ncand;
            for (;;) {
                r2 = seeded ? r2_seed : full_r2;
                overflowed = false;
                int sp = 0;
                ncand = 0;
                if (stored >= 1) {
                    if (lane == 0) { sh.stack_node[0] = 1u; sh.stack_bound[0] = 0.f; }
                    sp = 1;
                }
                __syncwarp();
                while (sp > 0) {
                    // NPL nodes per lane and iteration: that many loads in flight per round trip to L2.  Near the stack's
                    // capacity fall back to 32, then to one node per iteration (plain DFS grows by at most one entry per level).
                    const int m = sp > MIRO_GW_STACK - 32 * NPL - 96 ? 1 : (sp > MIRO_GW_STACK - 448 ? min(sp, 32) : min(sp, 32 * NPL));
                    uint32_t node[NPL];
                    float bound[NPL];
                    bool act[NPL];
#pragma unroll
                    for (int b = 0; b < NPL; ++b) {
                        const int e = (int)lane + 32 * b;
                        node[b] = 0u; bound[b] = 0.f;
                        act[b] = e < m;
                        if (act[b]) { node[b] = sh.stack_node[sp - 1 - e]; bound[b] = sh.stack_bound[sp - 1 - e]; act[b] = bound[b] < r2; }
                    }
                    sp -= m;
                    __syncwarp();
                    F8 ph[NPL];   // one 256-bit load per node: position, plane bits and direction share a sector
#pragma unroll
                    for (int b = 0; b < NPL; ++b) {
                        if (act[b]) ph[b] = ld256(search + 2 * (size_t)node[b]);
                        else ph[b].lo = ph[b].hi = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    bool push_near[NPL], push_far[NPL], accept[NPL];
                    uint32_t near_child[NPL], far_child[NPL];
                    float far_bound[NPL], dist2[NPL];
#pragma unroll
                    for (int b = 0; b < NPL; ++b) {
                        push_near[b] = push_far[b] = accept[b] = false;
                        near_child[b] = far_child[b] = 0u; far_bound[b] = dist2[b] = 0.f;
                        if (!act[b]) continue;
                        if ((int)node[b] < half_stored) {
                            const uint32_t plane = __float_as_uint(ph[b].lo.w) & 3u;
                            const float qc = plane == 0 ? px : (plane == 1 ? py : pz);
                            const float pc = plane == 0 ? ph[b].lo.x : (plane == 1 ? ph[b].lo.y : ph[b].lo.z);
                            const float dist1 = xsub(qc, pc);
                            near_child[b] = dist1 > 0.0f ? 2 * node[b] + 1 : 2 * node[b];
                            far_child[b] = dist1 > 0.0f ? 2 * node[b] : 2 * node[b] + 1;
                            push_near[b] = true;
                            const float p2 = xmul(dist1, dist1);
                            push_far[b] = p2 < r2;                    // PhotonMap.cpp:169,172
                            far_bound[b] = fmaxf(bound[b], p2);
                        }
                        float t = xsub(ph[b].lo.x, px);
                        dist2[b] = xmul(t, t);
                        t = xsub(ph[b].lo.y, py); dist2[b] = xadd(dist2[b], xmul(t, t));
                        t = xsub(ph[b].lo.z, pz); dist2[b] = xadd(dist2[b], xmul(t, t));
                        accept[b] = dist2[b] < r2 && xdot(ph[b].hi.x, ph[b].hi.y, ph[b].hi.z, nx, ny, nz) < 0.0f;   // PhotonMap.cpp:183-186
                    }
                    // children: far sides first, near sides on top of them (the near side of the topmost node ends on top)
#pragma unroll
                    for (int b = NPL - 1; b >= 0; --b) {
                        const unsigned fm = __ballot_sync(0xffffffffu, push_far[b]);
                        if (push_far[b]) { const int o = sp + __popc(fm & lt); sh.stack_node[o] = far_child[b]; sh.stack_bound[o] = far_bound[b]; }
                        sp += __popc(fm);
                    }
#pragma unroll
                    for (int b = NPL - 1; b >= 0; --b) {
                        const unsigned nm = __ballot_sync(0xffffffffu, push_near[b]);
                        if (push_near[b]) { const int o = sp + __popc(nm & lt); sh.stack_node[o] = near_child[b]; sh.stack_bound[o] = bound[b]; }
                        sp += __popc(nm);
                    }
#pragma unroll
                    for (int b = 0; b < NPL; ++b) {
                        const unsigned am = __ballot_sync(0xffffffffu, accept[b]);
                        if (accept[b]) { const int o = ncand + __popc(am & lt); sh.cand_d2[o] = dist2[b]; sh.cand_id[o] = node[b]; }
                        ncand += __popc(am);
                    }
                    __syncwarp();
                    if (ncand > MIRO_GW_CAND - 32 * NPL) {   // the next iteration could overflow the buffer: keep the k nearest
                        r2 = gather_select(sh, ncand, kmax, lane);
                        overflowed = true;
                        // the k-th itself stays in the buffer; later photons must be strictly closer than the radius, as in the reference
                    }
                }
                if (ncand > kmax) { r2 = gather_select(sh, ncand, kmax, lane); overflowed = true; }
                // A seeded walk that ends with exactly k photons cannot tell "exactly k facing photons exist" (the reference then
                // never builds its heap and divides by max_dist^2) from "more lie beyond the seed radius": walk again unseeded.
                if (seeded && !overflowed) { seeded = false; continue; }
                break;
            }
            float sx = 0.f, sy = 0.f, sz = 0.f;
            for (int i = (int)lane; i < ncand; i += 32) {
                const float4 pw = __ldg(photons + 2 * (size_t)sh.cand_id[i] + 1);
                sx += pw.x; sy += pw.y; sz += pw.z;
            }
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
            }
            // density estimate over the k-th nearest distance if the k-set overflowed, else over max_dist (PhotonMap.cpp:136)
            const float tmp = (float)(((double)1.0f / 3.14159265358979323846) / (double)(overflowed ? r2 : full_r2));
            if (lane == 0) { irr3[3 * q] = xmul(sx, tmp); irr3[3 * q + 1] = xmul(sy, tmp); irr3[3 * q + 2] = xmul(sz, tmp); }
            prev_valid = ncand == kmax;
            __syncwarp();
        }
    }
}

cudaError_t photon_gather_launch(const PhotonMapDevice& pm, const float* d_pos3, const float* d_normal3, size_t n, float max_dist,
                                 int k, float* d_irrad3, cudaStream_t st, const float4* active, const uint32_t* d_n)
{
    if (n == 0) return cudaSuccess;
    if (!pm.d_photons) return cudaMemsetAsync(d_irrad3, 0, n * 12, st);   // empty map: zero irradiance, like a map with no photons
    if (pm.exact || k + 128 > MIRO_GW_CAND || k < 1) {
        k_photon_gather<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pm.d_photons, pm.d_tables, pm.stored, pm.half_stored, d_pos3, d_normal3,
                                                                      active, n, d_n, max_dist, k, d_irrad3);
        return cudaGetLastError();
    }
    const size_t smem = sizeof(GatherWarpShared) * MIRO_GW_WARPS;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static std::atomic<bool> attr_set[64];   // the opt-in to > 48 KB of dynamic shared memory is per device
    if (dev < 0 || dev >= 64 || !attr_set[dev].load()) {
        cudaError_t e = cudaFuncSetAttribute(k_photon_gather_warp<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_photon_gather_warp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev].store(true);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // knobs (defaults measured on config 5): queries per claim, neighbour seeding, nodes per lane and iteration.  A chunk is the
    // unit of load balance and some queries cost 100x the median (a ceiling few photons face: 1400 iterations).  Since a warp
    // walks its home segment in order, seeding does not depend on the chunk size, and short chunks balance best: 1 / 2 / 4 / 8 /
    // 16 queries give 37.9 / 36.6 / 37.0 / 38.8 / 45.6 ms on the global map; 4 nodes per lane lose to 2.
    static const int chunk = std::max(1, env_int("MIROGPU_GATHER_CHUNK", 2));
    static const int seed_on = env_int("MIROGPU_GATHER_SEED", 1);
    static const int npl = env_int("MIROGPU_GATHER_NPL", 2);
    const size_t want = ((n + chunk - 1) / chunk + MIRO_GW_WARPS - 1) / MIRO_GW_WARPS;
    static std::atomic<unsigned> launch_seq{0};
    unsigned int* ticket = pm.d_tickets + (size_t)(launch_seq.fetch_add(1) % MIRO_GW_TICKET_SLOTS) * MIRO_GW_TICKET_SPAN;   // one cursor array per launch in flight
    cudaError_t e = cudaMemsetAsync(ticket, 0, MIRO_GW_TICKET_SPAN * sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    int occ = 1;
    if (npl == 4) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_photon_gather_warp<4>, 32 * MIRO_GW_WARPS, smem);
        const unsigned grid = (unsigned)std::min<size_t>(std::min<size_t>(want, (size_t)sms * std::max(occ, 1)), MIRO_GW_TICKET_SPAN / MIRO_GW_WARPS);
        k_photon_gather_warp<4><<<grid, 32 * MIRO_GW_WARPS, smem, st>>>(pm.d_photons, pm.d_search, pm.stored, pm.half_stored, d_pos3, d_normal3,
                                                                        active, n, d_n, max_dist, k, d_irrad3, ticket, chunk, seed_on);
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_photon_gather_warp<2>, 32 * MIRO_GW_WARPS, smem);
        const unsigned grid = (unsigned)std::min<size_t>(std::min<size_t>(want, (size_t)sms * std::max(occ, 1)), MIRO_GW_TICKET_SPAN / MIRO_GW_WARPS);   // persistent warps
        k_photon_gather_warp<2><<<grid, 32 * MIRO_GW_WARPS, smem, st>>>(pm.d_photons, pm.d_search, pm.stored, pm.half_stored, d_pos3, d_normal3,
                                                                        active, n, d_n, max_dist, k, d_irrad3, ticket, chunk, seed_on);
    }
    return cudaGetLastError();
}

}  // namespace mirogpu
#endif
