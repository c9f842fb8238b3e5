// kernels.cuh -- sm_100a kernels of the ray-intersection engine (included once by mirogpu.cu).
//
//   k_trace_hybrid      the default: persistent warps that vote every iteration between node steps and a leaf phase and hand
//                       finished lanes the next rays of the warp's pool (see the comment above the kernel)
//   k_trace_simple      one thread per ray, grid covers the batch (bring-up / comparison variant)
//   k_trace_persistent  persistent warps: grid = SMs x resident CTAs, every warp pulls 32-ray packets from
//                       a global ticket counter until the batch is drained (BVH2 coherent batches, CWBVH8)
//   k_gen_primary       Camera::eyeRay for a block of rows (Camera.cpp:104-161)
//   k_gen_bounce        Ray::diffuse at every hit (Ray.h:109-122, Utility.h:34-50)
//   k_resolve_hits      P, N, material from (prim, beta, gamma) (Triangle.cpp:160-166, Scene.cpp:262)
#ifndef MIROGPU_KERNELS_CUH
#define MIROGPU_KERNELS_CUH

#include <type_traits>
// 1: the hybrid kernel keeps a lane's prim / beta / gamma and ray index in shared memory (BestHitSm, traverse.cuh) instead of four
// registers.  Removes the node step's spills at the 48-register budget, and measured 1.3 % slower on the bench step (8.97 -> 8.85
// Grays/s: 2 KB of shared memory per CTA come out of L1, and the accesses are ALU-pipe address arithmetic) -- off by default.
#ifndef MIRO_BEST_SM
#define MIRO_BEST_SM 0
#endif
#include "traverse.cuh"
#include "rng.cuh"
#include "texture.cuh"

namespace mirogpu {

struct DeviceScene {
    const float4* nodes;   // BVH2: 4 float4 per node; BVH4: 8 float4 per node; QBVH4: 4 float4 per node; CWBVH8: 5 uint4 per node
    const float4* tris;    // 4 float4 (64 B) per triangle, leaf order: (A, prim) (B-A, n.x) (C-A, n.y) (n.z, -, -, -)
    const float4* shade;   // 6 float4 per primitive, prim-id order: (A,mat) (e1,kind) e2 nA nB nC; kind 1 sphere (centre,mat) (radius,1); 2 plane
    const float4* planes;  // 2 float4 per plane: (normal, prim id bits) (origin, 0) -- the unbounded objects of Scene::trace (Scene.cpp:219-230)
    uint32_t num_tris;     // all primitives (triangles, spheres, planes): the range of prim ids
    uint32_t num_planes;
    // scenes with TexturedPhong materials (texture.cuh): hit points are finished by resolve_hit_textured
    const mirogpu_material* mats;
    const float* uvs;      // 6 floats per triangle (tA, tB, tC of TriangleMesh::texCoords) or NULL: meshes without texture coordinates
    uint32_t textured;
};

struct CameraBasis {       // what Camera::eyeRay caches in its statics (Camera.cpp:106-125)
    float eye[3], w[3], u[3], v[3];
    float top, left, bottom, right;
};

#define MIRO_PI 3.1415926535897932384626433832795028841972f /* Miro.h:10 */

// Scene::trace's loop over the unbounded objects (Scene.cpp:219-230) with Plane::intersect (Plane.cpp:33-48) inlined: a plane
// replaces the tree's answer only if strictly closer, or -- nothing hit so far -- anywhere inside [tMin, tMax]; planes are
// tried in insertion order.  fabs(ndotd) < 1e-6 is a double comparison in the reference.  Out of line and by value (see
// sphere_hit_t): returns (t, prim id bits) of the answer after the planes.
__device__ __noinline__ float2 planes_hit(const float4* __restrict__ planes, const uint32_t nplanes, const mirogpu_ray r, float best_t, uint32_t best_prim)
{
    if (r.tmax >= r.tmin)
        for (uint32_t k = 0; k < nplanes; ++k) {
            const float4 n = __ldg(planes + 2 * k), o = __ldg(planes + 2 * k + 1);
            const float ndotd = xdot(n.x, n.y, n.z, r.dx, r.dy, r.dz);
            if ((double)fabsf(ndotd) < 1e-6) continue;
            const float t = xdiv(xdot(n.x, n.y, n.z, xsub(o.x, r.ox), xsub(o.y, r.oy), xsub(o.z, r.oz)), ndotd);
            if (t < r.tmin || t > r.tmax) continue;
            if (best_prim == MIROGPU_MISS || t < best_t) { best_t = t; best_prim = __float_as_uint(n.w); }
        }
    return make_float2(best_t, __uint_as_float(best_prim));
}
template <typename BEST>
__device__ __forceinline__ void planes_test(const DeviceScene& s, const mirogpu_ray& r, BEST& best)
{
    const float2 a = planes_hit(s.planes, s.num_planes, r, best.t, best.prim);
    if (__float_as_uint(a.y) != best.prim) { best.t = a.x; best.prim = __float_as_uint(a.y); best.beta = 0.f; best.gamma = 0.f; }
}

// NT: the scene holds non-triangle primitives (sphere slots in the leaf array, planes after the walk).
template <int LAYOUT, bool ANY, bool COUNT, bool NT = false>
__device__ __forceinline__ void trace_one(const DeviceScene& s, const mirogpu_ray& r, BestHit& best, TraceCounters* c)
{
    if (LAYOUT == MIROGPU_LAYOUT_BVH2) trace_bvh2<ANY, COUNT, NT>(s.nodes, s.tris, r, best, c);
    else if (LAYOUT == MIROGPU_LAYOUT_BVH4) trace_bvh4<ANY, COUNT, NT>(s.nodes, s.tris, r, best, c);
    else if (LAYOUT == MIROGPU_LAYOUT_QBVH4) trace_qbvh4<ANY, COUNT, NT>(s.nodes, s.tris, r, best, c);
    else trace_cwbvh8<ANY, COUNT, NT>(reinterpret_cast<const uint4*>(s.nodes), s.tris, r, best, c);
    if (NT && s.num_planes) planes_test(s, r, best);
}

// Rays and hits are streamed once per launch: loads/stores carry the evict-first hint so the batch does not push
// the scene (nodes + triangles, which every ray re-reads) out of L2.
__device__ __forceinline__ mirogpu_ray load_ray(const mirogpu_ray* rays, size_t i)
{
    const float4* p = reinterpret_cast<const float4*>(rays + i);
    const float4 a = __ldcs(p), b = __ldcs(p + 1);
    mirogpu_ray r;
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tmin = a.w; r.dx = b.x; r.dy = b.y; r.dz = b.z; r.tmax = b.w;
    return r;
}
template <typename BEST>
__device__ __forceinline__ void store_hit(mirogpu_hit* hits, size_t i, const BEST& b)
{
    float4 h; h.x = b.t; h.y = __uint_as_float(b.prim); h.z = b.beta; h.w = b.gamma;
    __stcs(reinterpret_cast<float4*>(hits + i), h);
}

template <int LAYOUT, bool ANY, bool COUNT, bool NT = false>
__global__ void __launch_bounds__(128) k_trace_simple(DeviceScene s, const mirogpu_ray* __restrict__ rays, size_t n,
                                                      mirogpu_hit* __restrict__ hits, unsigned long long* __restrict__ counters,
                                                      const uint32_t* __restrict__ d_n, uint32_t mult)
{
    if (d_n) n = min(n, (size_t)*d_n * mult);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    TraceCounters c = {0, 0, 0};
    uint32_t hit = 0;
    if (i < n) {
        const mirogpu_ray r = load_ray(rays, i);
        BestHit best;
        trace_one<LAYOUT, ANY, COUNT, NT>(s, r, best, &c);
        store_hit(hits, i, best);
        hit = best.prim != MIROGPU_MISS;
    }
    if (COUNT) {
        // warp-aggregate, then one atomic per warp per counter
        for (int o = 16; o > 0; o >>= 1) {
            c.nodes += __shfl_down_sync(0xffffffffu, c.nodes, o);
            c.boxes += __shfl_down_sync(0xffffffffu, c.boxes, o);
            c.tris += __shfl_down_sync(0xffffffffu, c.tris, o);
            hit += __shfl_down_sync(0xffffffffu, hit, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + 0, (unsigned long long)c.nodes);
            atomicAdd(counters + 1, (unsigned long long)c.boxes);
            atomicAdd(counters + 2, (unsigned long long)c.tris);
            atomicAdd(counters + 3, (unsigned long long)hit);
        }
    }
}

// Persistent warps.  The grid is sized to the machine (SM count x resident CTAs), not to the batch; each
// warp takes a ticket for the next 32 consecutive rays.  Consecutive tickets keep primary rays of one
// warp coherent; for incoherent batches the packet order is irrelevant and the ticket loop removes the
// tail effect of uneven ray costs across CTAs.
template <int LAYOUT, bool ANY, bool NT = false>
__global__ void __launch_bounds__(128) k_trace_persistent(DeviceScene s, const mirogpu_ray* __restrict__ rays, size_t n,
                                                          mirogpu_hit* __restrict__ hits, unsigned long long* __restrict__ ticket,
                                                          const uint32_t* __restrict__ d_n, uint32_t mult, uint32_t packets_per_ticket)
{
    if (d_n) n = min(n, (size_t)*d_n * mult);
    const unsigned lane = threadIdx.x & 31u;
    // A ticket may be worth several consecutive 32-ray packets (measured: 1, 2, 4, 8 packets per ticket are within 1.5 %
    // of each other -- the ticket atomic is not what bounds this kernel -- so the default stays 1 for the smallest tail).
    const unsigned long long grain = 32ull * packets_per_ticket;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(ticket, grain);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) return;
        for (uint32_t k = 0; k < packets_per_ticket; ++k) {
            const size_t i = (size_t)base + 32u * k + lane;
            if (i < n) {
                const mirogpu_ray r = load_ray(rays, i);
                BestHit best;
                trace_one<LAYOUT, ANY, false, NT>(s, r, best, nullptr);
                store_hit(hits, i, best);
            }
            __syncwarp();
        }
    }
}

// Persistent warps, hybrid step scheduling + ray replacement (BVH2, BVH4, QBVH4).  In the while-while kernel a round costs the
// warp max-over-lanes(descent length) node steps plus max-over-lanes(leaf size) triangle tests, and on incoherent
// rays the descent lengths differ so much that 7 of 32 lanes are busy on average (ncu; tools/simt_sim.cu replays
// the same rays and predicts the same figure).  Here every iteration the warp votes: it takes node steps while at
// least `nmin` lanes want one (or nobody waits at a leaf), otherwise the lanes waiting at leaves test their
// triangles.  Every `period` iterations, if at least `min_idle` lanes have finished their ray, those lanes pull the
// next rays of the warp's pool by ballot + popc-prefix (warp-level ray compaction); pools of `pool` consecutive rays
// are claimed from a global ticket one ahead of use and prefetched into L2 while the warp works on the current one.
//   PF    prefetch flags of bvh2_node_step; bit 3: prefetch the pool claimed ahead         MINB  resident CTAs per SM asked of the compiler
//   NREP  node steps per vote (a lane that leaves the inner nodes sits the rest out)
//   SHORT stack entries per lane kept in shared memory ([entry][thread], conflict-free), deeper ones in local memory (0: all local)
//   STAGE QBVH4 only: the first STAGE nodes (top levels, breadth-first numbering) copied into shared memory by every CTA
template <int LAYOUT, bool ANY, int PF, int MINB, int NREP, int SHORT = 0, int STAGE = 0, bool NT = false>
__global__ void __launch_bounds__(128, MINB) k_trace_hybrid(DeviceScene s, const mirogpu_ray* __restrict__ rays, size_t n,
                                                                 mirogpu_hit* __restrict__ hits, unsigned long long* __restrict__ ticket,
                                                                 int nmin, int period, int min_idle, uint32_t pool,
                                                                 const uint32_t* __restrict__ d_n, uint32_t mult, uint32_t num_nodes)
{
    constexpr int DEPTH = (LAYOUT == MIROGPU_LAYOUT_BVH2 ? MIRO_STACK : MIRO_STACK4) + 1;
    __shared__ int32_t s_stack[SHORT > 0 ? SHORT * 128 : 1];
    __shared__ float4 s_top[STAGE > 0 ? STAGE * 4 : 1];
    if (STAGE > 0) {
        // nodes past the end of a small tree are never addressed (links only name existing nodes)
        const uint32_t cnt = min((uint32_t)STAGE, num_nodes) * 4u;
        for (uint32_t k = threadIdx.x; k < cnt; k += 128) s_top[k] = __ldg(s.nodes + k);
        __syncthreads();
    }
    if (d_n) n = min(n, (size_t)*d_n * mult);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    // 32-bit ray indices (the launcher splits batches of 2^32 rays or more)
    const uint32_t n32 = (uint32_t)n;
#if MIRO_BEST_SM
    typename std::conditional<NT, uint32_t, BestSmSlot<3, uint32_t>>::type my;
    my = 0u;
#else
    uint32_t my = 0;
#endif
    uint32_t cur = 0, cur_end = 0, nxt;   // warp-uniform: the pool in use [cur, cur_end) and the base of the pool claimed ahead
    mirogpu_ray r;
    Bvh2Walk w;
#if MIRO_BEST_SM
    typename std::conditional<NT, BestHit, BestHitSm>::type best;
#else
    BestHit best;
#endif
    int32_t pleaf = MIRO_BVH2_DONE, pleaf2 = MIRO_BVH2_DONE;   // PF bits 6, 7: postponed leaves (none otherwise; the compiler drops them)
    SplitStack<SHORT, DEPTH, 128> stack;
    stack.sm = s_stack + threadIdx.x;
    r.ox = r.oy = r.oz = r.tmin = r.dx = r.dy = r.dz = r.tmax = 0.f;
    w.node = w.tos = MIRO_BVH2_DONE; w.sp = 0;
    w.idx = w.idy = w.idz = w.oodx = w.oody = w.oodz = 0.f;
    best.t = 0.f; best.prim = MIROGPU_MISS; best.beta = 0.f; best.gamma = 0.f;
    // Pools shrink near the end of the batch (guided self-scheduling): once fewer than two full pools per warp of the grid
    // remain, warps claim 16 rays at a time, so the last warps to finish are at most a quarter-pool apart.
    const uint32_t tail = 2u * pool * (gridDim.x * (blockDim.x >> 5));
    uint32_t nxt_len = 0;
    auto claim = [&](uint32_t seen) -> uint32_t {
        const uint32_t grain = (n32 - min(seen, n32) < tail) ? min(pool, 16u) : pool;
        unsigned long long base64 = 0;
        if (lane == 0) base64 = atomicAdd(ticket, (unsigned long long)grain);
        const uint32_t base = (uint32_t)min(base64, (unsigned long long)n32);
        const uint32_t b = __shfl_sync(0xffffffffu, base, 0);
        nxt_len = min(grain, n32 - b);
        // one 128-byte line holds four rays
        if (PF & 8) for (uint32_t k = 4 * lane; k < nxt_len; k += 128) prefetch_l2(rays + b + k);
        return b;
    };
    nxt = claim(0u);
    for (;;) {
        // ---- hand new rays to idle lanes (two passes, so a lane that drew a dead ray gets another) ----
        unsigned idle = __ballot_sync(0xffffffffu, w.node == MIRO_BVH2_DONE && pleaf == MIRO_BVH2_DONE);
        if (__popc(idle) >= min_idle) {
            for (int pass = 0; pass < 2 && idle; ++pass) {
                if (cur >= cur_end) {
                    if (nxt >= n32) break;   // drained
                    cur = nxt; cur_end = nxt + nxt_len;
                    nxt = claim(cur);
                }
                if (w.node == MIRO_BVH2_DONE && pleaf == MIRO_BVH2_DONE) {
                    const uint32_t i = cur + __popc(idle & lt_mask);
                    if (i < cur_end) {
                        r = load_ray(rays, i);
                        bvh2_begin(r, w, best);
                        if (w.node == MIRO_BVH2_DONE) store_hit(hits, i, best);   // empty interval: answered without traversal
                        else my = i;
                    }
                }
                cur = min(cur + (uint32_t)__popc(idle), cur_end);
                idle = __ballot_sync(0xffffffffu, w.node == MIRO_BVH2_DONE && pleaf == MIRO_BVH2_DONE);
            }
            if (idle == 0xffffffffu) {
                if (cur >= cur_end && nxt >= n32) return;
                continue;
            }
        }
        if (PF & 64) {
            // Postponed leaves: a lane whose descent reaches a leaf parks it in `pleaf` (PF bit 7: a second one in `pleaf2`) and walks
            // on (its best.t is stale until the parked leaf is tested, so it may enter a few nodes it would have culled); it only
            // stops for a leaf phase when one more leaf turns up than it can park.  More lanes take part in every node step, and
            // leaf phases start with more lanes holding a leaf.
            for (int it = 0; it < period; ++it) {
                const bool at_leaf = w.node < 0 && w.node != MIRO_BVH2_DONE;
                const unsigned mn = __ballot_sync(0xffffffffu, w.node >= 0);
                const unsigned ml = __ballot_sync(0xffffffffu, pleaf != MIRO_BVH2_DONE || at_leaf);
                if ((mn | ml) == 0u) break;
                if (mn != 0u && (__popc(mn) >= nmin || ml == 0u)) {
#pragma unroll
                    for (int rep = 0; rep < NREP; ++rep)
                        if (w.node >= 0) {
                            qbvh4_node_step<PF, decltype(stack), STAGE>(s.nodes, s.tris, r, w, stack, best, s_top);
                            if (w.node < 0 && w.node != MIRO_BVH2_DONE) {
                                if (pleaf == MIRO_BVH2_DONE) { pleaf = w.node; bvh2_pop(w, stack); }
                                else if ((PF & 128) && pleaf2 == MIRO_BVH2_DONE) { pleaf2 = w.node; bvh2_pop(w, stack); }
                            }
                        }
                } else {
                    if (pleaf == MIRO_BVH2_DONE && at_leaf) { pleaf = w.node; bvh2_pop(w, stack); }
                    if (pleaf != MIRO_BVH2_DONE) {
                        bool hit;
                        pleaf = leaf_ref_test_one<NT>(s.tris, r, pleaf, best, hit);
                        if (ANY && hit) { pleaf = pleaf2 = MIRO_BVH2_DONE; w.node = MIRO_BVH2_DONE; }
                        else if ((PF & 128) && pleaf == MIRO_BVH2_DONE) { pleaf = pleaf2; pleaf2 = MIRO_BVH2_DONE; }
                    }
                }
                if (w.node == MIRO_BVH2_DONE && pleaf == MIRO_BVH2_DONE && ((mn | ml) >> lane & 1u)) { if (NT && s.num_planes) planes_test(s, r, best); store_hit(hits, my, best); }
            }
            continue;
        }
        for (int it = 0; it < period; ++it) {
            const unsigned mn = __ballot_sync(0xffffffffu, w.node >= 0);
            const unsigned ml = __ballot_sync(0xffffffffu, w.node < 0 && w.node != MIRO_BVH2_DONE);
            if ((mn | ml) == 0u) break;
            if (mn != 0u && (__popc(mn) >= nmin || ml == 0u)) {
#pragma unroll
                for (int rep = 0; rep < NREP; ++rep)
                    if (w.node >= 0) {
                        if (LAYOUT == MIROGPU_LAYOUT_QBVH4) qbvh4_node_step<PF, decltype(stack), STAGE>(s.nodes, s.tris, r, w, stack, best, s_top);
                        else if (LAYOUT == MIROGPU_LAYOUT_BVH4) bvh4_node_step<PF>(s.nodes, s.tris, r, w, stack, best);
                        else bvh2_node_step<PF>(s.nodes, s.tris, r, w, stack, best);
                    }
            } else if (w.node < 0 && w.node != MIRO_BVH2_DONE) {
                if (PF & 32) bvh2_leaf_step_two<ANY, NT>(s.tris, r, w, stack, best);
                else if (PF & 16) bvh2_leaf_step_one<ANY, NT>(s.tris, r, w, stack, best);
                else bvh2_leaf_step<ANY, NT>(s.tris, r, w, stack, best);
            }
            if (w.node == MIRO_BVH2_DONE && ((mn | ml) >> lane & 1u)) { if (NT && s.num_planes) planes_test(s, r, best); store_hit(hits, my, best); }
        }
    }
}

// ---- Camera::eyeRay, Camera.cpp:127-160 (no DOF).  Same operand order, no FMA: bit-exact with the oracle. ----
__global__ void __launch_bounds__(256) k_gen_primary(CameraBasis cb, int width, int height, int first_row, int row_stride,
                                                      int nrows_local, int jitter, uint32_t seed, uint32_t sample_begin,
                                                      uint32_t sample_count, mirogpu_ray* __restrict__ rays)
{
    // grid: x over the shard's pixels, y over the samples -- 32-bit index arithmetic (one division), no 64-bit div / mod
    const uint32_t npix = (uint32_t)nrows_local * (uint32_t)width;
    const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= npix) return;
    const uint32_t s = blockIdx.y;
    const size_t i = (size_t)s * npix + lp;
    const uint32_t row = lp / (uint32_t)width;
    const int x = (int)(lp - row * (uint32_t)width), y = first_row + (int)row * row_stride;
    float dx = 0.5f, dy = 0.5f;
    if (jitter) uniform2(seed, (uint32_t)((size_t)y * width + x), sample_begin + s, RNG_DIM_PIXEL, dx, dy);
    const float U = xadd(cb.left, xmul(xsub(cb.right, cb.left), xdiv(xadd((float)x, dx), (float)width)));
    const float V = xadd(cb.bottom, xmul(xsub(cb.top, cb.bottom), xdiv(xadd((float)y, dy), (float)height)));
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = xsub(xadd(xmul(cb.u[k], U), xmul(cb.v[k], V)), cb.w[k]);
    const float inv = xdiv(1.0f, xsqrt(xdot(d[0], d[1], d[2], d[0], d[1], d[2])));
    float4 a, b;
    a.x = cb.eye[0]; a.y = cb.eye[1]; a.z = cb.eye[2]; a.w = 0.0f;
    b.x = xmul(d[0], inv); b.y = xmul(d[1], inv); b.z = xmul(d[2], inv); b.w = MIROGPU_TMAX;
    float4* o = reinterpret_cast<float4*>(rays + i);
    o[0] = a; o[1] = b;
}

struct SurfacePoint {
    float P[3], N[3];
    uint32_t material;
};

// Triangle.cpp:160-162 then Scene::trace's normalisation (Scene.cpp:262; all materials are UV-lookup Phong
// with zero bump height, so the perturbation term is exactly zero).
struct ShadeRecord {   // 96 bytes, 32-byte aligned: three 256-bit loads (three L1 wavefronts for a divergent warp instead of six)
    F8 r0, r1, r2;
};
__device__ __forceinline__ ShadeRecord load_shade_record(const DeviceScene& s, uint32_t prim)
{
    const float4* q = s.shade + 6 * (size_t)prim;
    ShadeRecord r;
    r.r0 = ld256(q); r.r1 = ld256(q + 2); r.r2 = ld256(q + 4);
    return r;
}
__device__ __forceinline__ uint32_t record_kind(const ShadeRecord& rec) { return __float_as_uint(rec.r0.hi.w); }   // 0 triangle, 1 sphere, 2 plane

__device__ __forceinline__ SurfacePoint resolve_hit(const ShadeRecord& rec, const mirogpu_hit& h)
{
    const float4 A = rec.r0.lo, e1 = rec.r0.hi, e2 = rec.r1.lo, nA = rec.r1.hi, nB = rec.r2.lo, nC = rec.r2.hi;
    SurfacePoint sp;
    sp.P[0] = xadd(xadd(A.x, xmul(e1.x, h.beta)), xmul(e2.x, h.gamma));
    sp.P[1] = xadd(xadd(A.y, xmul(e1.y, h.beta)), xmul(e2.y, h.gamma));
    sp.P[2] = xadd(xadd(A.z, xmul(e1.z, h.beta)), xmul(e2.z, h.gamma));
    const float alpha = xsub(xsub(1.0f, h.beta), h.gamma);
    float n[3];
    n[0] = xadd(xadd(xmul(nA.x, alpha), xmul(nB.x, h.beta)), xmul(nC.x, h.gamma));
    n[1] = xadd(xadd(xmul(nA.y, alpha), xmul(nB.y, h.beta)), xmul(nC.y, h.gamma));
    n[2] = xadd(xadd(xmul(nA.z, alpha), xmul(nB.z, h.beta)), xmul(nC.z, h.gamma));
    const float inv = xdiv(1.0f, xsqrt(xdot(n[0], n[1], n[2], n[0], n[1], n[2])));
    sp.N[0] = xmul(n[0], inv); sp.N[1] = xmul(n[1], inv); sp.N[2] = xmul(n[2], inv);
    sp.material = __float_as_uint(A.w);
    return sp;
}
// Spheres and planes: P = o + t d (Sphere.cpp:62, Plane.cpp:42); sphere N = (P - c).normalize() (Sphere.cpp:63-64), plane N =
// its normal as given (Plane.cpp:44); Scene::trace then normalises N once more (Scene.cpp:262).
__device__ __forceinline__ SurfacePoint resolve_hit_analytic(const ShadeRecord& rec, const mirogpu_hit& h, float ox, float oy, float oz, float dx, float dy, float dz)
{
    SurfacePoint sp;
    sp.P[0] = xadd(ox, xmul(dx, h.t)); sp.P[1] = xadd(oy, xmul(dy, h.t)); sp.P[2] = xadd(oz, xmul(dz, h.t));
    float n[3];
    if (record_kind(rec) == 1u) {
        n[0] = xsub(sp.P[0], rec.r0.lo.x); n[1] = xsub(sp.P[1], rec.r0.lo.y); n[2] = xsub(sp.P[2], rec.r0.lo.z);
        const float i0 = xdiv(1.0f, xsqrt(xdot(n[0], n[1], n[2], n[0], n[1], n[2])));
        n[0] = xmul(n[0], i0); n[1] = xmul(n[1], i0); n[2] = xmul(n[2], i0);
    } else { n[0] = rec.r1.hi.x; n[1] = rec.r1.hi.y; n[2] = rec.r1.hi.z; }
    const float inv = xdiv(1.0f, xsqrt(xdot(n[0], n[1], n[2], n[0], n[1], n[2])));
    sp.N[0] = xmul(n[0], inv); sp.N[1] = xmul(n[1], inv); sp.N[2] = xmul(n[2], inv);
    sp.material = __float_as_uint(rec.r0.lo.w);
    return sp;
}
// Any primitive kind; the ray is only read (through `ray`) for the analytic kinds.
__device__ __forceinline__ SurfacePoint resolve_hit(const ShadeRecord& rec, const mirogpu_hit& h, const mirogpu_ray* ray)
{
    if (record_kind(rec) == 0u) return resolve_hit(rec, h);
    const float4* p = reinterpret_cast<const float4*>(ray);
    const float4 a = p[0], b = p[1];
    return resolve_hit_analytic(rec, h, a.x, a.y, a.z, b.x, b.y, b.z);
}
__device__ __forceinline__ SurfacePoint resolve_hit(const ShadeRecord& rec, const mirogpu_hit& h, const mirogpu_ray& r)
{
    if (record_kind(rec) == 0u) return resolve_hit(rec, h);
    return resolve_hit_analytic(rec, h, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
}

// ---- scenes with textured materials ----------------------------------------------------------------------------------------
// Scene::trace's treatment of a hit (Scene.cpp:232-262) in full, and the diffuse colour Phong::shade / Scene::tracePhoton look up:
//   * the normal as the primitive left it -- a triangle's interpolated normal un-normalised (Triangle.cpp:162), a sphere's
//     normalised (Sphere.cpp:63-64), a plane's as given (Plane.cpp:44);
//   * materials with UV lookup coordinates (plain Phong, 2-D textures): Object::toUVCoordinates, central differences of the bump
//     height (non-zero for StoneTexture only), the tangent construction, N.normalize();
//   * materials with UVW lookup coordinates (3-D textures): the normal stays as it is;
//   * diffuse colour: Material::diffuse2D(uv) / diffuse3D(P).
// Out of line and by value: untextured scenes (every kernel's fast path) do not pay registers for the noise code.
struct TexturedPoint { SurfacePoint sp; float dc[3]; };

__device__ __forceinline__ void object_uv(const ShadeRecord& rec, const mirogpu_hit& h, const float P[3], const float* uvs, float uv[2])
{
    const uint32_t kind = record_kind(rec);
    if (kind == 2u) { uv[0] = P[0]; uv[1] = P[2]; return; }                 // Plane::toUVCoordinates, Plane.cpp:50-60
    if (kind == 1u) {                                                       // Sphere::toUVCoordinates, Sphere.cpp:83-95
        float d[3] = {xsub(P[0], rec.r0.lo.x), xsub(P[1], rec.r0.lo.y), xsub(P[2], rec.r0.lo.z)};
        const float inv = xdiv(1.0f, xsqrt(xdot(d[0], d[1], d[2], d[0], d[1], d[2])));
        d[0] = xmul(d[0], inv); d[1] = xmul(d[1], inv); d[2] = xmul(d[2], inv);
        uv[0] = (float)dadd((double)xdiv(atan2f(d[0], d[2]), xmul(2.0f, MIRO_PI)), 0.5);
        uv[1] = (float)dadd((double)xdiv(fmaxf(-1.0f, fminf(1.0f, asinf(d[1]))), MIRO_PI), 0.5);
        return;
    }
    uv[0] = uv[1] = 0.f;                                                    // a mesh without texture coordinates: tex_coord2d_t() (Triangle.cpp:174-175)
    if (!uvs) return;
    // Triangle::toUVCoordinates (Triangle.cpp:172-222): barycentrics by Cramer's rule in the plane that drops one axis
    const float* t = uvs + 6 * (size_t)h.prim_id;
    const float4 A = rec.r0.lo, e1 = rec.r0.hi, e2 = rec.r1.lo;
    const float B[3] = {e1.x, e1.y, e1.z}, C[3] = {e2.x, e2.y, e2.z};
    const float nx = xsub(xmul(B[1], C[2]), xmul(B[2], C[1])), ny = xsub(xmul(B[2], C[0]), xmul(B[0], C[2])), nz = xsub(xmul(B[0], C[1]), xmul(B[1], C[0]));
    int i = 0, j = 1;
    if (nx > nz) i = 2; else if (ny > nz) j = 2;
    const float p[3] = {xsub(P[0], A.x), xsub(P[1], A.y), xsub(P[2], A.z)};
    auto det = [](float a, float b, float c, float d) { return xsub(xmul(a, d), xmul(b, c)); };   // Det(a, b, c, d) = a d - b c (Utility.h)
    const float detPC = det(p[i], C[i], p[j], C[j]), detBP = det(B[i], p[i], B[j], p[j]), detBC = det(B[i], C[i], B[j], C[j]);
    const float beta = fmaxf(xdiv(detPC, detBC), 0.f), gamma = fmaxf(xdiv(detBP, detBC), 0.f);
    const float alpha = fmaxf(xsub(1.0f, xadd(beta, gamma)), 0.f);
    uv[0] = xadd(xadd(xmul(alpha, t[0]), xmul(beta, t[2])), xmul(gamma, t[4]));
    uv[1] = xadd(xadd(xmul(alpha, t[1]), xmul(beta, t[3])), xmul(gamma, t[5]));
}

__device__ __noinline__ TexturedPoint resolve_hit_textured(const ShadeRecord rec, const mirogpu_hit h, const float ox, const float oy, const float oz,
                                                            const float dx, const float dy, const float dz, const mirogpu_material* __restrict__ mats,
                                                            const float* __restrict__ uvs)
{
    TexturedPoint tp;
    SurfacePoint& sp = tp.sp;
    const uint32_t kind = record_kind(rec);
    float n[3];
    if (kind == 0u) {
        const float4 A = rec.r0.lo, e1 = rec.r0.hi, e2 = rec.r1.lo, nA = rec.r1.hi, nB = rec.r2.lo, nC = rec.r2.hi;
        sp.P[0] = xadd(xadd(A.x, xmul(e1.x, h.beta)), xmul(e2.x, h.gamma));
        sp.P[1] = xadd(xadd(A.y, xmul(e1.y, h.beta)), xmul(e2.y, h.gamma));
        sp.P[2] = xadd(xadd(A.z, xmul(e1.z, h.beta)), xmul(e2.z, h.gamma));
        const float alpha = xsub(xsub(1.0f, h.beta), h.gamma);
        n[0] = xadd(xadd(xmul(nA.x, alpha), xmul(nB.x, h.beta)), xmul(nC.x, h.gamma));
        n[1] = xadd(xadd(xmul(nA.y, alpha), xmul(nB.y, h.beta)), xmul(nC.y, h.gamma));
        n[2] = xadd(xadd(xmul(nA.z, alpha), xmul(nB.z, h.beta)), xmul(nC.z, h.gamma));
        sp.material = __float_as_uint(A.w);
    } else {
        sp.P[0] = xadd(ox, xmul(dx, h.t)); sp.P[1] = xadd(oy, xmul(dy, h.t)); sp.P[2] = xadd(oz, xmul(dz, h.t));
        if (kind == 1u) {
            n[0] = xsub(sp.P[0], rec.r0.lo.x); n[1] = xsub(sp.P[1], rec.r0.lo.y); n[2] = xsub(sp.P[2], rec.r0.lo.z);
            const float i0 = xdiv(1.0f, xsqrt(xdot(n[0], n[1], n[2], n[0], n[1], n[2])));
            n[0] = xmul(n[0], i0); n[1] = xmul(n[1], i0); n[2] = xmul(n[2], i0);
        } else { n[0] = rec.r1.hi.x; n[1] = rec.r1.hi.y; n[2] = rec.r1.hi.z; }
        sp.material = __float_as_uint(rec.r0.lo.w);
    }
    const mirogpu_material m = mats[sp.material];
    float uv[2];
    object_uv(rec, h, sp.P, uvs, uv);
    const bool uvw = m.texture == MIROGPU_TEX_PETAL || m.texture == MIROGPU_TEX_LEAF || m.texture == MIROGPU_TEX_FLOWER_CENTER;
    if (!uvw) {
        if (m.texture == MIROGPU_TEX_STONE) material_bump_normal(m, uv[0], uv[1], n);
        else {
            const float inv = xdiv(1.0f, xsqrt(xdot(n[0], n[1], n[2], n[0], n[1], n[2])));
            n[0] = xmul(n[0], inv); n[1] = xmul(n[1], inv); n[2] = xmul(n[2], inv);
        }
    }
    sp.N[0] = n[0]; sp.N[1] = n[1]; sp.N[2] = n[2];
    material_diffuse_color(m, uv, sp.P, tp.dc);
    return tp;
}

__global__ void __launch_bounds__(256) k_resolve_hits(DeviceScene s, const mirogpu_ray* __restrict__ rays, const mirogpu_hit* __restrict__ hits, size_t n,
                                                       float* __restrict__ P, float* __restrict__ N, uint32_t* __restrict__ mat)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 hv = __ldg(reinterpret_cast<const float4*>(hits + i));
    mirogpu_hit h; h.t = hv.x; h.prim_id = __float_as_uint(hv.y); h.beta = hv.z; h.gamma = hv.w;
    SurfacePoint sp;
    if (h.prim_id == MIROGPU_MISS) {
        sp.P[0] = sp.P[1] = sp.P[2] = 0.f; sp.N[0] = sp.N[1] = sp.N[2] = 0.f; sp.material = MIROGPU_MISS;
    } else {
        const ShadeRecord rec = load_shade_record(s, h.prim_id);
        if (s.textured) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (rays) { const float4* q = reinterpret_cast<const float4*>(rays + i); a = q[0]; b = q[1]; }
            sp = resolve_hit_textured(rec, h, a.x, a.y, a.z, b.x, b.y, b.z, s.mats, s.uvs).sp;
        } else sp = rays ? resolve_hit(rec, h, rays + i) : resolve_hit(rec, h);
    }
    if (P) { P[3 * i] = sp.P[0]; P[3 * i + 1] = sp.P[1]; P[3 * i + 2] = sp.P[2]; }
    if (N) { N[3 * i] = sp.N[0]; N[3 * i + 1] = sp.N[1]; N[3 * i + 2] = sp.N[2]; }
    if (mat) mat[i] = sp.material;
}

// alignHemisphereToVector, Utility.h:34-50.  Everything except sinf/cosf is evaluated in the reference's
// operand order without FMA; CUDA's sinf/cosf/asinf differ from glibc's in the last ulp, which is why
// generated bounce rays are compared with a tolerance and hit parity is checked on the dumped rays.
__device__ __forceinline__ void align_hemisphere(const float v[3], float theta, float phi, float out[3])
{
    float sp, cp, st, ct;
    sincosf(phi, &sp, &cp); sincosf(theta, &st, &ct);   // one range reduction per angle
    const float u1 = xmul(sp, ct), u2 = xmul(sp, st), u3 = cp;
    // t1 = cross((0,0,1), v)
    float t1[3] = {xsub(xmul(0.f, v[2]), xmul(1.f, v[1])), xsub(xmul(1.f, v[0]), xmul(0.f, v[2])), xsub(xmul(0.f, v[1]), xmul(0.f, v[0]))};
    if ((double)xdot(t1[0], t1[1], t1[2], t1[0], t1[1], t1[2]) < 1e-6) {
        // t1 = cross((0,1,0), v)
        t1[0] = xsub(xmul(1.f, v[2]), xmul(0.f, v[1]));
        t1[1] = xsub(xmul(0.f, v[0]), xmul(0.f, v[2]));
        t1[2] = xsub(xmul(0.f, v[1]), xmul(1.f, v[0]));
    }
    // cross(t1, v)
    const float c[3] = {xsub(xmul(t1[1], v[2]), xmul(t1[2], v[1])), xsub(xmul(t1[2], v[0]), xmul(t1[0], v[2])),
                        xsub(xmul(t1[0], v[1]), xmul(t1[1], v[0]))};
    float a[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] = xadd(xadd(xmul(t1[k], u1), xmul(c[k], u2)), xmul(v[k], u3));
    const float inv = xdiv(1.0f, xsqrt(xdot(a[0], a[1], a[2], a[0], a[1], a[2])));
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = xmul(a[k], inv);
}

// Ray::diffuse, Ray.h:109-122, with (u1,u2) from the counter RNG in place of rand().
#define MIRO_GENB_THREADS 128
// MIRO_GENB_ITEMS hits per thread (i, i + 128, ... of a tile; 4 by default): the kernel is a chain of two dependent fetches (the
// hit, then the 96-byte shading record it names) in front of ~300 instructions, so it runs at the latency of those fetches; all
// hits and then all records of the thread are requested before anything is computed.
template <int MIRO_GENB_ITEMS, bool TEXTURED = false>
__global__ void __launch_bounds__(MIRO_GENB_THREADS) k_gen_bounce(DeviceScene s, const mirogpu_ray* __restrict__ rays,
                                                                   const mirogpu_hit* __restrict__ hits, size_t n, uint32_t seed, uint32_t sample,
                                                                   uint32_t index_base, mirogpu_ray* __restrict__ out, unsigned long long* live_count)
{
    const size_t base = (size_t)blockIdx.x * (MIRO_GENB_THREADS * MIRO_GENB_ITEMS) + threadIdx.x;
    mirogpu_hit h[MIRO_GENB_ITEMS];
#pragma unroll
    for (int k = 0; k < MIRO_GENB_ITEMS; ++k) {
        const size_t i = base + (size_t)k * MIRO_GENB_THREADS;
        float4 hv = make_float4(0.f, __uint_as_float(MIROGPU_MISS), 0.f, 0.f);
        if (i < n) hv = __ldcs(reinterpret_cast<const float4*>(hits + i));
        h[k].t = hv.x; h[k].prim_id = __float_as_uint(hv.y); h[k].beta = hv.z; h[k].gamma = hv.w;
    }
    ShadeRecord rec[MIRO_GENB_ITEMS];
#pragma unroll
    for (int k = 0; k < MIRO_GENB_ITEMS; ++k) {
        if (h[k].prim_id != MIROGPU_MISS) rec[k] = load_shade_record(s, h[k].prim_id);
        else rec[k].r0.lo = rec[k].r0.hi = rec[k].r1.lo = rec[k].r1.hi = rec[k].r2.lo = rec[k].r2.hi = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (live_count) {   // one atomic per warp
        unsigned cnt = 0;
#pragma unroll
        for (int k = 0; k < MIRO_GENB_ITEMS; ++k) cnt += __popc(__ballot_sync(0xffffffffu, h[k].prim_id != MIROGPU_MISS));
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(live_count, (unsigned long long)cnt);
    }
#pragma unroll
    for (int k = 0; k < MIRO_GENB_ITEMS; ++k) {
        const size_t i = base + (size_t)k * MIRO_GENB_THREADS;
        if (i >= n) continue;
        float4 a, b;
        if (h[k].prim_id == MIROGPU_MISS) {
            a = make_float4(0.f, 0.f, 0.f, 0.f); b = make_float4(0.f, 0.f, 1.f, -1.0f);  // tmax < tmin: never hits
        } else {
            const SurfacePoint sp = TEXTURED ? resolve_hit_textured(rec[k], h[k], __ldg(&rays[i].ox), __ldg(&rays[i].oy), __ldg(&rays[i].oz), __ldg(&rays[i].dx), __ldg(&rays[i].dy),
                                                                      __ldg(&rays[i].dz), s.mats, s.uvs).sp
                                                : resolve_hit(rec[k], h[k], rays + i);
            float u1, u2;
            uniform2(seed, index_base + (uint32_t)i, sample, RNG_DIM_BOUNCE, u1, u2);
            const float phi = asinf(sqrtf(u1));
            const float theta = xmul(xmul(2.0f, MIRO_PI), u2);
            float d[3];
            align_hemisphere(sp.N, theta, phi, d);
            a.x = xadd(sp.P[0], xmul(d[0], MIRO_EPS)); a.y = xadd(sp.P[1], xmul(d[1], MIRO_EPS)); a.z = xadd(sp.P[2], xmul(d[2], MIRO_EPS));
            a.w = 0.0f;
            b.x = d[0]; b.y = d[1]; b.z = d[2]; b.w = MIROGPU_TMAX;
        }
        float4* o = reinterpret_cast<float4*>(out + i);
        __stcs(o, a); __stcs(o + 1, b);
    }
}

}  // namespace mirogpu
#endif
