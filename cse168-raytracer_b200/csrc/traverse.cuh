// traverse.cuh -- per-ray traversal cores and the reference's ray-triangle test, written once as
// host/device inline functions: the CUDA kernels in kernels.cu call them per thread on sm_100a, and the
// test-only emulation harness (tests/cpu_emu/emu.cu, never linked into the product) calls the very same
// code on the host so traversal logic can be checked against the oracle without a GPU.
//
// Arithmetic contract.  The triangle test reproduces Triangle::intersect (reference Triangle.cpp:136-169)
// operation for operation in IEEE binary32 with round-to-nearest and NO fused multiply-add: the device
// build uses __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn, which nvcc never contracts, so t, beta and gamma
// are bit-identical to the reference's scalar CPU build.  Box tests are free to use FMA: any conservative
// slab test yields the same closest hit.
#ifndef MIROGPU_TRAVERSE_CUH
#define MIROGPU_TRAVERSE_CUH

#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "../../include/mirogpu.h"
#include "qbvh4_config.h"

#define MIRO_HD __host__ __device__ __forceinline__

namespace mirogpu {

#define MIRO_EPS 1e-4f /* epsilon, Miro.h:9 */

// ---- exactly-rounded, never-contracted binary32 ops ----------------------------------------------------
MIRO_HD float xmul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
MIRO_HD float xadd(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
MIRO_HD float xsub(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
MIRO_HD float xdiv(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
MIRO_HD float xsqrt(float a)
{
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}
// dot(a, b) = a.x*b.x + a.y*b.y + a.z*b.z, left to right (Vector3.h:243-246)
MIRO_HD float xdot(float ax, float ay, float az, float bx, float by, float bz)
{
    return xadd(xadd(xmul(ax, bx), xmul(ay, by)), xmul(az, bz));
}

template <typename T>
MIRO_HD T ldg(const T* p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// 256-bit read-only load (LDG.E.ENL2.256.CONSTANT on sm_100a): one 32-byte sector per lane per instruction.  The
// traversal kernels are bound by L1 data-pipe wavefronts (one sector per wavefront for divergent lanes), so a 64-byte
// node or triangle record costs two of these instead of four 128-bit loads.  p must be 32-byte aligned.
struct F8 {
    float4 lo, hi;
};
MIRO_HD F8 ld256(const float4* p)
{
    F8 r;
#ifdef __CUDA_ARCH__
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
#else
    r.lo = p[0]; r.hi = p[1];
#endif
    return r;
}
// L2 prefetch of the 128-byte line holding p: issued when a far child is pushed on the stack (or a leaf is reached
// before the warp's leaf phase), so the DRAM/L2 latency of that fetch overlaps the walk of the near subtree.
MIRO_HD void prefetch_l2(const void* p)
{
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
MIRO_HD void prefetch_l1(const void* p)
{
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
MIRO_HD int popc32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
MIRO_HD int bfind32(uint32_t x)  // index of the highest set bit; x != 0
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
MIRO_HD float u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
MIRO_HD uint32_t f2u(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { uint32_t u; float f; } c; c.f = f; return c.u;
#endif
}

struct TraceCounters {
    uint32_t nodes, boxes, tris;
};

struct BestHit {
    float t;
    uint32_t prim;  // MIROGPU_MISS until something is accepted
    float beta, gamma;
};

#ifdef __CUDACC__
// The hybrid kernel's copy of BestHit for triangle-only scenes: only t is consulted by the walk (every box test), prim / beta /
// gamma are written when a triangle is accepted (once or twice per ray) and read when the ray is answered.  They live in the
// CTA's shared memory ([field][thread], conflict-free) instead of three registers per lane -- the kernel runs at a 48-register
// budget (10 CTAs of 128 threads per SM), and with them resident the node step spills.  Slot 3 is the lane's ray index.
#define MIRO_BEST_SM_THREADS 128
__shared__ uint32_t g_best_sm[4 * MIRO_BEST_SM_THREADS];
template <int K, typename T>
struct BestSmSlot {
    __device__ __forceinline__ operator T() const { const uint32_t v = g_best_sm[K * MIRO_BEST_SM_THREADS + threadIdx.x]; return *reinterpret_cast<const T*>(&v); }
    __device__ __forceinline__ BestSmSlot& operator=(T x) { g_best_sm[K * MIRO_BEST_SM_THREADS + threadIdx.x] = *reinterpret_cast<const uint32_t*>(&x); return *this; }
    __device__ __forceinline__ BestSmSlot& operator=(const BestSmSlot& o) { return *this = (T)o; }
};
struct BestHitSm {
    float t;
    BestSmSlot<0, uint32_t> prim;
    BestSmSlot<1, float> beta;
    BestSmSlot<2, float> gamma;
};
#endif

// Sphere::intersect, Sphere.cpp:28-69, on a sphere slot of the leaf array: v0 = (centre, prim id bits), radius separately.  Same
// operations in the same order (a = |d|^2, b = (2 d) . (o - c), c = |o - c|^2 - r r, discriminant b b - 4 a c, the two roots
// divided by 2 a), no FMA.  The reference takes the nearer root if it lies strictly inside (tMin, tMax) -- tMax being the best
// hit so far in its leaf loop (BVH.cpp:498) -- else the farther one; a hit must then be strictly closer than the best
// (BVH.cpp:500).  Equal t goes to the smaller primitive id here, as for triangles.
// Returns the accepted t, or NaN.  Not inlined and all by value: the traversal kernels run at a 48-register budget, and this
// path (scenes with spheres, slots a ray actually reaches) must not cost the triangle path registers or local memory.
#ifdef __CUDA_ARCH__
#define MIRO_RARE __device__ __noinline__
#else
#define MIRO_RARE inline
#endif
MIRO_RARE float sphere_hit_t(const float4 v0, const float radius, const mirogpu_ray r, const float best_t, const uint32_t best_prim)
{
    const float kNaN = u2f(0x7fc00000u);
    const float tox = xsub(r.ox, v0.x), toy = xsub(r.oy, v0.y), toz = xsub(r.oz, v0.z);
    const float a = xdot(r.dx, r.dy, r.dz, r.dx, r.dy, r.dz);
    const float b = xdot(xmul(r.dx, 2.0f), xmul(r.dy, 2.0f), xmul(r.dz, 2.0f), tox, toy, toz);
    const float c = xsub(xdot(tox, toy, toz, tox, toy, toz), xmul(radius, radius));
    const float discrim = xsub(xmul(b, b), xmul(xmul(4.0f, a), c));
    if (discrim < 0.f) return kNaN;
    const float sq = xsqrt(discrim), two_a = xmul(2.0f, a);
    const float t0 = xdiv(xsub(-b, sq), two_a), t1 = xdiv(xadd(-b, sq), two_a);
    const uint32_t prim = f2u(v0.w);
    float t;
    // (t0 > tMin && t0 < tMax) else (t1 > tMin && t1 < tMax), with tMax = the best hit so far; `closer` admits the tie rule
    if (t0 > r.tmin) { if (t0 < best_t || (t0 == best_t && prim < best_prim)) t = t0; else return kNaN; }
    else if (t1 > r.tmin && (t1 < best_t || (t1 == best_t && prim < best_prim))) t = t1;
    else return kNaN;
    if (!(t < r.tmax)) return kNaN;   // the sphere's range test is strict at the far end too
    return t;
}

// Triangle::intersect, Triangle.cpp:150-158.  v0 = (A, prim id bits), v1 = (B-A, n.x), v2 = (C-A, n.y), v3 = (n.z, kind, radius, -);
// kind != 0: the slot is a sphere (centre in v0, all of v1, v2, n.z zero).
// with n = cross(B-A, C-A) formed once on the host in the same binary32 operations the reference performs per call
// (make_tri_records), so the value is the one the reference computes.
// Acceptance on top of the reference's own reject line: the leaf keeps a hit only if it is strictly closer
// than the best so far (BVH.cpp:498-500); equal t goes to the smaller primitive id so the result does not
// depend on traversal order.  NaN t fails every comparison and is dropped, as in the reference.
template <bool NT, typename BEST>
MIRO_HD bool tri_test(const float4 v0, const float4 v1, const float4 v2, const float4 v3, const mirogpu_ray& r, BEST& best, const float4* __restrict__ rec)
{
    const float nx = v1.w, ny = v2.w, nz = v3.x;   // normal = cross(BmA, CmA)
    const float ndx = -r.dx, ndy = -r.dy, ndz = -r.dz;
    const float ddotn = xdot(ndx, ndy, ndz, nx, ny, nz);
    const float oax = xsub(r.ox, v0.x), oay = xsub(r.oy, v0.y), oaz = xsub(r.oz, v0.z);
    const float t = xdiv(xdot(oax, oay, oaz, nx, ny, nz), ddotn);
    // cross(o-A, CmA)
    const float c1x = xsub(xmul(oay, v2.z), xmul(oaz, v2.y));
    const float c1y = xsub(xmul(oaz, v2.x), xmul(oax, v2.z));
    const float c1z = xsub(xmul(oax, v2.y), xmul(oay, v2.x));
    const float beta = xdiv(xdot(ndx, ndy, ndz, c1x, c1y, c1z), ddotn);
    // cross(BmA, o-A)
    const float c2x = xsub(xmul(v1.y, oaz), xmul(v1.z, oay));
    const float c2y = xsub(xmul(v1.z, oax), xmul(v1.x, oaz));
    const float c2z = xsub(xmul(v1.x, oay), xmul(v1.y, oax));
    const float gamma = xdiv(xdot(ndx, ndy, ndz, c2x, c2y, c2z), ddotn);
    if (beta < -MIRO_EPS || gamma < -MIRO_EPS || xadd(beta, gamma) > 1 + MIRO_EPS || t < r.tmin || t > best.t) return false;
    // A sphere slot holds zero edge vectors and a zero normal: t, beta and gamma above are all 0 / 0 = NaN, every comparison of
    // the reject line is false, and the slot arrives here -- so triangles the ray misses (nearly all) never pay for the kind test.
    // NT: kernels instantiated for scenes that hold non-triangle primitives (the triangle-only instantiations, which run at a
    // 48-register budget, compile without any of this).  (kind, radius) are re-read from the record rather than kept live.
    if (NT) {
        const float2 kr = ldg(reinterpret_cast<const float2*>(rec) + 6);   // floats 12..13 of the record = (n.z, kind)
        if (f2u(kr.y) != 0u) {
            const float radius = ldg(reinterpret_cast<const float*>(rec) + 14);
            const float ts = sphere_hit_t(v0, radius, r, best.t, best.prim);
            if (ts != ts) return false;
            best.t = ts; best.prim = f2u(v0.w); best.beta = 0.f; best.gamma = 0.f;
            return true;
        }
    }
    const uint32_t prim = f2u(v0.w);
    if (t < best.t || (t == best.t && prim < best.prim)) {
        best.t = t; best.prim = prim; best.beta = beta; best.gamma = gamma;
        return true;
    }
    return false;
}

MIRO_HD float safe_rcp(float d)
{
    // 1/d with |d| clamped away from zero so that 0 * inf never produces a NaN in the slab test
    const float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
    return 1.0f / a;
}

#define MIRO_STACK 64
#define MIRO_STACK4 96   /* BVH4 pushes up to three entries per level; flatten_bvh4 computes the exact need and scene creation checks it */

// ---- BVH2 (64-byte nodes, two child boxes per fetch) ---------------------------------------------------
template <bool ANY, bool COUNT, bool NT = false>
MIRO_HD void trace_bvh2(const float4* __restrict__ nodes, const float4* __restrict__ tris, const mirogpu_ray& r,
                        BestHit& best, TraceCounters* cnt)
{
    const float idx = safe_rcp(r.dx), idy = safe_rcp(r.dy), idz = safe_rcp(r.dz);
    const float oodx = r.ox * idx, oody = r.oy * idy, oodz = r.oz * idz;
    int32_t stack[MIRO_STACK];
    int sp = 0;
    int32_t node = 0;
    best.t = r.tmax; best.prim = MIROGPU_MISS; best.beta = 0.f; best.gamma = 0.f;
    if (!(r.tmax >= r.tmin)) return;
    for (;;) {
        while (node >= 0) {
            const F8 na = ld256(nodes + 4 * (size_t)node), nb = ld256(nodes + 4 * (size_t)node + 2);
            const float4 n0 = na.lo, n1 = na.hi, nz = nb.lo, lk = nb.hi;
            if (COUNT) { cnt->nodes++; cnt->boxes += 2; }
            const float c0lox = n0.x * idx - oodx, c0hix = n0.y * idx - oodx;
            const float c0loy = n0.z * idy - oody, c0hiy = n0.w * idy - oody;
            const float c0loz = nz.x * idz - oodz, c0hiz = nz.y * idz - oodz;
            const float c1lox = n1.x * idx - oodx, c1hix = n1.y * idx - oodx;
            const float c1loy = n1.z * idy - oody, c1hiy = n1.w * idy - oody;
            const float c1loz = nz.z * idz - oodz, c1hiz = nz.w * idz - oodz;
            const float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), r.tmin));
            const float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), best.t));
            const float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), r.tmin));
            const float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), best.t));
            const bool h0 = t0n <= t0f, h1 = t1n <= t1f;
            const int32_t l0 = (int32_t)f2u(lk.x), l1 = (int32_t)f2u(lk.y);
            if (!h0 && !h1) {
                if (sp == 0) return;
                node = stack[--sp];
            } else {
                node = h0 ? l0 : l1;
                if (h0 && h1) {
                    int32_t other = l1;
                    if (t1n < t0n) { node = l1; other = l0; }
                    stack[sp++] = other;
                }
            }
        }
        // leaf: ~node = (first << 3) | (count - 1)
        {
            const uint32_t ref = (uint32_t)~node;
            const uint32_t first = ref >> 3, count = (ref & 7u) + 1u;
            for (uint32_t i = 0; i < count; ++i) {
                const F8 ta = ld256(tris + 4 * (size_t)(first + i)), tb = ld256(tris + 4 * (size_t)(first + i) + 2);
                if (COUNT) cnt->tris++;
                const bool acc = tri_test<NT>(ta.lo, ta.hi, tb.lo, tb.hi, r, best, tris + 4 * (size_t)(first + i));
                if (ANY && acc) return;
            }
            if (sp == 0) return;
            node = stack[--sp];
        }
    }
}

// ---- BVH2, single-step form (hybrid-scheduled kernel) ---------------------------------------------------------
// The same walk as trace_bvh2, cut into its two kinds of step so that a warp can decide per iteration which kind to
// run, with all per-ray state in Bvh2Walk:
//   bvh2_node_step  -- w.node >= 0: fetch the node, test both child boxes, descend / push / pop
//   bvh2_leaf_step  -- w.node < 0 and != MIRO_BVH2_DONE: test the leaf's triangles, then pop
// Either leaves w.node = next inner node, next leaf, or MIRO_BVH2_DONE when the walk is over.
// The top of the stack is cached in a register (w.tos): a pop hands out the register at once and the reload of the
// new top from local memory completes in the shadow of the next step.  The bottom-of-stack marker is an ordinary
// entry (MIRO_BVH2_DONE), so there is no emptiness test.
#define MIRO_BVH2_DONE ((int32_t)0x80000000)

struct Bvh2Walk {
    float idx, idy, idz, oodx, oody, oodz;
    int32_t node, tos;
    int sp;
};

MIRO_HD float fast_safe_rcp(float d)
{
    const float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));   // 1 ulp; box tests only (conservative boxes absorb it)
    return r;
#else
    return 1.0f / a;
#endif
}

template <typename BEST>
MIRO_HD void bvh2_begin(const mirogpu_ray& r, Bvh2Walk& w, BEST& best)
{
    w.idx = fast_safe_rcp(r.dx); w.idy = fast_safe_rcp(r.dy); w.idz = fast_safe_rcp(r.dz);
    w.oodx = r.ox * w.idx; w.oody = r.oy * w.idy; w.oodz = r.oz * w.idz;
    w.sp = 0;
    w.tos = MIRO_BVH2_DONE;
    w.node = (r.tmax >= r.tmin) ? 0 : MIRO_BVH2_DONE;   // an empty interval (or NaN bounds) never hits
    best.t = r.tmax; best.prim = MIROGPU_MISS; best.beta = 0.f; best.gamma = 0.f;
}

// Where the walk's stack lives.  LocalStack: a per-thread array (local memory).  SplitStack (hybrid kernel): the first SHORT
// entries of every lane in shared memory, laid out [entry][thread] so that a warp's access is conflict-free whatever the
// lanes' depths are (bank = lane), deeper entries in a local array -- a divergent local-memory access costs one L1
// wavefront per distinct line, the shared one always exactly one.
template <int N>
struct LocalStack {
    int32_t a[N];
    MIRO_HD void put(int i, int32_t v) { a[i] = v; }
    MIRO_HD int32_t get(int i) const { return a[i]; }
};
template <int SHORT, int N, int THREADS>
struct SplitStack {
    int32_t* sm;   // this thread's column of the CTA's [SHORT][THREADS] array
    int32_t a[N > SHORT ? N - SHORT : 1];
    // (SHORT > 0 &&: the index is a signed int the compiler cannot prove non-negative, so `i < 0` alone would keep the shared path alive)
    MIRO_HD void put(int i, int32_t v) { if (SHORT > 0 && i < SHORT) sm[i * THREADS] = v; else a[i - SHORT] = v; }
    MIRO_HD int32_t get(int i) const { return (SHORT > 0 && i < SHORT) ? sm[i * THREADS] : a[i - SHORT]; }
};

template <typename STK>
MIRO_HD void bvh2_pop(Bvh2Walk& w, const STK& stack)
{
    w.node = w.tos;
    if (w.tos != MIRO_BVH2_DONE) w.tos = stack.get(--w.sp);
}

// PF bit 0: prefetch the pushed far child (node or first triangle line) into L2; bit 1: into L1 instead;
// bit 2: prefetch the triangles of a leaf reached by descent (the lane usually waits for the warp's leaf phase).
template <int PF, typename STK, typename BEST>
MIRO_HD void bvh2_node_step(const float4* __restrict__ nodes, const float4* __restrict__ tris, const mirogpu_ray& r, Bvh2Walk& w,
                            STK& stack, const BEST& best)
{
    const F8 na = ld256(nodes + 4 * (size_t)w.node), nb = ld256(nodes + 4 * (size_t)w.node + 2);
    const float4 n0 = na.lo, n1 = na.hi, nz = nb.lo, lk = nb.hi;
    const float c0lox = n0.x * w.idx - w.oodx, c0hix = n0.y * w.idx - w.oodx;
    const float c0loy = n0.z * w.idy - w.oody, c0hiy = n0.w * w.idy - w.oody;
    const float c0loz = nz.x * w.idz - w.oodz, c0hiz = nz.y * w.idz - w.oodz;
    const float c1lox = n1.x * w.idx - w.oodx, c1hix = n1.y * w.idx - w.oodx;
    const float c1loy = n1.z * w.idy - w.oody, c1hiy = n1.w * w.idy - w.oody;
    const float c1loz = nz.z * w.idz - w.oodz, c1hiz = nz.w * w.idz - w.oodz;
    const float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), r.tmin));
    const float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), best.t));
    const float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), r.tmin));
    const float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), best.t));
    const bool h0 = t0n <= t0f, h1 = t1n <= t1f;
    const int32_t l0 = (int32_t)f2u(lk.x), l1 = (int32_t)f2u(lk.y);
    if (!h0 && !h1) {
        bvh2_pop(w, stack);
    } else {
        int32_t next = h0 ? l0 : l1;
        if (h0 && h1) {
            int32_t other = l1;
            if (t1n < t0n) { next = l1; other = l0; }
            stack.put(w.sp++, w.tos);
            w.tos = other;
            if (PF & 3) {
                const float4* a = other >= 0 ? nodes + 4 * (size_t)other : tris + 4 * (size_t)(((uint32_t)~other) >> 3);
                if (PF & 2) prefetch_l1(a); else prefetch_l2(a);
            }
        }
        if ((PF & 4) && next < 0) {
            const float4* a = tris + 4 * (size_t)(((uint32_t)~next) >> 3);
            if (PF & 2) prefetch_l1(a); else prefetch_l2(a);
        }
        w.node = next;
    }
}

template <bool ANY, bool NT = false, typename STK, typename BEST>
MIRO_HD void bvh2_leaf_step(const float4* __restrict__ tris, const mirogpu_ray& r, Bvh2Walk& w, const STK& stack,
                            BEST& best)
{
    const uint32_t ref = (uint32_t)~w.node;
    const uint32_t first = ref >> 3, count = (ref & 7u) + 1u;
    for (uint32_t i = 0; i < count; ++i) {
        const F8 ta = ld256(tris + 4 * (size_t)(first + i)), tb = ld256(tris + 4 * (size_t)(first + i) + 2);
        const bool acc = tri_test<NT>(ta.lo, ta.hi, tb.lo, tb.hi, r, best, tris + 4 * (size_t)(first + i));
        if (ANY && acc) { w.node = MIRO_BVH2_DONE; return; }
    }
    bvh2_pop(w, stack);
}

// One triangle of the leaf per call; a leaf with more stays the lane's node (first + 1, count - 1), so the warp's next vote
// decides again between node steps and another leaf phase and a leaf phase never loops over the longest leaf of the warp.
template <bool ANY, bool NT = false, typename STK, typename BEST>
MIRO_HD void bvh2_leaf_step_one(const float4* __restrict__ tris, const mirogpu_ray& r, Bvh2Walk& w, const STK& stack,
                                BEST& best)
{
    const uint32_t ref = (uint32_t)~w.node;
    const uint32_t first = ref >> 3;
    const F8 ta = ld256(tris + 4 * (size_t)first), tb = ld256(tris + 4 * (size_t)first + 2);
    const bool acc = tri_test<NT>(ta.lo, ta.hi, tb.lo, tb.hi, r, best, tris + 4 * (size_t)first);
    if (ANY && acc) { w.node = MIRO_BVH2_DONE; return; }
    if (ref & 7u) w.node = (int32_t)~(ref + 7u);   // first + 1 (ref + 8), count - 1 (ref - 1)
    else bvh2_pop(w, stack);
}

// Up to two triangles of the leaf per call (PF bit 5): fewer leaf phases -- and votes -- per ray than one at a time, at the
// price of lanes with a single triangle idling through the second test.
template <bool ANY, bool NT = false, typename STK, typename BEST>
MIRO_HD void bvh2_leaf_step_two(const float4* __restrict__ tris, const mirogpu_ray& r, Bvh2Walk& w, const STK& stack, BEST& best)
{
    const uint32_t ref = (uint32_t)~w.node;
    const uint32_t first = ref >> 3, left = ref & 7u;   // left = count - 1
    const F8 ta = ld256(tris + 4 * (size_t)first), tb = ld256(tris + 4 * (size_t)first + 2);
    F8 tc, td;
    if (left) { tc = ld256(tris + 4 * (size_t)first + 4); td = ld256(tris + 4 * (size_t)first + 6); }
    bool acc = tri_test<NT>(ta.lo, ta.hi, tb.lo, tb.hi, r, best, tris + 4 * (size_t)first);
    if (left && !(ANY && acc)) acc |= tri_test<NT>(tc.lo, tc.hi, td.lo, td.hi, r, best, tris + 4 * (size_t)first + 4);
    if (ANY && acc) { w.node = MIRO_BVH2_DONE; return; }
    if (left >= 2u) w.node = (int32_t)~(ref + 14u);   // first + 2 (ref + 16), count - 2 (ref - 2)
    else bvh2_pop(w, stack);
}

// One triangle of a postponed leaf reference (PF bit 6, see k_trace_hybrid); returns the reference of what is left of the leaf
// (MIRO_BVH2_DONE = nothing).  hit: a triangle was accepted.
template <bool NT = false, typename BEST>
MIRO_HD int32_t leaf_ref_test_one(const float4* __restrict__ tris, const mirogpu_ray& r, int32_t leaf, BEST& best, bool& hit)
{
    const uint32_t ref = (uint32_t)~leaf;
    const uint32_t first = ref >> 3;
    const F8 ta = ld256(tris + 4 * (size_t)first), tb = ld256(tris + 4 * (size_t)first + 2);
    hit = tri_test<NT>(ta.lo, ta.hi, tb.lo, tb.hi, r, best, tris + 4 * (size_t)first);
    return (ref & 7u) ? (int32_t)~(ref + 7u) : MIRO_BVH2_DONE;
}

// Child order of the four-wide layouts: a three-comparator tournament on (entry distance, link) finds the nearest hit
// child, which is descended next; the two first-round losers are pushed first and the runner-up of the final last, so
// the nearer of the remaining children tends to be popped earlier.  Misses carry distance +inf and are never pushed.
template <int PF, typename STK>
MIRO_HD void wide4_descend(const float d[4], const int32_t lk[4], const float4* __restrict__ tris, Bvh2Walk& w, STK& stack)
{
    const float kFar = u2f(0x7f800000u);
    // comparators written as min / max on the distances and selects on the links (no register shuffling)
    const bool p01 = d[1] < d[0], p23 = d[3] < d[2];
    const float w01 = fminf(d[0], d[1]), db = fmaxf(d[0], d[1]), w23 = fminf(d[2], d[3]), de = fmaxf(d[2], d[3]);
    const int32_t m01 = p01 ? lk[1] : lk[0], lb = p01 ? lk[0] : lk[1], m23 = p23 ? lk[3] : lk[2], le = p23 ? lk[2] : lk[3];
    const bool pf = w23 < w01;
    const float da = fminf(w01, w23), dc = fmaxf(w01, w23);
    const int32_t la = pf ? m23 : m01, lc = pf ? m01 : m23;
    if (!(da < kFar)) { bvh2_pop(w, stack); return; }
    if (db < kFar) { stack.put(w.sp++, w.tos); w.tos = lb; }
    if (de < kFar) { stack.put(w.sp++, w.tos); w.tos = le; }
    if (dc < kFar) { stack.put(w.sp++, w.tos); w.tos = lc; }
    if ((PF & 4) && la < 0) {
        const float4* a = tris + 4 * (size_t)(((uint32_t)~la) >> 3);
        if (PF & 2) prefetch_l1(a); else prefetch_l2(a);
    }
    w.node = la;
}

// f[i] = 2^-24 * byte i of w, exactly.  Device: two PRMTs build the four binary16 values 0x00bb -- subnormal halves,
// worth bb * 2^-24 -- and the (exact) half -> float conversions run on the FMA pipe, which the traversal leaves mostly
// idle.  The 2^24 is folded into the per-axis cell constant (a power of two, so exactly).
MIRO_HD void unpack_planes(uint32_t w, float f[4])
{
#if MIRO_QDIRECT
    // f[i] = 1 + 2^-15 * byte i of w: the byte dropped into bits 8..15 of the binary32 1.0 -- one PRMT per plane and no
    // conversion.  The builder shifts the grid origin down by 2^15 cells to match (bvh_build.h, qbvh4_cell_words).
#ifdef __CUDA_ARCH__
    f[0] = u2f(__byte_perm(w, 0x3f800000u, 0x7604)); f[1] = u2f(__byte_perm(w, 0x3f800000u, 0x7614));
    f[2] = u2f(__byte_perm(w, 0x3f800000u, 0x7624)); f[3] = u2f(__byte_perm(w, 0x3f800000u, 0x7634));
#else
    for (int i = 0; i < 4; ++i) f[i] = u2f(0x3f800000u | (((w >> (8 * i)) & 0xffu) << 8));
#endif
#elif defined(__CUDA_ARCH__)
    const uint32_t p01 = __byte_perm(w, 0u, 0x4140), p23 = __byte_perm(w, 0u, 0x4342);
    const __half2 h01 = *reinterpret_cast<const __half2*>(&p01), h23 = *reinterpret_cast<const __half2*>(&p23);
    f[0] = __low2float(h01); f[1] = __high2float(h01); f[2] = __low2float(h23); f[3] = __high2float(h23);
#else
    for (int i = 0; i < 4; ++i) f[i] = (float)((w >> (8 * i)) & 0xffu) * (1.0f / 16777216.0f);
#endif
}


#ifdef __CUDA_ARCH__
// (a, b) = (a, b) * s + c, each half rounded once (fma.rn.f32x2 -> FFMA2 with s and c as broadcast scalar operands).
__device__ __forceinline__ void fma2(float& a, float& b, const float s, const float c)
{
    asm("{\n\t.reg .b64 ra, rs, rc;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rs, {%2, %2};\n\tmov.b64 rc, {%3, %3};\n\t"
        "fma.rn.f32x2 ra, ra, rs, rc;\n\tmov.b64 {%0, %1}, ra;\n\t}" : "+f"(a), "+f"(b) : "f"(s), "f"(c));
}
#endif

// ---- BVH4 (128-byte nodes, four full-precision child boxes per fetch) ---------------------------------------------
// Node = 8 x float4: (lo.x[4]) (hi.x[4]) (lo.y[4]) (hi.y[4]) (lo.z[4]) (hi.z[4]) (link[4]) (pad) -- four 256-bit loads, one
// cache line.  Same walk state and leaf step as BVH2.  Child order: a three-comparator tournament finds the nearest
// hit child (descended next); the two first-round losers are pushed first, the runner-up of the final last, so the
// nearer of the remaining children tends to be popped earlier.  Misses carry distance +inf and are never pushed.
template <int PF, typename STK, typename BEST>
MIRO_HD void bvh4_node_step(const float4* __restrict__ nodes, const float4* __restrict__ tris, const mirogpu_ray& r, Bvh2Walk& w,
                            STK& stack, const BEST& best)
{
    const float4* p = nodes + 8 * (size_t)w.node;
    const F8 X = ld256(p), Y = ld256(p + 2), Z = ld256(p + 4), L = ld256(p + 6);
    const float lox[4] = {X.lo.x, X.lo.y, X.lo.z, X.lo.w}, hix[4] = {X.hi.x, X.hi.y, X.hi.z, X.hi.w};
    const float loy[4] = {Y.lo.x, Y.lo.y, Y.lo.z, Y.lo.w}, hiy[4] = {Y.hi.x, Y.hi.y, Y.hi.z, Y.hi.w};
    const float loz[4] = {Z.lo.x, Z.lo.y, Z.lo.z, Z.lo.w}, hiz[4] = {Z.hi.x, Z.hi.y, Z.hi.z, Z.hi.w};
    const int32_t lk[4] = {(int32_t)f2u(L.lo.x), (int32_t)f2u(L.lo.y), (int32_t)f2u(L.lo.z), (int32_t)f2u(L.lo.w)};
    const float kFar = u2f(0x7f800000u);
    float d[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float ax = lox[c] * w.idx - w.oodx, bx = hix[c] * w.idx - w.oodx;
        const float ay = loy[c] * w.idy - w.oody, by = hiy[c] * w.idy - w.oody;
        const float az = loz[c] * w.idz - w.oodz, bz = hiz[c] * w.idz - w.oodz;
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), r.tmin));
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), best.t));
        d[c] = tn <= tf ? tn : kFar;
    }
    wide4_descend<PF>(d, lk, tris, w, stack);
}

// Whole walk of one ray (packet / one-thread-per-ray kernels, the photon walker, the counting build).
template <bool ANY, bool COUNT, bool NT = false>
MIRO_HD void trace_bvh4(const float4* __restrict__ nodes, const float4* __restrict__ tris, const mirogpu_ray& r, BestHit& best,
                        TraceCounters* cnt)
{
    Bvh2Walk w;
    LocalStack<MIRO_STACK4 + 1> stack;
    bvh2_begin(r, w, best);
    for (;;) {
        while (w.node >= 0) {
            if (COUNT) { cnt->nodes++; cnt->boxes += 4; }
            bvh4_node_step<0>(nodes, tris, r, w, stack, best);
        }
        if (w.node == MIRO_BVH2_DONE) return;
        if (COUNT) cnt->tris += (((uint32_t)~w.node) & 7u) + 1u;
        bvh2_leaf_step<ANY, NT>(tris, r, w, stack, best);
        if (w.node == MIRO_BVH2_DONE) return;
    }
}

// ---- QBVH4 (64-byte nodes, four child boxes quantised to 8 bits per plane on a per-node grid) ------------------------
// Node = 4 x float4: (origin.xyz, ex|ey<<8|ez<<16) (qlo.x[4], qhi.x[4], qlo.y[4], qhi.y[4]) (qlo.z[4], qhi.z[4], link0, link1)
// (link2, link3, cell words: 2^24 cell of x | y as binary32 upper halves, of z whole): two 256-bit loads.  A plane is origin + q * cell with cell = 2^(e-127); along the ray
//   t = (origin + q cell - o) / d = (q 2^-24) * (2^24 cell / d) + (origin - o) / d,
// one FMA per plane once the two per-axis constants are formed.  The near / far plane words are picked by the sign of
// the direction (two selects per axis), so no per-plane min / max is needed.  The builder rounds q outward with a
// margin of 0.02 cell, far above the rounding of this decode (a few ulp of the plane distance).
// STAGE > 0 (device, hybrid kernel): nodes [0, STAGE) -- the top levels, numbered breadth-first by the flattener -- are read from
// the CTA's copy in shared memory (`staged`, 4 float4 per node) instead of through L1.
template <int PF, typename STK, int STAGE = 0, typename BEST = BestHit>
MIRO_HD void qbvh4_node_step(const float4* __restrict__ nodes, const float4* __restrict__ tris, const mirogpu_ray& r, Bvh2Walk& w,
                             STK& stack, const BEST& best, const float4* staged = nullptr)
{
    const float4* p = nodes + 4 * (size_t)w.node;
    F8 A, B;
    if (STAGE > 0 && w.node < STAGE) {
        const float4* q = staged + 4 * w.node;
        A.lo = q[0]; A.hi = q[1]; B.lo = q[2]; B.hi = q[3];
    } else {
        A = ld256(p); B = ld256(p + 2);
    }
    // cell * 2^MIRO_QSHIFT (the planes arrive as 1 + q * 2^-15 above an origin stored 2^15 cells low -- or, MIRO_QDIRECT 0, as
    // q * 2^-24): the builders leave the three powers of two ready-made in the node's last two words (bvh_build.h,
    // qbvh4_cell_words) -- x and y as the upper halves of their binary32, z whole
#if MIRO_QCELL
    const uint32_t cw = f2u(B.hi.z);
    const float cx = u2f(cw << 16) * w.idx, cy = u2f(cw & 0xffff0000u) * w.idy, cz = B.hi.w * w.idz;
#else
    const uint32_t ew = f2u(A.lo.w);
    const float cx = u2f(((ew & 0xffu) + 24u) << 23) * w.idx, cy = u2f((((ew >> 8) & 0xffu) + 24u) << 23) * w.idy,
                cz = u2f((((ew >> 16) & 0xffu) + 24u) << 23) * w.idz;
#endif
    const float bx = fmaf(A.lo.x, w.idx, -w.oodx), by = fmaf(A.lo.y, w.idy, -w.oody), bz = fmaf(A.lo.z, w.idz, -w.oodz);
    const bool px = w.idx >= 0.f, py = w.idy >= 0.f, pz = w.idz >= 0.f;
    const uint32_t qlx = f2u(A.hi.x), qhx = f2u(A.hi.y), qly = f2u(A.hi.z), qhy = f2u(A.hi.w), qlz = f2u(B.lo.x), qhz = f2u(B.lo.y);
    float nx[4], fx[4], ny[4], fy[4], nz[4], fz[4];
    unpack_planes(px ? qlx : qhx, nx); unpack_planes(px ? qhx : qlx, fx);
    unpack_planes(py ? qly : qhy, ny); unpack_planes(py ? qhy : qly, fy);
    unpack_planes(pz ? qlz : qhz, nz); unpack_planes(pz ? qhz : qlz, fz);
    const int32_t lk[4] = {(int32_t)f2u(B.lo.z), (int32_t)f2u(B.lo.w), (int32_t)f2u(B.hi.x), (int32_t)f2u(B.hi.y)};
    const float kFar = u2f(0x7f800000u);
    float d[4];
#if defined(__CUDA_ARCH__) && MIRO_FFMA2
    // Plane distances two children at a time: sm_100's packed binary32 FMA (FFMA2, the per-axis constants as broadcast scalar
    // operands) rounds each half like fmaf, so the distances -- and every box decision -- are the bits of the scalar form at
    // half the issue slots (24 FFMA -> 12 FFMA2 per node; the kernel is bound by issue slots, not by the FMA pipe).
    fma2(nx[0], nx[1], cx, bx); fma2(nx[2], nx[3], cx, bx); fma2(fx[0], fx[1], cx, bx); fma2(fx[2], fx[3], cx, bx);
    fma2(ny[0], ny[1], cy, by); fma2(ny[2], ny[3], cy, by); fma2(fy[0], fy[1], cy, by); fma2(fy[2], fy[3], cy, by);
    fma2(nz[0], nz[1], cz, bz); fma2(nz[2], nz[3], cz, bz); fma2(fz[0], fz[1], cz, bz); fma2(fz[2], fz[3], cz, bz);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float tn = fmaxf(fmaxf(nx[c], ny[c]), fmaxf(nz[c], r.tmin));
        const float tf = fminf(fminf(fx[c], fy[c]), fminf(fz[c], best.t));
        d[c] = tn <= tf ? tn : kFar;
    }
#else
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float tn = fmaxf(fmaxf(fmaf(nx[c], cx, bx), fmaf(ny[c], cy, by)), fmaxf(fmaf(nz[c], cz, bz), r.tmin));
        const float tf = fminf(fminf(fmaf(fx[c], cx, bx), fmaf(fy[c], cy, by)), fminf(fmaf(fz[c], cz, bz), best.t));
        d[c] = tn <= tf ? tn : kFar;
    }
#endif
    wide4_descend<PF>(d, lk, tris, w, stack);
}

template <bool ANY, bool COUNT, bool NT = false>
MIRO_HD void trace_qbvh4(const float4* __restrict__ nodes, const float4* __restrict__ tris, const mirogpu_ray& r, BestHit& best,
                         TraceCounters* cnt)
{
    Bvh2Walk w;
    LocalStack<MIRO_STACK4 + 1> stack;
    bvh2_begin(r, w, best);
    for (;;) {
        while (w.node >= 0) {
            if (COUNT) { cnt->nodes++; cnt->boxes += 4; }
            qbvh4_node_step<0>(nodes, tris, r, w, stack, best);
        }
        if (w.node == MIRO_BVH2_DONE) return;
        if (COUNT) cnt->tris += (((uint32_t)~w.node) & 7u) + 1u;
        bvh2_leaf_step<ANY, NT>(tris, r, w, stack, best);
        if (w.node == MIRO_BVH2_DONE) return;
    }
}

// ---- CWBVH8 (80-byte nodes, eight 8-bit-quantised child boxes per fetch) --------------------------------
// Node words (5 x uint4), see Cwbvh8Node in bvh_build.h:
//   w0 = px, py, pz, (ex | ey<<8 | ez<<16 | imask<<24)
//   w1 = child_base, tri_base, meta[0..3], meta[4..7]
//   w2 = qlox[0..3], qlox[4..7], qloy[0..3], qloy[4..7]
//   w3 = qloz[0..3], qloz[4..7], qhix[0..3], qhix[4..7]
//   w4 = qhiy[0..3], qhiy[4..7], qhiz[0..3], qhiz[4..7]
// Traversal state: a node group G = (child_base, hits<<24 | imask) whose bits 24..31 mark internal children
// still to visit, ordered so that the highest bit is the child nearest along the ray's octant; a triangle
// group T = (tri_base, 24-bit mask).  Slot s of a node gets bit 24 + (s ^ octinv).
MIRO_HD uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

template <bool ANY, bool COUNT, bool NT = false>
MIRO_HD void trace_cwbvh8(const uint4* __restrict__ nodes, const float4* __restrict__ tris, const mirogpu_ray& r,
                          BestHit& best, TraceCounters* cnt)
{
    const float idx = safe_rcp(r.dx), idy = safe_rcp(r.dy), idz = safe_rcp(r.dz);
    const uint32_t octinv = (r.dx >= 0.f ? 4u : 0u) | (r.dy >= 0.f ? 2u : 0u) | (r.dz >= 0.f ? 1u : 0u);
    uint2 stack[MIRO_STACK];
    int sp = 0;
    best.t = r.tmax; best.prim = MIROGPU_MISS; best.beta = 0.f; best.gamma = 0.f;
    if (!(r.tmax >= r.tmin)) return;
    uint2 G = make_uint2(0u, 0x80000000u);  // the root: base 0, one pending "child" whose relative index is 0
    for (;;) {
        uint2 T = make_uint2(0u, 0u);
        if (G.y & 0xff000000u) {
            const uint32_t hits_imask = G.y;
            const int bit = bfind32(hits_imask);
            G.y &= ~(1u << bit);
            if (G.y & 0xff000000u) stack[sp++] = G;
            const uint32_t slot = ((uint32_t)(bit - 24)) ^ octinv;
            const uint32_t rel = (uint32_t)popc32(hits_imask & ((1u << slot) - 1u) & 0xffu);
            const uint32_t ni = G.x + rel;
            const uint4 w0 = ldg(nodes + 5 * (size_t)ni + 0);
            const uint4 w1 = ldg(nodes + 5 * (size_t)ni + 1);
            const uint4 w2 = ldg(nodes + 5 * (size_t)ni + 2);
            const uint4 w3 = ldg(nodes + 5 * (size_t)ni + 3);
            const uint4 w4 = ldg(nodes + 5 * (size_t)ni + 4);
            if (COUNT) { cnt->nodes++; cnt->boxes += 8; }
            const uint32_t imask = w0.w >> 24;
            // cell size per axis * 1/d, and the node origin relative to the ray, in ray-parameter units
            const float ax = u2f((w0.w & 0xffu) << 23) * idx;
            const float ay = u2f(((w0.w >> 8) & 0xffu) << 23) * idy;
            const float az = u2f(((w0.w >> 16) & 0xffu) << 23) * idz;
            const float bx = (u2f(w0.x) - r.ox) * idx;
            const float by = (u2f(w0.y) - r.oy) * idy;
            const float bz = (u2f(w0.z) - r.oz) * idz;
            uint32_t hitmask = 0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t meta4 = half ? w1.w : w1.z;
                const uint32_t qlx = half ? w2.y : w2.x, qly = half ? w2.w : w2.z, qlz = half ? w3.y : w3.x;
                const uint32_t qhx = half ? w3.w : w3.z, qhy = half ? w4.y : w4.x, qhz = half ? w4.w : w4.z;
                // near / far plane per axis by the sign of the direction
                const uint32_t nxq = r.dx >= 0.f ? qlx : qhx, fxq = r.dx >= 0.f ? qhx : qlx;
                const uint32_t nyq = r.dy >= 0.f ? qly : qhy, fyq = r.dy >= 0.f ? qhy : qly;
                const uint32_t nzq = r.dz >= 0.f ? qlz : qhz, fzq = r.dz >= 0.f ? qhz : qlz;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float tnx = (float)byte_of(nxq, i) * ax + bx, tfx = (float)byte_of(fxq, i) * ax + bx;
                    const float tny = (float)byte_of(nyq, i) * ay + by, tfy = (float)byte_of(fyq, i) * ay + by;
                    const float tnz = (float)byte_of(nzq, i) * az + bz, tfz = (float)byte_of(fzq, i) * az + bz;
                    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
                    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, best.t));
                    const uint32_t meta = byte_of(meta4, i);
                    if (tn <= tf && meta != 0u) {
                        const uint32_t s = (uint32_t)(half * 4 + i);
                        const uint32_t internal = (imask >> s) & 1u;
                        const uint32_t bits = meta >> 5;
                        const uint32_t pos = (meta & 31u) ^ (internal ? octinv : 0u);
                        hitmask |= bits << pos;
                    }
                }
            }
            G = make_uint2(w1.x, (hitmask & 0xff000000u) | imask);
            T = make_uint2(w1.y, hitmask & 0x00ffffffu);
        }
        while (T.y) {
            const int bit = bfind32(T.y);
            T.y &= ~(1u << bit);
            const uint32_t ti = T.x + (uint32_t)bit;
            const F8 ta = ld256(tris + 4 * (size_t)ti), tb = ld256(tris + 4 * (size_t)ti + 2);
            if (COUNT) cnt->tris++;
            const bool acc = tri_test<NT>(ta.lo, ta.hi, tb.lo, tb.hi, r, best, tris + 4 * (size_t)ti);
            if (ANY && acc) return;
        }
        if ((G.y & 0xff000000u) == 0u) {
            if (sp == 0) return;
            G = stack[--sp];
        }
    }
}

}  // namespace mirogpu
#endif
