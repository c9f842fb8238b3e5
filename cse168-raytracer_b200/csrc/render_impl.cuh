// render_impl.cuh -- Scene::raytraceImage on the device as a wavefront (included by mirogpu.cu once
// mirogpu_scene is complete).
//
// The reference walks one pixel at a time through a recursive Scene::traceScene (Scene.cpp:270-346).  Here a
// frame is a sequence of waves over a queue of path items {ray, pixel, RGB weight, depth left}:
//     trace (closest hit)  ->  k_shade  ->  trace shadow rays  ->  k_shadow_accumulate  ->  next wave
// k_shade evaluates Phong::shade's light loop (Phong.cpp:61-158) up to the shadow test, appends the
// reflection / Fresnel / refraction / diffuse-bounce rays with their weights to the next queue, and the
// shadow pass applies the occlusion rule of Phong.cpp:97-114 (an occluder blocks the light unless it is
// refractive and faces the light, in which case the diffuse term is scaled by N_occ . l).
// Radiance is accumulated per pixel with float atomics, so sums differ from the reference's recursion order
// in the last bits; images are compared with a tolerance (tests/test_gpu_render.py).
#ifndef MIROGPU_RENDER_IMPL_CUH
#define MIROGPU_RENDER_IMPL_CUH

namespace {

using namespace mirogpu;

#define MIRO_MAX_LIGHTS 4

struct WaveParams {
    DeviceScene ds;
    const mirogpu_material* mats;
    const mirogpu_light* lights;
    uint32_t nlights;
    int mode, shadows, max_depth, use_pm;
    float bg[3];
    uint32_t seed, sample;
    uint32_t cap;          // queue capacity
    int width, first_row, row_stride;   // local pixel index -> frame pixel (for shard-independent random numbers)
    uint32_t npix;                      // local pixels; accumulation planes are [sample in batch][pixel][rgb]
    uint32_t sample_base;               // first sample of this batch
};

// weight.w carries (depth left | sample-in-batch << 8) as integer bits
__device__ __forceinline__ float pack_ds(int depth, uint32_t sample) { return __uint_as_float(((uint32_t)depth & 0xffu) | (sample << 8)); }
__device__ __forceinline__ int unpack_depth(float w) { return (int)(__float_as_uint(w) & 0xffu); }
__device__ __forceinline__ uint32_t unpack_sample(float w) { return __float_as_uint(w) >> 8; }

// queue layout (SoA)
struct Queue {
    mirogpu_ray* rays;
    mirogpu_hit* hits;
    uint32_t* pix;         // local pixel index
    float4* weight;        // rgb weight, w = depth left (as float)
};

__global__ void __launch_bounds__(256) k_render_primary(CameraBasis cb, int width, int height, int row_begin, int row_stride,
                                                         int nrows_local, int jitter, uint32_t seed, uint32_t sample_base, uint32_t nsamples,
                                                         int max_depth, Queue q)
{
    // grid: x over the shard's pixels, y over the samples of the batch (32-bit index arithmetic)
    const uint32_t npix = (uint32_t)nrows_local * (uint32_t)width;
    const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= npix) return;
    const uint32_t sb = blockIdx.y;
    const size_t i = (size_t)sb * npix + lp;
    const uint32_t row = lp / (uint32_t)width;
    const int x = (int)(lp - row * (uint32_t)width), y = row_begin + (int)row * row_stride;
    float dx = 0.5f, dy = 0.5f;
    if (jitter) uniform2(seed, (uint32_t)((size_t)y * width + x), sample_base + sb, RNG_DIM_PIXEL, dx, dy);
    const float U = xadd(cb.left, xmul(xsub(cb.right, cb.left), xdiv(xadd((float)x, dx), (float)width)));
    const float V = xadd(cb.bottom, xmul(xsub(cb.top, cb.bottom), xdiv(xadd((float)y, dy), (float)height)));
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = xsub(xadd(xmul(cb.u[k], U), xmul(cb.v[k], V)), cb.w[k]);
    const float inv = xdiv(1.0f, xsqrt(xdot(d[0], d[1], d[2], d[0], d[1], d[2])));
    float4* o = reinterpret_cast<float4*>(q.rays + i);
    o[0] = make_float4(cb.eye[0], cb.eye[1], cb.eye[2], 0.0f);
    o[1] = make_float4(xmul(d[0], inv), xmul(d[1], inv), xmul(d[2], inv), MIROGPU_TMAX);
    q.pix[i] = (uint32_t)lp;
    q.weight[i] = make_float4(1.f, 1.f, 1.f, pack_ds(max_depth, sb));
}

__device__ __forceinline__ float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// fixed_slot >= 0: the child takes that slot (one child per item, queue position = item position: no counter at all);
// otherwise slots are allocated from *next_count.  Returns true when a child was written.
__device__ __forceinline__ bool push_item(const WaveParams& p, Queue& next, uint32_t* next_count, uint32_t* dropped, const float o[3],
                                          const float d[3], uint32_t pixel, float wr, float wg, float wb, float packed_depth_sample,
                                          int64_t fixed_slot = -1)
{
    if (!(wr > 0.f || wg > 0.f || wb > 0.f)) return false;
    if (fixed_slot >= 0) {
        float4* r = reinterpret_cast<float4*>(next.rays + fixed_slot);
        r[0] = make_float4(o[0], o[1], o[2], 0.0f);
        r[1] = make_float4(d[0], d[1], d[2], MIROGPU_TMAX);
        next.pix[fixed_slot] = pixel;
        next.weight[fixed_slot] = make_float4(wr, wg, wb, packed_depth_sample);
        return true;
    }
    // warp-aggregated append: the lanes that reach this point together take consecutive slots with ONE atomic, which
    // also keeps the next wave's rays in the order of this wave's items (neighbouring pixels stay neighbours)
    const unsigned peers = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(next_count, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t slot = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    if (slot >= p.cap) { atomicAdd(dropped, 1u); return false; }
    float4* r = reinterpret_cast<float4*>(next.rays + slot);
    r[0] = make_float4(o[0], o[1], o[2], 0.0f);
    r[1] = make_float4(d[0], d[1], d[2], MIROGPU_TMAX);
    next.pix[slot] = pixel;
    next.weight[slot] = make_float4(wr, wg, wb, packed_depth_sample);
    return true;
}

// pow(float, int) as the reference's gnu++98 build evaluates it (Phong.cpp:152 -> __builtin_powif -> libgcc __powisf2): binary
// exponentiation in binary32, every product rounded.  m >= 0.
__device__ __forceinline__ float powi_ref(float x, int m)
{
    unsigned n = (unsigned)m;
    float y = (n & 1u) ? x : 1.0f;
    while (n >>= 1) {
        x = xmul(x, x);
        if (n & 1u) y = xmul(y, x);
    }
    return y;
}

// One light of Phong::shade's loop up to the shadow query (Phong.cpp:61-158), in the reference's operand order with separately
// rounded operations (no FMA contraction: every kernel that inlines this produces the same bits, and they are the bits of the
// reference's x86 build): unit vector to the light, its distance, the diffuse term colour * ((max(0, N.l * falloff * W) * kd) * kd)
// (Phong::diffuse2D returns m_diffuse, so kd enters twice, Phong.cpp:146, Phong.h:20) and the highlight (exponent fixed at 500,
// Phong.cpp:152).  false: the point lies outside a DirectionalAreaLight's beam.
__device__ __forceinline__ bool light_terms(const mirogpu_light L, const SurfacePoint& sp, const mirogpu_material& m, const float dc[3], const float rd[3],
                                            float l[3], float& dist, float cd[3], float& hl)
{
    if (L.kind == 1) { l[0] = -L.normal[0]; l[1] = -L.normal[1]; l[2] = -L.normal[2]; }
    else { l[0] = xsub(L.position[0], sp.P[0]); l[1] = xsub(L.position[1], sp.P[1]); l[2] = xsub(L.position[2], sp.P[2]); }
    float falloff = xdot(l[0], l[1], l[2], l[0], l[1], l[2]);
    dist = xsqrt(falloff);
    const float invd = xdiv(1.0f, dist);
    l[0] = xmul(l[0], invd); l[1] = xmul(l[1], invd); l[2] = xmul(l[2], invd);
    float nDotL;
    if (L.kind == 1) {
        nDotL = xdot(sp.N[0], sp.N[1], sp.N[2], -L.normal[0], -L.normal[1], -L.normal[2]);
        const float t = xdiv(xdot(L.normal[0], L.normal[1], L.normal[2], xsub(L.position[0], sp.P[0]), xsub(L.position[1], sp.P[1]), xsub(L.position[2], sp.P[2])), -1.0f);
        const float q0 = xsub(xsub(sp.P[0], xmul(L.normal[0], t)), L.position[0]), q1 = xsub(xsub(sp.P[1], xmul(L.normal[1], t)), L.position[1]),
                    q2 = xsub(xsub(sp.P[2], xmul(L.normal[2], t)), L.position[2]);
        if (xdot(q0, q1, q2, q0, q1, q2) > xmul(L.radius, L.radius)) return false;
        falloff = xdiv(1.0f, MIRO_PI);
    } else {
        nDotL = xdot(sp.N[0], sp.N[1], sp.N[2], l[0], l[1], l[2]);
        falloff = xdiv(1.0f, xmul(xmul(xmul(falloff, 4.0f), MIRO_PI), MIRO_PI));
    }
    const float dterm = fmaxf(0.0f, xmul(xmul(nDotL, falloff), L.wattage));
    // dc: the diffuse colour looked up for this point (Phong.cpp:50-55) -- m.kd itself for an untextured Phong
    cd[0] = xmul(L.color[0], xmul(xmul(dc[0], dterm), m.kd[0]));
    cd[1] = xmul(L.color[1], xmul(xmul(dc[1], dterm), m.kd[1]));
    cd[2] = xmul(L.color[2], xmul(xmul(dc[2], dterm), m.kd[2]));
    hl = 0.f;
    if (m.shininess < INFINITY) {
        const float ldn = xmul(2.0f, xdot(l[0], l[1], l[2], sp.N[0], sp.N[1], sp.N[2]));
        const float r0 = xadd(-l[0], xmul(sp.N[0], ldn)), r1 = xadd(-l[1], xmul(sp.N[1], ldn)), r2 = xadd(-l[2], xmul(sp.N[2], ldn));
        const float c = fmaxf(0.0f, fminf(1.f, xdot(-rd[0], -rd[1], -rd[2], r0, r1, r2)));
        hl = fmaxf(0.0f, xmul(xmul(powi_ref(c, 500), falloff), L.wattage));
    }
    return true;
}

// Returns true when the item wrote its (single) child in place -- diffuse-bounce mode only.
__device__ __forceinline__ bool shade_item(const WaveParams& p, const Queue& cur, uint32_t i, const float4 hv, const ShadeRecord& rec, Queue& next,
                                           uint32_t* next_count, uint32_t* dropped, mirogpu_ray* shadow_rays, float4* shadow_cd, float4* shadow_ch,
                                           float* accum, float* gather_pos, float* gather_nrm, float4* gather_w)
{
    const float4 w = cur.weight[i];
    const uint32_t lpix = cur.pix[i];
    const uint32_t sb = unpack_sample(w.w);
    const uint32_t pixel = sb * p.npix + lpix;     // accumulation plane of this item's sample
    const mirogpu_ray ray = load_ray(cur.rays, i);
    if (gather_w) gather_w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.shadows)
        for (uint32_t li = 0; li < p.nlights; ++li) shadow_rays[(size_t)i * p.nlights + li].tmax = -1.0f;   // inactive unless set below
    if (__float_as_uint(hv.y) == MIROGPU_MISS) {
        if (ray.tmax < ray.tmin) return false;   // dead item
        // Scene.cpp:338-342: environment / background colour
        atomicAdd(accum + 3 * (size_t)pixel + 0, xmul(w.x, p.bg[0]));
        atomicAdd(accum + 3 * (size_t)pixel + 1, xmul(w.y, p.bg[1]));
        atomicAdd(accum + 3 * (size_t)pixel + 2, xmul(w.z, p.bg[2]));
        return false;
    }
    mirogpu_hit h; h.t = hv.x; h.prim_id = __float_as_uint(hv.y); h.beta = hv.z; h.gamma = hv.w;
    SurfacePoint sp;
    float dc[3];
    if (p.ds.textured) {
        const TexturedPoint tp = resolve_hit_textured(rec, h, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, p.mats, p.ds.uvs);
        sp = tp.sp; dc[0] = tp.dc[0]; dc[1] = tp.dc[1]; dc[2] = tp.dc[2];
    } else sp = resolve_hit(rec, h, ray);
    const mirogpu_material m = p.mats[sp.material];
    if (!p.ds.textured) { dc[0] = m.kd[0]; dc[1] = m.kd[1]; dc[2] = m.kd[2]; }
    const float rd[3] = {ray.dx, ray.dy, ray.dz};
    float direct[3] = {0.f, 0.f, 0.f};

    // ---- Phong::shade light loop -------------------------------------------------------------------
    for (uint32_t li = 0; li < p.nlights; ++li) {
        float l[3], cd[3], dist, hl;
        if (!light_terms(p.lights[li], sp, m, dc, rd, l, dist, cd, hl)) continue;
        if (p.shadows) {
            if (cd[0] > 0.f || cd[1] > 0.f || cd[2] > 0.f || hl > 0.f) {
                const size_t s = (size_t)i * p.nlights + li;
                float4* r = reinterpret_cast<float4*>(shadow_rays + s);
                r[0] = make_float4(xadd(sp.P[0], xmul(l[0], MIRO_EPS)), xadd(sp.P[1], xmul(l[1], MIRO_EPS)), xadd(sp.P[2], xmul(l[2], MIRO_EPS)), 0.0f);
                r[1] = make_float4(l[0], l[1], l[2], dist);
                shadow_cd[s] = make_float4(w.x * cd[0], w.y * cd[1], w.z * cd[2], __uint_as_float(pixel));
                shadow_ch[s] = make_float4(w.x * hl, w.y * hl, w.z * hl, 0.f);
            }
        } else {
            direct[0] = xadd(xadd(direct[0], cd[0]), hl); direct[1] = xadd(xadd(direct[1], cd[1]), hl); direct[2] = xadd(xadd(direct[2], cd[2]), hl);   // L = L + diffuse; L = L + highlights (Phong.cpp:146-156)
        }
    }
    if (!p.shadows && (direct[0] != 0.f || direct[1] != 0.f || direct[2] != 0.f)) {
        atomicAdd(accum + 3 * (size_t)pixel + 0, xmul(w.x, direct[0]));
        atomicAdd(accum + 3 * (size_t)pixel + 1, xmul(w.y, direct[1]));
        atomicAdd(accum + 3 * (size_t)pixel + 2, xmul(w.z, direct[2]));
    }
    const bool diffuse = m.kd[0] > 0.f || m.kd[1] > 0.f || m.kd[2] > 0.f;
    // photon-map irradiance at diffuse hits (Scene.cpp:286-299): queued for the gather kernel
    if (p.use_pm && diffuse && gather_w) {
        gather_pos[3 * (size_t)i] = sp.P[0]; gather_pos[3 * (size_t)i + 1] = sp.P[1]; gather_pos[3 * (size_t)i + 2] = sp.P[2];
        gather_nrm[3 * (size_t)i] = sp.N[0]; gather_nrm[3 * (size_t)i + 1] = sp.N[1]; gather_nrm[3 * (size_t)i + 2] = sp.N[2];
        gather_w[i] = make_float4(w.x, w.y, w.z, 1.f);
    }
    // ---- secondary rays -------------------------------------------------------------------------------
    const int depth_now = unpack_depth(w.w);
    const int depth_next = depth_now - 1;          // --depth (Scene.cpp:282)
    if (depth_next < 0) return false;
    const float depth_left = pack_ds(depth_next, sb);
    if (p.mode == MIROGPU_RENDER_DIFFUSE_BOUNCE) {
        if (diffuse && depth_now == p.max_depth) {   // one cosine-weighted bounce from the first hit only
            float u1, u2;
            const uint32_t frame_pixel = (uint32_t)(p.first_row + (int)(lpix / (uint32_t)p.width) * p.row_stride) * (uint32_t)p.width + lpix % (uint32_t)p.width;
            uniform2(p.seed, frame_pixel, p.sample_base + sb, RNG_DIM_BOUNCE, u1, u2);
            float d[3];
            align_hemisphere(sp.N, xmul(xmul(2.0f, MIRO_PI), u2), asinf(sqrtf(u1)), d);
            const float o[3] = {xadd(sp.P[0], xmul(d[0], MIRO_EPS)), xadd(sp.P[1], xmul(d[1], MIRO_EPS)), xadd(sp.P[2], xmul(d[2], MIRO_EPS))};
            return push_item(p, next, next_count, dropped, o, d, lpix, xmul(w.x, m.kd[0]), xmul(w.y, m.kd[1]), xmul(w.z, m.kd[2]), pack_ds(0, sb), (int64_t)i);
        }
        return false;
    }
    if (p.mode != MIROGPU_RENDER_WHITTED) return false;
    const bool reflective = m.ks[0] > 0.f || m.ks[1] > 0.f || m.ks[2] > 0.f;
    const bool refractive = m.kt[0] > 0.f || m.kt[1] > 0.f || m.kt[2] > 0.f;
    if (!reflective && !refractive) return false;
    // Ray::reflect (Ray.h:160-163)
    const float dn = dot3(sp.N, rd);
    float dr[3] = {rd[0] - 2.f * dn * sp.N[0], rd[1] - 2.f * dn * sp.N[1], rd[2] - 2.f * dn * sp.N[2]};
    {
        const float inv = 1.0f / sqrtf(dot3(dr, dr));
        dr[0] *= inv; dr[1] *= inv; dr[2] *= inv;
    }
    const float orr[3] = {sp.P[0] + dr[0] * MIRO_EPS, sp.P[1] + dr[1] * MIRO_EPS, sp.P[2] + dr[2] * MIRO_EPS};
    if (reflective) push_item(p, next, next_count, dropped, orr, dr, lpix, w.x * m.ks[0], w.y * m.ks[1], w.z * m.ks[2], depth_left);
    if (refractive) {
        // Fresnel coefficient (Ray.h:168-200) and Snell refraction with TIR fallback (Ray.h:202-243)
        float n1, n2, n[3];
        if (dn < 0.f) { n1 = 1.0f; n2 = m.refract_index; n[0] = sp.N[0]; n[1] = sp.N[1]; n[2] = sp.N[2]; }
        else { n1 = m.refract_index; n2 = 1.0f; n[0] = -sp.N[0]; n[1] = -sp.N[1]; n[2] = -sp.N[2]; }
        const float nrd[3] = {-rd[0], -rd[1], -rd[2]};
        const float cosT = dot3(nrd, n);
        const float sinT = sinf(acosf(cosT));
        const float ps = (n1 / n2) * sinT * ((n1 / n2) * sinT);
        float Rs;
        if (ps > 1.f) Rs = 1.f;
        else {
            const float sq = sqrtf(1.f - ps);
            const float q = (n1 * cosT - sq) / (n1 * cosT + sq);
            Rs = q * q;
        }
        if (Rs > 0.01f) push_item(p, next, next_count, dropped, orr, dr, lpix, w.x * m.kt[0] * Rs, w.y * m.kt[1] * Rs, w.z * m.kt[2] * Rs, depth_left);
        const float ddn = dot3(rd, n);
        const float energy = 1.f - (n1 * n1 * (1.f - ddn * ddn) / (n2 * n2));
        const float tw = 1.f - Rs;
        if (energy < 0.f) {
            push_item(p, next, next_count, dropped, orr, dr, lpix, w.x * m.kt[0] * tw, w.y * m.kt[1] * tw, w.z * m.kt[2] * tw, depth_left);
        } else {
            const float se = sqrtf(energy);
            float dt[3];   // the refracted direction is NOT normalised in the reference (Ray.h:233)
#pragma unroll
            for (int k = 0; k < 3; ++k) dt[k] = n1 * (rd[k] - n[k] * ddn) / n2 - n[k] * se;
            const float ot[3] = {sp.P[0] + dt[0] * MIRO_EPS, sp.P[1] + dt[1] * MIRO_EPS, sp.P[2] + dt[2] * MIRO_EPS};
            push_item(p, next, next_count, dropped, ot, dt, lpix, w.x * m.kt[0] * tw, w.y * m.kt[1] * tw, w.z * m.kt[2] * tw, depth_left);
        }
    }
    return false;
}

// One wave of Scene::traceScene bodies.  shadow_* arrays have nlights slots per item.
// Diffuse-bounce mode: an item has at most one child, so the child takes the item's own queue position and no slot
// counter is touched (a single-address atomic per warp serialises at the L2 atomic unit: 520 k of them made this kernel
// as slow as the trace it feeds); items without a child leave a dead ray (tmax < tmin) there instead.
// MIRO_SHADE_ITEMS items per thread (i, i + 128, ... of a tile): all hits, then all 96-byte shading records are requested before
// any item is shaded.  Measured: 2 items per thread pay in k_gen_bounce (same fetch chain) but not here -- 96 registers and the
// queue / accumulation atomics of two items in one thread cost more than the overlap wins (e2e 6.93 -> 6.71 Grays/s) -- so 1.
#define MIRO_SHADE_THREADS 128
#define MIRO_SHADE_ITEMS 1
// counter block of a render (uint32 slots): [0..16] wave sizes, [17] dropped children, [18] tone-map maximum, [20..21] 64-bit
// secondary-ray total of the frame, [24..55] live children of in-place waves (spread by block)
#define MIRO_DROPPED_SLOT 17
#define MIRO_LIVE_SLOT0 24
__global__ void __launch_bounds__(MIRO_SHADE_THREADS, 8) k_shade(WaveParams p, Queue cur, uint32_t n, const uint32_t* __restrict__ d_n, Queue next,
                                                              uint32_t* next_count, uint32_t* dropped, mirogpu_ray* shadow_rays, float4* shadow_cd,
                                                              float4* shadow_ch, float* accum, float* gather_pos, float* gather_nrm, float4* gather_w)
{
    const uint32_t base = blockIdx.x * (MIRO_SHADE_THREADS * MIRO_SHADE_ITEMS) + threadIdx.x;
    if (d_n) n = min(n, *d_n);
    const bool in_place = p.mode == MIROGPU_RENDER_DIFFUSE_BOUNCE && d_n == nullptr;   // wave 0: the only one with children
    if (in_place && base == 0) *next_count = n;
    uint32_t* live_slots = dropped + (MIRO_LIVE_SLOT0 - MIRO_DROPPED_SLOT);
    bool any_child = false;
    float4 hv[MIRO_SHADE_ITEMS];
    ShadeRecord rec[MIRO_SHADE_ITEMS];
#pragma unroll
    for (int k = 0; k < MIRO_SHADE_ITEMS; ++k) {
        const uint32_t i = base + k * MIRO_SHADE_THREADS;
        hv[k] = i < n ? __ldg(reinterpret_cast<const float4*>(cur.hits + i)) : make_float4(0.f, __uint_as_float(MIROGPU_MISS), 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < MIRO_SHADE_ITEMS; ++k) {
        if (__float_as_uint(hv[k].y) != MIROGPU_MISS) rec[k] = load_shade_record(p.ds, __float_as_uint(hv[k].y));
        else rec[k].r0.lo = rec[k].r0.hi = rec[k].r1.lo = rec[k].r1.hi = rec[k].r2.lo = rec[k].r2.hi = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < MIRO_SHADE_ITEMS; ++k) {
        const uint32_t i = base + k * MIRO_SHADE_THREADS;
        if (i >= n) continue;
        const bool wrote = shade_item(p, cur, i, hv[k], rec[k], next, next_count, dropped, shadow_rays, shadow_cd, shadow_ch, accum, gather_pos,
                                      gather_nrm, gather_w);
        if (in_place && !wrote) {
            float4* r = reinterpret_cast<float4*>(next.rays + i);
            r[0] = make_float4(0.f, 0.f, 0.f, 0.0f);
            r[1] = make_float4(0.f, 0.f, 1.f, -1.0f);
        }
        any_child |= wrote;
    }
    // In-place waves fill every slot, dead or alive; the LIVE children are what the frame's ray count reports (a dead slot is
    // answered without traversal).  One non-returning add per warp, spread over 32 addresses by block.
    if (in_place && MIRO_SHADE_ITEMS == 1) {
        const unsigned live = __popc(__ballot_sync(0xffffffffu, any_child));
        if ((threadIdx.x & 31u) == 0u && live) atomicAdd(live_slots + (blockIdx.x & 31u), live);
    }
}

// Phong.cpp:97-114 applied to the traced shadow rays.
__global__ void __launch_bounds__(256) k_shadow_accumulate(WaveParams p, const mirogpu_ray* __restrict__ srays, const mirogpu_hit* __restrict__ shits,
                                                           const float4* __restrict__ cd, const float4* __restrict__ ch, size_t n,
                                                           const uint32_t* __restrict__ d_n, uint32_t mult, float* accum)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d_n) n = min(n, (size_t)*d_n * mult);
    if (i >= n) return;
    const mirogpu_ray r = load_ray(srays, i);
    if (r.tmax < r.tmin) return;   // no shadow ray in this slot
    const float4 hv = __ldg(reinterpret_cast<const float4*>(shits + i));
    float intensity = 1.f;
    if (__float_as_uint(hv.y) != MIROGPU_MISS) {
        mirogpu_hit h; h.t = hv.x; h.prim_id = __float_as_uint(hv.y); h.beta = hv.z; h.gamma = hv.w;
        const ShadeRecord orec = load_shade_record(p.ds, h.prim_id);
        const SurfacePoint sp = p.ds.textured ? resolve_hit_textured(orec, h, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, p.mats, p.ds.uvs).sp : resolve_hit(orec, h, r);
        const mirogpu_material m = p.mats[sp.material];
        if (!(m.kt[0] > 0.f || m.kt[1] > 0.f || m.kt[2] > 0.f)) return;      // opaque occluder
        const float l[3] = {r.dx, r.dy, r.dz};
        intensity = dot3(sp.N, l);
        if (intensity < 0.f || intensity < MIRO_EPS) return;
    }
    const float4 d = cd[i], s = ch[i];
    const uint32_t pixel = __float_as_uint(d.w);
    atomicAdd(accum + 3 * (size_t)pixel + 0, d.x * intensity + s.x);
    atomicAdd(accum + 3 * (size_t)pixel + 1, d.y * intensity + s.y);
    atomicAdd(accum + 3 * (size_t)pixel + 2, d.z * intensity + s.z);
}

__global__ void __launch_bounds__(256) k_gather_accumulate(const float4* __restrict__ gw, const uint32_t* __restrict__ pix,
                                                           const float* __restrict__ irr0, const float* __restrict__ irr1, uint32_t n,
                                                           const uint32_t* __restrict__ d_n, const float4* __restrict__ weight, uint32_t npix, float* accum)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (d_n) n = min(n, *d_n);
    if (i >= n) return;
    const float4 w = gw[i];
    if (w.w == 0.f) return;
    const uint32_t pixel = unpack_sample(weight[i].w) * npix + pix[i];
    for (int k = 0; k < 3; ++k) {
        const float e = (irr0 ? irr0[3 * (size_t)i + k] : 0.f) + (irr1 ? irr1[3 * (size_t)i + k] : 0.f);
        const float wk = k == 0 ? w.x : (k == 1 ? w.y : w.z);
        atomicAdd(accum + 3 * (size_t)pixel + k, wk * e);
    }
}

// Adds the per-sample planes of one batch, in sample order, into the frame accumulator (deterministic sums).
__global__ void __launch_bounds__(256) k_fold_planes(const float* __restrict__ planes, uint32_t nsamples, size_t nvals, float* __restrict__ frame)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvals) return;
    float a = frame[i];
    for (uint32_t s = 0; s < nsamples; ++s) a += planes[(size_t)s * nvals + i];
    frame[i] = a;
}

// frame accumulator (local pixels) -> rgb (full-frame layout), divided by the sample count like Scene.cpp:138.
__global__ void __launch_bounds__(256) k_resolve_frame(const float* __restrict__ accum, int width, int row_begin, int row_stride, int nrows_local,
                                                        float inv_spp, int spp, float* __restrict__ rgb, float* gmax)
{
    const size_t n = (size_t)nrows_local * width * 3;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = -INFINITY;
    if (i < n) {
        const size_t lp = i / 3; const int c = (int)(i % 3);
        const int x = (int)(lp % width), y = row_begin + (int)(lp / width) * row_stride;
        float a = accum[i];
        if (spp > 1) a *= inv_spp;
        rgb[3 * ((size_t)y * width + x) + c] = a;
        if (a == a) v = a;
    }
    // block max -> global max (for the NaN replacement of the tone map)
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) v = fmaxf(v, sm[k]);
        unsigned int* g = reinterpret_cast<unsigned int*>(gmax);   // float atomic max by CAS on the bit pattern
        unsigned int old = *g;
        while (v > __uint_as_float(old)) {
            const unsigned int assumed = old;
            old = atomicCAS(g, assumed, __float_as_uint(v));
            if (old == assumed) break;
        }
    }
}

// Scene.cpp:177-202: NaN -> maxIntensity, then sigmoid(6v - 3); optionally Image::Map()'s 8-bit truncation
// (Image.cpp:47-52).  Works on rows row_begin + j*row_stride of a full-frame float buffer.
__global__ void __launch_bounds__(256) k_tonemap(float* rgb, unsigned char* rgb8, int width, int row_begin, int row_stride, int nrows_local,
                                                  const float* gmax)
{
    const size_t n = (size_t)nrows_local * width * 3;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t lp = i / 3; const int c = (int)(i % 3);
    const int x = (int)(lp % width), y = row_begin + (int)(lp / width) * row_stride;
    const size_t o = 3 * ((size_t)y * width + x) + c;
    float v = rgb[o];
    if (v != v) v = *gmax;
    v = 1.0f / (1.0f + expf(-(6.0f * v - 3.0f)));
    if (rgb8) { const float m = 255.0f * v; rgb8[o] = m > 255.0f ? 255 : (unsigned char)m; }
    else rgb[o] = v;
}

// Largest non-NaN value over rows row_begin + j*row_stride of a full-frame buffer (atomic max into *gmax).
__global__ void __launch_bounds__(256) k_frame_max_rows(const float* __restrict__ rgb, int width, int row_begin, int row_stride, int nrows_local, float* gmax)
{
    const size_t n = (size_t)nrows_local * width * 3;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = -INFINITY;
    if (i < n) {
        const size_t per_row = (size_t)width * 3;
        const float a = rgb[((size_t)row_begin + (i / per_row) * (size_t)row_stride) * per_row + i % per_row];
        if (a == a) v = a;
    }
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) v = fmaxf(v, sm[k]);
        unsigned int* g = reinterpret_cast<unsigned int*>(gmax);
        unsigned int old = *g;
        while (v > __uint_as_float(old)) {
            const unsigned int assumed = old;
            old = atomicCAS(g, assumed, __float_as_uint(v));
            if (old == assumed) break;
        }
    }
}

__global__ void __launch_bounds__(256) k_frame_max(const float* __restrict__ rgb, size_t n, float* gmax)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = -INFINITY;
    if (i < n) { const float a = rgb[i]; if (a == a) v = a; }
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) v = fmaxf(v, sm[k]);
        unsigned int* g = reinterpret_cast<unsigned int*>(gmax);
        unsigned int old = *g;
        while (v > __uint_as_float(old)) {
            const unsigned int assumed = old;
            old = atomicCAS(g, assumed, __float_as_uint(v));
            if (old == assumed) break;
        }
    }
}


// ---- BASELINE config 3 as two fused waves (MIROGPU_RENDER_DIFFUSE_BOUNCE without shadow rays or photon maps) -----------------
// The general wavefront above carries {ray, pixel, weight, depth} per path item through queues.  In this mode none of that
// is data: item i IS (sample i / npix, pixel i % npix), wave 0's weight is 1 and its ray is a pure function of (pixel, sample)
// -- recomputed here (Camera::eyeRay, ~40 instructions) instead of being re-read -- and wave 1's weight is the kd of the
// material wave 0 hit (4 bytes).  Every item owns its slot of the per-sample plane, so radiance is stored, not accumulated
// atomically: wave 0 writes its direct term, wave 1 adds its own -- the same two floating-point additions, in the same order,
// as the general path's atomics, hence bit-identical frames.  The bounce ray overwrites the camera ray in place.
// Per item: wave 0 moves 16 (hit) + 96 (shading record) + 32 (bounce ray) + 4 + 12 bytes, wave 1 16 + 16 + 96 + 4 + 24.
#define MIRO_BW_THREADS 128
// READ_DIR: take the camera ray's direction from the ray buffer (16 bytes per item) instead of recomputing it (~150 instructions)
template <int ITEMS, bool READ_DIR>
__global__ void __launch_bounds__(MIRO_BW_THREADS) k_bounce_wave0(WaveParams p, CameraBasis cb, int height, int nrows_local, int jitter, uint32_t nitems,
                                                                  const mirogpu_hit* __restrict__ hits, mirogpu_ray* __restrict__ rays,
                                                                  uint32_t* __restrict__ parent_mat, float* __restrict__ planes, uint32_t* live_slots)
{
    const uint32_t base = blockIdx.x * (MIRO_BW_THREADS * ITEMS) + threadIdx.x;
    float4 hv[ITEMS];
    float4 dv[ITEMS];
    ShadeRecord rec[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = base + k * MIRO_BW_THREADS;
        hv[k] = i < nitems ? __ldcs(reinterpret_cast<const float4*>(hits + i)) : make_float4(0.f, __uint_as_float(MIROGPU_MISS), 0.f, 0.f);
        if (READ_DIR) dv[k] = i < nitems ? __ldcs(reinterpret_cast<const float4*>(rays + i) + 1) : make_float4(0.f, 0.f, 1.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (__float_as_uint(hv[k].y) != MIROGPU_MISS) rec[k] = load_shade_record(p.ds, __float_as_uint(hv[k].y));
        else rec[k].r0.lo = rec[k].r0.hi = rec[k].r1.lo = rec[k].r1.hi = rec[k].r2.lo = rec[k].r2.hi = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    unsigned live = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = base + k * MIRO_BW_THREADS;
        bool child = false;
        if (i < nitems) {
            const uint32_t sb = i / p.npix, lpix = i - sb * p.npix;
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.0f), r1 = make_float4(0.f, 0.f, 1.f, -1.0f);   // dead slot: tmax < tmin
            uint32_t mat = MIROGPU_MISS;
            float out[3];
            if (__float_as_uint(hv[k].y) == MIROGPU_MISS) {
                out[0] = p.bg[0]; out[1] = p.bg[1]; out[2] = p.bg[2];                                // Scene.cpp:338-342
            } else {
                // the camera ray of this item, as k_gen_primary made it (bit-exact: same operations)
                const uint32_t row = lpix / (uint32_t)p.width;
                const int x = (int)(lpix - row * (uint32_t)p.width), y = p.first_row + (int)row * p.row_stride;
                const uint32_t frame_pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
                float rd[3];
                if (READ_DIR) { rd[0] = dv[k].x; rd[1] = dv[k].y; rd[2] = dv[k].z; }
                else {
                    float dx = 0.5f, dy = 0.5f;
                    if (jitter) uniform2(p.seed, frame_pixel, p.sample_base + sb, RNG_DIM_PIXEL, dx, dy);
                    const float U = xadd(cb.left, xmul(xsub(cb.right, cb.left), xdiv(xadd((float)x, dx), (float)p.width)));
                    const float V = xadd(cb.bottom, xmul(xsub(cb.top, cb.bottom), xdiv(xadd((float)y, dy), (float)height)));
#pragma unroll
                    for (int c = 0; c < 3; ++c) rd[c] = xsub(xadd(xmul(cb.u[c], U), xmul(cb.v[c], V)), cb.w[c]);
                    const float inv = xdiv(1.0f, xsqrt(xdot(rd[0], rd[1], rd[2], rd[0], rd[1], rd[2])));
                    rd[0] = xmul(rd[0], inv); rd[1] = xmul(rd[1], inv); rd[2] = xmul(rd[2], inv);
                }
                mirogpu_hit h; h.t = hv[k].x; h.prim_id = __float_as_uint(hv[k].y); h.beta = hv[k].z; h.gamma = hv[k].w;
                const SurfacePoint sp = record_kind(rec[k]) == 0u ? resolve_hit(rec[k], h)
                                                                  : resolve_hit_analytic(rec[k], h, cb.eye[0], cb.eye[1], cb.eye[2], rd[0], rd[1], rd[2]);
                const mirogpu_material m = p.mats[sp.material];
                float direct[3] = {0.f, 0.f, 0.f};
                for (uint32_t li = 0; li < p.nlights; ++li) {
                    float l[3], cd[3], dist, hl;
                    if (!light_terms(p.lights[li], sp, m, m.kd, rd, l, dist, cd, hl)) continue;
                    direct[0] = xadd(xadd(direct[0], cd[0]), hl); direct[1] = xadd(xadd(direct[1], cd[1]), hl); direct[2] = xadd(xadd(direct[2], cd[2]), hl);   // L = L + diffuse; L = L + highlights (Phong.cpp:146-156)
                }
                out[0] = 1.f * direct[0]; out[1] = 1.f * direct[1]; out[2] = 1.f * direct[2];
                if (m.kd[0] > 0.f || m.kd[1] > 0.f || m.kd[2] > 0.f) {   // Ray::diffuse from the first hit (Ray.h:109-122)
                    float u1, u2;
                    uniform2(p.seed, frame_pixel, p.sample_base + sb, RNG_DIM_BOUNCE, u1, u2);
                    float d[3];
                    align_hemisphere(sp.N, xmul(xmul(2.0f, MIRO_PI), u2), asinf(sqrtf(u1)), d);
                    r0 = make_float4(xadd(sp.P[0], xmul(d[0], MIRO_EPS)), xadd(sp.P[1], xmul(d[1], MIRO_EPS)), xadd(sp.P[2], xmul(d[2], MIRO_EPS)), 0.0f);
                    r1 = make_float4(d[0], d[1], d[2], MIROGPU_TMAX);
                    mat = sp.material;
                    child = true;
                }
            }
            float4* r = reinterpret_cast<float4*>(rays + i);
            __stcs(r, r0); __stcs(r + 1, r1);
            parent_mat[i] = mat;
            float* o = planes + 3 * (size_t)i;     // plane of sample sb, pixel lpix: (sb * npix + lpix) = i
            o[0] = out[0]; o[1] = out[1]; o[2] = out[2];
        }
        live += __popc(__ballot_sync(0xffffffffu, child));
    }
    if ((threadIdx.x & 31u) == 0u && live) atomicAdd(live_slots + (blockIdx.x & 31u), live);
}

template <int ITEMS>
__global__ void __launch_bounds__(MIRO_BW_THREADS) k_bounce_wave1(WaveParams p, uint32_t nitems, const mirogpu_hit* __restrict__ hits,
                                                                  const mirogpu_ray* __restrict__ rays, const uint32_t* __restrict__ parent_mat,
                                                                  float* __restrict__ planes)
{
    const uint32_t base = blockIdx.x * (MIRO_BW_THREADS * ITEMS) + threadIdx.x;
    float4 hv[ITEMS];
    uint32_t pm[ITEMS];
    ShadeRecord rec[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = base + k * MIRO_BW_THREADS;
        pm[k] = i < nitems ? __ldcs(parent_mat + i) : MIROGPU_MISS;
        hv[k] = (i < nitems && pm[k] != MIROGPU_MISS) ? __ldcs(reinterpret_cast<const float4*>(hits + i)) : make_float4(0.f, __uint_as_float(MIROGPU_MISS), 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (__float_as_uint(hv[k].y) != MIROGPU_MISS) rec[k] = load_shade_record(p.ds, __float_as_uint(hv[k].y));
        else rec[k].r0.lo = rec[k].r0.hi = rec[k].r1.lo = rec[k].r1.hi = rec[k].r2.lo = rec[k].r2.hi = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = base + k * MIRO_BW_THREADS;
        if (i >= nitems || pm[k] == MIROGPU_MISS) continue;        // no bounce ray in this slot
        const mirogpu_material pmat = p.mats[pm[k]];
        const float w[3] = {1.f * pmat.kd[0], 1.f * pmat.kd[1], 1.f * pmat.kd[2]};
        float add[3];
        if (__float_as_uint(hv[k].y) == MIROGPU_MISS) {
            add[0] = __fmul_rn(w[0], p.bg[0]); add[1] = __fmul_rn(w[1], p.bg[1]); add[2] = __fmul_rn(w[2], p.bg[2]);
        } else {
            const float4 dv = __ldcs(reinterpret_cast<const float4*>(rays + i) + 1);
            const float rd[3] = {dv.x, dv.y, dv.z};
            mirogpu_hit h; h.t = hv[k].x; h.prim_id = __float_as_uint(hv[k].y); h.beta = hv[k].z; h.gamma = hv[k].w;
            const SurfacePoint sp = resolve_hit(rec[k], h, rays + i);
            const mirogpu_material m = p.mats[sp.material];
            float direct[3] = {0.f, 0.f, 0.f};
            for (uint32_t li = 0; li < p.nlights; ++li) {
                float l[3], cd[3], dist, hl;
                if (!light_terms(p.lights[li], sp, m, m.kd, rd, l, dist, cd, hl)) continue;
                direct[0] = xadd(xadd(direct[0], cd[0]), hl); direct[1] = xadd(xadd(direct[1], cd[1]), hl); direct[2] = xadd(xadd(direct[2], cd[2]), hl);   // L = L + diffuse; L = L + highlights (Phong.cpp:146-156)
            }
            if (direct[0] == 0.f && direct[1] == 0.f && direct[2] == 0.f) continue;
            add[0] = __fmul_rn(w[0], direct[0]); add[1] = __fmul_rn(w[1], direct[1]); add[2] = __fmul_rn(w[2], direct[2]);
        }
        // separately rounded product and sum, like the general path's atomicAdd(plane, w * direct) -- no FMA contraction
        float* o = planes + 3 * (size_t)i;
        o[0] = __fadd_rn(o[0], __fmul_rn(add[0], 1.0f)); o[1] = __fadd_rn(o[1], __fmul_rn(add[1], 1.0f)); o[2] = __fadd_rn(o[2], __fmul_rn(add[2], 1.0f));
    }
}

#define MIRO_MAX_WAVES 16
#define MIRO_SAMPLE_BATCH 8

// Adds this batch's secondary wave sizes (counters[1..16]) to the frame's running 64-bit total.
// in_place (diffuse-bounce frames): wave 1 holds a slot per item of wave 0, dead or alive; its rays are the live children
// k_shade counted.
__global__ void k_sum_wave_counters(const uint32_t* __restrict__ counters, unsigned long long* total, int in_place)
{
    unsigned long long s = 0;
    if (in_place) for (int k = 0; k < 32; ++k) s += counters[MIRO_LIVE_SLOT0 + k];
    else for (int w = 1; w <= MIRO_MAX_WAVES; ++w) s += counters[w];
    *total += s;
}

// d_rgb: full-frame float buffer (rows of this shard are written).  d_rgb8 (may be NULL): full-frame 8-bit
// tone-mapped output (requires rp.tonemap semantics; used by mirogpu_render_rgb8).
// One attempt at a frame.  *overflow (general wavefront, refractive scenes): a wave produced more children than the path queues
// hold (the reference recurses to TRACE_DEPTH with every child traced, Scene.cpp:301-335; a refractive hit spawns up to three) --
// nothing has been delivered, the caller retries with larger queues.
int render_device_once(mirogpu_scene* h, const mirogpu_camera& cam, const mirogpu_render_params& rp, float* d_rgb, unsigned char* d_rgb8,
                       cudaStream_t st, std::string& err, bool* overflow)
{
    *overflow = false;
#define RT(expr)                                                                     \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(_e); return _e == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA; } \
    } while (0)
    if (rp.width <= 0 || rp.height <= 0 || rp.spp < 1 || rp.row_stride < 1 || rp.row_phase < 0 || rp.row_phase >= rp.row_stride ||
        rp.row_begin < 0 || rp.row_end > rp.height || rp.row_begin > rp.row_end || rp.max_depth < 0 || rp.max_depth > 200) { err = "bad render parameters"; return MIROGPU_ERR_INVALID_ARG; }
    if (rp.mode != MIROGPU_RENDER_WHITTED && rp.mode != MIROGPU_RENDER_DIFFUSE_BOUNCE && rp.mode != MIROGPU_RENDER_PRIMARY_ONLY) { err = "unknown render mode"; return MIROGPU_ERR_INVALID_ARG; }
    if (h->nlights > MIRO_MAX_LIGHTS) { err = "too many lights (max 4)"; return MIROGPU_ERR_UNSUPPORTED; }
    std::lock_guard<std::mutex> lk(h->mtx);   // one render at a time per handle (scratch buffers)
    const int first_row = rp.row_begin + rp.row_phase;
    const int nrows = first_row < rp.row_end ? (rp.row_end - first_row + rp.row_stride - 1) / rp.row_stride : 0;
    const size_t npix = (size_t)nrows * rp.width;
    h->last_rays = 0; h->last_launches = 0;
    for (int k = 0; k < 64; ++k) h->h_stats[k] = 0;
    if (npix == 0) return MIROGPU_OK;
    const uint32_t batch = (uint32_t)std::min(rp.spp, MIRO_SAMPLE_BATCH);
    const size_t items0 = npix * batch;
    if (items0 >= (1ull << 30)) { err = "frame too large"; return MIROGPU_ERR_UNSUPPORTED; }

    const bool any_refractive = h->any_refractive, any_specular = h->any_specular;
    const bool shadows = rp.shadows != 0 && rp.mode != MIROGPU_RENDER_PRIMARY_ONLY && h->nlights > 0;
    const uint32_t nl = std::max<uint32_t>(h->nlights, 1);
    const bool whitted_secondary = rp.mode == MIROGPU_RENDER_WHITTED && (any_specular || any_refractive);
    // Wave plan: diffuse-only Whitted frames have exactly one wave, diffuse-bounce frames exactly two (no host
    // round trip between waves: counts stay on the device); specular scenes read the queue size back each wave.
    const int static_waves = rp.mode == MIROGPU_RENDER_DIFFUSE_BOUNCE ? 2 : (whitted_secondary ? -1 : 1);
    const size_t cap = items0 * (any_refractive && whitted_secondary ? (size_t)h->queue_mult : 1);
    const bool use_pm = rp.use_photon_maps && (h->pm[0].stored > 0 || h->pm[1].stored > 0);

    RenderScratch& sc = h->scratch;
    if (rp.mode == MIROGPU_RENDER_DIFFUSE_BOUNCE && !shadows && !use_pm && rp.max_depth >= 1 && !h->ds.textured && !getenv("MIROGPU_GENERAL_WAVEFRONT")) {
        // ---- two fused waves (see k_bounce_wave0) ------------------------------------------------------------------------
        const uint32_t fb = (uint32_t)std::min(rp.spp, 16);
        const size_t fitems0 = npix * fb;
        if (fitems0 >= (1ull << 31)) { err = "frame too large"; return MIROGPU_ERR_UNSUPPORTED; }
        // A batch of up to 16 samples is rendered as two halves on two streams (the caller's and a side stream of the handle): each
        // half is its own chain eye rays -> trace -> wave 0 -> trace -> wave 1 with its own ray / hit / parent-material buffers, so
        // one half's DRAM-bound wave kernels and the ragged end of its persistent trace launches overlap the other half's
        // issue-bound traversal.  Every sample owns its plane, and the planes are folded into the frame in sample order after both
        // halves are in -- the same additions in the same order as a single-stream batch: bit-identical frames.
        static const int two_streams = getenv("MIROGPU_RENDER_STREAMS") ? atoi(getenv("MIROGPU_RENDER_STREAMS")) : 2;
        const bool split = two_streams >= 2 && fb >= 2;
        const uint32_t fbA = split ? (fb + 1) / 2 : fb, fbB = fb - fbA;
        RT(sc.ensure(0, npix * fbA * sizeof(mirogpu_ray))); RT(sc.ensure(2, npix * fbA * sizeof(mirogpu_hit)));
        RT(sc.ensure(3, npix * fbA * 4)); RT(sc.ensure(12, fitems0 * 12));
        if (split) {
            RT(sc.ensure(16, npix * fbB * sizeof(mirogpu_ray))); RT(sc.ensure(17, npix * fbB * sizeof(mirogpu_hit))); RT(sc.ensure(18, npix * fbB * 4));
            RT(sc.ensure_side());
        }
        const size_t fbytes = (npix * 12 + 15) / 16 * 16;
        RT(sc.ensure(10, fbytes + 256));
        float* frame = reinterpret_cast<float*>(sc.buf[10]);
        uint32_t* counters = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sc.buf[10]) + fbytes);
        unsigned long long* d_total = reinterpret_cast<unsigned long long*>(counters + 20);
        float* planes = reinterpret_cast<float*>(sc.buf[12]);
        WaveParams wp;
        wp.ds = h->ds; wp.mats = h->d_materials; wp.lights = h->d_lights; wp.nlights = h->nlights;
        wp.mode = rp.mode; wp.shadows = 0; wp.max_depth = rp.max_depth; wp.use_pm = 0;
        for (int k = 0; k < 3; ++k) wp.bg[k] = rp.bg_color[k];
        wp.seed = rp.seed; wp.sample = 0; wp.cap = (uint32_t)fitems0;
        wp.width = rp.width; wp.first_row = first_row; wp.row_stride = rp.row_stride; wp.npix = (uint32_t)npix;
        CameraBasis cb;
        camera_basis(cam, rp.width, rp.height, cb);
        uint64_t launches = 0, host_rays = 0;
        RT(cudaMemsetAsync(frame, 0, fbytes + 256, st));
        // items per thread of the two wave kernels / camera direction re-read instead of recomputed (measured: profiles/r02_e2e_waves.json)
        // Measured on the bench frame (profiles/r02_e2e_waves.jsonl): 4 items per thread (144 registers, 18 % occupancy) 9.40 ms per
        // frame, 2 items 8.61, 1 item 8.61; with the direction re-read 9.25 / 8.56 / 8.78 -> 2 items, direction re-read.
        static const int bw_items = getenv("MIROGPU_BW_ITEMS") ? atoi(getenv("MIROGPU_BW_ITEMS")) : 2;
        static const int bw_readd = getenv("MIROGPU_BW_READD") ? atoi(getenv("MIROGPU_BW_READD")) : 1;
        const int BW = bw_items == 1 ? 1 : bw_items == 2 ? 2 : 4;
        // one half: samples [sb0, sb0 + nbh) of the frame into planes [plane0, plane0 + nbh), on stream s with its own buffers
        auto half = [&](cudaStream_t s, uint32_t sb0, uint32_t nbh, uint32_t plane0, mirogpu_ray* rays, mirogpu_hit* hits, uint32_t* pmat) -> int {
            const size_t items = npix * nbh;
            WaveParams w = wp;
            w.sample_base = sb0;
            float* pl = planes + (size_t)plane0 * npix * 3;
            k_gen_primary<<<dim3((unsigned)((npix + 255) / 256), nbh), 256, 0, s>>>(cb, rp.width, rp.height, first_row, rp.row_stride, nrows, rp.jitter, rp.seed, sb0, nbh, rays);
            RT(dispatch_trace(h, rays, items, hits, MIROGPU_CLOSEST_HIT | MIROGPU_HINT_COHERENT, s));
            const unsigned grid = (unsigned)((items + MIRO_BW_THREADS * BW - 1) / (MIRO_BW_THREADS * BW));
#define MIRO_W0(K, R) k_bounce_wave0<K, R><<<grid, MIRO_BW_THREADS, 0, s>>>(w, cb, rp.height, nrows, rp.jitter, (uint32_t)items, hits, rays, pmat, pl, counters + MIRO_LIVE_SLOT0)
            if (bw_readd) { if (BW == 1) MIRO_W0(1, true); else if (BW == 2) MIRO_W0(2, true); else MIRO_W0(4, true); }
            else { if (BW == 1) MIRO_W0(1, false); else if (BW == 2) MIRO_W0(2, false); else MIRO_W0(4, false); }
#undef MIRO_W0
            RT(dispatch_trace(h, rays, items, hits, MIROGPU_CLOSEST_HIT, s));
            if (BW == 1) k_bounce_wave1<1><<<grid, MIRO_BW_THREADS, 0, s>>>(w, (uint32_t)items, hits, rays, pmat, pl);
            else if (BW == 2) k_bounce_wave1<2><<<grid, MIRO_BW_THREADS, 0, s>>>(w, (uint32_t)items, hits, rays, pmat, pl);
            else k_bounce_wave1<4><<<grid, MIRO_BW_THREADS, 0, s>>>(w, (uint32_t)items, hits, rays, pmat, pl);
            launches += 5;
            host_rays += items;
            return MIROGPU_OK;
        };
        for (uint32_t s0 = 0; s0 < (uint32_t)rp.spp; s0 += fb) {
            const uint32_t nb = std::min<uint32_t>(fb, (uint32_t)rp.spp - s0);
            const uint32_t nA = (split && nb >= 2) ? (nb + 1) / 2 : nb, nB = nb - nA;
            if (nB) {
                // the side stream starts after everything queued so far on st (frame clear, the previous batch's fold of the planes)
                RT(cudaEventRecord(sc.ev_fork, st));
                RT(cudaStreamWaitEvent(sc.side, sc.ev_fork, 0));
            }
            int rc = half(st, s0, nA, 0, reinterpret_cast<mirogpu_ray*>(sc.buf[0]), reinterpret_cast<mirogpu_hit*>(sc.buf[2]), reinterpret_cast<uint32_t*>(sc.buf[3]));
            if (rc != MIROGPU_OK) return rc;
            if (nB) {
                rc = half(sc.side, s0 + nA, nB, nA, reinterpret_cast<mirogpu_ray*>(sc.buf[16]), reinterpret_cast<mirogpu_hit*>(sc.buf[17]), reinterpret_cast<uint32_t*>(sc.buf[18]));
                if (rc != MIROGPU_OK) return rc;
                RT(cudaEventRecord(sc.ev_join, sc.side));
                RT(cudaStreamWaitEvent(st, sc.ev_join, 0));
            }
            k_fold_planes<<<(unsigned)((npix * 3 + 255) / 256), 256, 0, st>>>(planes, nb, npix * 3, frame);
            launches += 1;
        }
        k_sum_wave_counters<<<1, 1, 0, st>>>(counters, d_total, 1);   // live bounce rays of all batches (the slots are never reset within a frame)
        float* gmax = reinterpret_cast<float*>(counters + 18);
        const float ninf = -INFINITY;
        RT(cudaMemcpyAsync(gmax, &ninf, 4, cudaMemcpyHostToDevice, st));
        const size_t nvals = npix * 3;
        k_resolve_frame<<<(unsigned)((nvals + 255) / 256), 256, 0, st>>>(frame, rp.width, first_row, rp.row_stride, nrows, 1.0f / (float)rp.spp, rp.spp, d_rgb, gmax);
        launches += 2;
        if (rp.tonemap || d_rgb8) {
            k_tonemap<<<(unsigned)((nvals + 255) / 256), 256, 0, st>>>(d_rgb, d_rgb8, rp.width, first_row, rp.row_stride, nrows, gmax);
            launches++;
        }
        RT(cudaGetLastError());
        RT(cudaMemcpyAsync(h->h_stats, d_total, 8, cudaMemcpyDeviceToHost, st));
        h->last_rays = host_rays; h->last_launches = launches;
        h->stats_batches = 1; h->stats_mult = 1;
        return MIROGPU_OK;
    }
    // 0,1: queue A/B rays  2: hits  3,4: pix A/B  5,6: weight A/B  7: shadow rays  8: shadow hits  9: shadow cd+ch
    // 10: frame accumulator + counters  11: gather  12: per-sample planes
    RT(sc.ensure(0, cap * sizeof(mirogpu_ray))); RT(sc.ensure(1, cap * sizeof(mirogpu_ray)));
    RT(sc.ensure(2, cap * sizeof(mirogpu_hit)));
    RT(sc.ensure(3, cap * 4)); RT(sc.ensure(4, cap * 4));
    RT(sc.ensure(5, cap * 16)); RT(sc.ensure(6, cap * 16));
    if (shadows) {
        RT(sc.ensure(7, cap * nl * sizeof(mirogpu_ray))); RT(sc.ensure(8, cap * nl * sizeof(mirogpu_hit))); RT(sc.ensure(9, cap * nl * 32));
    }
    const size_t frame_bytes = (npix * 12 + 15) / 16 * 16;
    RT(sc.ensure(10, frame_bytes + 256));
    RT(sc.ensure(12, items0 * 12));
    if (use_pm) RT(sc.ensure(11, cap * (12 + 12 + 16 + 12 + 12)));
    float* frame = reinterpret_cast<float*>(sc.buf[10]);
    uint32_t* counters = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sc.buf[10]) + frame_bytes);   // [0..16] wave sizes, [17] dropped, [18] gmax
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(counters + 20);                     // secondary items of the whole frame
    float* planes = reinterpret_cast<float*>(sc.buf[12]);
    Queue q[2];
    for (int k = 0; k < 2; ++k) {
        q[k].rays = reinterpret_cast<mirogpu_ray*>(sc.buf[0 + k]); q[k].hits = reinterpret_cast<mirogpu_hit*>(sc.buf[2]);
        q[k].pix = reinterpret_cast<uint32_t*>(sc.buf[3 + k]); q[k].weight = reinterpret_cast<float4*>(sc.buf[5 + k]);
    }
    mirogpu_ray* srays = reinterpret_cast<mirogpu_ray*>(sc.buf[7]);
    mirogpu_hit* shits = reinterpret_cast<mirogpu_hit*>(sc.buf[8]);
    float4* scd = reinterpret_cast<float4*>(sc.buf[9]);
    float4* sch = scd ? scd + cap * nl : nullptr;
    float *gpos = nullptr, *gnrm = nullptr, *girr0 = nullptr, *girr1 = nullptr; float4* gw = nullptr;
    if (use_pm) {
        char* b = reinterpret_cast<char*>(sc.buf[11]);
        gw = reinterpret_cast<float4*>(b); b += cap * 16;
        gpos = reinterpret_cast<float*>(b); b += cap * 12; gnrm = reinterpret_cast<float*>(b); b += cap * 12;
        girr0 = reinterpret_cast<float*>(b); b += cap * 12; girr1 = reinterpret_cast<float*>(b);
    }

    WaveParams wp;
    wp.ds = h->ds; wp.mats = h->d_materials; wp.lights = h->d_lights; wp.nlights = h->nlights;
    wp.mode = rp.mode; wp.shadows = shadows ? 1 : 0; wp.max_depth = rp.max_depth; wp.use_pm = use_pm ? 1 : 0;
    for (int k = 0; k < 3; ++k) wp.bg[k] = rp.bg_color[k];
    wp.seed = rp.seed; wp.sample = 0; wp.cap = (uint32_t)cap;
    wp.width = rp.width; wp.first_row = first_row; wp.row_stride = rp.row_stride; wp.npix = (uint32_t)npix;
    CameraBasis cb;
    camera_basis(cam, rp.width, rp.height, cb);

    uint64_t launches = 0;
    uint32_t* hs = h->h_stats;          // pinned: [0] rays traced (filled at the end), [1] launches
    uint64_t host_rays = 0;             // exact when counts are known on the host, else completed from device counters
    RT(cudaMemsetAsync(frame, 0, frame_bytes + 256, st));
    for (uint32_t s0 = 0; s0 < (uint32_t)rp.spp; s0 += batch) {
        const uint32_t nb = std::min<uint32_t>(batch, (uint32_t)rp.spp - s0);
        const size_t items = npix * nb;
        wp.sample_base = s0;
        RT(cudaMemsetAsync(planes, 0, items * 12, st));
        RT(cudaMemsetAsync(counters, 0, 18 * 4, st));
        if (rp.mode == MIROGPU_RENDER_DIFFUSE_BOUNCE) RT(cudaMemsetAsync(counters + MIRO_LIVE_SLOT0, 0, 32 * 4, st));
        int cur = 0;
        k_render_primary<<<dim3((unsigned)((npix + 255) / 256), nb), 256, 0, st>>>(cb, rp.width, rp.height, first_row, rp.row_stride, nrows, rp.jitter, rp.seed,
                                                                           s0, nb, rp.max_depth, q[cur]);
        launches++;
        size_t bound = items;
        for (int wave = 0; wave < MIRO_MAX_WAVES && wave <= rp.max_depth; ++wave) {
            const uint32_t* d_n = wave == 0 ? nullptr : counters + wave;
            RT(dispatch_trace(h, q[cur].rays, bound, q[cur].hits, MIROGPU_CLOSEST_HIT | (wave == 0 ? MIROGPU_HINT_COHERENT : 0), st, d_n, 1));
            k_shade<<<(unsigned)((bound + MIRO_SHADE_THREADS * MIRO_SHADE_ITEMS - 1) / (MIRO_SHADE_THREADS * MIRO_SHADE_ITEMS)), MIRO_SHADE_THREADS, 0, st>>>(wp, q[cur], (uint32_t)bound, d_n, q[cur ^ 1], counters + wave + 1, counters + 17,
                                                                      srays, scd, sch, planes, gpos, gnrm, gw);
            launches += 2;
            if (shadows) {
                const size_t ns = bound * nl;
                RT(dispatch_trace(h, srays, ns, shits, (any_refractive ? MIROGPU_CLOSEST_HIT : MIROGPU_ANY_HIT) | (wave == 0 ? MIROGPU_HINT_COHERENT : 0), st, d_n, nl));
                k_shadow_accumulate<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(wp, srays, shits, scd, sch, ns, d_n, nl, planes);
                launches += 2;
            }
            if (use_pm) {
                if (h->pm[0].stored > 0) { RT(photon_gather_launch(h->pm[0], gpos, gnrm, bound, 1e10f, 500, girr0, st, gw, d_n)); launches++; }
                if (h->pm[1].stored > 0) { RT(photon_gather_launch(h->pm[1], gpos, gnrm, bound, 1e10f, 500, girr1, st, gw, d_n)); launches++; }
                k_gather_accumulate<<<(unsigned)((bound + 255) / 256), 256, 0, st>>>(gw, q[cur].pix, h->pm[0].stored > 0 ? girr0 : nullptr,
                                                                                      h->pm[1].stored > 0 ? girr1 : nullptr, (uint32_t)bound, d_n,
                                                                                      q[cur].weight, (uint32_t)npix, planes);
                launches++;
            }
            cur ^= 1;
            if (static_waves > 0) {
                if (wave + 1 >= static_waves) break;
                bound = std::min(cap, bound);            // the next wave holds at most one item per item of this one
            } else {
                uint32_t next_n = 0;
                RT(cudaMemcpyAsync(&next_n, counters + wave + 1, 4, cudaMemcpyDeviceToHost, st));
                RT(cudaStreamSynchronize(st));
                if (next_n == 0) break;
                if (next_n > cap) { *overflow = true; return MIROGPU_OK; }   // children were dropped: this frame is void
                bound = std::min<size_t>(cap, next_n);
            }
        }
        k_fold_planes<<<(unsigned)((npix * 3 + 255) / 256), 256, 0, st>>>(planes, nb, npix * 3, frame);
        launches++;
        k_sum_wave_counters<<<1, 1, 0, st>>>(counters, d_total, rp.mode == MIROGPU_RENDER_DIFFUSE_BOUNCE ? 1 : 0);
        launches++;
        host_rays += items * (shadows ? 1 + nl : 1);
    }
    float* gmax = reinterpret_cast<float*>(counters + 18);
    const float ninf = -INFINITY;
    RT(cudaMemcpyAsync(gmax, &ninf, 4, cudaMemcpyHostToDevice, st));
    const size_t nvals = npix * 3;
    k_resolve_frame<<<(unsigned)((nvals + 255) / 256), 256, 0, st>>>(frame, rp.width, first_row, rp.row_stride, nrows, 1.0f / (float)rp.spp, rp.spp, d_rgb, gmax);
    launches++;
    if (rp.tonemap || d_rgb8) {
        k_tonemap<<<(unsigned)((nvals + 255) / 256), 256, 0, st>>>(d_rgb, d_rgb8, rp.width, first_row, rp.row_stride, nrows, gmax);
        launches++;
    }
    RT(cudaGetLastError());
    // rays: the primary wave is known here; the secondary waves were counted on the device and land in pinned
    // host memory (valid once the caller has synchronised the stream) for mirogpu_last_call_stats
    RT(cudaMemcpyAsync(hs, d_total, 8, cudaMemcpyDeviceToHost, st));
    h->last_rays = host_rays; h->last_launches = launches;
    h->stats_batches = 1;
    h->stats_mult = shadows ? 1 + nl : 1;
    return MIROGPU_OK;
#undef RT
}

// d_rgb: full-frame float buffer (rows of this shard are written).  d_rgb8 (may be NULL): full-frame 8-bit tone-mapped output.
// Queues of the general wavefront start at 4x the primary items for refractive scenes and double (kept for the handle's later
// frames) whenever a wave overflows them, so no child ray is ever dropped silently.
int render_device(mirogpu_scene* h, const mirogpu_camera& cam, const mirogpu_render_params& rp, float* d_rgb, unsigned char* d_rgb8,
                  cudaStream_t st, std::string& err)
{
    for (;;) {
        bool overflow = false;
        const int rc = render_device_once(h, cam, rp, d_rgb, d_rgb8, st, err, &overflow);
        if (rc != MIROGPU_OK || !overflow) return rc;
        if (h->queue_mult >= 256) { err = "path queues overflow even at 256x the primary rays"; return MIROGPU_ERR_UNSUPPORTED; }
        h->queue_mult *= 2;
    }
}

// ---- one handle, several devices: Scene::raytraceImage sharded by image rows -----------------------------------
// Replica k of N renders the rows row_begin + k + j N of the call's range on its own device and stream (the scene is replicated,
// random numbers are keyed by frame pixel, so the frame equals the one-device frame bit for bit).  The tone map's one frame-wide
// number -- the largest non-NaN value (Scene.cpp:157-164) -- is reduced per device, combined on the host (N floats), and
// each device maps its own rows.  Host framebuffer: every device copies its rows straight into the caller's buffer (N DMA
// engines in parallel).  Device framebuffer: the rows are gathered into the first device's buffer by peer copies (NVLink).
mirogpu_scene* replica_of(mirogpu_scene* h, size_t k) { return k == 0 ? h : h->replicas[k - 1]; }

cudaError_t replica_stream(mirogpu_scene* r)
{
    if (r->mstream) return cudaSuccess;
    cudaError_t e = cudaStreamCreateWithFlags(&r->mstream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->mevent, cudaEventDisableTiming);
    return e;
}

// Phase 1 on every replica: float rows into its own full-frame buffer (scratch 13), the maximum of its rows into pinned memory.
int render_multi_phase1(mirogpu_scene* h, const mirogpu_camera& cam, const mirogpu_render_params& rp, std::string& err)
{
    const size_t N = 1 + h->replicas.size();
    const size_t npx = (size_t)rp.width * rp.height;
    const bool dynamic_waves = rp.mode == MIROGPU_RENDER_WHITTED && (h->any_specular || h->any_refractive);
    std::vector<int> rcs(N, MIROGPU_OK);
    std::vector<std::string> errs(N);
    auto work = [&](size_t k) {
        mirogpu_scene* r = replica_of(h, k);
        cudaError_t e = cudaSetDevice(r->device);
        if (e == cudaSuccess) e = replica_stream(r);
        if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(r->mtx); e = r->scratch.ensure(13, npx * 12); if (e == cudaSuccess) e = r->scratch.ensure(14, npx * 3); if (e == cudaSuccess) e = r->scratch.ensure(15, 64); }
        if (e != cudaSuccess) { errs[k] = std::string("multi-device render setup: ") + cudaGetErrorString(e); rcs[k] = MIROGPU_ERR_CUDA; return; }
        mirogpu_render_params sub = rp;
        sub.row_stride = (int)N; sub.row_phase = (int)k; sub.tonemap = 0;
        float* d_rgb = reinterpret_cast<float*>(r->scratch.buf[13]);
        rcs[k] = render_device(r, cam, sub, d_rgb, nullptr, r->mstream, errs[k]);
        if (rcs[k] != MIROGPU_OK) return;
        // this replica's maximum (render_device's own reduction lives in its scratch; redo it over the rows into buf[15])
        float* d_max = reinterpret_cast<float*>(r->scratch.buf[15]);
        const float ninf = -INFINITY;
        const int first_row = rp.row_begin + (int)k;
        const int nrows = first_row < rp.row_end ? (rp.row_end - first_row + (int)N - 1) / (int)N : 0;
        e = cudaMemcpyAsync(d_max, &ninf, 4, cudaMemcpyHostToDevice, r->mstream);
        const size_t nvals = (size_t)nrows * rp.width * 3;
        if (e == cudaSuccess && nvals) k_frame_max_rows<<<(unsigned)((nvals + 255) / 256), 256, 0, r->mstream>>>(d_rgb, rp.width, first_row, (int)N, nrows, d_max);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->h_stats + 32, d_max, 4, cudaMemcpyDeviceToHost, r->mstream);
        if (e != cudaSuccess) { errs[k] = std::string("multi-device render: ") + cudaGetErrorString(e); rcs[k] = MIROGPU_ERR_CUDA; }
    };
    if (dynamic_waves) {   // wave sizes are read back per wave: one host thread per device so the round trips overlap
        std::vector<std::thread> th;
        for (size_t k = 0; k < N; ++k) th.emplace_back(work, k);
        for (auto& t : th) t.join();
    } else {
        for (size_t k = 0; k < N; ++k) work(k);   // everything is asynchronous: issue device after device
    }
    for (size_t k = 0; k < N; ++k) if (rcs[k] != MIROGPU_OK) { err = errs[k]; return rcs[k]; }
    return MIROGPU_OK;
}

int render_multi(mirogpu_scene* h, const mirogpu_camera& cam, const mirogpu_render_params& rp, float* rgb_out, unsigned char* rgb8_out,
                 float* d_rgb_primary, cudaStream_t st_primary, std::string& err)
{
    if (rp.width <= 0 || rp.height <= 0 || rp.row_begin < 0 || rp.row_end > rp.height || rp.row_begin > rp.row_end) { err = "bad render parameters"; return MIROGPU_ERR_INVALID_ARG; }
    const size_t N = 1 + h->replicas.size();
    int rc = render_multi_phase1(h, cam, rp, err);
    if (rc != MIROGPU_OK) return rc;
    cudaError_t e = cudaSuccess;
    float gmax = -INFINITY;
    uint64_t rays = 0, launches = 0;
    for (size_t k = 0; k < N && e == cudaSuccess; ++k) {
        mirogpu_scene* r = replica_of(h, k);
        cudaSetDevice(r->device);
        e = cudaStreamSynchronize(r->mstream);
        float m; memcpy(&m, r->h_stats + 32, 4);
        if (m > gmax) gmax = m;
        uint64_t a = 0, b = 0;
        mirogpu_last_call_stats(r, &a, &b);
        rays += a; launches += b;
    }
    const size_t rowf = (size_t)rp.width * 12, rowb = (size_t)rp.width * 3;
    for (size_t k = 0; k < N && e == cudaSuccess; ++k) {
        mirogpu_scene* r = replica_of(h, k);
        cudaSetDevice(r->device);
        const int first_row = rp.row_begin + (int)k;
        const int nrows = first_row < rp.row_end ? (rp.row_end - first_row + (int)N - 1) / (int)N : 0;
        if (nrows == 0) continue;
        float* d_rgb = reinterpret_cast<float*>(r->scratch.buf[13]);
        unsigned char* d_u8 = reinterpret_cast<unsigned char*>(r->scratch.buf[14]);
        float* d_max = reinterpret_cast<float*>(r->scratch.buf[15]);
        const size_t nvals = (size_t)nrows * rp.width * 3;
        if (rgb8_out || rp.tonemap) {
            e = cudaMemcpyAsync(d_max, &gmax, 4, cudaMemcpyHostToDevice, r->mstream);   // pageable 4-byte source: staged at once by the driver
            if (e != cudaSuccess) break;
            k_tonemap<<<(unsigned)((nvals + 255) / 256), 256, 0, r->mstream>>>(d_rgb, rgb8_out ? d_u8 : nullptr, rp.width, first_row, (int)N, nrows, d_max);
            launches++;
        }
        if (rgb8_out) e = cudaMemcpy2DAsync(rgb8_out + (size_t)first_row * rowb, rowb * N, d_u8 + (size_t)first_row * rowb, rowb * N, rowb, nrows, cudaMemcpyDeviceToHost, r->mstream);
        else if (rgb_out) e = cudaMemcpy2DAsync(reinterpret_cast<char*>(rgb_out) + (size_t)first_row * rowf, rowf * N, reinterpret_cast<char*>(d_rgb) + (size_t)first_row * rowf, rowf * N, rowf, nrows, cudaMemcpyDeviceToHost, r->mstream);
        else if (d_rgb_primary) {   // gather on the first device: peer copy over NVLink, ordered into the caller's stream by an event
            e = cudaMemcpy2DAsync(reinterpret_cast<char*>(d_rgb_primary) + (size_t)first_row * rowf, rowf * N, reinterpret_cast<char*>(d_rgb) + (size_t)first_row * rowf, rowf * N, rowf, nrows, cudaMemcpyDefault, r->mstream);
            if (e == cudaSuccess) e = cudaEventRecord(r->mevent, r->mstream);
        }
    }
    if (e == cudaSuccess && !rgb8_out && !rgb_out && d_rgb_primary) {
        cudaSetDevice(h->device);
        for (size_t k = 0; k < N && e == cudaSuccess; ++k) e = cudaStreamWaitEvent(st_primary, replica_of(h, k)->mevent, 0);
    } else {
        for (size_t k = 0; k < N; ++k) {
            mirogpu_scene* r = replica_of(h, k);
            cudaSetDevice(r->device);
            const cudaError_t e2 = cudaStreamSynchronize(r->mstream);
            if (e == cudaSuccess) e = e2;
        }
    }
    cudaSetDevice(h->device);
    if (e != cudaSuccess) { err = std::string("multi-device framebuffer gather: ") + cudaGetErrorString(e); return MIROGPU_ERR_CUDA; }
    {
        std::lock_guard<std::mutex> lk(h->mtx);
        h->last_rays = rays; h->last_launches = launches; h->stats_batches = 0;
    }
    return MIROGPU_OK;
}

int render_host(mirogpu_scene* h, const mirogpu_camera& cam, const mirogpu_render_params& rp, float* rgb_out, unsigned char* rgb8_out,
                std::string& err)
{
    if (rp.width <= 0 || rp.height <= 0) { err = "bad render parameters"; return MIROGPU_ERR_INVALID_ARG; }
    // a multi-device handle shards the rows itself -- unless the caller already asks for a shard (row_stride > 1)
    if (!h->replicas.empty() && rp.row_stride == 1) return render_multi(h, cam, rp, rgb_out, rgb8_out, nullptr, nullptr, err);
    const size_t npx = (size_t)rp.width * rp.height;
    cudaError_t e;
    {
        std::lock_guard<std::mutex> lk(h->mtx);
        e = h->scratch.ensure(13, npx * 12);
        if (e == cudaSuccess && rgb8_out) e = h->scratch.ensure(14, npx * 3);
    }
    if (e != cudaSuccess) { err = std::string("cudaMalloc framebuffer: ") + cudaGetErrorString(e); return MIROGPU_ERR_OOM; }
    float* d_rgb = reinterpret_cast<float*>(h->scratch.buf[13]);
    unsigned char* d_rgb8 = rgb8_out ? reinterpret_cast<unsigned char*>(h->scratch.buf[14]) : nullptr;
    cudaStream_t st = cudaStreamPerThread;
    int rc = render_device(h, cam, rp, d_rgb, d_rgb8, st, err);
    if (rc != MIROGPU_OK) return rc;
    // only this call's rows are defined on the device; copy them (one piece for a full frame)
    const bool full = rp.row_stride == 1 && rp.row_begin == 0 && rp.row_end == rp.height;
    const int first_row = rp.row_begin + rp.row_phase;
    const int nrows = first_row < rp.row_end ? (rp.row_end - first_row + rp.row_stride - 1) / rp.row_stride : 0;
    if (rgb8_out) {
        const size_t rowb = (size_t)rp.width * 3;
        if (full) e = cudaMemcpyAsync(rgb8_out, d_rgb8, npx * 3, cudaMemcpyDeviceToHost, st);
        else e = cudaMemcpy2DAsync(rgb8_out + (size_t)first_row * rowb, rowb * rp.row_stride, d_rgb8 + (size_t)first_row * rowb, rowb * rp.row_stride, rowb,
                                   nrows, cudaMemcpyDeviceToHost, st);
    } else {
        const size_t rowb = (size_t)rp.width * 12;
        if (full) e = cudaMemcpyAsync(rgb_out, d_rgb, npx * 12, cudaMemcpyDeviceToHost, st);
        else e = cudaMemcpy2DAsync(reinterpret_cast<char*>(rgb_out) + (size_t)first_row * rowb, rowb * rp.row_stride,
                                   reinterpret_cast<char*>(d_rgb) + (size_t)first_row * rowb, rowb * rp.row_stride, rowb, nrows, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { err = std::string("framebuffer copy: ") + cudaGetErrorString(e); return MIROGPU_ERR_CUDA; }
    return MIROGPU_OK;
}

}  // namespace
#endif
