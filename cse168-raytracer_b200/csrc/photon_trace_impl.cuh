// photon_trace_impl.cuh -- photon emission and tracing on the device: Scene::tracePhoton (reference Scene.cpp:526-641)
// for the emissions of Scene::tracePhotons / traceCausticPhotons (Scene.cpp:351-472).
//
// One thread walks one emitted photon to its end: emission from the DirectionalAreaLight's disc (DirectionalAreaLight.h:
// 20-35, sampleDisc Utility.h:82-95), then up to TRACE_DEPTH_PHOTONS + 1 = 6 segments of closest-hit traversal (the same
// per-ray core the trace kernels run) with the reference's roulette between diffuse / mirror / transmission / absorption,
// Fresnel coin, Ray::random, Ray::reflect and Ray::refract.  Each diffuse choice after the first hit records
// {power, position, incoming direction}; an emission records at most five.  Records land in per-emission slots, so the
// host consumes them in emission order and reproduces the reference's sequential stop rule ("emit while fewer than the
// target are stored", Scene.cpp:370-377) independently of how the walks were scheduled.
//
// Random numbers: the reference's rand() is unseeded and raced over by its OpenMP workers; here every frand() is
// Philox4x32-10 keyed by (seed, emission index, segment, purpose) -- the oracle draws the identical stream, so walks are
// compared emission by emission (tests/test_gpu_photon_trace.py) and the maps statistically against the real reference.
#ifndef MIROGPU_PHOTON_TRACE_IMPL_CUH
#define MIROGPU_PHOTON_TRACE_IMPL_CUH

namespace mirogpu {

#define MIRO_PHOTON_MAX_RECORDS 5   /* stores happen at depth 2..6 */
#define MIRO_TRACE_DEPTH_PHOTONS 5  /* Miro.h:14 */

struct PhotonEmitter {
    float pos[3], normal[3], t1[3], t2[3];
    float radius;
    float power[3];       // color * wattage * PI r^2 (/ 10 for the caustic pass)
    int caustic;
    uint32_t seed;
    unsigned long long first;
    uint32_t count;
};

__device__ __forceinline__ void uniform4(uint32_t seed, unsigned long long index, uint32_t sample, uint32_t purpose, float u[4])
{
    uint32_t r[4];
    philox4x32_10((uint32_t)index, sample, purpose, (uint32_t)(index >> 32), seed, 0x4D49524Fu, r);
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = (float)(r[k] >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float average3(const float v[3]) { return xdiv(xadd(xadd(v[0], v[1]), v[2]), 3.0f); }   // Vector3.h:226-229

template <int LAYOUT, bool NT = false>
__global__ void __launch_bounds__(128) k_photon_trace(DeviceScene s, const mirogpu_material* __restrict__ mats, PhotonEmitter em,
                                                      unsigned char* __restrict__ counts, float* __restrict__ records)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= em.count) return;
    const unsigned long long e = em.first + i;
    float* rec = records + (size_t)i * (9 * MIRO_PHOTON_MAX_RECORDS);
    float power[3] = {em.power[0], em.power[1], em.power[2]};
    float dir[3] = {em.normal[0], em.normal[1], em.normal[2]};
    float pos[3];
    {
        float x = 0.f, y = 0.f;
        for (uint32_t k = 0; k < 64; ++k) {
            float u[4];
            uniform4(em.seed, e, k, 4, u);
            x = xmul(xsub(xmul(2.0f, u[0]), 1.0f), em.radius);
            y = xmul(xsub(xmul(2.0f, u[1]), 1.0f), em.radius);
            if (!(xadd(xmul(x, x), xmul(y, y)) > xmul(em.radius, em.radius))) break;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) pos[k] = xadd(em.pos[k], xadd(xmul(em.t1[k], x), xmul(em.t2[k], y)));
    }
    int depth = 0, n = 0;
    for (;;) {
        if (depth > MIRO_TRACE_DEPTH_PHOTONS) break;
        mirogpu_ray r;
        r.ox = xadd(pos[0], xmul(dir[0], MIRO_EPS)); r.oy = xadd(pos[1], xmul(dir[1], MIRO_EPS)); r.oz = xadd(pos[2], xmul(dir[2], MIRO_EPS));
        r.dx = dir[0]; r.dy = dir[1]; r.dz = dir[2]; r.tmin = 0.0f; r.tmax = MIROGPU_TMAX;
        ++depth;
        BestHit best;
        trace_one<LAYOUT, false, false, NT>(s, r, best, nullptr);
        if (best.prim == MIROGPU_MISS) break;
        mirogpu_hit h; h.t = best.t; h.prim_id = best.prim; h.beta = best.beta; h.gamma = best.gamma;
        const ShadeRecord prec = load_shade_record(s, h.prim_id);
        SurfacePoint sp;
        float dc[3];
        if (s.textured) {
            const TexturedPoint tp = resolve_hit_textured(prec, h, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, mats, s.uvs);
            sp = tp.sp; dc[0] = tp.dc[0]; dc[1] = tp.dc[1]; dc[2] = tp.dc[2];
        } else sp = resolve_hit(prec, h, r);
        const mirogpu_material m = mats[sp.material];
        if (!s.textured) { dc[0] = m.kd[0]; dc[1] = m.kd[1]; dc[2] = m.kd[2]; }   // the diffuse colour looked up at the hit (Scene.cpp:546-551)
        float u[4];
        uniform4(em.seed, e, (uint32_t)depth, 3, u);
        const float rnd = u[0];
        const float p0 = average3(dc), p1 = xadd(p0, average3(m.ks)), p2 = xadd(p1, average3(m.kt));
        if (rnd > p2) break;                                   // absorbed
        if (rnd < p0) {
            if (depth > 1) {                                   // only indirect light is stored (Scene.cpp:567)
                float* q = rec + 9 * n;
                q[0] = power[0]; q[1] = power[1]; q[2] = power[2]; q[3] = sp.P[0]; q[4] = sp.P[1]; q[5] = sp.P[2];
                q[6] = dir[0]; q[7] = dir[1]; q[8] = dir[2];
                ++n;
            } else if (em.caustic) break;                      // caustic photons leave a specular surface first (:595)
            const float phi = asinf(sqrtf(u[2]));              // Ray::random, Ray.h:124-140
            const float theta = xmul(xmul(2.0f, MIRO_PI), u[3]);
            float d[3];
            align_hemisphere(sp.N, theta, phi, d);
            const float inv = xdiv(1.0f, p0);                  // Vector3 / float multiplies by the reciprocal (Vector3.h:125-129)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                pos[k] = xadd(sp.P[k], xmul(d[k], MIRO_EPS));
                dir[k] = d[k];
                power[k] = xmul(xmul(dc[k], power[k]), inv);
            }
        } else {
            const bool mirror = rnd < p1;
            if (!mirror && !(rnd < p2)) break;
            if (!em.caustic && depth == 1) break;              // only caustics count a specular first bounce (:604, :617)
            // Ray::reflect (Ray.h:160-163)
            const float dn = xdot(sp.N[0], sp.N[1], sp.N[2], r.dx, r.dy, r.dz);
            float dr[3] = {xsub(r.dx, xmul(xmul(2.0f, dn), sp.N[0])), xsub(r.dy, xmul(xmul(2.0f, dn), sp.N[1])), xsub(r.dz, xmul(xmul(2.0f, dn), sp.N[2]))};
            {
                const float inv = xdiv(1.0f, xsqrt(xdot(dr[0], dr[1], dr[2], dr[0], dr[1], dr[2])));
                dr[0] = xmul(dr[0], inv); dr[1] = xmul(dr[1], inv); dr[2] = xmul(dr[2], inv);
            }
            bool reflect = mirror;
            float dt[3] = {0.f, 0.f, 0.f};
            if (!mirror) {
                // Fresnel coefficient (Ray.h:168-200), then Snell with the TIR fallback (Ray.h:202-243)
                float n1, n2, nn[3];
                if (dn < 0.f) { n1 = 1.0f; n2 = m.refract_index; nn[0] = sp.N[0]; nn[1] = sp.N[1]; nn[2] = sp.N[2]; }
                else { n1 = m.refract_index; n2 = 1.0f; nn[0] = -sp.N[0]; nn[1] = -sp.N[1]; nn[2] = -sp.N[2]; }
                const float cosT = xdot(-r.dx, -r.dy, -r.dz, nn[0], nn[1], nn[2]);
                const float sinT = sinf(acosf(cosT));
                const float q0 = xmul(xdiv(n1, n2), sinT);
                const float ps = xmul(q0, q0);
                float Rs;
                if (ps > 1.f) Rs = 1.f;
                else {
                    const float sq = xsqrt(xsub(1.f, ps));
                    const float q = xdiv(xsub(xmul(n1, cosT), sq), xadd(xmul(n1, cosT), sq));
                    Rs = xmul(q, q);
                }
                if (u[1] < Rs) reflect = true;
                else {
                    const float ddn = xdot(r.dx, r.dy, r.dz, nn[0], nn[1], nn[2]);
                    const float energy = xsub(1.f, xdiv(xmul(xmul(n1, n1), xsub(1.f, xmul(ddn, ddn))), xmul(n2, n2)));
                    if (energy < 0.f) reflect = true;
                    else {
                        const float se = xsqrt(energy);
                        const float rdv[3] = {r.dx, r.dy, r.dz};
                        const float inv2 = xdiv(1.0f, n2);      // Vector3 / float (Vector3.h:125-129)
#pragma unroll
                        for (int k = 0; k < 3; ++k) dt[k] = xsub(xmul(xmul(n1, xsub(rdv[k], xmul(nn[k], ddn))), inv2), xmul(nn[k], se));
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { pos[k] = sp.P[k]; dir[k] = reflect ? dr[k] : dt[k]; }
        }
    }
    counts[i] = (unsigned char)n;
}

}  // namespace mirogpu
#endif
