// texture.cuh -- the reference's procedural textures behind TexturedPhong, for device-side material evaluation:
//   Phong::shade's diffuse colour lookup (Phong.cpp:50-55), Scene::tracePhoton's (Scene.cpp:546-551) and the bump height of
//   Scene::trace (Scene.cpp:232-266), for CheckerBoardTexture, StoneTexture, PetalTexture, StemTexture, LeafTexture and
//   FlowerCenterTexture (Texture.h:108-277, Texture.cpp:358-510) over Perlin's improved noise (lib/include/Perlin.h) and
//   Worley's cellular basis (lib/src/Worley.cpp, 2-D, order 3).
// Every function is __host__ __device__: the CPU tier evaluates the same code against the oracle (tests/cpu_emu).
// Arithmetic follows the reference expression by expression -- binary32 where its operands are float, double where a double
// literal promotes the expression -- with never-contracted operations, so host and device agree and differences to the
// reference come from libm only (powf / expf / acosf / sinf last-ulp).
#ifndef MIROGPU_TEXTURE_CUH
#define MIROGPU_TEXTURE_CUH

#include <cmath>
#include <cstdint>
#include "../../include/mirogpu.h"
#include "traverse.cuh"
#include "texture_tables.cuh"

#define MIRO_PI_TEX 3.1415926535897932384626433832795028841972f /* PI, Miro.h:10 */

namespace mirogpu {

__constant__ unsigned char c_perlin_perm[256] = {MIRO_PERLIN_PERM};
__constant__ unsigned char c_worley_poisson[256] = {MIRO_WORLEY_POISSON};
static const unsigned char h_perlin_perm[256] = {MIRO_PERLIN_PERM};
static const unsigned char h_worley_poisson[256] = {MIRO_WORLEY_POISSON};

MIRO_HD int perlin_p(int i)   // the reference's p[512] is the permutation twice
{
#ifdef __CUDA_ARCH__
    return c_perlin_perm[i & 255];
#else
    return h_perlin_perm[i & 255];
#endif
}
MIRO_HD int worley_count(uint32_t msb)
{
#ifdef __CUDA_ARCH__
    return c_worley_poisson[msb];
#else
    return h_worley_poisson[msb];
#endif
}

// double arithmetic the compiler must not contract either
MIRO_HD double dmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
MIRO_HD double dadd(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// ---- PerlinNoise::noise (lib/include/Perlin.h:16-56) ----------------------------------------------------------------
MIRO_HD float perlin_fade(float t) { return xmul(xmul(xmul(t, t), t), xadd(xmul(t, xsub(xmul(t, 6.0f), 15.0f)), 10.0f)); }
MIRO_HD float perlin_lerp(float t, float a, float b) { return xadd(a, xmul(t, xsub(b, a))); }
MIRO_HD float perlin_grad(int hash, float x, float y, float z)
{
    const int h = hash & 15;
    const float u = h < 8 ? x : y;
    const float v = h < 4 ? y : ((h == 12 || h == 14) ? x : z);
    return xadd((h & 1) == 0 ? u : -u, (h & 2) == 0 ? v : -v);
}
// int(floor(x)) as the reference's x86-64 build converts: out-of-range values (the 25-octave turbulence of PetalTexture reaches
// coordinates of 1e12) give the "integer indefinite" 0x80000000, where CUDA's conversion would saturate to INT_MAX
MIRO_HD int x86_int(float f) { return (f >= 2147483648.0f || f < -2147483648.0f || f != f) ? (int)0x80000000u : (int)f; }

MIRO_HD float perlin_noise(float x, float y, float z)
{
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int X = x86_int(fx) & 255, Y = x86_int(fy) & 255, Z = x86_int(fz) & 255;
    x = xsub(x, fx); y = xsub(y, fy); z = xsub(z, fz);
    const float u = perlin_fade(x), v = perlin_fade(y), w = perlin_fade(z);
    const int A = perlin_p(X) + Y, AA = perlin_p(A) + Z, AB = perlin_p(A + 1) + Z;
    const int B = perlin_p(X + 1) + Y, BA = perlin_p(B) + Z, BB = perlin_p(B + 1) + Z;
    const float x1 = xsub(x, 1.0f), y1 = xsub(y, 1.0f), z1 = xsub(z, 1.0f);
    return perlin_lerp(w,
                       perlin_lerp(v, perlin_lerp(u, perlin_grad(perlin_p(AA), x, y, z), perlin_grad(perlin_p(BA), x1, y, z)),
                                   perlin_lerp(u, perlin_grad(perlin_p(AB), x, y1, z), perlin_grad(perlin_p(BB), x1, y1, z))),
                       perlin_lerp(v, perlin_lerp(u, perlin_grad(perlin_p(AA + 1), x, y, z1), perlin_grad(perlin_p(BA + 1), x1, y, z1)),
                                   perlin_lerp(u, perlin_grad(perlin_p(AB + 1), x, y1, z1), perlin_grad(perlin_p(BB + 1), x1, y1, z1))));
}

// generateNoise (Texture.h:20-38): turbulence normalised by the summed amplitudes
MIRO_HD float turbulence(float x, float y, float z, float frequency, float frequency_increase, float amplitude_falloff, int iterations)
{
    float amplitude = 1.0f, value = 0.0f, max_val = 0.0f;
    for (int i = 0; i < iterations; ++i) {
        value = xadd(value, xmul(amplitude, perlin_noise(xmul(x, frequency), xmul(y, frequency), xmul(z, frequency))));
        max_val = xadd(max_val, amplitude);
        frequency = xmul(frequency, frequency_increase);
        amplitude = xmul(amplitude, amplitude_falloff);
    }
    return xdiv(value, max_val);
}

// ---- WorleyNoise::noise2D, order 3 (lib/src/Worley.cpp:95-173, 366-436) ----------------------------------------------
struct Worley3 { float F[3]; uint32_t id[3]; };

MIRO_HD void worley_cell(int xi, int yi, const float at[2], Worley3& w)
{
    // the cube's LCG stream: seed from the cube's coordinates (32-bit wrap-around), count from its top byte
    uint32_t seed = 702395077u * (uint32_t)xi + 915488749u * (uint32_t)yi;
    const int count = worley_count(seed >> 24);
    seed = 1402024253u * seed + 586950981u;
    for (int j = 0; j < count; ++j) {
        const uint32_t this_id = seed;
        seed = 1402024253u * seed + 586950981u;
        const float fx = (float)dmul(dadd((double)seed, 0.5), 1.0 / 4294967296.0);
        seed = 1402024253u * seed + 586950981u;
        const float fy = (float)dmul(dadd((double)seed, 0.5), 1.0 / 4294967296.0);
        seed = 1402024253u * seed + 586950981u;
        const float dx = xsub(xadd((float)xi, fx), at[0]), dy = xsub(xadd((float)yi, fy), at[1]);
        const float d2 = xadd(xmul(dx, dx), xmul(dy, dy));
        if (d2 < w.F[2]) {      // insertion into the sorted three
            int index = 3;
            while (index > 0 && d2 < w.F[index - 1]) --index;
            for (int i = 1; i >= index; --i) { w.F[i + 1] = w.F[i]; w.id[i + 1] = w.id[i]; }
            w.F[index] = d2; w.id[index] = this_id;
        }
    }
}

MIRO_HD Worley3 worley2(float u, float v)
{
    const double density = 0.398150;
    Worley3 w;
    for (int i = 0; i < 3; ++i) { w.F[i] = 999999.9f; w.id[i] = 0u; }
    const float at[2] = {(float)dmul(density, (double)u), (float)dmul(density, (double)v)};
    const int ix = (int)floorf(at[0]), iy = (int)floorf(at[1]);
    worley_cell(ix, iy, at, w);
    float x2 = xsub(at[0], (float)ix), y2 = xsub(at[1], (float)iy);
    const float mx2 = (float)dmul(dadd(1.0, -(double)x2), dadd(1.0, -(double)x2)), my2 = (float)dmul(dadd(1.0, -(double)y2), dadd(1.0, -(double)y2));
    x2 = xmul(x2, x2); y2 = xmul(y2, y2);
    // neighbours in the reference's order, each only if it can still hold a closer feature point
    if (x2 < w.F[2]) worley_cell(ix - 1, iy, at, w);
    if (y2 < w.F[2]) worley_cell(ix, iy - 1, at, w);
    if (mx2 < w.F[2]) worley_cell(ix + 1, iy, at, w);
    if (my2 < w.F[2]) worley_cell(ix, iy + 1, at, w);
    if (xadd(x2, y2) < w.F[2]) worley_cell(ix - 1, iy - 1, at, w);
    if (xadd(mx2, my2) < w.F[2]) worley_cell(ix + 1, iy + 1, at, w);
    if (xadd(x2, my2) < w.F[2]) worley_cell(ix - 1, iy + 1, at, w);
    if (xadd(mx2, y2) < w.F[2]) worley_cell(ix + 1, iy - 1, at, w);
    for (int i = 0; i < 3; ++i) w.F[i] = (float)dmul((double)xsqrt(w.F[i]), 1.0 / density);
    return w;
}

// ---- texture lookups ------------------------------------------------------------------------------------------------
// tex[] parameter layout per kind: see mirogpu.h (mirogpu_material).
MIRO_HD void tex_checker(const float* tp, float cu, float cv, float out[3])   // Texture.h:125-132
{
    const float scale = tp[6];
    float u = fabsf(xmul(scale, cu)), v = fabsf(xmul(scale, cv));
    if (cu < 0.f) u = xadd(u, scale);
    if (cv < 0.f) v = xadd(v, scale);
    const bool first = ((int)u + (int)v) % 2 == 0;
    out[0] = first ? tp[0] : tp[3]; out[1] = first ? tp[1] : tp[4]; out[2] = first ? tp[2] : tp[5];
}

MIRO_HD void tex_stone(const float* tp, float cu, float cv, float out[3])     // Texture.cpp:396-440
{
    const float u = xmul(cu, tp[0]), v = xmul(cv, tp[0]);
    const Worley3 w = worley2(u, v);
    const float f1f0 = (float)dmul((double)xsub(1.0f, powf(xsub(w.F[1], w.F[0]), 0.8f)), 1.5);
    float base = fminf(fmaxf(xsub(powf(xadd(xsub(w.F[2], w.F[1]), w.F[0]), 0.1f), f1f0), 0.f), 0.5f);
    const float cell10 = (float)(w.id[0] % 10u), cell5 = (float)(w.id[0] % 5u);
    base = (float)dmul((double)base, dadd((double)xdiv(cell10, 20.0f), 0.5));
    const float turb = turbulence(u, v, 0.f, 3.0f, 2.0f, 0.8f, 5);
    base = fmaxf(0.0f, base);
    base = (float)dadd((double)base, dmul(0.8, (double)fabsf(turb)));
    if ((double)f1f0 > 1.1) {
        const float edges = fminf(xsub(xmul(f1f0, f1f0), 1.0f), 0.75f);
        out[0] = out[1] = out[2] = (float)dadd((double)edges, dmul(0.25, (double)fabsf(turb)));
    } else {
        out[0] = xadd(base, xdiv(cell10, 10.0f));
        out[1] = xadd(base, xmul(xdiv(cell10, 10.0f), 0.5f));
        out[2] = xadd(base, xmul(xdiv(cell5, 5.0f), 0.25f));
    }
}

MIRO_HD float tex_stone_bump(const float* tp, float cu, float cv)            // Texture.cpp:358-393
{
    const float u = xmul(cu, tp[0]), v = xmul(cv, tp[0]);
    const float height_factor = 0.3f;
    const Worley3 w = worley2(u, v);
    const float d10 = xsub(w.F[1], w.F[0]);
    float f1f0 = (float)dmul((double)xsub(1.0f, powf(d10, 0.8f)), 1.5);
    f1f0 = xmul(f1f0, -1.0f);
    const float height = (float)(1.0 / dadd(1.0, exp(dmul(-20.0, dadd((double)d10, -0.3)))));
    if ((double)f1f0 > -1.1) {
        const float cellturb = (float)dadd((double)xdiv(turbulence(u, v, 0.f, 0.5f, 2.0f, 0.5f, (int)(w.id[0] % 3u) + 5), 5.0f), 0.5);
        return xadd(xmul(0.8f, cellturb), xmul(height_factor, height));
    }
    const float turb = (float)dadd((double)xdiv(turbulence(u, v, 0.f, 1.0f, 2.0f, 0.5f, 3), 10.0f), 0.5);
    return xadd(xmul(1.0f, turb), xmul(height_factor, height));
}

MIRO_HD float tex_turb_ramp(float turb) { return fminf(xmul(powf(xdiv(turb, 0.1f), 0.85f), 1.5f), 1.0f); }

MIRO_HD void tex_petal(const float* tp, const float P[3], float out[3])       // Texture.cpp:447-510
{
    const float base_hl[3] = {0.2f, 0.f, 0.8f}, tip_hl[3] = {0.8f, 0.5f, 1.f}, base_dep[3] = {0.2f, 0.0f, 0.5f}, tip_dep[3] = {0.3f, 0.15f, 0.75f};
    const float base_col[3] = {0.1f, 0.0f, 0.6f}, tip_col[3] = {0.6f, 0.3f, 1.0f};
    float pos[3] = {xsub(P[0], tp[0]), xsub(P[1], tp[1]), xsub(P[2], tp[2])};
    const float radius = xsqrt(xdot(pos[0], pos[1], pos[2], pos[0], pos[1], pos[2]));
    const float dist = xdiv(radius, tp[3]), omd = xsub(1.0f, dist);
    float col[3], hl[3], dep[3];
    for (int k = 0; k < 3; ++k) {
        col[k] = xadd(xmul(omd, base_col[k]), xmul(dist, tip_col[k]));
        hl[k] = xadd(xmul(omd, base_hl[k]), xmul(dist, tip_hl[k]));
        dep[k] = xadd(xmul(omd, base_dep[k]), xmul(dist, tip_dep[k]));
    }
    const float inv = xdiv(1.0f, radius);                   // position.normalize() normalises in place (Vector3.h:205-208)
    pos[0] = xmul(pos[0], inv); pos[1] = xmul(pos[1], inv); pos[2] = xmul(pos[2], inv);
    const float phi = acosf(-xdot(0.f, 1.f, 0.f, pos[0], pos[1], pos[2]));
    const float v = xdiv(phi, MIRO_PI_TEX);
    const float theta = xdiv(acosf(xdot(pos[0], pos[1], pos[2], 1.f, 0.f, 0.f)), xmul(2.0f, MIRO_PI_TEX));
    // dot(cross(north, equator), position) with north = (0,1,0), equator = (1,0,0): cross = (0, 0, -1)
    const float side = xdot(xsub(xmul(1.f, 0.f), xmul(0.f, 0.f)), xsub(xmul(0.f, 1.f), xmul(0.f, 0.f)), xsub(xmul(0.f, 0.f), xmul(1.f, 1.f)), pos[0], pos[1], pos[2]);
    const float u = side > 0.f ? theta : xsub(1.0f, theta);
    const float high = tex_turb_ramp(fabsf(turbulence(u, (float)dmul((double)v, 0.25), 0.f, 4.0f, 2.0f, 0.9f, 10)));
    const float low = tex_turb_ramp(fabsf(turbulence(u, v, 0.f, 4.0f, 3.0f, 0.9f, 25)));
    for (int k = 0; k < 3; ++k)
        out[k] = xadd(xmul(0.5f, xadd(xmul(high, col[k]), xmul(xsub(1.0f, high), hl[k]))), xmul(0.5f, xadd(xmul(low, col[k]), xmul(xsub(1.0f, low), dep[k]))));
}

MIRO_HD void tex_stem_leaf(float scale, float cu, float cv, float out[3])     // Texture.h:193-215, 233-255 (same body)
{
    const float u = xmul(cu, scale), v = xmul(cv, scale);
    const Worley3 w = worley2(u, v);
    const float noise = turbulence(u, v, 0.f, 10.0f, 1.5f, 0.8f, 10);
    const float cells = xsub(w.F[0], w.F[1]);
    out[0] = 0.f; out[2] = 0.f;
    out[1] = (float)dadd(dadd(0.5, dmul(0.5, (double)xadd(noise, 1.0f)) / 2.0), -dmul(0.3, (double)cells));
}

MIRO_HD void tex_flower_center(const float* tp, const float P[3], float out[3])   // Texture.h:266-281
{
    const float d[3] = {xsub(P[0], tp[0]), xsub(P[1], tp[1]), xsub(P[2], tp[2])};
    const float dist = xsqrt(xdot(d[0], d[1], d[2], d[0], d[1], d[2]));
    const float fraction = fmaxf(fminf(powf(xdiv(dist, tp[3]), 30.0f), 1.0f), 0.0f);
    const float omf = xsub(1.0f, fraction);
    out[0] = fminf(xadd(xmul(omf, 0.31f), xmul(fraction, 0.92f)), 1.0f);
    out[1] = fminf(xadd(xmul(omf, 0.18f), xmul(fraction, 0.71f)), 1.0f);
    out[2] = 0.1f;
}

// GetLookupCoordinates() == UV (Texture2D) or UVW (Texture3D), Texture.h:74-84
MIRO_HD bool tex_is_uv(int kind) { return kind == MIROGPU_TEX_CHECKER || kind == MIROGPU_TEX_STONE || kind == MIROGPU_TEX_STEM; }

// Material::diffuse2D / diffuse3D (Phong.h:20-21, Texture.cpp:519-527): the colour Phong::shade and Scene::tracePhoton look up.
// uv: Object::toUVCoordinates(P) of the hit object (used by the UV kinds), P: the hit point (used by the others).
MIRO_HD void material_diffuse_color(const mirogpu_material& m, const float uv[2], const float P[3], float out[3])
{
    switch (m.texture) {
    case MIROGPU_TEX_CHECKER: tex_checker(m.tex, uv[0], uv[1], out); break;
    case MIROGPU_TEX_STONE: tex_stone(m.tex, uv[0], uv[1], out); break;
    case MIROGPU_TEX_STEM: tex_stem_leaf(m.tex[0], uv[0], uv[1], out); break;
    case MIROGPU_TEX_LEAF: tex_stem_leaf(m.tex[0], P[0], P[1], out); break;
    case MIROGPU_TEX_PETAL: tex_petal(m.tex, P, out); break;
    case MIROGPU_TEX_FLOWER_CENTER: tex_flower_center(m.tex, P, out); break;
    default: out[0] = m.kd[0]; out[1] = m.kd[1]; out[2] = m.kd[2]; break;   // Phong::diffuse2D returns m_diffuse
    }
}

MIRO_HD float material_bump_height(const mirogpu_material& m, float u, float v)
{
    return m.texture == MIROGPU_TEX_STONE ? tex_stone_bump(m.tex, u, v) : 0.0f;   // the other textures return 0 (Texture.h, Texture.cpp:442-445)
}

// Scene::trace's bump mapping for UV materials (Scene.cpp:232-262): central differences of the bump height, two tangents from
// the largest normal component, N += dx (N x t1) - dy (N x (N x t1)), normalised.  N: the un-normalised interpolated normal.
MIRO_HD void material_bump_normal(const mirogpu_material& m, float u, float v, float N[3])
{
    const float delta = 0.0001f;
    const float u1 = material_bump_height(m, xsub(u, delta), v), u2 = material_bump_height(m, xadd(u, delta), v);
    const float v1 = material_bump_height(m, u, xsub(v, delta)), v2 = material_bump_height(m, u, xadd(v, delta));
    const float dx = xdiv(xsub(u2, u1), xmul(2.0f, delta)), dy = xdiv(xsub(v2, v1), xmul(2.0f, delta));
    int k = 0;
    if (N[1] > N[0]) k = 1;
    if (N[2] > N[k]) k = 2;
    const float r[3] = {k == 2 ? -N[2] : 0.f, k == 0 ? -N[0] : 0.f, k == 1 ? -N[1] : 0.f};
    auto cross = [](const float a[3], const float b[3], float o[3]) {
        o[0] = xsub(xmul(a[1], b[2]), xmul(a[2], b[1])); o[1] = xsub(xmul(a[2], b[0]), xmul(a[0], b[2])); o[2] = xsub(xmul(a[0], b[1]), xmul(a[1], b[0]));
    };
    float t1[3], a[3], b[3], c[3];
    cross(N, r, t1);
    cross(N, t1, a);          // N x t1
    cross(N, a, b);           // N x (N x t1)
    for (int i = 0; i < 3; ++i) c[i] = xsub(xmul(dx, a[i]), xmul(dy, b[i]));
    for (int i = 0; i < 3; ++i) N[i] = xadd(N[i], c[i]);
    const float inv = xdiv(1.0f, xsqrt(xdot(N[0], N[1], N[2], N[0], N[1], N[2])));
    N[0] = xmul(N[0], inv); N[1] = xmul(N[1], inv); N[2] = xmul(N[2], inv);
}

}  // namespace mirogpu
#endif
