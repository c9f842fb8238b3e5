// tools/simt_sim.cu -- design-space tool (host only, never shipped): replays the BVH2 traversal of the bench's
// bounce rays on the CPU, records every ray's schedule (node steps / triangle tests per while-while round) and
// evaluates how many warp issue slots different SIMT scheduling policies would spend on them.  The traversal
// kernels are issue-bound at ~7 of 32 lanes active (profiles/ncu_summary.json), so lane utilisation is the
// figure to optimise; this lets a policy be judged without a GPU.
//
//   nvcc -O3 -std=c++17 -Xcompiler -fopenmp tools/simt_sim.cu build/bvh_build.o -o /tmp/simt_sim -lgomp
//   /tmp/simt_sim /tmp/bunny20_verts.bin [rows]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../cse168-raytracer_b200/csrc/bvh_build.h"
#include "../cse168-raytracer_b200/csrc/rng.cuh"
#include "../cse168-raytracer_b200/csrc/traverse.cuh"

using namespace mirogpu;

struct Round { uint8_t nn, nt; };
struct Trace { std::vector<Round> r; };

static int g_maxleaf = 4;

// trace_bvh2 with schedule recording
static void trace_record(const float4* nodes, const float4* tris, const mirogpu_ray& r, BestHit& best, Trace* tr)
{
    const float idx = safe_rcp(r.dx), idy = safe_rcp(r.dy), idz = safe_rcp(r.dz);
    const float oodx = r.ox * idx, oody = r.oy * idy, oodz = r.oz * idz;
    int32_t stack[MIRO_STACK];
    int sp = 0;
    int32_t node = 0;
    best.t = r.tmax; best.prim = MIROGPU_MISS; best.beta = 0.f; best.gamma = 0.f;
    if (!(r.tmax >= r.tmin)) return;
    for (;;) {
        int nn = 0;
        bool done = false;
        while (node >= 0) {
            ++nn;
            const float4 n0 = nodes[4 * node + 0], n1 = nodes[4 * node + 1], nz = nodes[4 * node + 2], lk = nodes[4 * node + 3];
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
                if (sp == 0) { done = true; break; }
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
        if (done) { tr->r.push_back({(uint8_t)std::min(nn, 255), 0}); return; }
        const uint32_t ref = (uint32_t)~node;
        const uint32_t first = ref >> 3, count = (ref & 7u) + 1u;
        for (uint32_t i = 0; i < count; ++i) tri_test(tris[4 * (first + i)], tris[4 * (first + i) + 1], tris[4 * (first + i) + 2], tris[4 * (first + i) + 3], r, best);
        tr->r.push_back({(uint8_t)std::min(nn, 255), (uint8_t)count});
        if (sp == 0) return;
        node = stack[--sp];
    }
}

struct Cost { double Cn = 50, Ct = 115, Cround = 6, Crefill = 40; };
static Cost C;

struct Result { double slots = 0, useful = 0; };

static double ray_work(const Trace& t)
{
    double w = 0;
    for (auto& r : t.r) w += r.nn * C.Cn + r.nt * C.Ct;
    return w;
}

// Policy A: fixed 32-ray packets, while-while
static Result sim_packet(const std::vector<Trace>& T, const std::vector<uint32_t>& order)
{
    Result R;
    const size_t n = order.size();
    for (size_t b = 0; b < n; b += 32) {
        const size_t e = std::min(n, b + 32);
        size_t maxr = 0;
        for (size_t i = b; i < e; ++i) maxr = std::max(maxr, T[order[i]].r.size());
        for (size_t k = 0; k < maxr; ++k) {
            int mn = 0, mt = 0;
            for (size_t i = b; i < e; ++i) {
                const auto& tr = T[order[i]].r;
                if (k < tr.size()) { mn = std::max(mn, (int)tr[k].nn); mt = std::max(mt, (int)tr[k].nt); }
            }
            R.slots += mn * C.Cn + mt * C.Ct + C.Cround;
        }
        for (size_t i = b; i < e; ++i) R.useful += ray_work(T[order[i]]);
        R.slots += C.Crefill;
    }
    return R;
}

// Policy B: persistent warps with replacement every `rounds` rounds (lanes pull the next rays of `order`);
// nwarps warps share the queue round-robin (approximates the global pool; pool granularity `pool` rays)
static Result sim_replace(const std::vector<Trace>& T, const std::vector<uint32_t>& order, int rounds, int pool, int min_idle = 1)
{
    Result R;
    const size_t n = order.size();
    size_t next = 0;
    // process pool by pool (a warp owns a pool of `pool` consecutive rays, as the kernel does)
    while (next < n) {
        const size_t pe = std::min(n, next + (size_t)pool);
        size_t pn = next;
        next = pe;
        const Trace* cur[32];
        size_t pos[32];
        for (int l = 0; l < 32; ++l) cur[l] = nullptr;
        for (;;) {
            // refill
            int idle = 0;
            for (int l = 0; l < 32; ++l) if (!cur[l]) ++idle;
            if (idle >= min_idle || idle == 32) {
                for (int pass = 0; pass < 2; ++pass)
                    for (int l = 0; l < 32; ++l)
                        if (!cur[l] && pn < pe) {
                            const Trace* t = &T[order[pn++]];
                            if (!t->r.empty()) { cur[l] = t; pos[l] = 0; R.useful += ray_work(*t); }
                        }
                R.slots += C.Crefill;
            }
            bool any = false;
            for (int l = 0; l < 32; ++l) any |= cur[l] != nullptr;
            if (!any) { if (pn >= pe) break; else continue; }
            for (int k = 0; k < rounds; ++k) {
                int mn = 0, mt = 0;
                bool live = false;
                for (int l = 0; l < 32; ++l)
                    if (cur[l]) {
                        const Round& r = cur[l]->r[pos[l]];
                        mn = std::max(mn, (int)r.nn); mt = std::max(mt, (int)r.nt);
                        if (++pos[l] >= cur[l]->r.size()) cur[l] = nullptr; else live = true;
                    }
                R.slots += mn * C.Cn + mt * C.Ct + C.Cround;
                if (!live) break;
            }
        }
    }
    return R;
}

// Policy E: single-step loop.  Each iteration every lane does ONE step, node or triangle; the warp pays for each
// kind that at least one lane needs.  Replacement every `period` iterations.
static Result sim_stepwise(const std::vector<Trace>& T, const std::vector<uint32_t>& order, int period, int pool, double Cn, double Ct)
{
    Result R;
    const size_t n = order.size();
    size_t next = 0;
    while (next < n) {
        const size_t pe = std::min(n, next + (size_t)pool);
        size_t pn = next;
        next = pe;
        const Trace* cur[32];
        size_t pos[32];
        int remn[32], remt[32];
        for (int l = 0; l < 32; ++l) cur[l] = nullptr;
        for (;;) {
            for (int pass = 0; pass < 2; ++pass)
                for (int l = 0; l < 32; ++l)
                    if (!cur[l] && pn < pe) {
                        const Trace* t = &T[order[pn++]];
                        if (!t->r.empty()) { cur[l] = t; pos[l] = 0; remn[l] = t->r[0].nn; remt[l] = t->r[0].nt; R.useful += ray_work(*t); }
                    }
            R.slots += C.Crefill;
            bool any = false;
            for (int l = 0; l < 32; ++l) any |= cur[l] != nullptr;
            if (!any) { if (pn >= pe) break; else continue; }
            for (int k = 0; k < period; ++k) {
                bool an = false, at = false, live = false;
                for (int l = 0; l < 32; ++l)
                    if (cur[l]) {
                        if (remn[l] > 0) { an = true; --remn[l]; }
                        else if (remt[l] > 0) { at = true; --remt[l]; }
                        if (remn[l] == 0 && remt[l] == 0) {
                            if (++pos[l] >= cur[l]->r.size()) cur[l] = nullptr;
                            else { remn[l] = cur[l]->r[pos[l]].nn; remt[l] = cur[l]->r[pos[l]].nt; }
                        }
                        if (cur[l]) live = true;
                    }
                R.slots += (an ? Cn : 0) + (at ? Ct : 0) + 4;
                if (!live) break;
            }
        }
    }
    return R;
}


// Policy H: hybrid.  One loop; an iteration is EITHER a node step (taken while at least `nmin` lanes want one, or
// no lane waits at a leaf) OR a leaf phase (every waiting lane tests its whole leaf).  Replacement every `period`
// iterations, only if at least `min_idle` lanes are idle.
static Result sim_hybrid(const std::vector<Trace>& T, const std::vector<uint32_t>& order, int nmin, int period, int pool, int min_idle, bool leaf_one_tri = false)
{
    Result R;
    const size_t n = order.size();
    size_t next = 0;
    while (next < n) {
        const size_t pe = std::min(n, next + (size_t)pool);
        size_t pn = next;
        next = pe;
        const Trace* cur[32];
        size_t pos[32];
        int remn[32], remt[32];
        for (int l = 0; l < 32; ++l) cur[l] = nullptr;
        for (;;) {
            int idle = 0;
            for (int l = 0; l < 32; ++l) if (!cur[l]) ++idle;
            if (idle >= min_idle || idle == 32) {
                for (int pass = 0; pass < 2; ++pass)
                    for (int l = 0; l < 32; ++l)
                        if (!cur[l] && pn < pe) {
                            const Trace* t = &T[order[pn++]];
                            if (!t->r.empty()) { cur[l] = t; pos[l] = 0; remn[l] = t->r[0].nn; remt[l] = t->r[0].nt; R.useful += ray_work(*t); }
                        }
                R.slots += C.Crefill;
            }
            bool any = false;
            for (int l = 0; l < 32; ++l) any |= cur[l] != nullptr;
            if (!any) { if (pn >= pe) break; else continue; }
            for (int k = 0; k < period; ++k) {
                int cn = 0, ct = 0;
                for (int l = 0; l < 32; ++l) if (cur[l]) { if (remn[l] > 0) ++cn; else ++ct; }
                if (cn + ct == 0) break;
                auto advance = [&](int l) {
                    if (++pos[l] >= cur[l]->r.size()) cur[l] = nullptr;
                    else { remn[l] = cur[l]->r[pos[l]].nn; remt[l] = cur[l]->r[pos[l]].nt; }
                };
                if (cn > 0 && (cn >= nmin || ct == 0)) {
                    for (int l = 0; l < 32; ++l)
                        if (cur[l] && remn[l] > 0) { if (--remn[l] == 0 && remt[l] == 0) advance(l); }
                    R.slots += C.Cn + 4;
                } else {
                    int mt = 0;
                    for (int l = 0; l < 32; ++l)
                        if (cur[l] && remn[l] == 0) {
                            if (leaf_one_tri) { mt = 1; if (--remt[l] == 0) advance(l); }
                            else { mt = std::max(mt, remt[l]); advance(l); }
                        }
                    R.slots += mt * C.Ct + 6;
                }
            }
        }
    }
    return R;
}

static void report(const char* name, const Result& r, double base)
{
    printf("%-58s slots/ray-unit %10.3e  efficiency %5.1f%%  speedup vs A %.2fx\n", name, r.slots, 100.0 * r.useful / (32.0 * r.slots), base / r.slots);
}

int main(int argc, char** argv)
{
    const char* path = argc > 1 ? argv[1] : "/tmp/bunny20_verts.bin";
    const int rows = argc > 2 ? atoi(argv[2]) : 64;
    if (getenv("MAXLEAF")) g_maxleaf = atoi(getenv("MAXLEAF"));
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); return 1; }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<float> verts(bytes / 4);
    if (fread(verts.data(), 1, bytes, f) != (size_t)bytes) return 1;
    fclose(f);
    const uint32_t ntris = (uint32_t)(verts.size() / 9);
    BinaryBvh bin = build_binary_sah(verts.data(), ntris, g_maxleaf, 32);
    FlatBvh flat;
    flatten_bvh2(bin, flat);
    std::vector<TriRecord> trec;
    make_tri_records(verts.data(), flat.order, trec);
    const float4* nodes = reinterpret_cast<const float4*>(flat.nodes2.data());
    const float4* tris = reinterpret_cast<const float4*>(trec.data());
    printf("tris %u nodes %zu\n", ntris, flat.nodes2.size());

    // camera of the bench: eye (0,5,15) look-at 0, up y, fov 45, 1920x1080; a band of `rows` rows in the middle
    const int W = 1920, H = 1080;
    const float eye[3] = {0, 5, 15};
    float w[3] = {0 - 0, 5 - 0, 15 - 0};
    float wl = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (float& x : w) x /= wl;
    float up[3] = {0, 1, 0};
    float u[3] = {up[1] * w[2] - up[2] * w[1], up[2] * w[0] - up[0] * w[2], up[0] * w[1] - up[1] * w[0]};
    float ul = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (float& x : u) x /= ul;
    float v[3] = {w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0]};
    const float top = tanf(45.f * 3.14159265f / 360.f), right = top * W / H;
    const int row0 = getenv("ROW0") ? atoi(getenv("ROW0")) : (H - rows) / 2;
    const size_t n = (size_t)rows * W;
    std::vector<mirogpu_ray> prim(n), bounce(n);
    std::vector<Trace> TP(n), TB(n);
    std::vector<uint8_t> live(n);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; ++i) {
        const int x = (int)(i % W), y = row0 + (int)(i / W);
        float j1, j2;
        uniform2(168, (uint32_t)(y * W + x), 0, 0, j1, j2);
        const float U = -right + 2 * right * (x + j1) / W, V = -top + 2 * top * (y + j2) / H;
        float d[3];
        for (int k = 0; k < 3; ++k) d[k] = u[k] * U + v[k] * V - w[k];
        const float dl = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        mirogpu_ray r = {eye[0], eye[1], eye[2], 0.f, d[0] / dl, d[1] / dl, d[2] / dl, 1e12f};
        prim[i] = r;
        BestHit best;
        trace_record(nodes, tris, r, best, &TP[i]);
        mirogpu_ray b = {0, 0, 0, 0, 0, 0, 1, -1};
        if (best.prim != MIROGPU_MISS) {
            live[i] = 1;
            const float* q = &verts[9 * (size_t)best.prim];
            float e1[3] = {q[3] - q[0], q[4] - q[1], q[5] - q[2]}, e2[3] = {q[6] - q[0], q[7] - q[1], q[8] - q[2]};
            float N[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            float nl = sqrtf(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
            if (N[0] * r.dx + N[1] * r.dy + N[2] * r.dz > 0) nl = -nl;
            for (float& c : N) c /= nl;
            float P[3] = {r.ox + best.t * r.dx, r.oy + best.t * r.dy, r.oz + best.t * r.dz};
            float u1, u2;
            uniform2(168, (uint32_t)(y * W + x), 0, 1, u1, u2);
            const float phi = asinf(sqrtf(u1)), th = 6.2831853f * u2;
            float t1[3] = {-N[1], N[0], 0};
            if (t1[0] * t1[0] + t1[1] * t1[1] < 1e-6f) { t1[0] = N[2]; t1[1] = 0; t1[2] = -N[0]; }
            float c[3] = {t1[1] * N[2] - t1[2] * N[1], t1[2] * N[0] - t1[0] * N[2], t1[0] * N[1] - t1[1] * N[0]};
            float a[3];
            for (int k = 0; k < 3; ++k) a[k] = t1[k] * sinf(phi) * cosf(th) + c[k] * sinf(phi) * sinf(th) + N[k] * cosf(phi);
            const float al = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            for (float& k : a) k /= al;
            b = {P[0] + 1e-4f * a[0], P[1] + 1e-4f * a[1], P[2] + 1e-4f * a[2], 0.f, a[0], a[1], a[2], 1e12f};
        }
        bounce[i] = b;
        BestHit bb;
        trace_record(nodes, tris, b, bb, &TB[i]);
    }
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<Trace>& T = pass ? TB : TP;
        const std::vector<mirogpu_ray>& rays = pass ? bounce : prim;
        double nn = 0, nt = 0, nr = 0, nlive = 0;
        std::vector<double> works;
        for (auto& t : T) { if (!t.r.empty()) { ++nlive; works.push_back(ray_work(t)); } for (auto& r : t.r) { nn += r.nn; nt += r.nt; nr += 1; } }
        std::sort(works.begin(), works.end());
        printf("\n==== %s rays: %zu (%.0f live); per live ray: node steps %.2f, tri tests %.2f, rounds %.2f; work p50 %.0f p90 %.0f p99 %.0f max %.0f mean %.0f\n",
               pass ? "bounce" : "primary", n, nlive, nn / nlive, nt / nlive, nr / nlive, works[works.size() / 2], works[works.size() * 9 / 10],
               works[works.size() * 99 / 100], works.back(), std::accumulate(works.begin(), works.end(), 0.0) / works.size());
        std::vector<uint32_t> ident(n);
        std::iota(ident.begin(), ident.end(), 0u);
        const Result A = sim_packet(T, ident);
        report("A  packets of 32, while-while (today)", A, A.slots);
        for (int rounds : {1, 2, 4, 8}) {
            char nm[128];
            snprintf(nm, sizeof nm, "B  replacement every %d rounds, pool 256", rounds);
            report(nm, sim_replace(T, ident, rounds, 256), A.slots);
        }
        report("B  replacement every 2 rounds, pool 1024", sim_replace(T, ident, 2, 1024), A.slots);
        report("B  replacement every 2 rounds, pool 256, refill when >=8 idle", sim_replace(T, ident, 2, 256, 8), A.slots);
        for (int period : {1, 4, 16}) {
            char nm[128];
            snprintf(nm, sizeof nm, "E  single-step loop, replacement every %d steps", period);
            report(nm, sim_stepwise(T, ident, period, 256, C.Cn, C.Ct), A.slots);
        }

        for (int nmin : {4, 8, 12, 16, 20, 24})
            for (int period : {4, 16}) {
                char nm[160];
                snprintf(nm, sizeof nm, "H  hybrid nmin %d, replacement every %d iters (idle>=4)", nmin, period);
                report(nm, sim_hybrid(T, ident, nmin, period, 256, 4), A.slots);
            }
        report("H  hybrid nmin 12, every 8, pool 1024", sim_hybrid(T, ident, 12, 8, 1024, 4), A.slots);
        report("H  hybrid nmin 12, every 8, one tri per leaf iteration", sim_hybrid(T, ident, 12, 8, 256, 4, true), A.slots);
        // ---- sorted orders ----
        auto octant = [&](const mirogpu_ray& r) { return (r.dx < 0 ? 1u : 0u) | (r.dy < 0 ? 2u : 0u) | (r.dz < 0 ? 4u : 0u); };
        auto dirbin = [&](const mirogpu_ray& r, int res) {   // cube-map face + res x res cells
            const float ax = fabsf(r.dx), ay = fabsf(r.dy), az = fabsf(r.dz);
            int face; float a, b2, m;
            if (ax >= ay && ax >= az) { face = r.dx < 0; a = r.dy; b2 = r.dz; m = ax; }
            else if (ay >= az) { face = 2 + (r.dy < 0); a = r.dx; b2 = r.dz; m = ay; }
            else { face = 4 + (r.dz < 0); a = r.dx; b2 = r.dy; m = az; }
            const int ia = std::min(res - 1, (int)((a / m * 0.5f + 0.5f) * res)), ib = std::min(res - 1, (int)((b2 / m * 0.5f + 0.5f) * res));
            return (uint32_t)((face * res + ia) * res + ib);
        };
        for (int tile : {4096}) {
            for (int mode = 0; mode < 3; ++mode) {
                std::vector<uint32_t> ord(n);
                std::iota(ord.begin(), ord.end(), 0u);
                for (size_t b = 0; b < n; b += tile) {
                    const size_t e = std::min(n, b + (size_t)tile);
                    std::stable_sort(ord.begin() + b, ord.begin() + e, [&](uint32_t x, uint32_t y) {
                        const bool dx_ = rays[x].tmax < 0, dy_ = rays[y].tmax < 0;   // dead rays last
                        if (dx_ != dy_) return dy_;
                        const uint32_t kx = mode == 0 ? octant(rays[x]) : dirbin(rays[x], mode == 1 ? 2 : 4);
                        const uint32_t ky = mode == 0 ? octant(rays[y]) : dirbin(rays[y], mode == 1 ? 2 : 4);
                        return kx < ky;
                    });
                }
                char nm[160];
                snprintf(nm, sizeof nm, "D  tile %5d sorted by %s, packets", tile, mode == 0 ? "octant" : mode == 1 ? "cube 2x2" : "cube 4x4");
                report(nm, sim_packet(T, ord), A.slots);
                snprintf(nm, sizeof nm, "D+B tile %5d sorted by %s, replacement/2", tile, mode == 0 ? "octant" : mode == 1 ? "cube 2x2" : "cube 4x4");
                report(nm, sim_replace(T, ord, 2, 256), A.slots);
            }
        }
        // ideal: sort by work (upper bound for any reordering that balances termination)
        {
            std::vector<uint32_t> ord(n);
            std::iota(ord.begin(), ord.end(), 0u);
            std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return T[x].r.size() < T[y].r.size(); });
            report("X  oracle order: sorted by #rounds, packets", sim_packet(T, ord), A.slots);
        }
    }
    return 0;
}
