// mirogpu.cu -- the C ABI of include/mirogpu.h: scene construction (host SAH build -> flat layout -> HBM),
// batched intersection, device-side ray generation, render and photon gather entry points.
// No CPU fallback: every compute entry point needs a CUDA device and says so loudly when there is none.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mirogpu.h"
#include "bvh_build.h"
#include "kernels.cuh"
#include "render.cuh"
#include "photon.cuh"

using namespace mirogpu;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess) {                                                                                \
            const int code = (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? MIROGPU_ERR_NO_DEVICE \
                             : (_e == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA);          \
            return fail(code, std::string(#expr) + ": " + cudaGetErrorString(_e));                              \
        }                                                                                                       \
    } while (0)

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// Device allocation released on scope exit (error paths included).
struct DevBuf {
    void* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// Staging of the host-buffer batch API: a stream, device ray / hit buffers that grow on demand, and a small page-locked
// pair for tiny batches (the reference's call pattern is ONE ray per BVH::intersect call: no allocation, no stream
// creation and no pageable-copy staging on that path).  Slots belong to the scene handle; a call borrows one.
#define MIRO_SMALL_BATCH 1024
struct BatchSlot {
    cudaStream_t st = nullptr;
    mirogpu_ray* d_r = nullptr; mirogpu_hit* d_h = nullptr; size_t cap = 0;
    mirogpu_ray* h_r = nullptr; mirogpu_hit* h_h = nullptr;   // page-locked, MIRO_SMALL_BATCH entries, device-visible (UVA)
    mirogpu_ray* pin_r = nullptr; mirogpu_hit* pin_h = nullptr; size_t pin_cap = 0;   // page-locked staging of one chunk, for callers with pageable buffers
    unsigned long long* d_c = nullptr;                          // 4 counters of the instrumented kernel
    bool busy = false;
    cudaError_t init()
    {
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaHostAlloc(&h_r, MIRO_SMALL_BATCH * sizeof(mirogpu_ray), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostAlloc(&h_h, MIRO_SMALL_BATCH * sizeof(mirogpu_hit), cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaMalloc(&d_c, 4 * sizeof(unsigned long long));
        return e;
    }
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        cudaFree(d_r); cudaFree(d_h); d_r = nullptr; d_h = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&d_r, n * sizeof(mirogpu_ray));
        if (e == cudaSuccess) e = cudaMalloc(&d_h, n * sizeof(mirogpu_hit));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    cudaError_t ensure_pinned(size_t n)
    {
        if (n <= pin_cap) return cudaSuccess;
        if (pin_r) cudaFreeHost(pin_r);
        if (pin_h) cudaFreeHost(pin_h);
        pin_r = nullptr; pin_h = nullptr; pin_cap = 0;
        cudaError_t e = cudaHostAlloc(&pin_r, n * sizeof(mirogpu_ray), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaHostAlloc(&pin_h, n * sizeof(mirogpu_hit), cudaHostAllocDefault);
        if (e == cudaSuccess) pin_cap = n;
        return e;
    }
    void release()
    {
        if (pin_r) cudaFreeHost(pin_r);
        if (pin_h) cudaFreeHost(pin_h);
        pin_r = nullptr; pin_h = nullptr; pin_cap = 0;
        cudaFree(d_r); cudaFree(d_h); cudaFree(d_c);
        if (h_r) cudaFreeHost(h_r);
        if (h_h) cudaFreeHost(h_h);
        if (st) cudaStreamDestroy(st);
        d_r = nullptr; d_h = nullptr; d_c = nullptr; h_r = nullptr; h_h = nullptr; st = nullptr; cap = 0;
    }
};

#define MIRO_TICKETS 1024u
struct mirogpu_scene {
    std::vector<std::unique_ptr<BatchSlot>> slots;   // guarded by slot_mtx
    std::mutex slot_mtx;
    int device = 0;
    int layout = MIROGPU_LAYOUT_CWBVH8;
    int variant = -1;   // mirogpu_set_kernel_variant
    int sm_count = 148;
    // hybrid kernel (variant 2) knobs; env MIROGPU_NMIN / _PERIOD / _MINIDLE / _POOL / _PF / _MINB / _NREP override (tuning)
    int hyb_nmin = 16, hyb_period = 4, hyb_min_idle = 8, hyb_pool = 64, hyb_nrep = 2;
    int packets_per_ticket = 1;                            // packet kernel: 32-ray packets per ticket; env MIROGPU_PPT
    size_t node_bytes_dev = 0, tri_bytes_dev = 0;
    int hyb_pf = 0, hyb_minb = 10;                         // step flags (traverse.cuh / k_trace_hybrid), min resident CTAs (10: 48 registers)
    int hyb_pf_inc = 0, hyb_nmin_inc = 16;                 // the same two knobs for batches NOT flagged coherent
    int hyb_short = 0, hyb_stage = 0;                      // shared-memory stack entries per lane / staged top nodes (QBVH4); env MIROGPU_SHORT / _STAGE
    DeviceScene ds{};
    void* d_nodes = nullptr;
    void* d_tris = nullptr;
    void* d_shade = nullptr;
    void* d_planes = nullptr;              // 2 float4 per plane (unbounded objects, tested after the walk)
    uint32_t non_triangles = 0;            // spheres + planes: their hit points need the ray (resolve)
    int queue_mult = 4;                    // general wavefront, refractive scenes: path-queue capacity in units of the primary items (grows on overflow)
    std::vector<mirogpu_scene*> replicas;  // the other devices of a multi-device handle (this struct is the first device's replica)
    cudaStream_t mstream = nullptr;        // multi-device render: this replica's stream and its "rows are in place" event
    cudaEvent_t mevent = nullptr;
    mirogpu_material* d_materials = nullptr;
    float* d_uvs = nullptr;
    uint32_t nmaterials = 0;
    mirogpu_light* d_lights = nullptr;
    uint32_t nlights = 0;
    std::vector<mirogpu_light> h_lights;   // host copy (photon emission parameters are derived on the host)
    unsigned long long* d_ticket = nullptr;  // persistent-kernel ticket counters (ring of MIRO_TICKETS: a slot is reused that many launches later -- callers that queue launches on several streams far ahead of the device must not wrap onto a launch still running)
    std::atomic<uint32_t> ticket_slot{0};
    mirogpu_scene_info info{};
    std::vector<uint8_t> h_nodes;     // host copies kept for mirogpu_debug_copy_*
    std::vector<TriRecord> h_tris;
    PhotonMapDevice pm[2];
    RenderScratch scratch;
    std::mutex mtx;                   // guards scratch and the last-call stats
    uint64_t last_rays = 0, last_launches = 0;
    bool any_refractive = false, any_specular = false;
    uint32_t* h_stats = nullptr;      // pinned: device wave counters of the last render land here
    uint32_t stats_batches = 0, stats_mult = 1;
};

namespace {

// n is the number of rays, or -- when d_n is given -- an upper bound on it: the kernels then read the real count
// *d_n * mult from device memory (wavefront queues whose size the host never sees).
template <int LAYOUT, bool ANY, int PF, int MINB, int NREP, int SHORT = 0, int STAGE = 0>
cudaError_t launch_hybrid_inst(mirogpu_scene* h, const mirogpu_ray* d_rays, size_t n, mirogpu_hit* d_hits, unsigned long long* ticket,
                               cudaStream_t st, const uint32_t* d_n, uint32_t mult, int nmin)
{
    static std::atomic<int> cached_occ{0};   // per instantiation; the answer depends only on the kernel and the device type
    int occ = cached_occ.load(std::memory_order_relaxed);
    if (occ == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_hybrid<LAYOUT, ANY, PF, MINB, NREP, SHORT, STAGE>, 128, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        cached_occ.store(occ, std::memory_order_relaxed);
    }
    size_t grid = (size_t)h->sm_count * occ;
    const size_t need = (n + 127) / 128;
    if (grid > need) grid = need;
    k_trace_hybrid<LAYOUT, ANY, PF, MINB, NREP, SHORT, STAGE><<<(unsigned)grid, 128, 0, st>>>(h->ds, d_rays, n, d_hits, ticket, nmin, h->hyb_period,
                                                                            h->hyb_min_idle, (uint32_t)h->hyb_pool, d_n, mult, h->info.num_nodes);
    return cudaGetLastError();
}

template <int LAYOUT, bool ANY>
cudaError_t launch_hybrid(mirogpu_scene* h, const mirogpu_ray* d_rays, size_t n, mirogpu_hit* d_hits, unsigned long long* ticket,
                          cudaStream_t st, const uint32_t* d_n, uint32_t mult, bool coherent)
{
    // per-kind knobs: batches flagged coherent (camera rays in pixel order) walk without postponed leaves
    const int pf = coherent ? h->hyb_pf : h->hyb_pf_inc, nrep = h->hyb_nrep, minb = h->hyb_minb, sh = h->hyb_short, stg = h->hyb_stage;
    const int nmin = coherent ? h->hyb_nmin : h->hyb_nmin_inc;
#define MIRO_HYB(PF, MINB, NREP) if (pf == PF && minb == MINB && nrep == NREP && sh == 0 && stg == 0) return launch_hybrid_inst<LAYOUT, ANY, PF, MINB, NREP>(h, d_rays, n, d_hits, ticket, st, d_n, mult, nmin);
    MIRO_HYB(0, 9, 2) MIRO_HYB(0, 8, 2) MIRO_HYB(0, 9, 1) MIRO_HYB(0, 9, 3) MIRO_HYB(0, 10, 2) MIRO_HYB(16, 9, 2) MIRO_HYB(16, 9, 3) MIRO_HYB(16, 10, 3)
#undef MIRO_HYB
    if (LAYOUT == MIROGPU_LAYOUT_QBVH4) {
        // shared-memory short stack / staged top levels (north star; measured and rejected): kept selectable
#define MIRO_HYB2(MINB, SHORT, STAGE) if (pf == 16 && nrep == 3 && minb == MINB && sh == SHORT && stg == STAGE) return launch_hybrid_inst<MIROGPU_LAYOUT_QBVH4, ANY, 16, MINB, 3, SHORT, STAGE>(h, d_rays, n, d_hits, ticket, st, d_n, mult, nmin);
        MIRO_HYB2(9, 8, 0) MIRO_HYB2(9, 16, 0) MIRO_HYB2(10, 8, 0) MIRO_HYB2(10, 16, 0) MIRO_HYB2(9, 0, 21) MIRO_HYB2(9, 0, 85) MIRO_HYB2(9, 8, 85) MIRO_HYB2(10, 0, 85)
#undef MIRO_HYB2
        // leaf-phase variants: two triangles per phase (PF 48), one postponed leaf per lane (PF 80, the default for incoherent batches), two (PF 208)
#define MIRO_HYB3(PF, MINB, NREP, SHORT) if (pf == PF && minb == MINB && nrep == NREP && sh == SHORT && stg == 0) return launch_hybrid_inst<MIROGPU_LAYOUT_QBVH4, ANY, PF, MINB, NREP, SHORT, 0>(h, d_rays, n, d_hits, ticket, st, d_n, mult, nmin);
        MIRO_HYB3(48, 9, 3, 0) MIRO_HYB3(80, 9, 3, 0) MIRO_HYB3(80, 10, 3, 8) MIRO_HYB3(80, 10, 3, 0) MIRO_HYB3(80, 10, 2, 0) MIRO_HYB3(208, 10, 3, 0) MIRO_HYB3(208, 9, 3, 0) MIRO_HYB3(208, 10, 2, 0)
#undef MIRO_HYB3
        if (getenv("MIROGPU_STRICT")) return cudaErrorInvalidValue;   // measurement runs: an uninstantiated knob combination must not silently time the default
        return coherent ? launch_hybrid_inst<LAYOUT, ANY, 16, 10, 3>(h, d_rays, n, d_hits, ticket, st, d_n, mult, nmin)
                        : launch_hybrid_inst<MIROGPU_LAYOUT_QBVH4, ANY, 80, 10, 3, 0, 0>(h, d_rays, n, d_hits, ticket, st, d_n, mult, nmin);
    }
    return launch_hybrid_inst<LAYOUT, ANY, 0, 10, 2>(h, d_rays, n, d_hits, ticket, st, d_n, mult, nmin);
}

template <int LAYOUT, bool ANY>
cudaError_t launch_trace(mirogpu_scene* h, const mirogpu_ray* d_rays, size_t n, mirogpu_hit* d_hits, cudaStream_t st,
                         const uint32_t* d_n, uint32_t mult, bool coherent)
{
    if (n == 0) return cudaSuccess;
    if (h->variant == 1) {
        const unsigned grid = (unsigned)((n + 127) / 128);
        if (h->non_triangles) k_trace_simple<LAYOUT, ANY, false, true><<<grid, 128, 0, st>>>(h->ds, d_rays, n, d_hits, nullptr, d_n, mult);
        else k_trace_simple<LAYOUT, ANY, false><<<grid, 128, 0, st>>>(h->ds, d_rays, n, d_hits, nullptr, d_n, mult);
        return cudaGetLastError();
    }
    unsigned long long* ticket = h->d_ticket + (h->ticket_slot.fetch_add(1) & (MIRO_TICKETS - 1u));
    cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    if (h->non_triangles) {
        // scenes with spheres / planes: the NT instantiations (sphere slots in the leaf array, planes after the walk) with the
        // default scheduling of each layout -- the triangle-only kernels below compile without that code and keep their registers
        const bool hybrid = LAYOUT != MIROGPU_LAYOUT_CWBVH8 && h->variant != 0 && n < 0xFF000000ull;
        auto grid_of = [&](int occ) { size_t g = (size_t)h->sm_count * std::max(occ, 1); const size_t need = (n + 127) / 128; return (unsigned)std::min(g, need); };
        int occ = 0;
        if (hybrid) {
            constexpr int L = LAYOUT == MIROGPU_LAYOUT_CWBVH8 ? MIROGPU_LAYOUT_BVH2 : LAYOUT;
            constexpr int PFK = L == MIROGPU_LAYOUT_QBVH4 ? 16 : 0, NR = L == MIROGPU_LAYOUT_QBVH4 ? 3 : 2;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_hybrid<L, ANY, PFK, 9, NR, 0, 0, true>, 128, 0);
            if (e != cudaSuccess) return e;
            k_trace_hybrid<L, ANY, PFK, 9, NR, 0, 0, true><<<grid_of(occ), 128, 0, st>>>(h->ds, d_rays, n, d_hits, ticket, h->hyb_nmin, h->hyb_period, h->hyb_min_idle,
                                                                                     (uint32_t)h->hyb_pool, d_n, mult, h->info.num_nodes);
        } else {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_persistent<LAYOUT, ANY, true>, 128, 0);
            if (e != cudaSuccess) return e;
            k_trace_persistent<LAYOUT, ANY, true><<<grid_of(occ), 128, 0, st>>>(h->ds, d_rays, n, d_hits, ticket, d_n, mult, (uint32_t)h->packets_per_ticket);
        }
        return cudaGetLastError();
    }
    // automatic choice: the hybrid-scheduled kernel, except coherent batches on BVH2, where 32-ray packets walking in
    // lockstep are cheaper (QBVH4: hybrid 10.8 vs packets 9.2 Grays/s on camera rays; BVH2: 10.5 vs 11.5)
    if (LAYOUT != MIROGPU_LAYOUT_CWBVH8 && (h->variant == 2 || (h->variant < 0 && !(coherent && LAYOUT == MIROGPU_LAYOUT_BVH2))) && n < 0xFF000000ull)
        return launch_hybrid<LAYOUT == MIROGPU_LAYOUT_CWBVH8 ? MIROGPU_LAYOUT_BVH2 : LAYOUT, ANY>(h, d_rays, n, d_hits, ticket, st, d_n, mult, coherent);
    static std::atomic<int> cached_occ{0};
    int occ = cached_occ.load(std::memory_order_relaxed);
    if (occ == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_persistent<LAYOUT, ANY>, 128, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        cached_occ.store(occ, std::memory_order_relaxed);
    }
    size_t grid = (size_t)h->sm_count * occ;
    const size_t need = (n + 127) / 128;
    if (grid > need) grid = need;
    k_trace_persistent<LAYOUT, ANY><<<(unsigned)grid, 128, 0, st>>>(h->ds, d_rays, n, d_hits, ticket, d_n, mult, (uint32_t)h->packets_per_ticket);
    return cudaGetLastError();
}

cudaError_t dispatch_trace(mirogpu_scene* h, const mirogpu_ray* d_rays, size_t n, mirogpu_hit* d_hits, int mode, cudaStream_t st,
                           const uint32_t* d_n = nullptr, uint32_t mult = 1)
{
    const bool any = (mode & 0xff) == MIROGPU_ANY_HIT;
    const bool coherent = (mode & MIROGPU_HINT_COHERENT) != 0;
    if (h->layout == MIROGPU_LAYOUT_BVH2)
        return any ? launch_trace<MIROGPU_LAYOUT_BVH2, true>(h, d_rays, n, d_hits, st, d_n, mult, coherent)
                   : launch_trace<MIROGPU_LAYOUT_BVH2, false>(h, d_rays, n, d_hits, st, d_n, mult, coherent);
    if (h->layout == MIROGPU_LAYOUT_BVH4)
        return any ? launch_trace<MIROGPU_LAYOUT_BVH4, true>(h, d_rays, n, d_hits, st, d_n, mult, coherent)
                   : launch_trace<MIROGPU_LAYOUT_BVH4, false>(h, d_rays, n, d_hits, st, d_n, mult, coherent);
    if (h->layout == MIROGPU_LAYOUT_QBVH4)
        return any ? launch_trace<MIROGPU_LAYOUT_QBVH4, true>(h, d_rays, n, d_hits, st, d_n, mult, coherent)
                   : launch_trace<MIROGPU_LAYOUT_QBVH4, false>(h, d_rays, n, d_hits, st, d_n, mult, coherent);
    return any ? launch_trace<MIROGPU_LAYOUT_CWBVH8, true>(h, d_rays, n, d_hits, st, d_n, mult, coherent)
               : launch_trace<MIROGPU_LAYOUT_CWBVH8, false>(h, d_rays, n, d_hits, st, d_n, mult, coherent);
}

// Camera::eyeRay's cached basis (Camera.cpp:113-124), computed on the host in the reference's operand order.
void camera_basis(const mirogpu_camera& c, int W, int H, CameraBasis& b)
{
    auto norm3 = [](float* v) {
        const float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const float inv = float(1) / len;
        v[0] *= inv; v[1] *= inv; v[2] *= inv;
    };
    auto cross3 = [](const float* a, const float* bb, float* o) {
        o[0] = a[1] * bb[2] - a[2] * bb[1]; o[1] = a[2] * bb[0] - a[0] * bb[2]; o[2] = a[0] * bb[1] - a[1] * bb[0];
    };
    for (int k = 0; k < 3; ++k) { b.eye[k] = c.eye[k]; b.w[k] = -c.view_dir[k]; }
    norm3(b.w);
    cross3(c.up, b.w, b.u); norm3(b.u);
    cross3(b.w, b.u, b.v);
    const float PI = MIRO_PI, DegToRad = PI / 180.0f, HalfDegToRad = DegToRad / 2.0f;  // Miro.h:10-11, Camera.cpp:15
    const float aspect = (float)W / (float)H;
    b.top = tanf(c.fov_degrees * HalfDegToRad);
    b.right = aspect * b.top; b.bottom = -b.top; b.left = -b.right;
}

}  // namespace

// RenderScratch / render + photon implementations need the pieces above.
#include "photon_impl.cuh"
#include "render_impl.cuh"
#include "photon_trace_impl.cuh"
#include "lbvh_impl.cuh"

namespace {
// Borrow / return a staging slot of the handle.
BatchSlot* slot_acquire(mirogpu_scene* h, cudaError_t& e)
{
    e = cudaSuccess;
    std::lock_guard<std::mutex> lk(h->slot_mtx);
    for (auto& s : h->slots) if (!s->busy) { s->busy = true; return s.get(); }
    std::unique_ptr<BatchSlot> s(new (std::nothrow) BatchSlot);
    if (!s) { e = cudaErrorMemoryAllocation; return nullptr; }
    e = s->init();
    if (e != cudaSuccess) { s->release(); (void)cudaGetLastError(); return nullptr; }
    s->busy = true;
    h->slots.push_back(std::move(s));
    return h->slots.back().get();
}
void slot_release(mirogpu_scene* h, BatchSlot* s)
{
    if (!s) return;
    std::lock_guard<std::mutex> lk(h->slot_mtx);
    s->busy = false;
}
struct SlotGuard {
    mirogpu_scene* h; BatchSlot* s;
    ~SlotGuard() { slot_release(h, s); }
};
int cuda_code(cudaError_t e) { return e == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA; }

template <bool COUNT>
cudaError_t launch_simple(mirogpu_scene* h, const mirogpu_ray* d_r, size_t n, mirogpu_hit* d_h, bool any, unsigned long long* d_c, cudaStream_t st)
{
    const unsigned grid = (unsigned)((n + 127) / 128);
#define MIRO_SIMPLE(L)                                                                                         \
    {                                                                                                          \
        if (h->non_triangles) {                                                                                \
            if (any) k_trace_simple<L, true, COUNT, true><<<grid, 128, 0, st>>>(h->ds, d_r, n, d_h, d_c, nullptr, 1);  \
            else k_trace_simple<L, false, COUNT, true><<<grid, 128, 0, st>>>(h->ds, d_r, n, d_h, d_c, nullptr, 1);     \
        } else if (any) k_trace_simple<L, true, COUNT><<<grid, 128, 0, st>>>(h->ds, d_r, n, d_h, d_c, nullptr, 1);     \
        else k_trace_simple<L, false, COUNT><<<grid, 128, 0, st>>>(h->ds, d_r, n, d_h, d_c, nullptr, 1);               \
    }
    if (h->layout == MIROGPU_LAYOUT_BVH4) MIRO_SIMPLE(MIROGPU_LAYOUT_BVH4)
    else if (h->layout == MIROGPU_LAYOUT_QBVH4) MIRO_SIMPLE(MIROGPU_LAYOUT_QBVH4)
    else if (h->layout == MIROGPU_LAYOUT_BVH2) MIRO_SIMPLE(MIROGPU_LAYOUT_BVH2)
    else MIRO_SIMPLE(MIROGPU_LAYOUT_CWBVH8)
#undef MIRO_SIMPLE
    return cudaGetLastError();
}
}  // namespace

extern "C" {

int mirogpu_version(void) { return MIROGPU_VERSION; }

const char* mirogpu_last_error(void) { return g_last_error.c_str(); }

int mirogpu_device_count(int* count)
{
    if (!count) return fail(MIROGPU_ERR_INVALID_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; (void)cudaGetLastError(); return fail(MIROGPU_ERR_NO_DEVICE, cudaGetErrorString(e)); }
    *count = n;
    return MIROGPU_OK;
}

}  // extern "C"

namespace {

// What one build produces on the host; uploaded once per device of the handle.
struct HostBuild {
    mirogpu_build_options o;
    const void* node_src = nullptr; size_t node_bytes = 0;
    std::vector<TriRecord> tris;
    std::unique_ptr<float4[]> shade; size_t shade_count = 0;
    std::vector<mirogpu_material> mats;
    std::vector<float4> planes;   // 2 per plane: (normal, prim id bits) (origin, 0)
    std::vector<float> uvs;       // 6 per triangle, prim-id order (empty: no texture coordinates)
    bool textured = false;        // a material carries a procedural texture
    BinaryBvh bin; FlatBvh flat;
    uint32_t nprims = 0, ntris = 0, nspheres = 0, nplanes = 0;
    mirogpu_scene_info info{};
    bool any_refractive = false, any_specular = false;
};

void apply_tuning_knobs(mirogpu_scene* h, int layout)
{
    if (layout == MIROGPU_LAYOUT_QBVH4 || layout == MIROGPU_LAYOUT_BVH4) { h->hyb_period = 2; h->hyb_min_idle = 6; }   // measured optimum of the four-wide steps
    // QBVH4: one triangle per leaf phase, three node steps per vote, node steps while >= 20 lanes want one (7.28 -> 7.53 Grays/s);
    // incoherent batches: one postponed leaf per lane, node steps while >= 24 lanes want one (bounce rays 7.17 -> 8.02 Grays/s with 48 registers)
    if (layout == MIROGPU_LAYOUT_QBVH4) { h->hyb_pf = 16; h->hyb_nrep = 3; h->hyb_nmin = 20; h->hyb_pf_inc = 80; h->hyb_nmin_inc = 24; }
    else h->hyb_nmin_inc = h->hyb_nmin;
    if (const char* e = getenv("MIROGPU_POOL")) { const int v = atoi(e); if (v >= 32 && v <= 65536) h->hyb_pool = v; }
    if (const char* e = getenv("MIROGPU_NREP")) h->hyb_nrep = atoi(e);
    if (const char* e = getenv("MIROGPU_PPT")) { const int v = atoi(e); if (v >= 1 && v <= 64) h->packets_per_ticket = v; }
    if (const char* e = getenv("MIROGPU_NMIN")) { const int v = atoi(e); if (v >= 1 && v <= 32) h->hyb_nmin = h->hyb_nmin_inc = v; }
    if (const char* e = getenv("MIROGPU_NMIN_INC")) { const int v = atoi(e); if (v >= 1 && v <= 32) h->hyb_nmin_inc = v; }
    if (const char* e = getenv("MIROGPU_PERIOD")) { const int v = atoi(e); if (v >= 1 && v <= 100000) h->hyb_period = v; }
    if (const char* e = getenv("MIROGPU_PF")) h->hyb_pf = h->hyb_pf_inc = atoi(e);
    if (const char* e = getenv("MIROGPU_PF_INC")) h->hyb_pf_inc = atoi(e);
    if (const char* e = getenv("MIROGPU_MINB")) h->hyb_minb = atoi(e);
    if (const char* e = getenv("MIROGPU_SHORT")) h->hyb_short = atoi(e);
    if (const char* e = getenv("MIROGPU_STAGE")) h->hyb_stage = atoi(e);
    if (const char* e = getenv("MIROGPU_MINIDLE")) { const int v = atoi(e); if (v >= 1 && v <= 32) h->hyb_min_idle = v; }
    if (const char* e = getenv("MIROGPU_QUEUE_MULT")) { const int v = atoi(e); if (v >= 1 && v <= 256) h->queue_mult = v; }   // tests: start small to exercise the overflow retry
}

// Uploads one replica of the scene to `dev`.  lb: a tree built on that device (device builders), or NULL.
int upload_replica(const HostBuild& hb, int dev, const LbvhOut* lb, mirogpu_scene** out)
{
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(dev));
    mirogpu_scene* h = new (std::nothrow) mirogpu_scene;
    if (!h) return fail(MIROGPU_ERR_OOM, "host allocation failed");
    h->device = dev; h->layout = hb.o.layout;
    {
        int sms = 0;
        const cudaError_t pe = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (pe != cudaSuccess) { delete h; return fail(MIROGPU_ERR_CUDA, std::string("cudaDeviceGetAttribute: ") + cudaGetErrorString(pe)); }
        h->sm_count = sms;
    }
    apply_tuning_knobs(h, hb.o.layout);
    const bool device_built = lb != nullptr;
    const size_t node_bytes = device_built ? lb->node_bytes : hb.node_bytes;
    const size_t tri_bytes = device_built ? lb->tri_bytes : hb.tris.size() * sizeof(TriRecord), shade_bytes = hb.shade_count * sizeof(float4);
    auto bail = [&](cudaError_t e, const char* what) {
        std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
        mirogpu_scene_destroy(h);
        return fail(e == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA, msg);
    };
    cudaError_t e;
    // nodes and triangles share one allocation (the traversal working set is one address range)
    const size_t node_span = device_built ? lb->node_span : (std::max<size_t>(node_bytes, 16) + 255) & ~(size_t)255;
    if (device_built) h->d_nodes = lb->d_geom;
    else if ((e = cudaMalloc(&h->d_nodes, node_span + std::max<size_t>(tri_bytes, 16))) != cudaSuccess) return bail(e, "cudaMalloc nodes + triangles");
    h->d_tris = static_cast<char*>(h->d_nodes) + node_span;
    h->node_bytes_dev = node_bytes; h->tri_bytes_dev = tri_bytes;
    if ((e = cudaMalloc(&h->d_shade, std::max<size_t>(shade_bytes, 16))) != cudaSuccess) return bail(e, "cudaMalloc shading records");
    if ((e = cudaMalloc(&h->d_materials, hb.mats.size() * sizeof(mirogpu_material))) != cudaSuccess) return bail(e, "cudaMalloc materials");
    if ((e = cudaMalloc(&h->d_ticket, MIRO_TICKETS * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc tickets");
    if (!hb.planes.empty() && (e = cudaMalloc(&h->d_planes, hb.planes.size() * sizeof(float4))) != cudaSuccess) return bail(e, "cudaMalloc planes");
    if (!device_built && node_bytes && (e = cudaMemcpy(h->d_nodes, hb.node_src, node_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload nodes");
    if (!device_built && tri_bytes && (e = cudaMemcpy(h->d_tris, hb.tris.data(), tri_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload triangles");
    if (shade_bytes && (e = cudaMemcpy(h->d_shade, hb.shade.get(), shade_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload shading records");
    if ((e = cudaMemcpy(h->d_materials, hb.mats.data(), hb.mats.size() * sizeof(mirogpu_material), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload materials");
    if (!hb.planes.empty() && (e = cudaMemcpy(h->d_planes, hb.planes.data(), hb.planes.size() * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload planes");
    if (hb.textured && !hb.uvs.empty()) {
        if ((e = cudaMalloc(&h->d_uvs, hb.uvs.size() * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc texture coordinates");
        if ((e = cudaMemcpy(h->d_uvs, hb.uvs.data(), hb.uvs.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "upload texture coordinates");
    }
    if ((e = cudaMemset(h->d_ticket, 0, MIRO_TICKETS * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "memset tickets");
    h->nmaterials = (uint32_t)hb.mats.size();
    h->any_refractive = hb.any_refractive; h->any_specular = hb.any_specular;
    if ((e = cudaHostAlloc(&h->h_stats, 64 * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) return bail(e, "cudaHostAlloc stats");
    memset(h->h_stats, 0, 64 * sizeof(uint32_t));
    h->ds.nodes = reinterpret_cast<const float4*>(h->d_nodes);
    h->ds.tris = reinterpret_cast<const float4*>(h->d_tris);
    h->ds.shade = reinterpret_cast<const float4*>(h->d_shade);
    h->ds.planes = reinterpret_cast<const float4*>(h->d_planes);
    h->ds.num_tris = hb.nprims;
    h->ds.num_planes = hb.nplanes;
    h->ds.mats = h->d_materials; h->ds.uvs = h->d_uvs; h->ds.textured = hb.textured ? 1u : 0u;
    h->non_triangles = hb.nspheres + hb.nplanes;
    h->info = hb.info;
    *out = h;
    return MIROGPU_OK;
}

int scene_create_impl(const mirogpu_scene_desc& d, const mirogpu_build_options* opt, mirogpu_handle* out)
{
    *out = nullptr;
    const uint32_t ntris = d.ntris, nspheres = d.nspheres, nplanes = d.nplanes;
    if (ntris && !d.tri_vertices) return fail(MIROGPU_ERR_INVALID_ARG, "tri_vertices is NULL");
    if ((nspheres && !d.spheres) || (nplanes && !d.planes)) return fail(MIROGPU_ERR_INVALID_ARG, "sphere / plane array is NULL");
    if ((uint64_t)ntris + nspheres + nplanes >= (1u << 28) - 16u) return fail(MIROGPU_ERR_INVALID_ARG, "too many primitives (limit 2^28 - 16)");
    HostBuild hb;
    mirogpu_build_options& o = hb.o;
    o.layout = MIROGPU_LAYOUT_QBVH4; o.max_leaf = 0; o.sah_bins = 32; o.device = -1; o.builder = MIROGPU_BUILDER_SAH_HOST;
    if (opt) o = *opt;
    if (const char* e = getenv("MIROGPU_BUILDER")) {   // tuning knob
        if (!strcmp(e, "lbvh")) o.builder = MIROGPU_BUILDER_LBVH_DEVICE;
        else if (!strcmp(e, "ploc")) o.builder = MIROGPU_BUILDER_PLOC_DEVICE;
        if (o.builder != MIROGPU_BUILDER_SAH_HOST && o.layout != MIROGPU_LAYOUT_QBVH4) o.builder = MIROGPU_BUILDER_SAH_HOST;   // the knob only applies where it can
    }
    bool device_builder = o.builder == MIROGPU_BUILDER_LBVH_DEVICE || o.builder == MIROGPU_BUILDER_PLOC_DEVICE;
    if (o.builder != MIROGPU_BUILDER_SAH_HOST && !device_builder) return fail(MIROGPU_ERR_INVALID_ARG, "unknown builder");
    if (device_builder && o.layout != MIROGPU_LAYOUT_QBVH4) return fail(MIROGPU_ERR_INVALID_ARG, "the device builders emit the QBVH4 layout only");
    if (o.layout != MIROGPU_LAYOUT_BVH2 && o.layout != MIROGPU_LAYOUT_CWBVH8 && o.layout != MIROGPU_LAYOUT_BVH4 && o.layout != MIROGPU_LAYOUT_QBVH4)
        return fail(MIROGPU_ERR_INVALID_ARG, "unknown layout");
    if (o.max_leaf <= 0) if (const char* e = getenv("MIROGPU_MAX_LEAF")) o.max_leaf = atoi(e);   // tuning knob
    // The device builders make every subtree of <= max_leaf triangles a leaf.  Unless the caller asks otherwise they build
    // single-triangle leaves: the agglomerated tree has no leaf cost model, and four-triangle leaves cost the hybrid kernel (one
    // triangle per lane and leaf phase) 13 % of its camera-ray rate on the bench scene -- 5.75 / 6.34 / 6.34 / 6.57 Grays/s for
    // 4 / 3 / 2 / 1 triangles per leaf, against 6.9 for the host SAH tree, whose builder decides leaf sizes by cost.
    const int device_max_leaf = o.max_leaf > 0 ? std::min(o.max_leaf, 4) : 1;
    if (o.max_leaf <= 0) o.max_leaf = (o.layout == MIROGPU_LAYOUT_CWBVH8) ? 3 : 4;
    if (o.layout == MIROGPU_LAYOUT_CWBVH8 && o.max_leaf > 3) o.max_leaf = 3;
    if (o.layout != MIROGPU_LAYOUT_CWBVH8 && o.max_leaf > 8) o.max_leaf = 8;
    if (o.sah_bins <= 0) o.sah_bins = 32;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(MIROGPU_ERR_NO_DEVICE, "no CUDA device: mirogpu has no CPU path");
    }
    std::vector<int> devs;
    if (d.devices && d.ndevices) {
        for (uint32_t i = 0; i < d.ndevices; ++i) {
            if (d.devices[i] < 0 || d.devices[i] >= ndev) return fail(MIROGPU_ERR_INVALID_ARG, "device ordinal out of range");
            for (int q : devs) if (q == d.devices[i]) return fail(MIROGPU_ERR_INVALID_ARG, "device listed twice");
            devs.push_back(d.devices[i]);
        }
    } else {
        int dev = o.device;
        if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= ndev) return fail(MIROGPU_ERR_INVALID_ARG, "device ordinal out of range");
        devs.push_back(dev);
    }
    if (devs.size() > 16) return fail(MIROGPU_ERR_INVALID_ARG, "more than 16 devices");
    CUDA_TRY(cudaSetDevice(devs[0]));
    // spheres enter the builder as proxy triangles whose bounding box is the sphere's ([c - r, c + r] exactly; the builder
    // pads it further); the device builders make their own triangle records, so a scene with spheres is built on the host
    const uint32_t nleafprims = ntris + nspheres;
    hb.ntris = ntris; hb.nspheres = nspheres; hb.nplanes = nplanes; hb.nprims = nleafprims + nplanes;
    if (nspheres && device_builder) { device_builder = false; o.builder = MIROGPU_BUILDER_SAH_HOST; }
    if (devs.size() > 1 && device_builder) { device_builder = false; o.builder = MIROGPU_BUILDER_SAH_HOST; }   // one host build, N uploads
    std::vector<float> verts_all;
    const float* verts = d.tri_vertices;
    if (nspheres) {
        verts_all.resize((size_t)nleafprims * 9);
        if (ntris) memcpy(verts_all.data(), d.tri_vertices, (size_t)ntris * 9 * sizeof(float));
        for (uint32_t i = 0; i < nspheres; ++i) {
            const mirogpu_sphere& sp = d.spheres[i];
            float* v = verts_all.data() + (size_t)(ntris + i) * 9;
            for (int k = 0; k < 3; ++k) { v[k] = sp.center[k] - sp.radius; v[3 + k] = sp.center[k] + sp.radius; }
            v[6] = sp.center[0] - sp.radius; v[7] = sp.center[1] + sp.radius; v[8] = sp.center[2] - sp.radius;
        }
        verts = verts_all.data();
    }

    // ---- device build (LBVH / PLOC; single device, triangles only) ------------------------------------------------------
    double t0 = now_s();
    LbvhOut lb;
    bool device_built = false;
    if (device_builder) {
        const cudaError_t be = build_lbvh_device(verts, ntris, device_max_leaf, lb, o.builder == MIROGPU_BUILDER_PLOC_DEVICE);
        if (be == cudaErrorNotSupported) lb.d_geom = nullptr;   // the device builder gave up on this input (see lbvh_impl.cuh): use the host builder
        else if (be != cudaSuccess) return fail(be == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA, std::string("device BVH build: ") + cudaGetErrorString(be));
        else if (lb.max_stack <= MIRO_STACK4) device_built = true;
        else { cudaFree(lb.d_geom); lb.d_geom = nullptr; }   // a tree too deep for the kernels' stacks: use the host builder
    }
    // ---- host build ------------------------------------------------------------------------------
    BinaryBvh& bin = hb.bin;
    FlatBvh& flat = hb.flat;
    double t1 = now_s(), t2 = t1;
    const bool wide4 = o.layout == MIROGPU_LAYOUT_BVH4 || o.layout == MIROGPU_LAYOUT_QBVH4;
    if (!device_built) {
        t0 = now_s();
        bin = build_binary_sah(verts, nleafprims, o.max_leaf, o.sah_bins);
        t1 = now_s();
        if (o.layout == MIROGPU_LAYOUT_BVH2) flatten_bvh2(bin, flat);
        else if (o.layout == MIROGPU_LAYOUT_BVH4) flatten_bvh4(bin, flat);
        else if (o.layout == MIROGPU_LAYOUT_QBVH4) flatten_qbvh4(bin, flat);
        else flatten_cwbvh8(bin, flat);
        // a walk pushes at most one entry per level (BVH2 / CWBVH8 groups) resp. flat.max_stack entries (BVH4): refuse a tree the
        // kernels' fixed per-thread stacks cannot hold rather than overrun them (the builder's depth cap makes this unreachable
        // below ~16 M triangles)
        if (!wide4 && bin.max_depth > MIRO_STACK) return fail(MIROGPU_ERR_INVALID_ARG, "tree deeper than the kernels' traversal stack");
        if (wide4 && flat.max_stack > MIRO_STACK4) return fail(MIROGPU_ERR_INVALID_ARG, "BVH4 tree needs a deeper traversal stack than the kernels carry");
        make_tri_records(verts, flat.order, hb.tris);
        // sphere slots: (centre, prim id) (0, 0, 0, 0) (0, 0, 0, 0) (0, kind = 1, radius, 0) -- see tri_test in traverse.cuh
        if (nspheres)
            for (TriRecord& r : hb.tris)
                if (r.prim_id >= ntris) {
                    const mirogpu_sphere& sp = d.spheres[r.prim_id - ntris];
                    const uint32_t id = r.prim_id;
                    memset(&r, 0, sizeof r);
                    r.ax = sp.center[0]; r.ay = sp.center[1]; r.az = sp.center[2]; r.prim_id = id;
                    r.pad[0] = 1.0f; r.pad[1] = sp.radius;
                }
        t2 = now_s();
        if (o.layout == MIROGPU_LAYOUT_BVH2) { hb.node_src = flat.nodes2.data(); hb.node_bytes = flat.nodes2.size() * sizeof(Bvh2Node); }
        else if (o.layout == MIROGPU_LAYOUT_BVH4) { hb.node_src = flat.nodes4.data(); hb.node_bytes = flat.nodes4.size() * sizeof(Bvh4Node); }
        else if (o.layout == MIROGPU_LAYOUT_QBVH4) { hb.node_src = flat.nodesq.data(); hb.node_bytes = flat.nodesq.size() * sizeof(Qbvh4Node); }
        else { hb.node_src = flat.nodes8.data(); hb.node_bytes = flat.nodes8.size() * sizeof(Cwbvh8Node); }
    }

    // shading records in prim-id order: (A, material) (e1, kind) e2 nA nB nC -- 96 bytes per primitive, written by all host threads
    // (133 MB for the bench scene: a single thread spends longer here than the device builders spend on the whole tree).
    // kind (bits of e1.w): 0 triangle, 1 sphere ((centre, material) (radius, 1)), 2 plane ((origin, material) (-, 2) - normal in the nA slot)
    const uint32_t nmat_eff = std::max(d.materials ? d.nmaterials : 0u, 1u);   // no table = one white Lambert
    hb.shade_count = (size_t)hb.nprims * 6;
    hb.shade.reset(new float4[std::max<size_t>(hb.shade_count, 1)]);
    float4* shade = hb.shade.get();
    bool bad_material = false;
    const float* tri_vertices = d.tri_vertices; const float* tri_normals = d.tri_normals; const uint32_t* material_ids = d.tri_material_ids;
#pragma omp parallel for schedule(static) reduction(|| : bad_material)
    for (int64_t ii = 0; ii < (int64_t)ntris; ++ii) {
        const size_t i = (size_t)ii;
        const float* v = tri_vertices + 9 * i;
        uint32_t m = material_ids ? material_ids[i] : 0u;
        if (m >= nmat_eff) { bad_material = true; m = 0u; }
        float4 A = make_float4(v[0], v[1], v[2], 0.f);
        memcpy(&A.w, &m, 4);
        shade[6 * i + 0] = A;
        shade[6 * i + 1] = make_float4(v[3] - v[0], v[4] - v[1], v[5] - v[2], 0.f);
        shade[6 * i + 2] = make_float4(v[6] - v[0], v[7] - v[1], v[8] - v[2], 0.f);
        for (int k = 0; k < 3; ++k) {
            if (tri_normals) shade[6 * i + 3 + k] = make_float4(tri_normals[9 * i + 3 * k], tri_normals[9 * i + 3 * k + 1], tri_normals[9 * i + 3 * k + 2], 0.f);
            else shade[6 * i + 3 + k] = make_float4(0.f, 1.f, 0.f, 0.f);
        }
    }
    auto kind_bits = [](uint32_t k) { float f; memcpy(&f, &k, 4); return f; };
    for (uint32_t i = 0; i < nspheres; ++i) {
        const mirogpu_sphere& sp = d.spheres[i];
        uint32_t m = sp.material_id;
        if (m >= nmat_eff) { bad_material = true; m = 0u; }
        float4* r = shade + 6 * (size_t)(ntris + i);
        for (int k = 0; k < 6; ++k) r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        r[0] = make_float4(sp.center[0], sp.center[1], sp.center[2], 0.f); memcpy(&r[0].w, &m, 4);
        r[1] = make_float4(sp.radius, 0.f, 0.f, kind_bits(1u));
    }
    for (uint32_t i = 0; i < nplanes; ++i) {
        const mirogpu_plane& pl = d.planes[i];
        uint32_t m = pl.material_id;
        if (m >= nmat_eff) { bad_material = true; m = 0u; }
        const uint32_t id = nleafprims + i;
        float4* r = shade + 6 * (size_t)id;
        for (int k = 0; k < 6; ++k) r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        r[0] = make_float4(pl.origin[0], pl.origin[1], pl.origin[2], 0.f); memcpy(&r[0].w, &m, 4);
        r[1] = make_float4(0.f, 0.f, 0.f, kind_bits(2u));
        r[3] = make_float4(pl.normal[0], pl.normal[1], pl.normal[2], 0.f);
        float4 n4 = make_float4(pl.normal[0], pl.normal[1], pl.normal[2], 0.f); memcpy(&n4.w, &id, 4);
        hb.planes.push_back(n4);
        hb.planes.push_back(make_float4(pl.origin[0], pl.origin[1], pl.origin[2], 0.f));
    }
    if (bad_material) {
        if (device_built && lb.d_geom) cudaFree(lb.d_geom);
        return fail(MIROGPU_ERR_INVALID_ARG, "material id out of range");
    }
    if (d.materials && d.nmaterials) hb.mats.assign(d.materials, d.materials + d.nmaterials);
    else {
        mirogpu_material m; memset(&m, 0, sizeof m);
        m.kd[0] = m.kd[1] = m.kd[2] = 1.f; m.shininess = 1.f; m.refract_index = 1.f;  // Lambert(Vector3(1)) = Phong defaults
        hb.mats.push_back(m);
    }
    for (const mirogpu_material& m : hb.mats) {
        hb.any_refractive |= m.kt[0] > 0.f || m.kt[1] > 0.f || m.kt[2] > 0.f;
        hb.any_specular |= m.ks[0] > 0.f || m.ks[1] > 0.f || m.ks[2] > 0.f;
        if (m.texture < MIROGPU_TEX_NONE || m.texture > MIROGPU_TEX_FLOWER_CENTER) {
            if (device_built && lb.d_geom) cudaFree(lb.d_geom);
            return fail(MIROGPU_ERR_INVALID_ARG, "unknown texture kind");
        }
        hb.textured |= m.texture != MIROGPU_TEX_NONE;
    }
    if (hb.textured && d.tri_texcoords && ntris) hb.uvs.assign(d.tri_texcoords, d.tri_texcoords + (size_t)ntris * 6);

    mirogpu_scene_info& in = hb.info;
    in.num_triangles = ntris;
    in.num_nodes = device_built ? lb.num_nodes : (uint32_t)(o.layout == MIROGPU_LAYOUT_BVH2 ? flat.nodes2.size() : wide4 ? flat.nodes4.size() : flat.nodes8.size());
    in.num_binary_nodes = device_built ? (ntris ? 2 * ntris - 1 : 0) : (uint32_t)bin.nodes.size();
    in.num_binary_leaves = device_built ? ntris : bin.num_leaves;
    in.max_depth = device_built ? lb.max_depth : (o.layout == MIROGPU_LAYOUT_BVH2 ? bin.max_depth : flat.max_depth);
    in.builder = device_built ? o.builder : MIROGPU_BUILDER_SAH_HOST;
    in.layout = o.layout;
    in.node_bytes = device_built ? lb.node_bytes : hb.node_bytes;
    in.triangle_bytes = device_built ? lb.tri_bytes : hb.tris.size() * sizeof(TriRecord); in.shading_bytes = hb.shade_count * sizeof(float4);
    in.build_seconds = device_built ? lb.seconds : t1 - t0; in.flatten_seconds = device_built ? 0.0 : t2 - t1;
    for (int k = 0; k < 3; ++k) { in.bounds_min[k] = device_built ? lb.lo[k] : flat.root.lo[k]; in.bounds_max[k] = device_built ? lb.hi[k] : flat.root.hi[k]; }

    // ---- upload: one replica per device ---------------------------------------------------------------------------------
    const double tu = now_s();
    mirogpu_scene* primary = nullptr;
    int rc = upload_replica(hb, devs[0], device_built ? &lb : nullptr, &primary);
    if (rc != MIROGPU_OK) { if (device_built && lb.d_geom && !primary) cudaFree(lb.d_geom); return rc; }
    if (!device_built) { primary->h_nodes.assign((const uint8_t*)hb.node_src, (const uint8_t*)hb.node_src + hb.node_bytes); primary->h_tris = hb.tris; }
    for (size_t k = 1; k < devs.size(); ++k) {
        mirogpu_scene* rep = nullptr;
        rc = upload_replica(hb, devs[k], nullptr, &rep);
        if (rc != MIROGPU_OK) { const std::string msg = g_last_error; mirogpu_scene_destroy(primary); return fail(rc, msg); }
        primary->replicas.push_back(rep);
    }
    // peer access between the replicas' devices (framebuffer gather over NVLink); failure only means staged copies
    for (size_t a2 = 0; a2 < devs.size(); ++a2)
        for (size_t b2 = 0; b2 < devs.size(); ++b2)
            if (a2 != b2) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devs[a2], devs[b2]) == cudaSuccess && can) {
                    cudaSetDevice(devs[a2]);
                    const cudaError_t pe = cudaDeviceEnablePeerAccess(devs[b2], 0);
                    if (pe != cudaSuccess) (void)cudaGetLastError();   // already enabled is fine
                }
            }
    cudaSetDevice(devs[0]);
    primary->info.upload_seconds = now_s() - tu;
    *out = primary;
    return MIROGPU_OK;
}

}  // namespace

extern "C" {

int mirogpu_scene_create(const float* tri_vertices, const float* tri_normals, const uint32_t* material_ids, uint32_t ntris,
                         const mirogpu_material* materials, uint32_t nmaterials, const mirogpu_build_options* opt,
                         mirogpu_handle* out)
{
    if (!out) return fail(MIROGPU_ERR_INVALID_ARG, "out handle is NULL");
    mirogpu_scene_desc d; memset(&d, 0, sizeof d);
    d.tri_vertices = tri_vertices; d.tri_normals = tri_normals; d.tri_material_ids = material_ids; d.ntris = ntris;
    d.materials = materials; d.nmaterials = nmaterials;
    return scene_create_impl(d, opt, out);
}

int mirogpu_scene_create_ex(const mirogpu_scene_desc* desc, const mirogpu_build_options* opt, mirogpu_handle* out)
{
    if (!out) return fail(MIROGPU_ERR_INVALID_ARG, "out handle is NULL");
    *out = nullptr;
    if (!desc) return fail(MIROGPU_ERR_INVALID_ARG, "desc is NULL");
    return scene_create_impl(*desc, opt, out);
}

int mirogpu_scene_devices(mirogpu_handle h, int32_t* devices, uint32_t capacity, uint32_t* ndevices)
{
    if (!h || !ndevices) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    *ndevices = 1u + (uint32_t)h->replicas.size();
    if (devices) {
        if (capacity > 0) devices[0] = h->device;
        for (size_t k = 0; k < h->replicas.size() && k + 1 < capacity; ++k) devices[k + 1] = h->replicas[k]->device;
    }
    return MIROGPU_OK;
}

int mirogpu_scene_destroy(mirogpu_handle h)
{
    if (!h) return MIROGPU_OK;
    for (mirogpu_scene* r : h->replicas) mirogpu_scene_destroy(r);
    h->replicas.clear();
    cudaSetDevice(h->device);
    cudaFree(h->d_nodes); cudaFree(h->d_shade); cudaFree(h->d_materials); cudaFree(h->d_uvs);   // d_tris lives inside d_nodes's allocation
    cudaFree(h->d_lights); cudaFree(h->d_ticket); cudaFree(h->d_planes);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    for (int i = 0; i < 2; ++i) h->pm[i].release();
    for (auto& sl : h->slots) sl->release();
    h->scratch.release();
    if (h->mevent) cudaEventDestroy(h->mevent);
    if (h->mstream) cudaStreamDestroy(h->mstream);
    (void)cudaGetLastError();
    delete h;
    return MIROGPU_OK;
}

int mirogpu_scene_info_get(mirogpu_handle h, mirogpu_scene_info* info)
{
    if (!h || !info) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    *info = h->info;
    return MIROGPU_OK;
}

int mirogpu_scene_set_lights(mirogpu_handle h, const mirogpu_light* lights, uint32_t nlights)
{
    if (!h || (nlights && !lights)) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    for (mirogpu_scene* r : h->replicas) { const int rc = mirogpu_scene_set_lights(r, lights, nlights); if (rc != MIROGPU_OK) return rc; }
    CUDA_TRY(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> lk(h->mtx);
    cudaFree(h->d_lights); h->d_lights = nullptr; h->nlights = 0; h->h_lights.clear();
    if (nlights) {
        CUDA_TRY(cudaMalloc(&h->d_lights, nlights * sizeof(mirogpu_light)));
        CUDA_TRY(cudaMemcpy(h->d_lights, lights, nlights * sizeof(mirogpu_light), cudaMemcpyHostToDevice));
        h->nlights = nlights;
    }
    h->h_lights.assign(lights, lights + nlights);
    return MIROGPU_OK;
}

int mirogpu_debug_copy_nodes(mirogpu_handle h, void* out, uint64_t* bytes)
{
    if (!h || !bytes) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    const uint64_t need = h->node_bytes_dev;
    if (out && *bytes >= need) {
        if (h->h_nodes.size() == need) memcpy(out, h->h_nodes.data(), need);
        else { CUDA_TRY(cudaSetDevice(h->device)); CUDA_TRY(cudaMemcpy(out, h->d_nodes, need, cudaMemcpyDeviceToHost)); }   // built on the device
    }
    *bytes = need;
    return MIROGPU_OK;
}

int mirogpu_debug_copy_triangles(mirogpu_handle h, void* out, uint64_t* bytes)
{
    if (!h || !bytes) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    const uint64_t need = h->tri_bytes_dev;
    if (out && *bytes >= need) {
        if (h->h_tris.size() * sizeof(TriRecord) == need) memcpy(out, h->h_tris.data(), need);
        else { CUDA_TRY(cudaSetDevice(h->device)); CUDA_TRY(cudaMemcpy(out, h->d_tris, need, cudaMemcpyDeviceToHost)); }
    }
    *bytes = need;
    return MIROGPU_OK;
}

int mirogpu_set_kernel_variant(mirogpu_handle h, int variant)
{
    if (!h) return fail(MIROGPU_ERR_INVALID_ARG, "NULL handle");
    if (variant < -1 || variant > 2)
        return fail(MIROGPU_ERR_INVALID_ARG, "variant must be -1 (automatic), 0 (persistent warps, 32-ray tickets, while-while), 1 (one thread per ray) or 2 (persistent warps, hybrid step scheduling + ray replacement; BVH2, BVH4, QBVH4)");
    h->variant = variant;
    return MIROGPU_OK;
}

int mirogpu_intersect_batch_device(mirogpu_handle h, const mirogpu_ray* d_rays, size_t n, mirogpu_hit* d_hits, int mode,
                                   void* cuda_stream)
{
    if (!h || (n && (!d_rays || !d_hits))) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if ((mode & ~MIROGPU_HINT_COHERENT) != MIROGPU_CLOSEST_HIT && (mode & ~MIROGPU_HINT_COHERENT) != MIROGPU_ANY_HIT)
        return fail(MIROGPU_ERR_INVALID_ARG, "unknown mode");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(dispatch_trace(h, d_rays, n, d_hits, mode, (cudaStream_t)cuda_stream));
    {
        std::lock_guard<std::mutex> lk(h->mtx);
        h->last_rays = n; h->last_launches = n ? 1 : 0; h->stats_batches = 0;
    }
    return MIROGPU_OK;
}

// Host-buffer query.  Tiny batches (the reference calls BVH::intersect one ray at a time): the rays are copied into the
// slot's page-locked buffer, ONE kernel (a thread per ray) reads them and writes the hits through the mapping, one stream
// synchronisation -- no device allocation, no copy engine.  Large batches: chunks of up to 4 Mi rays double-buffered over two
// slots so the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the traversal of chunk k.
int mirogpu_intersect_batch(mirogpu_handle h, const mirogpu_ray* rays, size_t n, mirogpu_hit* hits, int mode)
{
    if (!h || (n && (!rays || !hits))) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if ((mode & ~MIROGPU_HINT_COHERENT) != MIROGPU_CLOSEST_HIT && (mode & ~MIROGPU_HINT_COHERENT) != MIROGPU_ANY_HIT)
        return fail(MIROGPU_ERR_INVALID_ARG, "unknown mode");
    if (n == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaError_t e;
    SlotGuard g0{h, slot_acquire(h, e)};
    if (!g0.s) return fail(cuda_code(e), std::string("intersect_batch staging: ") + cudaGetErrorString(e));
    uint64_t launches = 0;
    if (n <= MIRO_SMALL_BATCH) {
        memcpy(g0.s->h_r, rays, n * sizeof(mirogpu_ray));
        e = launch_simple<false>(h, g0.s->h_r, n, g0.s->h_h, (mode & 0xff) == MIROGPU_ANY_HIT, nullptr, g0.s->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g0.s->st);
        if (e != cudaSuccess) return fail(MIROGPU_ERR_CUDA, std::string("intersect_batch: ") + cudaGetErrorString(e));
        memcpy(hits, g0.s->h_h, n * sizeof(mirogpu_hit));
        launches = 1;
    } else {
        // Pageable caller buffers (the usual case for a C++ caller's std::vector): cudaMemcpyAsync would stage them through the
        // driver's own bounce buffer on this thread (measured: 0.23 Grays/s against 1.2 from page-locked memory).  Instead the
        // chunks go through page-locked staging of the slots, filled and drained by all host threads, two chunks in flight.
        cudaPointerAttributes pa_r, pa_h;
        const bool pageable = (cudaPointerGetAttributes(&pa_r, rays) != cudaSuccess || pa_r.type == cudaMemoryTypeUnregistered) ||
                              (cudaPointerGetAttributes(&pa_h, hits) != cudaSuccess || pa_h.type == cudaMemoryTypeUnregistered);
        (void)cudaGetLastError();
        const size_t chunk = std::min<size_t>(n, pageable ? (size_t)1 << 20 : (size_t)4 << 20);
        const int nbuf = n > chunk ? 2 : 1;
        SlotGuard g1{h, nbuf == 2 ? slot_acquire(h, e) : nullptr};
        if (nbuf == 2 && !g1.s) return fail(cuda_code(e), std::string("intersect_batch staging: ") + cudaGetErrorString(e));
        BatchSlot* sl[2] = {g0.s, g1.s};
        for (int b = 0; b < nbuf; ++b) {
            e = sl[b]->ensure(chunk);
            if (e == cudaSuccess && pageable) e = sl[b]->ensure_pinned(chunk);
            if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(cuda_code(e), std::string("intersect_batch staging: ") + cudaGetErrorString(e)); }
        }
        auto par_copy = [](void* dst, const void* src, size_t bytes) {
            const size_t piece = (size_t)1 << 20;
            const long pieces = (long)((bytes + piece - 1) / piece);
#pragma omp parallel for schedule(static)
            for (long p = 0; p < pieces; ++p) {
                const size_t o = (size_t)p * piece;
                memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, std::min(piece, bytes - o));
            }
        };
        int rc = MIROGPU_OK;
        size_t k = 0;
        size_t pend_off[2] = {0, 0}, pend_m[2] = {0, 0};   // chunk whose hits still sit in the slot's staging buffer
        for (size_t off = 0; off < n && rc == MIROGPU_OK; off += chunk, ++k) {
            BatchSlot* s = sl[k % nbuf];
            const size_t m = std::min(chunk, n - off);
            if (pageable) {
                if (pend_m[k % nbuf]) {   // the slot's previous chunk: wait for it, hand its hits to the caller
                    e = cudaStreamSynchronize(s->st);
                    if (e != cudaSuccess) { rc = fail(MIROGPU_ERR_CUDA, std::string("intersect_batch: ") + cudaGetErrorString(e)); break; }
                    par_copy(hits + pend_off[k % nbuf], s->pin_h, pend_m[k % nbuf] * sizeof(mirogpu_hit));
                    pend_m[k % nbuf] = 0;
                }
                par_copy(s->pin_r, rays + off, m * sizeof(mirogpu_ray));
                e = cudaMemcpyAsync(s->d_r, s->pin_r, m * sizeof(mirogpu_ray), cudaMemcpyHostToDevice, s->st);
                if (e == cudaSuccess) e = dispatch_trace(h, s->d_r, m, s->d_h, mode, s->st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(s->pin_h, s->d_h, m * sizeof(mirogpu_hit), cudaMemcpyDeviceToHost, s->st);
                pend_off[k % nbuf] = off; pend_m[k % nbuf] = m;
            } else {
                e = cudaMemcpyAsync(s->d_r, rays + off, m * sizeof(mirogpu_ray), cudaMemcpyHostToDevice, s->st);
                if (e == cudaSuccess) e = dispatch_trace(h, s->d_r, m, s->d_h, mode, s->st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(hits + off, s->d_h, m * sizeof(mirogpu_hit), cudaMemcpyDeviceToHost, s->st);
            }
            if (e != cudaSuccess) rc = fail(MIROGPU_ERR_CUDA, std::string("intersect_batch: ") + cudaGetErrorString(e));
            launches++;
        }
        for (int b = 0; b < nbuf; ++b) {
            e = cudaStreamSynchronize(sl[b]->st);
            if (e != cudaSuccess && rc == MIROGPU_OK) rc = fail(MIROGPU_ERR_CUDA, std::string("intersect_batch sync: ") + cudaGetErrorString(e));
            if (rc == MIROGPU_OK && pageable && pend_m[b]) par_copy(hits + pend_off[b], sl[b]->pin_h, pend_m[b] * sizeof(mirogpu_hit));
        }
        if (rc != MIROGPU_OK) return rc;
    }
    {
        std::lock_guard<std::mutex> lk(h->mtx);
        h->last_rays = n; h->last_launches = launches; h->stats_batches = 0;
    }
    return MIROGPU_OK;
}

int mirogpu_intersect_batch_counted(mirogpu_handle h, const mirogpu_ray* rays, size_t n, mirogpu_hit* hits, int mode,
                                    mirogpu_counters* c)
{
    if (!h || !c || (n && (!rays || !hits))) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if ((mode & ~MIROGPU_HINT_COHERENT) != MIROGPU_CLOSEST_HIT && (mode & ~MIROGPU_HINT_COHERENT) != MIROGPU_ANY_HIT)
        return fail(MIROGPU_ERR_INVALID_ARG, "unknown mode");
    if (n == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaError_t e;
    SlotGuard g{h, slot_acquire(h, e)};
    if (!g.s) return fail(cuda_code(e), std::string("intersect_batch_counted staging: ") + cudaGetErrorString(e));
    BatchSlot* s = g.s;
    unsigned long long hc[4] = {0, 0, 0, 0};
    e = s->ensure(n);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_c, 0, 4 * sizeof(unsigned long long), s->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_r, rays, n * sizeof(mirogpu_ray), cudaMemcpyHostToDevice, s->st);
    if (e == cudaSuccess) e = launch_simple<true>(h, s->d_r, n, s->d_h, (mode & 0xff) == MIROGPU_ANY_HIT, s->d_c, s->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hc, s->d_c, sizeof hc, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hits, s->d_h, n * sizeof(mirogpu_hit), cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(cuda_code(e), std::string("intersect_batch_counted: ") + cudaGetErrorString(e)); }
    const uint64_t node_size = h->layout == MIROGPU_LAYOUT_BVH2 ? 64 : h->layout == MIROGPU_LAYOUT_BVH4 ? 128 : h->layout == MIROGPU_LAYOUT_QBVH4 ? 64 : 80;
    c->rays += n; c->node_visits += hc[0]; c->box_tests += hc[1]; c->triangle_tests += hc[2]; c->hits += hc[3];
    c->bytes_fetched += hc[0] * node_size + hc[2] * sizeof(TriRecord);
    return MIROGPU_OK;
}

int mirogpu_resolve_hits_device(mirogpu_handle h, const mirogpu_hit* d_hits, size_t n, float* d_P, float* d_N,
                                uint32_t* d_material, void* cuda_stream)
{
    if (!h || (n && !d_hits)) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if (h->non_triangles) return fail(MIROGPU_ERR_INVALID_ARG, "the scene holds spheres / planes: their hit point is o + t d, use mirogpu_resolve_hits_rays_device");
    if (n == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    k_resolve_hits<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(h->ds, nullptr, d_hits, n, d_P, d_N, d_material);
    CUDA_TRY(cudaGetLastError());
    return MIROGPU_OK;
}

int mirogpu_resolve_hits_rays_device(mirogpu_handle h, const mirogpu_ray* d_rays, const mirogpu_hit* d_hits, size_t n, float* d_P,
                                     float* d_N, uint32_t* d_material, void* cuda_stream)
{
    if (!h || (n && (!d_hits || !d_rays))) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if (n == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    k_resolve_hits<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(h->ds, d_rays, d_hits, n, d_P, d_N, d_material);
    CUDA_TRY(cudaGetLastError());
    return MIROGPU_OK;
}

int mirogpu_generate_primary_device(mirogpu_handle h, const mirogpu_camera* cam, int width, int height, int row_begin,
                                    int row_end, int row_stride, int row_phase, int jitter, uint32_t seed, uint32_t sample_begin,
                                    uint32_t sample_count, mirogpu_ray* d_rays, void* cuda_stream)
{
    if (!h || !cam || !d_rays) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if (width <= 0 || height <= 0 || row_begin < 0 || row_end > height || row_begin > row_end || row_stride < 1 || row_phase < 0 ||
        row_phase >= row_stride)
        return fail(MIROGPU_ERR_INVALID_ARG, "bad image or row range");
    const int first_row = row_begin + row_phase;
    const int nrows = first_row < row_end ? (row_end - first_row + row_stride - 1) / row_stride : 0;
    const size_t total = (size_t)nrows * width * sample_count;
    if (total == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    CameraBasis cb;
    camera_basis(*cam, width, height, cb);
    const size_t npix = (size_t)nrows * width;
    if (npix >= (1ull << 31)) return fail(MIROGPU_ERR_UNSUPPORTED, "frame too large");
    for (uint32_t s0 = 0; s0 < sample_count; s0 += 65535u) {   // grid y carries the sample index
        const uint32_t ns = std::min(sample_count - s0, 65535u);
        k_gen_primary<<<dim3((unsigned)((npix + 255) / 256), ns), 256, 0, (cudaStream_t)cuda_stream>>>(cb, width, height, first_row, row_stride, nrows, jitter, seed,
                                                                                                     sample_begin + s0, ns, d_rays + (size_t)s0 * npix);
    }
    CUDA_TRY(cudaGetLastError());
    return MIROGPU_OK;
}

int mirogpu_generate_bounce_device(mirogpu_handle h, const mirogpu_ray* d_rays, const mirogpu_hit* d_hits, size_t n,
                                   uint32_t seed, uint32_t sample, uint32_t index_base, mirogpu_ray* d_out,
                                   unsigned long long* d_live_count, void* cuda_stream)
{
    if (!h || (n && (!d_rays || !d_hits || !d_out))) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if (n == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    static const int items = getenv("MIROGPU_GENB_ITEMS") ? atoi(getenv("MIROGPU_GENB_ITEMS")) : 4;   // tuning knob: hits per thread (1 / 2 / 3 / 4: 7.51 / 7.81 / 7.98 / 8.01 Grays/s on the bench)
    const cudaStream_t st = (cudaStream_t)cuda_stream;
#define MIRO_GENB(K)                                                                                                              \
    {                                                                                                                             \
        const size_t tile = (size_t)MIRO_GENB_THREADS * K;                                                                        \
        if (h->ds.textured) k_gen_bounce<K, true><<<(unsigned)((n + tile - 1) / tile), MIRO_GENB_THREADS, 0, st>>>(h->ds, d_rays, d_hits, n, seed, sample, \
                                                                                                              index_base, d_out, d_live_count); \
        else k_gen_bounce<K><<<(unsigned)((n + tile - 1) / tile), MIRO_GENB_THREADS, 0, st>>>(h->ds, d_rays, d_hits, n, seed, sample, \
                                                                                              index_base, d_out, d_live_count);      \
    }
    if (items == 1) MIRO_GENB(1) else if (items == 2) MIRO_GENB(2) else if (items == 3) MIRO_GENB(3) else MIRO_GENB(4)
#undef MIRO_GENB
    CUDA_TRY(cudaGetLastError());
    return MIROGPU_OK;
}

int mirogpu_rng_uniforms(uint32_t seed, uint32_t sample, uint32_t dimension, size_t first, size_t n, float* out)
{
    if (n && !out) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    for (size_t i = 0; i < n; ++i) uniform2(seed, (uint32_t)(first + i), sample, dimension, out[2 * i], out[2 * i + 1]);
    return MIROGPU_OK;
}

void mirogpu_release_build_scratch(void)
{
    BuildScratch& s = build_scratch();
    std::lock_guard<std::mutex> lk(s.mtx);
    s.release_locked();
}

void* mirogpu_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void mirogpu_host_free(void* p) { if (p) cudaFreeHost(p); }

int mirogpu_last_call_stats(mirogpu_handle h, uint64_t* rays_traced, uint64_t* kernel_launches)
{
    if (!h) return fail(MIROGPU_ERR_INVALID_ARG, "NULL handle");
    std::lock_guard<std::mutex> lk(h->mtx);
    uint64_t rays = h->last_rays;
    // secondary waves of the last render batches were counted on the device (valid after the stream was synchronised)
    if (h->stats_batches) { uint64_t secondary; memcpy(&secondary, h->h_stats, 8); rays += secondary * h->stats_mult; }
    if (rays_traced) *rays_traced = rays;
    if (kernel_launches) *kernel_launches = h->last_launches;
    return MIROGPU_OK;
}

int mirogpu_render(mirogpu_handle h, const mirogpu_camera* cam, const mirogpu_render_params* p, float* rgb_out)
{
    if (!h || !cam || !p || !rgb_out) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    std::string err;
    const int rc = render_host(h, *cam, *p, rgb_out, nullptr, err);
    return rc == MIROGPU_OK ? rc : fail(rc, err);
}

int mirogpu_render_rgb8(mirogpu_handle h, const mirogpu_camera* cam, const mirogpu_render_params* p, uint8_t* rgb8_out)
{
    if (!h || !cam || !p || !rgb8_out) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    std::string err;
    const int rc = render_host(h, *cam, *p, nullptr, rgb8_out, err);
    return rc == MIROGPU_OK ? rc : fail(rc, err);
}

int mirogpu_render_device(mirogpu_handle h, const mirogpu_camera* cam, const mirogpu_render_params* p, float* d_rgb,
                          void* cuda_stream)
{
    if (!h || !cam || !p || !d_rgb) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    std::string err;
    const int rc = (!h->replicas.empty() && p->row_stride == 1) ? render_multi(h, *cam, *p, nullptr, nullptr, d_rgb, (cudaStream_t)cuda_stream, err)
                                                                 : render_device(h, *cam, *p, d_rgb, nullptr, (cudaStream_t)cuda_stream, err);
    return rc == MIROGPU_OK ? rc : fail(rc, err);
}

int mirogpu_tonemap_rgb8_device(mirogpu_handle h, const float* d_rgb, int width, int height, uint8_t* d_rgb8, void* cuda_stream)
{
    if (!h || !d_rgb || !d_rgb8 || width <= 0 || height <= 0) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    std::lock_guard<std::mutex> lk(h->mtx);
    CUDA_TRY(h->scratch.ensure(15, 64));
    float* gmax = reinterpret_cast<float*>(h->scratch.buf[15]);
    const float ninf = -INFINITY;
    CUDA_TRY(cudaMemcpyAsync(gmax, &ninf, 4, cudaMemcpyHostToDevice, st));
    const size_t nvals = (size_t)width * height * 3;
    k_frame_max<<<(unsigned)((nvals + 255) / 256), 256, 0, st>>>(d_rgb, nvals, gmax);
    k_tonemap<<<(unsigned)((nvals + 255) / 256), 256, 0, st>>>(const_cast<float*>(d_rgb), d_rgb8, width, 0, 1, height, gmax);
    CUDA_TRY(cudaGetLastError());
    return MIROGPU_OK;
}

static bool rows_ok(int width, int height, int row_begin, int row_end, int row_stride, int row_phase)
{
    return width > 0 && height > 0 && row_stride >= 1 && row_phase >= 0 && row_phase < row_stride && row_begin >= 0 && row_end <= height && row_begin <= row_end;
}

int mirogpu_frame_max_device(mirogpu_handle h, const float* d_rgb, int width, int height, int row_begin, int row_end, int row_stride,
                             int row_phase, float* d_max, void* cuda_stream)
{
    if (!h || !d_rgb || !d_max || !rows_ok(width, height, row_begin, row_end, row_stride, row_phase)) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const float ninf = -INFINITY;
    CUDA_TRY(cudaMemcpyAsync(d_max, &ninf, 4, cudaMemcpyHostToDevice, st));
    const int first_row = row_begin + row_phase;
    const int nrows = first_row < row_end ? (row_end - first_row + row_stride - 1) / row_stride : 0;
    const size_t nvals = (size_t)nrows * width * 3;
    if (nvals) k_frame_max_rows<<<(unsigned)((nvals + 255) / 256), 256, 0, st>>>(d_rgb, width, first_row, row_stride, nrows, d_max);
    CUDA_TRY(cudaGetLastError());
    return MIROGPU_OK;
}

int mirogpu_tonemap_rows_rgb8_device(mirogpu_handle h, const float* d_rgb, int width, int height, int row_begin, int row_end,
                                     int row_stride, int row_phase, const float* d_max, uint8_t* d_rgb8, void* cuda_stream)
{
    if (!h || !d_rgb || !d_max || !d_rgb8 || !rows_ok(width, height, row_begin, row_end, row_stride, row_phase)) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const int first_row = row_begin + row_phase;
    const int nrows = first_row < row_end ? (row_end - first_row + row_stride - 1) / row_stride : 0;
    const size_t nvals = (size_t)nrows * width * 3;
    if (nvals) k_tonemap<<<(unsigned)((nvals + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(const_cast<float*>(d_rgb), d_rgb8, width, first_row, row_stride, nrows, d_max);
    CUDA_TRY(cudaGetLastError());
    return MIROGPU_OK;
}

namespace {
// k_photon_trace for `count` emissions of one DirectionalAreaLight into device buffers (counts: count bytes, records:
// count * 45 floats; only the first counts[i] records of an emission are written).
int photon_trace_launch(mirogpu_scene* h, int light_index, int caustic, uint32_t seed, uint64_t first_emission, uint32_t count,
                        unsigned char* d_counts, float* d_records, cudaStream_t st)
{
    if (light_index < 0 || (size_t)light_index >= h->h_lights.size()) return fail(MIROGPU_ERR_INVALID_ARG, "light index out of range");
    const mirogpu_light& L = h->h_lights[(size_t)light_index];
    if (L.kind != 1) return fail(MIROGPU_ERR_INVALID_ARG, "photons are emitted from DirectionalAreaLights only (Scene.cpp:368)");
    PhotonEmitter em; memset(&em, 0, sizeof em);
    // getTangents (Utility.h:25-31), as SquareLight::preCalc calls it on the light normal
    const float n[3] = {L.normal[0], L.normal[1], L.normal[2]};
    float t1[3] = {0.f * n[2] - 1.f * n[1], 1.f * n[0] - 0.f * n[2], 0.f * n[1] - 0.f * n[0]};                 // cross((0,0,1), n)
    if ((double)(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]) < 1e-6) {
        t1[0] = 1.f * n[2] - 0.f * n[1]; t1[1] = 0.f * n[0] - 0.f * n[2]; t1[2] = 0.f * n[1] - 1.f * n[0];        // cross((0,1,0), n)
    }
    const float t2[3] = {t1[1] * n[2] - t1[2] * n[1], t1[2] * n[0] - t1[0] * n[2], t1[0] * n[1] - t1[1] * n[0]};   // cross(t1, n)
    // power = color * wattage, then *= PI r r (/ 10.f for the caustic pass): Scene.cpp:379-385, 431-434
    const float pi = 3.1415926535897932384626433832795028841972f;
    const float area = caustic ? pi * L.radius * L.radius / 10.f : pi * L.radius * L.radius;
    for (int k = 0; k < 3; ++k) {
        em.pos[k] = L.position[k]; em.normal[k] = n[k]; em.t1[k] = t1[k]; em.t2[k] = t2[k];
        em.power[k] = L.color[k] * L.wattage * area;
    }
    em.radius = L.radius; em.caustic = caustic ? 1 : 0; em.seed = seed; em.first = first_emission; em.count = count;
    const unsigned grid = (count + 127) / 128;
#define MIRO_PT(L)                                                                                                      \
    {                                                                                                                       \
        if (h->non_triangles) k_photon_trace<L, true><<<grid, 128, 0, st>>>(h->ds, h->d_materials, em, d_counts, d_records);       \
        else k_photon_trace<L><<<grid, 128, 0, st>>>(h->ds, h->d_materials, em, d_counts, d_records);                              \
    }
    if (h->layout == MIROGPU_LAYOUT_BVH2) MIRO_PT(MIROGPU_LAYOUT_BVH2)
    else if (h->layout == MIROGPU_LAYOUT_BVH4) MIRO_PT(MIROGPU_LAYOUT_BVH4)
    else if (h->layout == MIROGPU_LAYOUT_QBVH4) MIRO_PT(MIROGPU_LAYOUT_QBVH4)
    else MIRO_PT(MIROGPU_LAYOUT_CWBVH8)
#undef MIRO_PT
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MIROGPU_ERR_CUDA, std::string("photon_trace: ") + cudaGetErrorString(e));
    return MIROGPU_OK;
}
}  // namespace

int mirogpu_photon_trace(mirogpu_handle h, int light_index, int caustic, uint32_t seed, uint64_t first_emission, uint32_t count,
                         uint8_t* counts, float* records)
{
    if (!h || (count && (!counts || !records))) return fail(MIROGPU_ERR_INVALID_ARG, "NULL argument");
    if (light_index < 0 || (size_t)light_index >= h->h_lights.size()) return fail(MIROGPU_ERR_INVALID_ARG, "light index out of range");
    if (h->h_lights[(size_t)light_index].kind != 1) return fail(MIROGPU_ERR_INVALID_ARG, "photons are emitted from DirectionalAreaLights only (Scene.cpp:368)");
    if (count == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    DevBuf bc, br;
    const size_t rec_bytes = (size_t)count * 9 * MIRO_PHOTON_MAX_RECORDS * sizeof(float);
    cudaError_t e = bc.alloc(count);
    if (e == cudaSuccess) e = br.alloc(rec_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(br.p, 0, rec_bytes, cudaStreamPerThread);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(cuda_code(e), std::string("photon_trace staging: ") + cudaGetErrorString(e)); }
    const int rc = photon_trace_launch(h, light_index, caustic, seed, first_emission, count, bc.as<unsigned char>(), br.as<float>(), cudaStreamPerThread);
    if (rc != MIROGPU_OK) return rc;
    e = cudaMemcpyAsync(counts, bc.p, count, cudaMemcpyDeviceToHost, cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaMemcpyAsync(records, br.p, rec_bytes, cudaMemcpyDeviceToHost, cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
    if (e != cudaSuccess) return fail(MIROGPU_ERR_CUDA, std::string("photon_trace: ") + cudaGetErrorString(e));
    {
        std::lock_guard<std::mutex> lk(h->mtx);
        h->last_rays = 0; h->last_launches = 1; h->stats_batches = 0;
    }
    return MIROGPU_OK;
}

// Scene::tracePhotons / traceCausticPhotons (Scene.cpp:351-472) without leaving the device: emit in batches, store in
// emission order under the reference's stop rule, scale by 1 / emissions, balance, lay out the gather records.
int mirogpu_photon_pass(mirogpu_handle h, int which, int caustic, uint32_t seed, int target, long long max_emissions,
                        long long* emissions_out, int* stored_out)
{
    if (!h || which < 0 || which > 1 || target < 0) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    if (max_emissions <= 0) max_emissions = 1ll << 28;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = cudaStreamPerThread;
    long long total = 0;
    int stored = 0;
    const bool timing = getenv("MIROGPU_TIMING") != nullptr;
    const double t_begin = now_s();
    double t_trace_store = 0.0, t_alloc = 0.0;
    PhotonBuild b;
    CUDA_TRY(b.alloc(target + 8));
    DevBuf bc, br;
    uint32_t buf_cap = 0;
    for (size_t l = 0; l < h->h_lights.size() && target > 0; ++l) {
        if (h->h_lights[l].kind != 1) continue;                // "Temporary hack" of the reference: only directional area lights emit
        uint64_t next = 0;                                     // emission index of this light = its random stream
        uint32_t batch = 65536;
        while (stored < target && total < max_emissions) {
            if (batch > buf_cap) {
                const double t0 = now_s();
                cudaFree(bc.p); cudaFree(br.p); bc.p = br.p = nullptr;
                CUDA_TRY(bc.alloc(batch));
                CUDA_TRY(br.alloc((size_t)batch * 9 * MIRO_PHOTON_MAX_RECORDS * sizeof(float)));
                buf_cap = batch;
                t_alloc += now_s() - t0;
            }
            const double t1 = now_s();
            const int rc = photon_trace_launch(h, (int)l, caustic, seed, ((uint64_t)l << 40) | next, batch, bc.as<unsigned char>(), br.as<float>(), st);
            if (rc != MIROGPU_OK) return rc;
            int used = 0, after = stored;
            CUDA_TRY(b.store_batch(bc.as<unsigned char>(), br.as<float>(), batch, stored, target, st, &used, &after));
            const int before = stored;
            stored = after; total += used; next += (uint64_t)used;
            t_trace_store += now_s() - t1;
            if (timing) fprintf(stderr, "mirogpu_photon_pass: light %zu batch %u used %d stored %d\n", l, batch, used, stored);
            // size the next batch from the observed yield (photons per emission), with some slack
            const double yield = std::max(1e-4, (double)(stored - before) / (double)std::max(used, 1));
            const double want = (double)(target - stored) / yield * 1.05 + 1024.0;
            batch = (uint32_t)std::min(4194304.0, std::max(4096.0, want));
        }
    }
    if (total > 0) CUDA_TRY(b.scale(1, stored, 1.0f / (float)total, st));   // Scene.cpp:400
    std::string err;
    int rc;
    {
        std::lock_guard<std::mutex> lk(h->mtx);
        rc = h->pm[which].build_from_store(b, stored, true, st, err);
        h->last_rays = 0; h->last_launches = 1; h->stats_batches = 0;
    }
    if (rc != MIROGPU_OK) return fail(rc, err);
    if (timing) fprintf(stderr, "mirogpu_photon_pass: total %.2f ms (buffers %.2f, trace + store %.2f, scale + balance + records %.2f)\n",
                        (now_s() - t_begin) * 1e3, t_alloc * 1e3, t_trace_store * 1e3, (now_s() - t_begin - t_alloc - t_trace_store) * 1e3);
    for (mirogpu_scene* r : h->replicas) {
        CUDA_TRY(cudaSetDevice(r->device));
        std::lock_guard<std::mutex> lk(r->mtx);
        rc = r->pm[which].clone_from(h->pm[which], h->device, r->device, err);
        if (rc != MIROGPU_OK) { cudaSetDevice(h->device); return fail(rc, err); }
    }
    CUDA_TRY(cudaSetDevice(h->device));
    if (emissions_out) *emissions_out = total;
    if (stored_out) *stored_out = stored;
    return MIROGPU_OK;
}

int mirogpu_photon_download(mirogpu_handle h, int which, void* photons, int capacity, int* stored)
{
    if (!h || which < 0 || which > 1) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> lk(h->mtx);
    const int n = h->pm[which].stored;
    if (stored) *stored = n;
    if (!photons) return MIROGPU_OK;                           // size query
    if (capacity < n) return fail(MIROGPU_ERR_INVALID_ARG, "photon_download: capacity below the stored count");
    CUDA_TRY(h->pm[which].export28(photons, cudaStreamPerThread));
    return MIROGPU_OK;
}

int mirogpu_photon_balance(int device, void* photons, int stored, const float bbox_min[3], const float bbox_max[3])
{
    if (stored < 0 || (stored && (!photons || !bbox_min || !bbox_max))) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(photon_balance_host_array(photons, stored, bbox_min, bbox_max, cudaStreamPerThread));
    return MIROGPU_OK;
}

int mirogpu_photon_upload(mirogpu_handle h, int which, const void* photons, int stored)
{
    if (!h || which < 0 || which > 1 || stored < 0 || (stored && !photons)) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    for (mirogpu_scene* r : h->replicas) { const int rc = mirogpu_photon_upload(r, which, photons, stored); if (rc != MIROGPU_OK) return rc; }
    CUDA_TRY(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> lk(h->mtx);
    std::string err;
    const int rc = h->pm[which].upload(photons, stored, err);
    return rc == MIROGPU_OK ? rc : fail(rc, err);
}

int mirogpu_texture_lookup(int kind, const float* tex, float u, float v, float w, float* rgb3)
{
    if (!tex || !rgb3 || kind <= MIROGPU_TEX_NONE || kind > MIROGPU_TEX_FLOWER_CENTER) return fail(MIROGPU_ERR_INVALID_ARG, "bad texture query");
    mirogpu_material m; memset(&m, 0, sizeof m);
    m.texture = kind; memcpy(m.tex, tex, sizeof m.tex);
    const float uv[2] = {u, v}, P[3] = {u, v, w};
    material_diffuse_color(m, uv, P, rgb3);
    return MIROGPU_OK;
}

int mirogpu_texture_bump(int kind, const float* tex, float u, float v, float* height)
{
    if (!tex || !height || kind <= MIROGPU_TEX_NONE || kind > MIROGPU_TEX_FLOWER_CENTER) return fail(MIROGPU_ERR_INVALID_ARG, "bad texture query");
    mirogpu_material m; memset(&m, 0, sizeof m);
    m.texture = kind; memcpy(m.tex, tex, sizeof m.tex);
    *height = material_bump_height(m, u, v);
    return MIROGPU_OK;
}

int mirogpu_photon_set_exact(mirogpu_handle h, int which, int exact)
{
    if (!h || which < 0 || which > 1) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    for (mirogpu_scene* r : h->replicas) mirogpu_photon_set_exact(r, which, exact);
    std::lock_guard<std::mutex> lk(h->mtx);
    h->pm[which].exact = exact != 0;
    return MIROGPU_OK;
}

int mirogpu_photon_gather_device(mirogpu_handle h, int which, const float* d_pos3, const float* d_normal3, size_t n,
                                 float max_dist, int k, float* d_irrad3, void* cuda_stream)
{
    if (!h || which < 0 || which > 1 || (n && (!d_pos3 || !d_normal3 || !d_irrad3))) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    if (k < 1 || k > MIRO_PHOTON_KMAX) return fail(MIROGPU_ERR_INVALID_ARG, "k out of range (1..512)");
    if (n == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(photon_gather_launch(h->pm[which], d_pos3, d_normal3, n, max_dist, k, d_irrad3, (cudaStream_t)cuda_stream));
    return MIROGPU_OK;
}

int mirogpu_photon_gather(mirogpu_handle h, int which, const float* pos3, const float* normal3, size_t n, float max_dist,
                          int k, float* irrad3)
{
    if (!h || which < 0 || which > 1 || (n && (!pos3 || !normal3 || !irrad3))) return fail(MIROGPU_ERR_INVALID_ARG, "bad argument");
    if (k < 1 || k > MIRO_PHOTON_KMAX) return fail(MIROGPU_ERR_INVALID_ARG, "k out of range (1..512)");
    if (n == 0) return MIROGPU_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    if (n <= MIRO_SMALL_BATCH) {
        // The reference's own call pattern is ONE query per irradiance_estimate call (Scene.cpp:292-293): no allocation and no
        // staging copies on that path -- the queries sit in the slot's page-locked, device-visible buffers and the kernel reads
        // them over the bus (32 KB hold 1024 positions + normals, 16 KB the irradiances).
        cudaError_t es;
        SlotGuard g{h, slot_acquire(h, es)};
        if (!g.s) return fail(cuda_code(es), std::string("photon_gather staging: ") + cudaGetErrorString(es));
        float* hp = reinterpret_cast<float*>(g.s->h_r);
        float* hn = hp + 3 * MIRO_SMALL_BATCH;
        float* hi = reinterpret_cast<float*>(g.s->h_h);
        memcpy(hp, pos3, n * 12); memcpy(hn, normal3, n * 12);
        es = photon_gather_launch(h->pm[which], hp, hn, n, max_dist, k, hi, g.s->st);
        if (es == cudaSuccess) es = cudaStreamSynchronize(g.s->st);
        if (es != cudaSuccess) return fail(MIROGPU_ERR_CUDA, std::string("photon_gather: ") + cudaGetErrorString(es));
        memcpy(irrad3, hi, n * 12);
        return MIROGPU_OK;
    }
    DevBuf bp, bn, bi;
    cudaError_t e = bp.alloc(n * 12);
    if (e == cudaSuccess) e = bn.alloc(n * 12);
    if (e == cudaSuccess) e = bi.alloc(n * 12);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(cuda_code(e), std::string("photon_gather staging: ") + cudaGetErrorString(e)); }
    float *d_p = bp.as<float>(), *d_n = bn.as<float>(), *d_i = bi.as<float>();
    e = cudaMemcpyAsync(d_p, pos3, n * 12, cudaMemcpyHostToDevice, cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, normal3, n * 12, cudaMemcpyHostToDevice, cudaStreamPerThread);
    if (e == cudaSuccess) e = photon_gather_launch(h->pm[which], d_p, d_n, n, max_dist, k, d_i, cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaMemcpyAsync(irrad3, d_i, n * 12, cudaMemcpyDeviceToHost, cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
    if (e != cudaSuccess) return fail(MIROGPU_ERR_CUDA, std::string("photon_gather: ") + cudaGetErrorString(e));
    return MIROGPU_OK;
}

}  // extern "C"
