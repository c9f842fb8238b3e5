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

#include <cmath>
#include <vector>

namespace mirogpu {

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
    if (e == cudaSuccess) e = cudaMemcpy(d_photons, packed.data(), packed.size() * sizeof(float4), cudaMemcpyHostToDevice);
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

cudaError_t photon_gather_launch(const PhotonMapDevice& pm, const float* d_pos3, const float* d_normal3, size_t n, float max_dist,
                                 int k, float* d_irrad3, cudaStream_t st, const float4* active, const uint32_t* d_n)
{
    if (n == 0) return cudaSuccess;
    if (!pm.d_photons) return cudaMemsetAsync(d_irrad3, 0, n * 12, st);   // empty map: zero irradiance, like a map with no photons
    k_photon_gather<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pm.d_photons, pm.d_tables, pm.stored, pm.half_stored, d_pos3, d_normal3,
                                                                  active, n, d_n, max_dist, k, d_irrad3);
    return cudaGetLastError();
}

}  // namespace mirogpu
#endif
