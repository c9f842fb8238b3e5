// photon_build.cu -- the photon pass behind the walks, on the device: Photon_map::store with the
// reference's sequential stop rule, scale_photon_power and balance (reference PhotonMap.cpp:246-466), then the gather
// records of photon_impl.cuh -- emitted photons never leave HBM between k_photon_trace and the kNN gather.
//
// balance() must end in the SAME heap array as the reference's, and that array depends on more than the photon
// positions: Jensen's median_split is a Hoare quickselect whose outcome among equal keys (every photon on an
// axis-aligned wall shares a coordinate) follows from the order the two scan pointers meet the elements.  A sort-based
// kd build gets the split planes right and the tie-breaks wrong.  So the quickselect itself runs here, round for round,
// but each partition round is evaluated in parallel from a closed form of what the two pointers do:
//
//   round on [left, right], pivot v = key[right]:  G = positions in [left, right) whose key is not < v, ascending;
//   S = positions in [left, right) whose key is not > v, descending.  The left pointer stops exactly at g_1, g_2, ...,
//   the right pointer at s_1, s_2, ... (between the pointers the array is still the original, so the stops are decided
//   by the original keys), the k-th exchange swaps g_k with s_k, and exchanges stop at the first k with g_k >= s_k.
//   With K = #{k : g_k < s_k} the left pointer ends at i = min(g_{K+1}, s_K) (s_0 = right: the pivot is its own
//   sentinel; after a swap s_K holds a key >= v), and the pivot is swapped into i.  (The `j > left` guard of the
//   right pointer only ever fires when no exchange is left, PhotonMap.cpp:385-386.)
//
// G / S ranks are two prefix counts over the range, K is a search on a monotone predicate, the K exchanges are
// independent -- one pass of a thread group per round.  Large segments are processed by one 1024-thread CTA each, level
// by level (k_pb_level); a segment of <= PB_SMALL photons is finished, whole subtree, by one warp in shared memory
// (k_pb_small).  tests/test_photon_build_model.py replays this closed form in numpy against the oracle's balance();
// tests/test_gpu_photon_build.py compares the device array with the oracle's bit for bit (ties included).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "photon.cuh"

namespace mirogpu {

#define PB_SMALL 512
#define PB_WARPS 4

struct PbSeg { int start, end; float lo[3], hi[3]; };   // an open segment of the balance: positions [start, end] (1-based), narrowed box

namespace {

__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu)); }

// ---- store (PhotonMap.cpp:246-288) with the stop rule of Scene::tracePhotons (Scene.cpp:370-377) ----------------------------
// Emission i of the batch is consumed iff fewer than `target` photons were stored before it, in emission order; a consumed
// emission stores all of its records (so a map ends with target .. target + 4 photons, as the reference's does).
__global__ void __launch_bounds__(256) k_photon_store(const unsigned char* __restrict__ counts, const uint32_t* __restrict__ excl,
                                                      const float* __restrict__ records, uint32_t batch, int base, int target,
                                                      float4* __restrict__ pos, float4* __restrict__ pow, uint32_t* __restrict__ bbox,
                                                      int* __restrict__ ctl)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    if (i < batch) {
        const int before = base + (int)excl[i], c = counts[i];
        if (before < target) {
            const float* r = records + (size_t)i * 45;
            for (int j = 0; j < c; ++j) {
                const float* q = r + 9 * j;
                // direction quantised to two bytes (PhotonMap.cpp:275-287): double acos / atan2 of the binary32 components
                const int theta = int(acos((double)q[8]) * (256.0 / 3.14159265358979323846));
                const int phi = int(atan2((double)q[7], (double)q[6]) * (256.0 / (2.0 * 3.14159265358979323846)));
                const uint32_t tb = theta > 255 ? 255u : (uint32_t)(unsigned char)theta;
                const uint32_t pb = phi > 255 ? 255u : (phi < 0 ? (uint32_t)(unsigned char)(phi + 256) : (uint32_t)(unsigned char)phi);
                const size_t slot = (size_t)before + j + 1;
                pos[slot] = make_float4(q[3], q[4], q[5], __uint_as_float((tb << 16) | (pb << 24)));
                pow[slot] = make_float4(q[0], q[1], q[2], 0.f);
#pragma unroll
                for (int k = 0; k < 3; ++k) { const uint32_t o = f2ord(q[3 + k]); lo[k] = min(lo[k], o); hi[k] = max(hi[k], o); }
            }
            if (i + 1 == batch || base + (int)excl[i + 1] >= target) { ctl[0] = (int)i + 1; ctl[1] = before + c; }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t l = __reduce_min_sync(0xffffffffu, lo[k]), h = __reduce_max_sync(0xffffffffu, hi[k]);
        if ((threadIdx.x & 31) == 0) {
            if (l != 0xffffffffu) atomicMin(bbox + k, l);
            if (h != 0u) atomicMax(bbox + 3 + k, h);
        }
    }
}

__global__ void k_pb_bbox_init(uint32_t* bbox)
{
    // Photon_map::Photon_map (PhotonMap.cpp:36-39): min = 1e8, max = -1e8
    if (threadIdx.x < 3) bbox[threadIdx.x] = f2ord(1e8f);
    else if (threadIdx.x < 6) bbox[threadIdx.x] = f2ord(-1e8f);
}

__global__ void k_photon_scale(float4* __restrict__ pow, int first, int last, float scale)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > last) return;
    float4 p = pow[i];
    p.x = __fmul_rn(p.x, scale); p.y = __fmul_rn(p.y, scale); p.z = __fmul_rn(p.z, scale);
    pow[i] = p;
}

// ---- balance ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pb_median(int start, int end)   // PhotonMap.cpp:416-425
{
    const int count = end - start + 1;
    int median = 1;
    while (4 * median <= count) median += median;
    if (3 * median <= count) { median += median; median += start - 1; }
    else median = end - median + 1;
    return median;
}

__device__ __forceinline__ int pb_axis(const float* lo, const float* hi)   // PhotonMap.cpp:431-436
{
    const float ex = __fsub_rn(hi[0], lo[0]), ey = __fsub_rn(hi[1], lo[1]), ez = __fsub_rn(hi[2], lo[2]);
    if (ex > ey && ex > ez) return 0;
    if (ey > ez) return 1;
    return 2;
}

template <int NT> __device__ __forceinline__ void group_sync() { if (NT == 32) __syncwarp(); else __syncthreads(); }

// Exclusive ranks of two flags over a group of NT threads (NT = 32: a warp; NT = 1024: a CTA of 32 warps, sh = 64 words).
template <int NT>
__device__ __forceinline__ void flag_ranks(bool g, bool s, int tid, uint32_t* sh, int& rg, int& rs, int& tg, int& ts)
{
    const unsigned bg = __ballot_sync(0xffffffffu, g), bs = __ballot_sync(0xffffffffu, s);
    const int lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    rg = __popc(bg & lt); rs = __popc(bs & lt);
    if (NT == 32) { tg = __popc(bg); ts = __popc(bs); return; }
    const int w = tid >> 5;
    if (lane == 0) { sh[w] = __popc(bg); sh[32 + w] = __popc(bs); }
    __syncthreads();
    unsigned cg = sh[lane], cs = sh[32 + lane];
    const unsigned mine_g = sh[w], mine_s = sh[32 + w];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, cg, o), y = __shfl_up_sync(0xffffffffu, cs, o);
        if (lane >= o) { cg += x; cs += y; }
    }
    tg = (int)__shfl_sync(0xffffffffu, cg, 31); ts = (int)__shfl_sync(0xffffffffu, cs, 31);
    rg += (int)(__shfl_sync(0xffffffffu, cg, w) - mine_g);
    rs += (int)(__shfl_sync(0xffffffffu, cs, w) - mine_s);
    __syncthreads();
}

// One partition round of median_split (PhotonMap.cpp:380-394) on [left, right]; returns where the pivot lands.
// key(q): key at position q; swp(a, b): exchange positions a and b; posG / posS: scratch indexed by position.
template <int NT, typename IDX, typename KEY, typename SWP>
__device__ __forceinline__ int hoare_round(int left, int right, KEY key, SWP swp, IDX* posG, IDX* posS, int tid, uint32_t* sh)
{
    const float v = key(right);
    int cg = 0, cs = 0;
    for (int base = left; base < right; base += NT) {
        const int q = base + tid;
        const bool in = q < right;
        const float kq = in ? key(q) : 0.f;
        const bool g = in && !(kq < v), s = in && !(kq > v);
        int rg, rs, tg, ts;
        flag_ranks<NT>(g, s, tid, sh, rg, rs, tg, ts);
        if (g) posG[left + cg + rg] = (IDX)q;
        if (s) posS[left + cs + rs] = (IDX)q;
        cg += tg; cs += ts;
    }
    group_sync<NT>();
    // g_k = posG[left + k - 1], s_k = posS[left + cs - k];  K = #{k >= 1 : g_k < s_k}, the predicate is monotone in k
    const int lane = tid & 31;
    int lo = 0, hi = min(cg, cs);
    while (lo < hi) {
        const int step = (hi - lo + 31) >> 5;
        const int k = min(lo + (lane + 1) * step, hi);
        const bool ok = (int)posG[left + k - 1] < (int)posS[left + cs - k];
        const int c = __popc(__ballot_sync(0xffffffffu, ok));
        const int nlo = min(lo + c * step, hi);
        const int nhi = c == 32 ? hi : min(lo + (c + 1) * step, hi) - 1;
        lo = nlo; hi = max(nhi, nlo);
    }
    const int K = lo;
    for (int k = 1 + tid; k <= K; k += NT) swp((int)posG[left + k - 1], (int)posS[left + cs - k]);
    const int gi = K + 1 <= cg ? (int)posG[left + K] : INT_MAX;
    const int si = K >= 1 ? (int)posS[left + cs - K] : right;
    const int i = min(gi, si);
    group_sync<NT>();
    if (tid == 0 && i != right) swp(i, right);
    group_sync<NT>();
    return i;
}

__global__ void k_pb_init(int n, const uint32_t* __restrict__ bbox, float3 lo, float3 hi, uint32_t* __restrict__ ids, PbSeg* __restrict__ segs,
                          uint32_t* __restrict__ heap, uint32_t* __restrict__ small_list, uint32_t* __restrict__ small_count)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= n) ids[q] = (uint32_t)q;
    if (q == 0) {
        *small_count = 0;
        heap[0] = 0;
        if (n == 1) heap[1] = 1;   // "if (stored_photons > 1)" (PhotonMap.cpp:321): a single photon stays where it is
        if (n > 1) {
            PbSeg s; s.start = 1; s.end = n;
            if (bbox) { for (int k = 0; k < 3; ++k) { s.lo[k] = ord2f(bbox[k]); s.hi[k] = ord2f(bbox[3 + k]); } }
            else { s.lo[0] = lo.x; s.lo[1] = lo.y; s.lo[2] = lo.z; s.hi[0] = hi.x; s.hi[1] = hi.y; s.hi[2] = hi.z; }
            segs[1] = s;
            if (n <= PB_SMALL) { small_list[0] = 1; *small_count = 1; }
        }
    }
}

__device__ __forceinline__ float comp(const float4& p, int axis) { return axis == 0 ? p.x : (axis == 1 ? p.y : p.z); }

// The node's children (PhotonMap.cpp:449-466): a segment of one photon is placed, a longer one is opened with the box
// narrowed at the split plane; small open segments go to the list k_pb_small works off.
__device__ __forceinline__ void pb_open_child(uint32_t child, int start, int end, const PbSeg& parent, int axis, float split, bool left_side,
                                              PbSeg* segs, uint32_t* small_list, uint32_t* small_count)
{
    PbSeg c = parent;
    c.start = start; c.end = end;
    if (left_side) c.hi[axis] = split; else c.lo[axis] = split;
    segs[child] = c;
    if (end - start + 1 <= PB_SMALL) small_list[atomicAdd(small_count, 1u)] = child;
}

__global__ void __launch_bounds__(1024) k_pb_level(int level, int n, const float4* __restrict__ pos, uint32_t* __restrict__ ids,
                                                   float* __restrict__ keys, uint32_t* __restrict__ posG, uint32_t* __restrict__ posS,
                                                   PbSeg* __restrict__ segs, uint32_t* __restrict__ heap, uint32_t* __restrict__ small_list,
                                                   uint32_t* __restrict__ small_count)
{
    __shared__ uint32_t sh[64];
    const uint32_t index = (1u << level) + blockIdx.x;
    if (index > (uint32_t)n) return;
    const PbSeg sg = segs[index];
    if (sg.start == 0 || sg.end - sg.start + 1 <= PB_SMALL) return;   // not open at this level / listed for k_pb_small
    const int tid = threadIdx.x;
    const int start = sg.start, end = sg.end, median = pb_median(start, end), axis = pb_axis(sg.lo, sg.hi);
    for (int q = start + tid; q <= end; q += 1024) keys[q] = comp(pos[ids[q]], axis);
    __syncthreads();
    auto key = [&](int q) { return keys[q]; };
    auto swp = [&](int a, int b) {
        const uint32_t ia = ids[a], ib = ids[b]; ids[a] = ib; ids[b] = ia;
        const float ka = keys[a], kb = keys[b]; keys[a] = kb; keys[b] = ka;
    };
    int left = start, right = end;
    while (right > left) {
        const int i = hoare_round<1024, uint32_t>(left, right, key, swp, posG, posS, tid, sh);
        if (i >= median) right = i - 1;
        if (i <= median) left = i + 1;
    }
    if (tid == 0) {
        const uint32_t id = ids[median];
        const float split = keys[median];
        heap[index] = id | ((uint32_t)axis << 30);
        if (median > start) {
            if (start < median - 1) pb_open_child(2 * index, start, median - 1, sg, axis, split, true, segs, small_list, small_count);
            else heap[2 * index] = ids[start];
        }
        if (median < end) {
            if (median + 1 < end) pb_open_child(2 * index + 1, median + 1, end, sg, axis, split, false, segs, small_list, small_count);
            else heap[2 * index + 1] = ids[end];
        }
    }
}

__global__ void __launch_bounds__(32 * PB_WARPS) k_pb_small(const float4* __restrict__ pos, const uint32_t* __restrict__ ids,
                                                            const PbSeg* __restrict__ segs, uint32_t* __restrict__ heap,
                                                            const uint32_t* __restrict__ small_list, const uint32_t* __restrict__ small_count)
{
    __shared__ float sk[PB_WARPS][3][PB_SMALL];
    __shared__ uint32_t sgid[PB_WARPS][PB_SMALL];
    __shared__ uint16_t sperm[PB_WARPS][PB_SMALL], sG[PB_WARPS][PB_SMALL], sS[PB_WARPS][PB_SMALL];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t task = blockIdx.x * PB_WARPS + w;
    if (task >= *small_count) return;
    const uint32_t root = small_list[task];
    const PbSeg sg = segs[root];
    const int count = sg.end - sg.start + 1;
    for (int s = lane; s < count; s += 32) {
        const uint32_t id = ids[sg.start + s];
        const float4 p = pos[id];
        sk[w][0][s] = p.x; sk[w][1][s] = p.y; sk[w][2][s] = p.z;
        sgid[w][s] = id; sperm[w][s] = (uint16_t)s;
    }
    __syncwarp();
    // explicit stack of open segments (local positions); every lane holds the same copy
    uint32_t st_index[16]; int st_start[16], st_end[16]; float st_lo[16][3], st_hi[16][3];
    int sp = 0;
    st_index[0] = root; st_start[0] = 0; st_end[0] = count - 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) { st_lo[0][k] = sg.lo[k]; st_hi[0][k] = sg.hi[k]; }
    sp = 1;
    uint16_t* perm = sperm[w];
    while (sp > 0) {
        --sp;
        const uint32_t index = st_index[sp];
        const int start = st_start[sp], end = st_end[sp];
        float lo[3] = {st_lo[sp][0], st_lo[sp][1], st_lo[sp][2]}, hi[3] = {st_hi[sp][0], st_hi[sp][1], st_hi[sp][2]};
        const int median = pb_median(start, end), axis = pb_axis(lo, hi);
        const float* kx = sk[w][axis];
        auto key = [&](int q) { return kx[perm[q]]; };
        auto swp = [&](int a, int b) { const uint16_t t = perm[a]; perm[a] = perm[b]; perm[b] = t; };
        int left = start, right = end;
        while (right > left) {
            const int i = hoare_round<32, uint16_t>(left, right, key, swp, sG[w], sS[w], lane, nullptr);
            if (i >= median) right = i - 1;
            if (i <= median) left = i + 1;
        }
        const int node = perm[median];
        const float split = kx[node];
        if (lane == 0) heap[index] = sgid[w][node] | ((uint32_t)axis << 30);
        // right child first onto the stack, so the left one is balanced next (the order does not change the result)
        if (median < end) {
            if (median + 1 < end) {
                st_index[sp] = 2 * index + 1; st_start[sp] = median + 1; st_end[sp] = end;
#pragma unroll
                for (int k = 0; k < 3; ++k) { st_lo[sp][k] = k == axis ? split : lo[k]; st_hi[sp][k] = hi[k]; }
                ++sp;
            } else if (lane == 0) heap[2 * index + 1] = sgid[w][perm[end]];
        }
        if (median > start) {
            if (start < median - 1) {
                st_index[sp] = 2 * index; st_start[sp] = start; st_end[sp] = median - 1;
#pragma unroll
                for (int k = 0; k < 3; ++k) { st_lo[sp][k] = lo[k]; st_hi[sp][k] = k == axis ? split : hi[k]; }
                ++sp;
            } else if (lane == 0) heap[2 * index] = sgid[w][perm[start]];
        }
    }
}

// ---- records out of the heap ---------------------------------------------------------------------------------------
// pos[].w carries word 3 of the reference's 28-byte Photon: plane (short) | theta << 16 | phi << 24.
__global__ void k_pb_pack_gather(int n, const uint32_t* __restrict__ heap, const float4* __restrict__ pos, const float4* __restrict__ pow,
                                 const float* __restrict__ tables, float4* __restrict__ photons, float4* __restrict__ search)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == 0) { photons[0] = photons[1] = search[0] = search[1] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const uint32_t hw = heap[i], id = hw & 0x3fffffffu, axis = hw >> 30;
    const float4 p = pos[id], e = pow[id];
    const uint32_t w3 = __float_as_uint(p.w), theta = (w3 >> 16) & 0xffu, phi = w3 >> 24;
    const uint32_t plane = 2 * i <= n ? axis : (w3 & 3u);   // leaves keep whatever they carried; the search never reads it
    const float4 rec = make_float4(p.x, p.y, p.z, __uint_as_float(plane | (theta << 8) | (phi << 16)));
    photons[2 * (size_t)i] = rec; photons[2 * (size_t)i + 1] = make_float4(e.x, e.y, e.z, 0.f);
    const float st = tables[256 + theta];
    search[2 * (size_t)i] = rec;
    search[2 * (size_t)i + 1] = make_float4(__fmul_rn(st, tables[512 + phi]), __fmul_rn(st, tables[768 + phi]), tables[theta], 0.f);
}

__global__ void k_pb_pack28(int n, const uint32_t* __restrict__ heap, const float4* __restrict__ pos, const float4* __restrict__ pow,
                            uint32_t* __restrict__ out7)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t* o = out7 + 7 * (size_t)i;
    uint32_t id, plane_valid = 0, axis = 0;
    if (i == 0) id = 0;
    else { const uint32_t hw = heap[i]; id = hw & 0x3fffffffu; axis = hw >> 30; plane_valid = 2 * i <= n; }
    const float4 p = pos[id], e = pow[id];
    uint32_t w3 = __float_as_uint(p.w);
    if (plane_valid) w3 = (w3 & 0xffff0000u) | axis;   // pbal[index]->plane = axis (PhotonMap.cpp:443); others keep theirs
    o[0] = __float_as_uint(p.x); o[1] = __float_as_uint(p.y); o[2] = __float_as_uint(p.z); o[3] = w3;
    o[4] = __float_as_uint(e.x); o[5] = __float_as_uint(e.y); o[6] = __float_as_uint(e.z);
}

__global__ void k_pm_export28(int n, const float4* __restrict__ photons, uint32_t* __restrict__ out7)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const float4 p = photons[2 * (size_t)i], e = photons[2 * (size_t)i + 1];
    const uint32_t bits = __float_as_uint(p.w);
    uint32_t* o = out7 + 7 * (size_t)i;
    o[0] = __float_as_uint(p.x); o[1] = __float_as_uint(p.y); o[2] = __float_as_uint(p.z);
    o[3] = (bits & 3u) | (((bits >> 8) & 0xffu) << 16) | (((bits >> 16) & 0xffu) << 24);
    o[4] = __float_as_uint(e.x); o[5] = __float_as_uint(e.y); o[6] = __float_as_uint(e.z);
}

__global__ void k_pb_unpack28(int n, const uint32_t* __restrict__ in7, float4* __restrict__ pos, float4* __restrict__ pow)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t* s = in7 + 7 * (size_t)i;
    pos[i] = make_float4(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]), __uint_as_float(s[3]));
    pow[i] = make_float4(__uint_as_float(s[4]), __uint_as_float(s[5]), __uint_as_float(s[6]), 0.f);
}

__global__ void k_pb_identity_heap(int n, uint32_t* heap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) heap[i] = (uint32_t)i;
}

struct Free { void* p = nullptr; ~Free() { if (p) cudaFree(p); } };

}  // namespace

// ---- PhotonBuild ----------------------------------------------------------------------------------------------------
cudaError_t PhotonBuild::alloc(int capacity)
{
    release();
    cap = capacity;
    // one allocation: pos, pow (cap + 1 float4 each), bbox (8 words), ctl (2 ints + padding)
    const size_t recs = (size_t)cap + 1;
    cudaError_t e = cudaMalloc(&pos, 2 * recs * sizeof(float4) + 64);
    if (e != cudaSuccess) { pos = nullptr; cap = 0; return e; }
    pow = pos + recs;
    bbox = reinterpret_cast<uint32_t*>(pow + recs);
    ctl = reinterpret_cast<int*>(bbox + 8);
    e = cudaMemsetAsync(pos, 0, sizeof(float4), cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaMemsetAsync(pow, 0, sizeof(float4), cudaStreamPerThread);
    if (e == cudaSuccess) { k_pb_bbox_init<<<1, 32, 0, cudaStreamPerThread>>>(bbox); e = cudaGetLastError(); }
    if (e != cudaSuccess) release();
    return e;
}

void PhotonBuild::release()
{
    cudaFree(pos); cudaFree(excl); cudaFree(scan_tmp);
    pos = pow = nullptr; bbox = nullptr; ctl = nullptr; excl = nullptr; scan_tmp = nullptr; cap = 0; excl_cap = 0; scan_bytes = 0;
}

cudaError_t PhotonBuild::store_batch(const unsigned char* d_counts, const float* d_records, uint32_t batch, int base, int target,
                                     cudaStream_t st, int* used, int* stored_after)
{
    *used = 0; *stored_after = base;
    if (batch == 0 || base >= target) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (batch > excl_cap) {
        cudaFree(excl); cudaFree(scan_tmp); excl = nullptr; scan_tmp = nullptr; excl_cap = 0;
        e = cudaMalloc(&excl, (size_t)batch * sizeof(uint32_t));
        if (e != cudaSuccess) return e;
        scan_bytes = 0;
        e = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_counts, excl, (int)batch, st);
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&scan_tmp, scan_bytes ? scan_bytes : 16);
        if (e != cudaSuccess) return e;
        excl_cap = batch;
    }
    size_t bytes = scan_bytes;
    e = cub::DeviceScan::ExclusiveSum(scan_tmp, bytes, d_counts, excl, (int)batch, st);
    if (e != cudaSuccess) return e;
    k_photon_store<<<(batch + 255) / 256, 256, 0, st>>>(d_counts, excl, d_records, batch, base, target, pos, pow, bbox, ctl);
    e = cudaGetLastError();
    int h[2] = {0, base};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, ctl, sizeof h, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) { *used = h[0]; *stored_after = h[1]; }
    return e;
}

cudaError_t PhotonBuild::scale(int first, int last, float s, cudaStream_t st)
{
    if (last < first) return cudaSuccess;
    k_photon_scale<<<(unsigned)((last - first + 256) / 256), 256, 0, st>>>(pow, first, last, s);
    return cudaGetLastError();
}

// balance(): d_heap[1..n] = (store index | split axis << 30) in the reference's heap order.  box: NULL = the store's own box.
cudaError_t photon_balance_device(const float4* d_pos, int n, const uint32_t* d_bbox, const float* lo3, const float* hi3, uint32_t* d_heap,
                                  cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    if (n >= (1 << 30)) return cudaErrorInvalidValue;
    // one allocation for all scratch: ids, keys, posG, posS (m words each), small-task list (m + 1 words), open segments
    const size_t m = (size_t)n + 2;
    const size_t words = 4 * m + (m + 2);
    Free scratch;
    cudaError_t e = cudaMalloc(&scratch.p, words * 4 + m * sizeof(PbSeg));
    if (e != cudaSuccess) return e;
    uint32_t* base = static_cast<uint32_t*>(scratch.p);
    struct { void* p; } ids{base}, keys{base + m}, pg{base + 2 * m}, ps{base + 3 * m}, small{base + 4 * m}, segs{base + words};
    e = cudaMemsetAsync(segs.p, 0, m * sizeof(PbSeg), st);
    if (e != cudaSuccess) return e;
    uint32_t* small_list = static_cast<uint32_t*>(small.p) + 1;
    uint32_t* small_count = static_cast<uint32_t*>(small.p);
    const float3 lo = lo3 ? make_float3(lo3[0], lo3[1], lo3[2]) : make_float3(0.f, 0.f, 0.f);
    const float3 hi = hi3 ? make_float3(hi3[0], hi3[1], hi3[2]) : make_float3(0.f, 0.f, 0.f);
    k_pb_init<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(n, d_bbox, lo, hi, static_cast<uint32_t*>(ids.p), static_cast<PbSeg*>(segs.p), d_heap,
                                                           small_list, small_count);
    int height = 0;
    while (((1ll << height) - 1) < n) ++height;                 // levels of the left-balanced tree
    // a segment of level L holds at most 2^(height - L) - 1 photons: levels are processed by k_pb_level while that exceeds PB_SMALL
    int levels_run = 0;
    for (int level = 0; level < height && ((1ll << (height - level)) - 1) > PB_SMALL; ++level, ++levels_run)
        k_pb_level<<<1u << level, 1024, 0, st>>>(level, n, d_pos, static_cast<uint32_t*>(ids.p), static_cast<float*>(keys.p),
                                                 static_cast<uint32_t*>(pg.p), static_cast<uint32_t*>(ps.p), static_cast<PbSeg*>(segs.p), d_heap,
                                                 small_list, small_count);
    // small segments are opened by the (fewer than 2^levels_run) nodes of the levels above, two each at most, or are the root itself;
    // the kernel reads the real count
    const unsigned max_tasks = (unsigned)std::min<size_t>((size_t)n, (size_t)2 << levels_run);
    k_pb_small<<<(max_tasks + PB_WARPS - 1) / PB_WARPS, 32 * PB_WARPS, 0, st>>>(d_pos, static_cast<uint32_t*>(ids.p), static_cast<PbSeg*>(segs.p), d_heap,
                                                                               small_list, small_count);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // the scratch above is freed on return
    return e;
}

static void fill_tables(std::vector<float>& tab)
{
    // direction tables exactly as the reference fills them (PhotonMap.cpp:47-53): double trig, stored as float
    tab.resize(1024);
    for (int i = 0; i < 256; ++i) {
        const double angle = double(i) * (1.0 / 256.0) * M_PI;
        tab[i] = (float)cos(angle); tab[256 + i] = (float)sin(angle);
        tab[512 + i] = (float)cos(2.0 * angle); tab[768 + i] = (float)sin(2.0 * angle);
    }
}

// Balance the stored photons and make them the map the gather kernels read -- device to device.
int PhotonMapDevice::build_from_store(const PhotonBuild& b, int n, bool balance, cudaStream_t st, std::string& err)
{
    release();
    if (n <= 0) return MIROGPU_OK;
    std::vector<float> tab; fill_tables(tab);
    const size_t recs = (size_t)2 * (n + 1);
    cudaError_t e = cudaMalloc(&d_photons, recs * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&d_tables, tab.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_search, recs * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&d_tickets, (size_t)MIRO_GW_TICKET_SLOTS * MIRO_GW_TICKET_SPAN * sizeof(unsigned int));
    Free heap;
    if (e == cudaSuccess) e = cudaMalloc(&heap.p, ((size_t)n + 1) * sizeof(uint32_t));
    uint32_t* d_heap = static_cast<uint32_t*>(heap.p);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tables, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        if (balance) e = photon_balance_device(b.pos, n, b.bbox, nullptr, nullptr, d_heap, st);
        else { k_pb_identity_heap<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(n, d_heap); e = cudaGetLastError(); }
    }
    if (e == cudaSuccess) {
        k_pb_pack_gather<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(n, d_heap, b.pos, b.pow, d_tables, d_photons, d_search);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { err = std::string("photon map build: ") + cudaGetErrorString(e); release(); return e == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA; }
    stored = n;
    half_stored = n / 2 - 1;   // PhotonMap.cpp:358
    return MIROGPU_OK;
}

// The map as the reference's 28-byte Photon array (index 0 unused), for the host mirror of Photon_map.
cudaError_t PhotonMapDevice::export28(void* photons28, cudaStream_t st) const
{
    if (stored <= 0) return cudaSuccess;
    Free out;
    cudaError_t e = cudaMalloc(&out.p, ((size_t)stored + 1) * 28);
    if (e != cudaSuccess) return e;
    k_pm_export28<<<(unsigned)((stored + 256) / 256), 256, 0, st>>>(stored, d_photons, static_cast<uint32_t*>(out.p));
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(photons28, out.p, ((size_t)stored + 1) * 28, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
}

int PhotonMapDevice::clone_from(const PhotonMapDevice& src, int src_device, int dst_device, std::string& err)
{
    release();
    if (src.stored <= 0) return MIROGPU_OK;
    const size_t recs = (size_t)2 * (src.stored + 1) * sizeof(float4);
    cudaError_t e = cudaMalloc(&d_photons, recs);
    if (e == cudaSuccess) e = cudaMalloc(&d_tables, 1024 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_search, recs);
    if (e == cudaSuccess) e = cudaMalloc(&d_tickets, (size_t)MIRO_GW_TICKET_SLOTS * MIRO_GW_TICKET_SPAN * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemcpyPeer(d_photons, dst_device, src.d_photons, src_device, recs);
    if (e == cudaSuccess) e = cudaMemcpyPeer(d_search, dst_device, src.d_search, src_device, recs);
    if (e == cudaSuccess) e = cudaMemcpyPeer(d_tables, dst_device, src.d_tables, src_device, 1024 * sizeof(float));
    if (e != cudaSuccess) { err = std::string("photon map replicate: ") + cudaGetErrorString(e); release(); return e == cudaErrorMemoryAllocation ? MIROGPU_ERR_OOM : MIROGPU_ERR_CUDA; }
    stored = src.stored; half_stored = src.half_stored; exact = src.exact;
    return MIROGPU_OK;
}

static cudaError_t photon_download28(const PhotonBuild& b, const uint32_t* d_heap, int n, void* photons28, cudaStream_t st)
{
    Free out;
    cudaError_t e = cudaMalloc(&out.p, ((size_t)n + 1) * 28);
    if (e != cudaSuccess) return e;
    k_pb_pack28<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(n, d_heap, b.pos, b.pow, static_cast<uint32_t*>(out.p));
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(photons28, out.p, ((size_t)n + 1) * 28, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
}

// Photon_map::balance() for a map the caller filled on the host: records up, balance, heap-ordered records back.
cudaError_t photon_balance_host_array(void* photons28, int n, const float* lo3, const float* hi3, cudaStream_t st)
{
    if (n <= 1) return cudaSuccess;
    PhotonBuild b;
    Free raw, heap;
    cudaError_t e = b.alloc(n);
    if (e == cudaSuccess) e = cudaMalloc(&raw.p, ((size_t)n + 1) * 28);
    if (e == cudaSuccess) e = cudaMalloc(&heap.p, ((size_t)n + 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(raw.p, photons28, ((size_t)n + 1) * 28, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_pb_unpack28<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(n, static_cast<const uint32_t*>(raw.p), b.pos, b.pow);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = photon_balance_device(b.pos, n, nullptr, lo3, hi3, static_cast<uint32_t*>(heap.p), st);
    if (e == cudaSuccess) e = photon_download28(b, static_cast<uint32_t*>(heap.p), n, photons28, st);
    b.release();
    return e;
}

}  // namespace mirogpu
