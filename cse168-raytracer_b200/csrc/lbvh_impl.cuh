// lbvh_impl.cuh -- acceleration-structure construction ON THE DEVICE (replaces BVH::build, reference
// BVH.cpp:60-339, for callers that rebuild often): a linear BVH over the Morton order of the triangles' box centres,
// collapsed four-wide and quantised into the QBVH4 layout the traversal kernels read, plus the 64-byte triangle records
// in the same leaf order.  Selected with mirogpu_build_options.builder = MIROGPU_BUILDER_LBVH_DEVICE; the default stays
// the host binned-SAH builder (bvh_build.cpp), whose trees trace faster.
//
//   k_lbvh_bounds      conservative triangle bounds (the same formula as triangle_bounds, bvh_build.cpp) + scene bounds
//   k_lbvh_morton      63-bit Morton code of the box centre (21 bits per axis); bit 63 sets the few huge triangles apart
//   cub radix sort     (key, triangle) pairs -- a library sort, as for any plain sort
//   k_lbvh_hierarchy   Karras 2012: every internal node finds its key range and split in parallel
//   k_lbvh_refit       bottom-up union of boxes (second arrival at a node continues upwards)
//   k_ploc_*           MIROGPU_BUILDER_PLOC_DEVICE: instead of the two steps above, bottom-up merges of Morton neighbours by
//                      smallest union area (parallel locally-ordered clustering), then the triangles are put in the depth-
//                      first order of that tree
//   k_lbvh_collapse    one launch per level of the wide tree: a node takes its binary subtree's children, opening the
//                      largest-area child until four; subtrees of <= max_leaf triangles become leaves (their triangles are
//                      contiguous in Morton order)
//   k_lbvh_quantise    full-precision four-wide node -> Qbvh4Node (same rule as flatten_qbvh4)
//   k_lbvh_tri_records 64-byte records in leaf order, edge vectors and plane normal in the reference's operations
//
// Correctness does not depend on tree quality: every box is the union of conservative triangle bounds, so the walk finds
// exactly what the exhaustive search finds (tests: hits identical to the SAH-built tree, bit for bit).
#ifndef MIROGPU_LBVH_IMPL_CUH
#define MIROGPU_LBVH_IMPL_CUH

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace mirogpu {

struct LbvhOut {
    void* d_geom = nullptr;      // one allocation: nodes (padded to 256 B) then triangle records
    size_t node_span = 0, node_bytes = 0, tri_bytes = 0;
    uint32_t num_nodes = 0, max_stack = 0, max_depth = 0;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    double seconds = 0;
};

__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__global__ void __launch_bounds__(256) k_lbvh_bounds(const float* __restrict__ v, uint32_t n, float4* __restrict__ blo, float4* __restrict__ bhi,
                                                      uint32_t* __restrict__ scene /* 6 ordered uints: lo xyz, hi xyz */)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (i < n) {
        const float* p = v + 9 * (size_t)i;
        const float kSlop = 2e-4f, kAbsPad = 1e-4f, kRelPad = 2e-6f;   // bvh_build.cpp
        float maxabs = 0.f;
        for (int k = 0; k < 9; ++k) maxabs = fmaxf(maxabs, fabsf(p[k]));
        const float pad = kAbsPad + kRelPad * maxabs;
        for (int k = 0; k < 3; ++k) {
            const float a = p[k], b = p[3 + k], c = p[6 + k], e1 = b - a, e2 = c - a;
            const float q0 = a - kSlop * e1 - kSlop * e2, q1 = a + (1.f + 2.f * kSlop) * e1 - kSlop * e2, q2 = a - kSlop * e1 + (1.f + 2.f * kSlop) * e2;
            lo[k] = fminf(fminf(fminf(q0, q1), fminf(q2, a)), fminf(b, c)) - pad;
            hi[k] = fmaxf(fmaxf(fmaxf(q0, q1), fmaxf(q2, a)), fmaxf(b, c)) + pad;
        }
        blo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        bhi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    // warp reduce, one atomic per warp and component
    for (int k = 0; k < 3; ++k) {
        float l = lo[k], h = hi[k];
        for (int o = 16; o > 0; o >>= 1) { l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o)); h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o)); }
        if ((threadIdx.x & 31) == 0 && l <= h) { atomicMin(scene + k, f2ord(l)); atomicMax(scene + 3 + k, f2ord(h)); }
    }
}

__device__ __forceinline__ unsigned long long spread21(uint32_t x)   // 21 bits -> every third bit of 63
{
    unsigned long long v = x & 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void __launch_bounds__(256) k_lbvh_morton(uint32_t n, const float4* __restrict__ blo, const float4* __restrict__ bhi,
                                                      const uint32_t* __restrict__ scene, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t q[3];
    const float c[3] = {0.5f * (blo[i].x + bhi[i].x), 0.5f * (blo[i].y + bhi[i].y), 0.5f * (blo[i].z + bhi[i].z)};
    const float e[3] = {bhi[i].x - blo[i].x, bhi[i].y - blo[i].y, bhi[i].z - blo[i].z};
    bool big = false;   // a triangle spanning more than 1/8 of the scene along some axis
    for (int k = 0; k < 3; ++k) {
        const float lo = ord2f(scene[k]), hi = ord2f(scene[3 + k]);
        const float ext = hi - lo;
        const float t = ext > 0.f ? (c[k] - lo) / ext : 0.f;
        q[k] = (uint32_t)fminf(fmaxf(t * 2097152.0f, 0.f), 2097151.0f);
        big |= e[k] > 0.125f * ext;
    }
    // Bit 63 separates the few huge triangles from the rest, so the root's split isolates them: left among their Morton
    // neighbours, one ground-plane triangle would blow up the boxes of a whole root-to-leaf path that every ray then walks.
    keys[i] = (big ? 0ull : 0x8000000000000000ull) | (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    idx[i] = i;
}

// Node numbering of the binary tree: internal nodes 0 .. n-2 (root 0), leaf p (Morton position) = n-1+p.
__device__ __forceinline__ int lbvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);   // equal codes: fall back to the position, which is unique
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) k_lbvh_hierarchy(int n, const unsigned long long* __restrict__ keys, int2* __restrict__ child, int* __restrict__ parent,
                                                         int2* __restrict__ range)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t > 0; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = (first == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
    child[i] = make_int2(left, right);
    range[i] = make_int2(first, last);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void __launch_bounds__(256) k_lbvh_refit(int n, const uint32_t* __restrict__ idx, const float4* __restrict__ blo, const float4* __restrict__ bhi,
                                                     const int2* __restrict__ child, const int* __restrict__ parent, float4* __restrict__ nlo,
                                                     float4* __restrict__ nhi, uint32_t* __restrict__ arrived, int2* __restrict__ range)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t t = idx[p];
    int node = n - 1 + p;
    nlo[node] = blo[t]; nhi[node] = bhi[t];
    range[node] = make_int2(p, p);
    __threadfence();
    node = parent[node];
    while (node >= 0) {
        if (atomicAdd(arrived + node, 1u) == 0u) return;   // first arrival: the sibling subtree is not finished yet
        __threadfence();
        const int2 c = child[node];
        const float4 a = nlo[c.x], b = nlo[c.y], e = nhi[c.x], f = nhi[c.y];
        nlo[node] = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f);
        nhi[node] = make_float4(fmaxf(e.x, f.x), fmaxf(e.y, f.y), fmaxf(e.z, f.z), 0.f);
        __threadfence();
        node = parent[node];
    }
}

// ---- PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): bottom-up agglomeration along the Morton order ----
// The clusters (at first the triangles) stay in Morton order; every round each cluster looks `radius` neighbours to either
// side for the partner whose union box has the smallest area, mutual choices merge into a new binary node, and the
// survivors are compacted.  The result approaches agglomerative clustering -- far better boxes than Morton median splits
// where meshes overlap -- in ~30 rounds of three small kernels and a scan.  Node numbering as above (leaf p = n-1+p);
// merges take ids n-2, n-3, ... so that the last one, the root, is node 0.
#define MIRO_PLOC_RADIUS 64          /* neighbours examined to either side (tuning knob MIROGPU_PLOC_RADIUS, at most 64); camera rays on the bench scene: 8 -> 5.20, 16 -> 5.48, 32 -> 5.68, 64 -> 5.77 Grays/s, build 9-10 ms throughout */
#define MIRO_PLOC_MAX_RADIUS 64
#define MIRO_PLOC_THREADS 256

__global__ void __launch_bounds__(256) k_ploc_init(int n, const uint32_t* __restrict__ idx, const float4* __restrict__ blo, const float4* __restrict__ bhi,
                                                    int* __restrict__ cid, float4* __restrict__ clo, float4* __restrict__ chi, float4* __restrict__ nlo,
                                                    float4* __restrict__ nhi, int* __restrict__ size, int* __restrict__ parent)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t t = idx[p];
    const float4 a = blo[t], b = bhi[t];
    cid[p] = n - 1 + p; clo[p] = a; chi[p] = b;
    nlo[n - 1 + p] = a; nhi[n - 1 + p] = b; size[n - 1 + p] = 1; parent[n - 1 + p] = -1;
}

__global__ void __launch_bounds__(MIRO_PLOC_THREADS) k_ploc_nearest(int m, int radius, const float4* __restrict__ clo, const float4* __restrict__ chi,
                                                                    int* __restrict__ nn)
{
    __shared__ float4 slo[MIRO_PLOC_THREADS + 2 * MIRO_PLOC_MAX_RADIUS], shi[MIRO_PLOC_THREADS + 2 * MIRO_PLOC_MAX_RADIUS];
    const int b0 = blockIdx.x * MIRO_PLOC_THREADS - radius;
    for (int k = threadIdx.x; k < MIRO_PLOC_THREADS + 2 * radius; k += MIRO_PLOC_THREADS) {
        const int j = b0 + k;
        if (j >= 0 && j < m) { slo[k] = clo[j]; shi[k] = chi[j]; }
    }
    __syncthreads();
    const int i = blockIdx.x * MIRO_PLOC_THREADS + threadIdx.x;
    if (i >= m) return;
    const float4 a = slo[threadIdx.x + radius], b = shi[threadIdx.x + radius];
    float best = INFINITY; int bj = -1;
    for (int k = 0; k <= 2 * radius; ++k) {        // ascending j: ties keep the smaller index
        const int j = b0 + (int)threadIdx.x + k;
        if (j < 0 || j >= m || j == i) continue;
        const float4 c = slo[threadIdx.x + k], d = shi[threadIdx.x + k];
        const float dx = fmaxf(b.x, d.x) - fminf(a.x, c.x), dy = fmaxf(b.y, d.y) - fminf(a.y, c.y), dz = fmaxf(b.z, d.z) - fminf(a.z, c.z);
        const float area = dx * dy + dy * dz + dz * dx;
        if (area < best) { best = area; bj = j; }
    }
    nn[i] = bj;
}

// flags[i] = 1 when slot i survives the round (unmerged, or the lower index of a merged pair, which then holds the new node)
__global__ void __launch_bounds__(256) k_ploc_merge(int m, int n, const int* __restrict__ nn, int* __restrict__ cid, float4* __restrict__ clo,
                                                     float4* __restrict__ chi, uint32_t* __restrict__ merges, int2* __restrict__ child,
                                                     int* __restrict__ parent, float4* __restrict__ nlo, float4* __restrict__ nhi, int* __restrict__ size,
                                                     int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i];
    int keep = 1;
    if (j >= 0 && nn[j] == i) {
        if (i < j) {
            const int id = n - 2 - (int)atomicAdd(merges, 1u);
            const int a = cid[i], b = cid[j];
            const float4 la = clo[i], lb = clo[j], ha = chi[i], hb = chi[j];
            const float4 lo = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.f);
            const float4 hi = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.f);
            child[id] = make_int2(a, b); parent[a] = id; parent[b] = id; parent[id] = -1;
            nlo[id] = lo; nhi[id] = hi; size[id] = size[a] + size[b];
            // slot i is read by nobody else in this kernel after nn (its partner j only reads nn): safe to overwrite in place
            cid[i] = id; clo[i] = lo; chi[i] = hi;
        } else keep = 0;
    }
    flags[i] = keep;
}

__global__ void __launch_bounds__(256) k_ploc_compact(int m, const int* __restrict__ flags, const int* __restrict__ pos, const int* __restrict__ cid,
                                                       const float4* __restrict__ clo, const float4* __restrict__ chi, int* __restrict__ ocid,
                                                       float4* __restrict__ oclo, float4* __restrict__ ochi, int* __restrict__ new_m)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (flags[i]) { const int o = pos[i]; ocid[o] = cid[i]; oclo[o] = clo[i]; ochi[o] = chi[i]; }
    if (i == m - 1) *new_m = pos[i] + flags[i];
}

// Position of every node's first triangle in the depth-first leaf order: the sizes of the left siblings on the way up.
__global__ void __launch_bounds__(256) k_ploc_ranges(int total, const int2* __restrict__ child, const int* __restrict__ parent, const int* __restrict__ size,
                                                      int2* __restrict__ range, uint32_t* __restrict__ too_deep)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= total) return;
    int first = 0, u = v, steps = 0;
    for (int p = parent[u]; p >= 0; p = parent[u]) {
        const int2 c = child[p];
        if (c.y == u) first += size[c.x];
        u = p;
        if (++steps > 8192) { atomicExch(too_deep, 1u); break; }
    }
    range[v] = make_int2(first, first + size[v] - 1);
}

__global__ void __launch_bounds__(256) k_ploc_order(int n, const uint32_t* __restrict__ idx, const int2* __restrict__ range, uint32_t* __restrict__ order)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    order[range[n - 1 + p].x] = idx[p];
}

struct LbvhFrontier { int bnode; uint32_t slot, pending, depth; };

// counters: [0] wide nodes allocated, [1] next frontier size, [2] max stack need, [3] max depth
__global__ void __launch_bounds__(128) k_lbvh_collapse(int n, int max_leaf, const int2* __restrict__ child, const int2* __restrict__ range,
                                                        const float4* __restrict__ nlo, const float4* __restrict__ nhi,
                                                        const LbvhFrontier* __restrict__ cur, uint32_t ncur, LbvhFrontier* __restrict__ next,
                                                        uint32_t* __restrict__ counters, Bvh4Node* __restrict__ out, uint32_t capacity)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncur) return;
    const LbvhFrontier fr = cur[t];
    auto size_of = [&](int node) { return range[node].y - range[node].x + 1; };   // range: positions in the leaf order, all 2n-1 nodes
    auto first_of = [&](int node) { return range[node].x; };
    auto area_of = [&](int node) {
        const float4 a = nlo[node], b = nhi[node];
        const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
        return dx * dy + dy * dz + dz * dx;
    };
    int ch[4];
    int k = 0;
    ch[k++] = child[fr.bnode].x; ch[k++] = child[fr.bnode].y;
    while (k < 4) {
        int best = -1; float best_area = -1.f;
        for (int i = 0; i < k; ++i) {
            if (size_of(ch[i]) <= max_leaf) continue;     // a leaf of the flat tree
            const float a = area_of(ch[i]);
            if (a > best_area) { best_area = a; best = i; }
        }
        if (best < 0) break;
        const int open = ch[best];
        ch[best] = child[open].x;
        ch[k++] = child[open].y;
    }
    Bvh4Node nd;
    for (int c = 0; c < 4; ++c) {
        nd.lox[c] = nd.hix[c] = nd.loy[c] = nd.hiy[c] = nd.loz[c] = nd.hiz[c] = INFINITY;
        nd.link[c] = (int32_t)0x80000000; nd.pad[c] = 0;
    }
    const uint32_t pending = fr.pending + (uint32_t)(k - 1);
    atomicMax(counters + 2, pending);
    atomicMax(counters + 3, fr.depth);
    for (int c = 0; c < k; ++c) {
        const float4 a = nlo[ch[c]], b = nhi[ch[c]];
        nd.lox[c] = a.x; nd.hix[c] = b.x; nd.loy[c] = a.y; nd.hiy[c] = b.y; nd.loz[c] = a.z; nd.hiz[c] = b.z;
        const int sz = size_of(ch[c]);
        if (sz <= max_leaf) nd.link[c] = ~(int32_t)(((uint32_t)first_of(ch[c]) << 3) | (uint32_t)(sz - 1));
        else {
            const uint32_t slot = atomicAdd(counters + 0, 1u);
            nd.link[c] = (int32_t)slot;
            if (slot < capacity) {
                const uint32_t q = atomicAdd(counters + 1, 1u);
                next[q] = LbvhFrontier{ch[c], slot, pending, fr.depth + 1};
            }
        }
    }
    if (fr.slot < capacity) out[fr.slot] = nd;
}

__global__ void __launch_bounds__(256) k_lbvh_quantise(uint32_t nn, const Bvh4Node* __restrict__ in, Qbvh4Node* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const Bvh4Node w = in[i];
    Qbvh4Node q;
    memset(&q, 0, sizeof q);
    const float* lo[3] = {w.lox, w.loy, w.loz};
    const float* hi[3] = {w.hix, w.hiy, w.hiz};
    uint8_t* qlo[3] = {q.qlox, q.qloy, q.qloz};
    uint8_t* qhi[3] = {q.qhix, q.qhiy, q.qhiz};
    const double kMargin = 0.02;   // as flatten_qbvh4
    for (int a = 0; a < 3; ++a) {
        float mn = INFINITY, mx = -INFINITY;
        for (int c = 0; c < 4; ++c) if (isfinite(w.lox[c])) { mn = fminf(mn, lo[a][c]); mx = fmaxf(mx, hi[a][c]); }
        if (!(mn <= mx)) { mn = 0.f; mx = 0.f; }
        int e = (int)ceil(log2(fmax((double)mx - (double)mn, 1e-30) / (255.0 - 2.0 * kMargin)));
        e = max(-100, min(100, e));
        for (;;) {
            const double cell = ldexp(1.0, e);
            q.origin[a] = qbvh4_stored_origin(mn, e);
            const double g = qbvh4_grid_origin(q.origin[a], e);   // the grid the traversal decodes (<= mn)
            bool ok = true;
            for (int c = 0; c < 4 && ok; ++c) {
                if (!isfinite(w.lox[c])) { qlo[a][c] = 255; qhi[a][c] = 0; continue; }
                const double l = floor(((double)lo[a][c] - g) / cell - kMargin);
                const double h = ceil(((double)hi[a][c] - g) / cell + kMargin);
                if (h > 255.0) { ok = false; break; }
                qlo[a][c] = (uint8_t)fmax(0.0, l);
                qhi[a][c] = (uint8_t)h;
            }
            if (ok) break;
            ++e;
        }
        q.e[a] = (uint8_t)(e + 127);
    }
    for (int c = 0; c < 4; ++c) q.link[c] = w.link[c];
    qbvh4_cell_words(q);
    out[i] = q;
}

__global__ void __launch_bounds__(256) k_lbvh_tri_records(uint32_t n, const float* __restrict__ v, const uint32_t* __restrict__ idx, TriRecord* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = idx[i];
    const float* a = v + 9 * (size_t)p;
    TriRecord r;
    r.ax = a[0]; r.ay = a[1]; r.az = a[2]; r.prim_id = p;
    r.e1x = __fsub_rn(a[3], a[0]); r.e1y = __fsub_rn(a[4], a[1]); r.e1z = __fsub_rn(a[5], a[2]);
    r.e2x = __fsub_rn(a[6], a[0]); r.e2y = __fsub_rn(a[7], a[1]); r.e2z = __fsub_rn(a[8], a[2]);
    r.nx = __fsub_rn(__fmul_rn(r.e1y, r.e2z), __fmul_rn(r.e1z, r.e2y));
    r.ny = __fsub_rn(__fmul_rn(r.e1z, r.e2x), __fmul_rn(r.e1x, r.e2z));
    r.nz = __fsub_rn(__fmul_rn(r.e1x, r.e2y), __fmul_rn(r.e1y, r.e2x));
    r.pad[0] = r.pad[1] = r.pad[2] = 0.f;
    out[i] = r;
}

// The builders' scratch memory (~400-500 bytes per triangle) is ONE allocation that stays cached for the next build on the same
// device: allocating and, above all, freeing half a gigabyte costs the driver far more than the build itself (measured: build
// 19 ms, cudaFree of its scratch 60-630 ms).  mirogpu_release_build_scratch() returns it.
struct BuildScratch {
    std::mutex mtx;
    int device = -1;
    char* ptr = nullptr;
    size_t bytes = 0;
    void release_locked()
    {
        if (ptr) {
            int cur = 0;
            cudaGetDevice(&cur);
            if (cur != device) cudaSetDevice(device);
            cudaFree(ptr);
            if (cur != device) cudaSetDevice(cur);
        }
        ptr = nullptr; bytes = 0; device = -1;
    }
};
inline BuildScratch& build_scratch() { static BuildScratch s; return s; }

// Builds QBVH4 nodes + triangle records on the current device from HOST vertices.  Returns cudaSuccess and fills `o`;
// o.max_stack may exceed what the kernels carry -- the caller checks and falls back to the host builder.
inline cudaError_t build_lbvh_device(const float* tri_vertices, uint32_t ntris, int max_leaf, LbvhOut& o, bool ploc = false)
{
    if (max_leaf < 1) max_leaf = 1;
    if (max_leaf > 4) max_leaf = 4;
    const uint32_t n = ntris;
    cudaError_t e = cudaSuccess;
    // all temporaries come out of ONE allocation (cudaMalloc / cudaFree of two dozen 20-90 MB buffers cost more than the build)
    const size_t per_tri = 36 + 32 + 16 + 8 + 8 + 16 + 8 + 64 + 4 + 64 + 16 + 16 /* sort scratch */ + (ploc ? 8 + 64 + 12 + 8 : 0);
    // wide nodes: a node has at least two children, so with leaves of up to max_leaf >= 2 triangles the bottom nodes cover three
    // triangles or more (n / 2 + 2 nodes is safe); with single-triangle leaves a bottom node may cover just two (up to ~2n / 3)
    const size_t wide_capacity = max_leaf == 1 ? (size_t)n + 2 : (size_t)n / 2 + 2;
    const size_t pool_bytes = (size_t)n * per_tri + wide_capacity * (sizeof(Bvh4Node) + 2 * 16) + (8u << 20);
    BuildScratch& scratch = build_scratch();
    std::lock_guard<std::mutex> scratch_lock(scratch.mtx);   // one device build at a time per process
    char* pool = nullptr;
    size_t pool_used = 0;
    std::vector<void*> tmp;
    auto dalloc = [&](void** p, size_t bytes) {
        bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
        if (!pool) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (scratch.device != dev || scratch.bytes < pool_bytes) {
                scratch.release_locked();
                const cudaError_t r = cudaMalloc((void**)&scratch.ptr, pool_bytes);
                if (r != cudaSuccess) { scratch.ptr = nullptr; return r; }
                scratch.device = dev; scratch.bytes = pool_bytes;
            }
            pool = scratch.ptr;
        }
        if (pool_used + bytes <= pool_bytes) { *p = pool + pool_used; pool_used += bytes; return cudaSuccess; }
        const cudaError_t r = cudaMalloc(p, bytes);     // the estimate fell short (library scratch larger than expected)
        if (r == cudaSuccess) tmp.push_back(*p);
        return r;
    };
    auto cleanup = [&]() { for (void* p : tmp) cudaFree(p); tmp.clear(); pool = nullptr; };
#define LB(x) do { e = (x); if (e != cudaSuccess) { cleanup(); if (o.d_geom) { cudaFree(o.d_geom); o.d_geom = nullptr; } return e; } } while (0)
    cudaStream_t st = cudaStreamPerThread;
    const auto t0 = std::chrono::steady_clock::now();
    float* d_v = nullptr; float4 *blo = nullptr, *bhi = nullptr, *nlo = nullptr, *nhi = nullptr;
    uint32_t *scene = nullptr, *idx = nullptr, *idx2 = nullptr, *arrived = nullptr, *counters = nullptr;
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    int2 *child = nullptr, *range = nullptr; int* parent = nullptr;
    LB(dalloc((void**)&d_v, (size_t)n * 36));
    LB(cudaMemcpyAsync(d_v, tri_vertices, (size_t)n * 36, cudaMemcpyHostToDevice, st));
    LB(dalloc((void**)&blo, (size_t)n * 16)); LB(dalloc((void**)&bhi, (size_t)n * 16));
    LB(dalloc((void**)&scene, 32));
    const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    LB(cudaMemcpyAsync(scene, init, sizeof init, cudaMemcpyHostToDevice, st));
    const unsigned g256 = (n + 255) / 256;
    // the final geometry allocation is sized for that worst case
    const uint32_t capacity = (uint32_t)wide_capacity;
    o.node_span = (((size_t)capacity * sizeof(Qbvh4Node)) + 255) & ~(size_t)255;
    o.tri_bytes = (size_t)n * sizeof(TriRecord);
    LB(cudaMalloc(&o.d_geom, o.node_span + std::max<size_t>(o.tri_bytes, 16)));
    if (n > 0) k_lbvh_bounds<<<g256, 256, 0, st>>>(d_v, n, blo, bhi, scene);
    uint32_t h_scene[6];
    LB(cudaMemcpyAsync(h_scene, scene, sizeof h_scene, cudaMemcpyDeviceToHost, st));
    Bvh4Node* wide = nullptr;
    LB(dalloc((void**)&wide, (size_t)capacity * sizeof(Bvh4Node)));
    uint32_t num_wide = 1;
    if (n <= (uint32_t)max_leaf) {
        // the whole scene is one leaf (or empty): a root whose first child is that leaf -- assembled on the host
        LB(cudaStreamSynchronize(st));
        Bvh4Node root;
        for (int c = 0; c < 4; ++c) { root.lox[c] = root.hix[c] = root.loy[c] = root.hiy[c] = root.loz[c] = root.hiz[c] = INFINITY; root.link[c] = (int32_t)0x80000000; root.pad[c] = 0; }
        if (n > 0) {
            auto of = [](uint32_t u) { uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &b, 4); return f; };
            root.lox[0] = of(h_scene[0]); root.loy[0] = of(h_scene[1]); root.loz[0] = of(h_scene[2]);
            root.hix[0] = of(h_scene[3]); root.hiy[0] = of(h_scene[4]); root.hiz[0] = of(h_scene[5]);
            root.link[0] = ~(int32_t)((0u << 3) | (n - 1));
        }
        LB(cudaMemcpyAsync(wide, &root, sizeof root, cudaMemcpyHostToDevice, st));
        LB(dalloc((void**)&idx, (size_t)std::max<uint32_t>(n, 1) * 4));
        std::vector<uint32_t> ident(std::max<uint32_t>(n, 1));
        for (uint32_t i = 0; i < n; ++i) ident[i] = i;
        LB(cudaMemcpyAsync(idx, ident.data(), (size_t)std::max<uint32_t>(n, 1) * 4, cudaMemcpyHostToDevice, st));
        LB(cudaStreamSynchronize(st));
        o.max_stack = 0; o.max_depth = 1;
    } else {
        LB(dalloc((void**)&keys, (size_t)n * 8)); LB(dalloc((void**)&keys2, (size_t)n * 8));
        LB(dalloc((void**)&idx2, (size_t)n * 4)); LB(dalloc((void**)&idx, (size_t)n * 4));
        k_lbvh_morton<<<g256, 256, 0, st>>>(n, blo, bhi, scene, keys2, idx2);
        size_t sort_bytes = 0;
        LB(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys2, keys, idx2, idx, (int)n, 0, 64, st));
        void* sort_tmp = nullptr;
        LB(dalloc(&sort_tmp, sort_bytes));
        LB(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys2, keys, idx2, idx, (int)n, 0, 64, st));
        LB(dalloc((void**)&child, (size_t)n * 8)); LB(dalloc((void**)&range, (size_t)2 * n * 8));
        LB(dalloc((void**)&parent, (size_t)2 * n * 4));
        LB(dalloc((void**)&nlo, (size_t)2 * n * 16)); LB(dalloc((void**)&nhi, (size_t)2 * n * 16));
        if (!ploc) {
            LB(dalloc((void**)&arrived, (size_t)n * 4));
            LB(cudaMemsetAsync(arrived, 0, (size_t)n * 4, st));
            k_lbvh_hierarchy<<<(n - 1 + 255) / 256, 256, 0, st>>>((int)n, keys, child, parent, range);
            k_lbvh_refit<<<g256, 256, 0, st>>>((int)n, idx, blo, bhi, child, parent, nlo, nhi, arrived, range);
        } else {
            int *cid[2] = {nullptr, nullptr}, *nn = nullptr, *flags = nullptr, *pos = nullptr, *size = nullptr, *d_m = nullptr;
            float4 *clo[2] = {nullptr, nullptr}, *chi[2] = {nullptr, nullptr};
            uint32_t* pc = nullptr;   // [0] merges so far, [1] a walk to the root gave up
            for (int k = 0; k < 2; ++k) { LB(dalloc((void**)&cid[k], (size_t)n * 4)); LB(dalloc((void**)&clo[k], (size_t)n * 16)); LB(dalloc((void**)&chi[k], (size_t)n * 16)); }
            LB(dalloc((void**)&nn, (size_t)n * 4)); LB(dalloc((void**)&flags, (size_t)n * 4)); LB(dalloc((void**)&pos, (size_t)n * 4));
            LB(dalloc((void**)&size, (size_t)2 * n * 4)); LB(dalloc((void**)&d_m, 4)); LB(dalloc((void**)&pc, 8));
            LB(cudaMemsetAsync(pc, 0, 8, st));
            size_t scan_bytes = 0;
            LB(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags, pos, (int)n, st));
            void* scan_tmp = nullptr;
            LB(dalloc(&scan_tmp, scan_bytes));
            k_ploc_init<<<g256, 256, 0, st>>>((int)n, idx, blo, bhi, cid[0], clo[0], chi[0], nlo, nhi, size, parent);
            int m = (int)n, cur = 0;
            int radius = MIRO_PLOC_RADIUS;
            if (const char* ev = getenv("MIROGPU_PLOC_RADIUS")) radius = std::min(std::max(atoi(ev), 1), MIRO_PLOC_MAX_RADIUS);
            const bool dbg = getenv("MIROGPU_DEBUG_BUILD") != nullptr;
            const auto tp0 = std::chrono::steady_clock::now();
            int rounds = 0, slow = 0;
            for (int round = 0; m > 1 && round < 4096; ++round) {
                ++rounds;
                if (dbg && (round < 40 || round % 50 == 0)) fprintf(stderr, "ploc round %d m %d t %.3f ms\n", round, m, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
                const unsigned gm = (unsigned)((m + 255) / 256);
                k_ploc_nearest<<<(unsigned)((m + MIRO_PLOC_THREADS - 1) / MIRO_PLOC_THREADS), MIRO_PLOC_THREADS, 0, st>>>(m, radius, clo[cur], chi[cur], nn);
                k_ploc_merge<<<gm, 256, 0, st>>>(m, (int)n, nn, cid[cur], clo[cur], chi[cur], pc, child, parent, nlo, nhi, size, flags);
                LB(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flags, pos, m, st));
                k_ploc_compact<<<gm, 256, 0, st>>>(m, flags, pos, cid[cur], clo[cur], chi[cur], cid[cur ^ 1], clo[cur ^ 1], chi[cur ^ 1], d_m);
                int new_m = 0;
                LB(cudaMemcpyAsync(&new_m, d_m, 4, cudaMemcpyDeviceToHost, st));
                LB(cudaStreamSynchronize(st));
                if (new_m >= m || new_m < 1) { cleanup(); cudaFree(o.d_geom); o.d_geom = nullptr; return cudaErrorNotSupported; }   // cannot happen: a mutual pair always exists
                // thousands of coincident boxes merge one pair per round: give up after 24 rounds in a row that merged less
                // than a thousandth of the clusters instead of grinding on to the round cap (the caller falls back to the host build)
                slow = (long long)(m - new_m) * 1024 < m ? slow + 1 : 0;
                m = new_m; cur ^= 1;
                if (slow >= 24) break;
            }
            // cudaErrorNotSupported = "this builder gives up on this input": the caller falls back to the host builder.  An unfinished
            // agglomeration (round cap, or the early exit above) leaves child[] / parent[] partly unwritten: nothing may walk it.
            if (m != 1) { cleanup(); cudaFree(o.d_geom); o.d_geom = nullptr; return cudaErrorNotSupported; }
            k_ploc_ranges<<<(unsigned)((2 * n - 1 + 255) / 256), 256, 0, st>>>((int)(2 * n - 1), child, parent, size, range, pc + 1);
            uint32_t h_pc[2] = {0, 0};
            LB(cudaMemcpyAsync(h_pc, pc, 8, cudaMemcpyDeviceToHost, st));
            LB(cudaStreamSynchronize(st));
            if (dbg) fprintf(stderr, "ploc rounds %d, clustering + ranges %.3f ms\n", rounds, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
            // (a chain deeper than the walk cap gives up the same way)
            if (h_pc[0] != n - 1 || h_pc[1] != 0) { cleanup(); cudaFree(o.d_geom); o.d_geom = nullptr; return cudaErrorNotSupported; }
            // triangles in the depth-first order of the agglomerated tree: every subtree is a contiguous run again
            k_ploc_order<<<g256, 256, 0, st>>>((int)n, idx, range, idx2);
            std::swap(idx, idx2);
        }
        // level-by-level collapse into four-wide nodes
        LbvhFrontier *fa = nullptr, *fb = nullptr;
        LB(dalloc((void**)&fa, (size_t)capacity * sizeof(LbvhFrontier))); LB(dalloc((void**)&fb, (size_t)capacity * sizeof(LbvhFrontier)));
        LB(dalloc((void**)&counters, 16));
        const uint32_t c0[4] = {1u, 0u, 0u, 0u};
        LB(cudaMemcpyAsync(counters, c0, sizeof c0, cudaMemcpyHostToDevice, st));
        const LbvhFrontier rootf = {0, 0u, 0u, 1u};
        LB(cudaMemcpyAsync(fa, &rootf, sizeof rootf, cudaMemcpyHostToDevice, st));
        uint32_t ncur = 1;
        const auto tc0 = std::chrono::steady_clock::now();
        int levels = 0;
        for (int level = 0; ncur > 0 && level < 4096; ++level) {
            ++levels;
            k_lbvh_collapse<<<(ncur + 127) / 128, 128, 0, st>>>((int)n, max_leaf, child, range, nlo, nhi, fa, ncur, fb, counters, wide, capacity);
            uint32_t hc[4];
            LB(cudaMemcpyAsync(hc, counters, sizeof hc, cudaMemcpyDeviceToHost, st));
            LB(cudaStreamSynchronize(st));
            ncur = hc[1]; num_wide = hc[0]; o.max_stack = hc[2]; o.max_depth = hc[3];
            if (num_wide > capacity) { cleanup(); cudaFree(o.d_geom); o.d_geom = nullptr; return cudaErrorMemoryAllocation; }
            const uint32_t zero = 0;
            LB(cudaMemcpyAsync(counters + 1, &zero, 4, cudaMemcpyHostToDevice, st));
            std::swap(fa, fb);
        }
        if (getenv("MIROGPU_DEBUG_BUILD")) fprintf(stderr, "collapse: %d levels, %.3f ms, %u wide nodes, max stack %u; since start %.3f ms\n", levels, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count(), num_wide, o.max_stack, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    o.num_nodes = num_wide;
    o.node_bytes = (size_t)num_wide * sizeof(Qbvh4Node);
    k_lbvh_quantise<<<(num_wide + 255) / 256, 256, 0, st>>>(num_wide, wide, reinterpret_cast<Qbvh4Node*>(o.d_geom));
    if (n > 0) k_lbvh_tri_records<<<g256, 256, 0, st>>>(n, d_v, idx, reinterpret_cast<TriRecord*>(static_cast<char*>(o.d_geom) + o.node_span));
    LB(cudaGetLastError());
    LB(cudaStreamSynchronize(st));
    auto of = [](uint32_t u) { uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &b, 4); return f; };
    for (int k = 0; k < 3; ++k) { o.lo[k] = n ? of(h_scene[k]) : 0.f; o.hi[k] = n ? of(h_scene[3 + k]) : 0.f; }
    const double before_cleanup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    cleanup();
    o.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("MIROGPU_DEBUG_BUILD")) fprintf(stderr, "device build %.3f ms (%.3f ms before freeing the scratch pool of %.0f MB, %.0f MB used)\n", 1e3 * o.seconds, 1e3 * before_cleanup, pool_bytes / 1048576.0, pool_used / 1048576.0);
#undef LB
    return cudaSuccess;
}

}  // namespace mirogpu
#endif
