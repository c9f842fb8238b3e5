// bvh_build.cpp -- binned-SAH binary build + flattening to the BVH2 / CWBVH8 GPU layouts.
// See bvh_build.h for the role of this file and the box contract.  Own design; the reference's builder
// (BVH.cpp:60-339) is restated only in oracle/miro_oracle.cpp.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace mirogpu {

namespace {

const float kInf = std::numeric_limits<float>::infinity();
// The reference accepts hits with beta, gamma >= -1e-4 and beta+gamma <= 1+1e-4 (Triangle.cpp:158,
// Miro.h:9).  Grow by twice that so rounding in the barycentrics cannot leave the box.
const float kSlop = 2e-4f;
const float kAbsPad = 1e-4f;   // the reference pads every node by epsilon (BVH.cpp:75-79)
const float kRelPad = 2e-6f;   // covers o + t*d rounding at large coordinates

inline void box_reset(Aabb& b)
{
    for (int k = 0; k < 3; ++k) { b.lo[k] = kInf; b.hi[k] = -kInf; }
}
inline void box_grow(Aabb& b, const Aabb& o)
{
    for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(b.lo[k], o.lo[k]); b.hi[k] = std::max(b.hi[k], o.hi[k]); }
}
inline float box_half_area(const Aabb& b)
{
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0.f) || !(dy >= 0.f) || !(dz >= 0.f)) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}

struct Builder {
    // The triangle list is partitioned as 32-byte records (bounds + id), not as indices into per-triangle arrays: every pass
    // of every level then streams through contiguous memory instead of gathering (the build is a memory walk, nothing else).
    struct Prim { Aabb b; uint32_t id, pad; };
    static float centre(const Prim& p, int k) { return 0.5f * (p.b.lo[k] + p.b.hi[k]); }
    Prim* prims;                // the records, partitioned in place
    BinaryNode* nodes;
    std::atomic<uint32_t> next_node{1};
    std::atomic<uint32_t> num_leaves{0};
    std::atomic<uint32_t> max_depth{0};
    int max_leaf, bins;
    float trav_cost = 1.0f;     // cost of one node visit relative to one triangle test
    Prim* tmp = nullptr;        // scratch records for the parallel partition of big nodes (ntris entries)

    // Nodes above this size are processed by several tasks per pass (bounds, binning, partition): near the root a node's
    // triangle list is the whole scene, and one thread walking it five times is most of the build.  Partial results are
    // combined with min / max / integer sums and the partition is stable, so the tree is the one the serial passes give.
    static const uint32_t kBig = 65536, kChunk = 32768;
    struct BinSet { Aabb bb[3][256]; uint32_t bc[3][256]; };

    template <typename F>
    void for_chunks(uint32_t begin, uint32_t end, F f)
    {
        const uint32_t nchunks = (end - begin + kChunk - 1) / kChunk;
        for (uint32_t c = 0; c < nchunks; ++c) {
#pragma omp task default(shared) firstprivate(c)
            f(c, begin + c * kChunk, std::min(end, begin + (c + 1) * kChunk));
        }
#pragma omp taskwait
    }

    void note_depth(uint32_t d)
    {
        uint32_t cur = max_depth.load();
        while (d > cur && !max_depth.compare_exchange_weak(cur, d)) {}
    }

    void make_leaf(uint32_t ni, uint32_t begin, uint32_t end, uint32_t depth)
    {
        nodes[ni].left = nodes[ni].right = -1;
        nodes[ni].first = begin;
        nodes[ni].count = end - begin;
        num_leaves++;
        note_depth(depth);
    }

    void build(uint32_t ni, uint32_t begin, uint32_t end, uint32_t depth)
    {
        Aabb box, cbox;
        box_reset(box); box_reset(cbox);
        const uint32_t n = end - begin;
        const bool big = n > kBig;
        const uint32_t nchunks = big ? (n + kChunk - 1) / kChunk : 0;
        auto bounds_of = [&](uint32_t b, uint32_t e, Aabb& bx, Aabb& cx) {
            for (uint32_t i = b; i < e; ++i) {
                const Prim& p = prims[i];
                box_grow(bx, p.b);
                for (int k = 0; k < 3; ++k) { const float c = centre(p, k); cx.lo[k] = std::min(cx.lo[k], c); cx.hi[k] = std::max(cx.hi[k], c); }
            }
        };
        if (big) {
            std::vector<Aabb> part(2 * (size_t)nchunks);
            for (auto& a : part) box_reset(a);
            for_chunks(begin, end, [&](uint32_t c, uint32_t b, uint32_t e) { bounds_of(b, e, part[2 * c], part[2 * c + 1]); });
            for (uint32_t c = 0; c < nchunks; ++c) { box_grow(box, part[2 * c]); box_grow(cbox, part[2 * c + 1]); }
        } else bounds_of(begin, end, box, cbox);
        nodes[ni].box = box;
        if (n <= 1) { make_leaf(ni, begin, end, depth); return; }

        // --- binned SAH over the three axes -----------------------------------------------------
        const int B = bins;
        int best_axis = -1, best_split = -1;
        float best_cost = kInf;
        // Past depth 40 only median splits are taken, which bounds the depth by 40 + log2(n).
        if (depth < 40) {
            // one pass over the triangles fills the bins of all three axes (the lists are index-permuted, so every pass is a
            // gather over the per-triangle arrays: three passes cost three times the cache misses)
            float lo3[3], scale3[3];
            bool use[3];
            for (int axis = 0; axis < 3; ++axis) {
                const float ext = cbox.hi[axis] - cbox.lo[axis];
                lo3[axis] = cbox.lo[axis]; use[axis] = ext > 0.f; scale3[axis] = use[axis] ? (float)B / ext : 0.f;
            }
            auto fill = [&](BinSet& S, uint32_t b0, uint32_t e0) {
                for (int axis = 0; axis < 3; ++axis) for (int b = 0; b < B; ++b) { box_reset(S.bb[axis][b]); S.bc[axis][b] = 0; }
                for (uint32_t i = b0; i < e0; ++i) {
                    const Prim& p = prims[i];
                    const Aabb& pbx = p.b;
                    for (int axis = 0; axis < 3; ++axis) {
                        if (!use[axis]) continue;
                        int b = (int)((centre(p, axis) - lo3[axis]) * scale3[axis]);
                        b = b < 0 ? 0 : (b >= B ? B - 1 : b);
                        S.bc[axis][b]++; box_grow(S.bb[axis][b], pbx);
                    }
                }
            };
            static thread_local BinSet local;     // used between here and the split decision only: no task switch in between
            std::vector<BinSet> sets;
            if (big) {
                sets.resize(nchunks);
                for_chunks(begin, end, [&](uint32_t c, uint32_t b0, uint32_t e0) { fill(sets[c], b0, e0); });
                for (uint32_t c = 1; c < nchunks; ++c)
                    for (int axis = 0; axis < 3; ++axis)
                        for (int b = 0; b < B; ++b) { sets[0].bc[axis][b] += sets[c].bc[axis][b]; box_grow(sets[0].bb[axis][b], sets[c].bb[axis][b]); }
            } else fill(local, begin, end);
            const BinSet& S = big ? sets[0] : local;
            float right_area[256];
            for (int axis = 0; axis < 3; ++axis) {
                if (!use[axis]) continue;
                const Aabb* bb = S.bb[axis];
                const uint32_t* bc = S.bc[axis];
                Aabb acc; box_reset(acc);
                for (int b = B - 1; b > 0; --b) { box_grow(acc, bb[b]); right_area[b] = box_half_area(acc); }
                box_reset(acc);
                uint32_t nl = 0;
                for (int b = 0; b < B - 1; ++b) {
                    box_grow(acc, bb[b]); nl += bc[b];
                    const uint32_t nr = n - nl;
                    if (nl == 0 || nr == 0) continue;
                    const float cost = box_half_area(acc) * (float)nl + right_area[b + 1] * (float)nr;
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
                }
            }
        }
        const float parent_area = box_half_area(box);
        if ((int)n <= max_leaf) {
            // leaf cost n*Ct vs split cost Ctrav + (sum A_i n_i)/A * Ct, with Ct = 1
            const float leaf_cost = (float)n;
            const float split_cost = (best_axis >= 0 && parent_area > 0.f) ? trav_cost + best_cost / parent_area : kInf;
            if (leaf_cost <= split_cost) { make_leaf(ni, begin, end, depth); return; }
        }
        uint32_t mid;
        if (best_axis >= 0) {
            const float lo = cbox.lo[best_axis], scale = (float)B / (cbox.hi[best_axis] - cbox.lo[best_axis]);
            const int axis = best_axis, split = best_split;
            auto goes_left = [=](const Prim& p) {
                int b = (int)((centre(p, axis) - lo) * scale);
                b = b < 0 ? 0 : (b >= B ? B - 1 : b);
                return b <= split;
            };
            if (big) {
                // stable partition in two parallel passes: count per chunk, then scatter to the scratch permutation and copy back
                std::vector<uint32_t> nleft(nchunks + 1, 0);
                for_chunks(begin, end, [&](uint32_t ch, uint32_t b0, uint32_t e0) {
                    uint32_t k = 0;
                    for (uint32_t i = b0; i < e0; ++i) k += goes_left(prims[i]) ? 1u : 0u;
                    nleft[ch + 1] = k;
                });
                for (uint32_t ch = 0; ch < nchunks; ++ch) nleft[ch + 1] += nleft[ch];
                const uint32_t total_left = nleft[nchunks];
                for_chunks(begin, end, [&](uint32_t ch, uint32_t b0, uint32_t e0) {
                    uint32_t l = begin + nleft[ch], r = begin + total_left + ((b0 - begin) - nleft[ch]);
                    for (uint32_t i = b0; i < e0; ++i) { const Prim& p = prims[i]; if (goes_left(p)) tmp[l++] = p; else tmp[r++] = p; }
                });
                for_chunks(begin, end, [&](uint32_t, uint32_t b0, uint32_t e0) { memcpy(prims + b0, tmp + b0, (size_t)(e0 - b0) * sizeof(Prim)); });
                mid = begin + total_left;
            } else {
                Prim* m = std::partition(prims + begin, prims + end, goes_left);
                mid = (uint32_t)(m - prims);
            }
        } else {
            mid = begin;
        }
        if (mid == begin || mid == end) {
            // degenerate (all centroids equal, or depth cap): median split on the widest centroid axis
            int axis = 0;
            float e0 = cbox.hi[0] - cbox.lo[0], e1 = cbox.hi[1] - cbox.lo[1], e2 = cbox.hi[2] - cbox.lo[2];
            if (e1 > e0) axis = 1;
            if (e2 > std::max(e0, e1)) axis = 2;
            mid = begin + n / 2;
            std::nth_element(prims + begin, prims + mid, prims + end, [=](const Prim& a, const Prim& b) {
                const float ca = centre(a, axis), cb = centre(b, axis);
                return ca < cb || (ca == cb && a.id < b.id);
            });
        }
        const uint32_t l = next_node.fetch_add(2), r = l + 1;
        nodes[ni].left = (int32_t)l; nodes[ni].right = (int32_t)r;
        nodes[ni].first = 0; nodes[ni].count = 0;
        if (n > 65536) {
#pragma omp task default(shared) firstprivate(l, begin, mid, depth)
            build(l, begin, mid, depth + 1);
#pragma omp task default(shared) firstprivate(r, mid, end, depth)
            build(r, mid, end, depth + 1);
#pragma omp taskwait
        } else {
            build(l, begin, mid, depth + 1);
            build(r, mid, end, depth + 1);
        }
    }
};

}  // namespace

void triangle_bounds(const float* v, Aabb& out)
{
    // Corners of the slop-grown triangle in barycentric space: (-s,-s), (1+2s,-s), (-s,1+2s).
    float e1[3], e2[3], q[3][3];
    float maxabs = 0.f;
    for (int k = 0; k < 3; ++k) {
        e1[k] = v[3 + k] - v[k]; e2[k] = v[6 + k] - v[k];
        maxabs = std::max(maxabs, std::max(std::fabs(v[k]), std::max(std::fabs(v[3 + k]), std::fabs(v[6 + k]))));
    }
    for (int k = 0; k < 3; ++k) {
        q[0][k] = v[k] - kSlop * e1[k] - kSlop * e2[k];
        q[1][k] = v[k] + (1.f + 2.f * kSlop) * e1[k] - kSlop * e2[k];
        q[2][k] = v[k] - kSlop * e1[k] + (1.f + 2.f * kSlop) * e2[k];
    }
    const float pad = kAbsPad + kRelPad * maxabs;
    for (int k = 0; k < 3; ++k) {
        float lo = std::min(std::min(q[0][k], q[1][k]), std::min(q[2][k], std::min(v[k], std::min(v[3 + k], v[6 + k]))));
        float hi = std::max(std::max(q[0][k], q[1][k]), std::max(q[2][k], std::max(v[k], std::max(v[3 + k], v[6 + k]))));
        out.lo[k] = lo - pad; out.hi[k] = hi + pad;
    }
}

BinaryBvh build_binary_sah(const float* tri_vertices, uint32_t ntris, int max_leaf, int bins)
{
    BinaryBvh out;
    if (max_leaf < 1) max_leaf = 1;
    if (max_leaf > 8) max_leaf = 8;
    if (bins < 4) bins = 4;
    if (bins > 256) bins = 256;
    out.order.resize(ntris);
    for (uint32_t i = 0; i < ntris; ++i) out.order[i] = i;
    if (ntris == 0) {
        BinaryNode n; box_reset(n.box); n.left = n.right = -1; n.first = 0; n.count = 0;
        out.nodes.push_back(n); out.num_leaves = 1;
        return out;
    }
    std::vector<Builder::Prim> prims(ntris), scratch(ntris);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)ntris; ++i) {
        triangle_bounds(tri_vertices + 9 * i, prims[i].b);
        prims[i].id = (uint32_t)i; prims[i].pad = 0;
    }
    out.nodes.resize((size_t)2 * ntris);
    Builder b;
    b.prims = prims.data(); b.nodes = out.nodes.data();
    b.max_leaf = max_leaf; b.bins = bins;
    b.tmp = scratch.data();
    if (const char* e = getenv("MIROGPU_CTRAV")) { const float v = (float)atof(e); if (v > 0.f && v < 100.f) b.trav_cost = v; }   // tuning knob
#pragma omp parallel
#pragma omp single nowait
    b.build(0, 0, ntris, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)ntris; ++i) out.order[i] = prims[i].id;
    out.nodes.resize(b.next_node.load());
    out.num_leaves = b.num_leaves.load();
    out.max_depth = b.max_depth.load();
    return out;
}

// ---------------------------------------------------------------------------------------------------
void flatten_bvh2(const BinaryBvh& b, FlatBvh& out)
{
    out.layout = 0;
    out.order = b.order;
    out.root = b.nodes[0].box;
    out.max_depth = b.max_depth;
    out.nodes2.clear();
    auto leaf_ref = [](const BinaryNode& n) -> int32_t { return ~(int32_t)((n.first << 3) | (n.count - 1)); };
    auto put_child = [](Bvh2Node& fn, int c, const Aabb& bx) {
        fn.f[4 * c + 0] = bx.lo[0]; fn.f[4 * c + 1] = bx.hi[0]; fn.f[4 * c + 2] = bx.lo[1]; fn.f[4 * c + 3] = bx.hi[1];
        fn.f[8 + 2 * c + 0] = bx.lo[2]; fn.f[8 + 2 * c + 1] = bx.hi[2];
    };
    // Empty slot: both planes at +inf.  The slab test takes min/max of the two plane distances per axis, so an
    // inverted box would read as all of space; with lo = hi = +inf every axis yields +-inf for both planes and
    // the interval is empty for any finite ray.
    Aabb empty;
    for (int k = 0; k < 3; ++k) { empty.lo[k] = kInf; empty.hi[k] = kInf; }

    if (b.nodes[0].left < 0) {
        // The whole scene is one leaf (or empty): a root whose first child is that leaf.
        Bvh2Node fn; std::memset(&fn, 0, sizeof fn);
        if (b.nodes[0].count > 0) { put_child(fn, 0, b.nodes[0].box); fn.link[0] = leaf_ref(b.nodes[0]); }
        else { put_child(fn, 0, empty); fn.link[0] = ~0; }
        put_child(fn, 1, empty); fn.link[1] = fn.link[0];
        out.nodes2.push_back(fn);
        return;
    }
    // Depth-first numbering of internal nodes so a subtree is contiguous in HBM.
    std::vector<int32_t> flat_index(b.nodes.size(), -1);
    std::vector<uint32_t> stack; stack.push_back(0);
    std::vector<uint32_t> internal_order;
    while (!stack.empty()) {
        uint32_t n = stack.back(); stack.pop_back();
        flat_index[n] = (int32_t)internal_order.size();
        internal_order.push_back(n);
        const BinaryNode& bn = b.nodes[n];
        if (b.nodes[bn.right].left >= 0) stack.push_back((uint32_t)bn.right);
        if (b.nodes[bn.left].left >= 0) stack.push_back((uint32_t)bn.left);
    }
    out.nodes2.resize(internal_order.size());
    for (size_t i = 0; i < internal_order.size(); ++i) {
        const BinaryNode& bn = b.nodes[internal_order[i]];
        Bvh2Node fn; std::memset(&fn, 0, sizeof fn);
        const int32_t ch[2] = {bn.left, bn.right};
        for (int c = 0; c < 2; ++c) {
            const BinaryNode& cn = b.nodes[ch[c]];
            put_child(fn, c, cn.box);
            fn.link[c] = cn.left >= 0 ? flat_index[ch[c]] : leaf_ref(cn);
        }
        out.nodes2[i] = fn;
    }
}


// ---------------------------------------------------------------------------------------------------
void flatten_bvh4(const BinaryBvh& b, FlatBvh& out)
{
    out.layout = 2;
    out.order = b.order;
    out.root = b.nodes[0].box;
    out.max_depth = 0;
    out.nodes4.clear();
    auto leaf_ref = [](const BinaryNode& n) -> int32_t { return ~(int32_t)((n.first << 3) | (n.count - 1)); };
    auto blank = []() {
        Bvh4Node fn; std::memset(&fn, 0, sizeof fn);
        for (int c = 0; c < 4; ++c) {   // empty slot: both planes at +inf on every axis (see flatten_bvh2)
            fn.lox[c] = fn.hix[c] = fn.loy[c] = fn.hiy[c] = fn.loz[c] = fn.hiz[c] = kInf;
            fn.link[c] = (int32_t)0x80000000;
        }
        return fn;
    };
    auto put_child = [](Bvh4Node& fn, int c, const Aabb& bx) {
        fn.lox[c] = bx.lo[0]; fn.hix[c] = bx.hi[0]; fn.loy[c] = bx.lo[1]; fn.hiy[c] = bx.hi[1]; fn.loz[c] = bx.lo[2]; fn.hiz[c] = bx.hi[2];
    };
    if (b.nodes[0].left < 0) {
        Bvh4Node fn = blank();
        if (b.nodes[0].count > 0) { put_child(fn, 0, b.nodes[0].box); fn.link[0] = leaf_ref(b.nodes[0]); }
        out.nodes4.push_back(fn);
        return;
    }
    // children of a wide node: start from the two binary children, keep opening the internal child with the largest
    // surface area until there are four (or nothing left to open)
    auto collect = [&](uint32_t n, int32_t ch[4]) -> int {
        int k = 0;
        ch[k++] = b.nodes[n].left; ch[k++] = b.nodes[n].right;
        while (k < 4) {
            int best = -1; float best_area = -1.f;
            for (int i = 0; i < k; ++i) {
                const BinaryNode& c = b.nodes[ch[i]];
                if (c.left < 0) continue;
                const float a = box_half_area(c.box);
                if (a > best_area) { best_area = a; best = i; }
            }
            if (best < 0) break;
            const int32_t open = ch[best];
            ch[best] = b.nodes[open].left;
            ch[k++] = b.nodes[open].right;
        }
        return k;
    };
    // numbering of the wide nodes: a short breadth-first prefix, then depth-first (a subtree is contiguous in HBM)
    struct Item { uint32_t bnode; uint32_t depth; uint32_t pending; };   // pending: stack entries left behind by the ancestors
    std::vector<int32_t> flat_index(b.nodes.size(), -1);
    out.max_stack = 0;
    // The first MIRO_TOP_NODES wide nodes are numbered breadth-first (the top levels occupy indices [0, 85): levels of 1 + 4 +
    // 16 + 64 nodes when full -- the hybrid kernel can stage that prefix in shared memory), everything below depth-first.
    std::vector<uint32_t> wide_order;
    std::vector<Item> top; top.push_back({0u, 1u, 0u});
    size_t head = 0;
    auto visit = [&](const Item& it, int32_t ch[4]) -> int {
        flat_index[it.bnode] = (int32_t)wide_order.size();
        wide_order.push_back(it.bnode);
        out.max_depth = std::max(out.max_depth, it.depth);
        const int k = collect(it.bnode, ch);
        out.max_stack = std::max(out.max_stack, it.pending + (uint32_t)(k - 1));
        return k;
    };
    while (head < top.size() && wide_order.size() < MIRO_TOP_NODES) {
        const Item it = top[head++];
        int32_t ch[4];
        const int k = visit(it, ch);
        for (int i = 0; i < k; ++i)
            if (b.nodes[ch[i]].left >= 0) top.push_back({(uint32_t)ch[i], it.depth + 1, it.pending + (uint32_t)(k - 1)});
    }
    std::vector<Item> stack;
    for (size_t r = top.size(); r > head; --r) stack.push_back(top[r - 1]);   // the subtrees below the prefix, first one on top
    while (!stack.empty()) {
        const Item it = stack.back(); stack.pop_back();
        int32_t ch[4];
        const int k = visit(it, ch);
        for (int i = k - 1; i >= 0; --i)
            if (b.nodes[ch[i]].left >= 0) stack.push_back({(uint32_t)ch[i], it.depth + 1, it.pending + (uint32_t)(k - 1)});
    }
    out.nodes4.resize(wide_order.size());
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)wide_order.size(); ++i) {   // every wide node is written from the binary tree alone
        Bvh4Node fn = blank();
        int32_t ch[4];
        const int k = collect(wide_order[i], ch);
        for (int c = 0; c < k; ++c) {
            const BinaryNode& cn = b.nodes[ch[c]];
            put_child(fn, c, cn.box);
            fn.link[c] = cn.left >= 0 ? flat_index[ch[c]] : leaf_ref(cn);
        }
        out.nodes4[i] = fn;
    }
}

// ---------------------------------------------------------------------------------------------------
void flatten_qbvh4(const BinaryBvh& b, FlatBvh& out)
{
    flatten_bvh4(b, out);
    out.layout = 3;
    out.nodesq.resize(out.nodes4.size());
    const double kMargin = 0.02;   // cells; covers the few-ulp rounding of the device decode (traverse.cuh, qbvh4_node_step)
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)out.nodes4.size(); ++i) {
        const Bvh4Node& w = out.nodes4[i];
        Qbvh4Node q; std::memset(&q, 0, sizeof q);
        const float* lo[3] = {w.lox, w.loy, w.loz};
        const float* hi[3] = {w.hix, w.hiy, w.hiz};
        uint8_t* qlo[3] = {q.qlox, q.qloy, q.qloz};
        uint8_t* qhi[3] = {q.qhix, q.qhiy, q.qhiz};
        bool used[4];
        for (int c = 0; c < 4; ++c) used[c] = std::isfinite(w.lox[c]);
        for (int a = 0; a < 3; ++a) {
            float mn = kInf, mx = -kInf;
            for (int c = 0; c < 4; ++c) if (used[c]) { mn = std::min(mn, lo[a][c]); mx = std::max(mx, hi[a][c]); }
            if (!(mn <= mx)) { mn = 0.f; mx = 0.f; }
            // smallest power-of-two cell with origin + (255 - margin) * cell >= max; retried one notch coarser if a plane
            // still lands above 255 after outward rounding
            int e = (int)std::ceil(std::log2(std::max((double)mx - (double)mn, 1e-30) / (255.0 - 2.0 * kMargin)));
            e = std::max(-100, std::min(100, e));
            for (;;) {
                const double cell = std::ldexp(1.0, e);
                q.origin[a] = qbvh4_stored_origin(mn, e);
                const double g = qbvh4_grid_origin(q.origin[a], e);   // the grid the traversal decodes (<= mn)
                bool ok = true;
                for (int c = 0; c < 4 && ok; ++c) {
                    if (!used[c]) { qlo[a][c] = 255; qhi[a][c] = 0; continue; }
                    const double l = std::floor(((double)lo[a][c] - g) / cell - kMargin);
                    const double h = std::ceil(((double)hi[a][c] - g) / cell + kMargin);
                    if (h > 255.0) { ok = false; break; }
                    qlo[a][c] = (uint8_t)std::max(0.0, l);
                    qhi[a][c] = (uint8_t)h;
                }
                if (ok) break;
                ++e;
            }
            q.e[a] = (uint8_t)(e + 127);
        }
        for (int c = 0; c < 4; ++c) q.link[c] = w.link[c];
        qbvh4_cell_words(q);
        out.nodesq[(size_t)i] = q;
    }
}

// ---------------------------------------------------------------------------------------------------
namespace {
struct WideChild { int32_t bnode; };
}

void flatten_cwbvh8(const BinaryBvh& b, FlatBvh& out)
{
    out.layout = 1;
    out.root = b.nodes[0].box;
    out.nodes8.clear();
    out.order.clear();
    out.order.reserve(b.order.size());
    out.max_depth = 0;

    if (b.nodes[0].left < 0 && b.nodes[0].count == 0) {
        // empty scene: one node whose eight slots are all empty
        Cwbvh8Node fn; std::memset(&fn, 0, sizeof fn);
        fn.e[0] = fn.e[1] = fn.e[2] = 127;
        for (int s = 0; s < 8; ++s) { fn.qlox[s] = fn.qloy[s] = fn.qloz[s] = 255; }
        out.nodes8.push_back(fn);
        out.max_depth = 1;
        return;
    }
    struct Task { int32_t bnode; uint32_t depth; };
    std::vector<Task> tasks;  // tasks[i] describes flat node i
    tasks.push_back({0, 1});
    out.nodes8.resize(1);

    for (size_t ti = 0; ti < tasks.size(); ++ti) {
        const Task task = tasks[ti];
        out.max_depth = std::max(out.max_depth, task.depth);
        const BinaryNode& root = b.nodes[task.bnode];
        // ---- collapse: open the internal child of largest area until 8 children ----------------
        int32_t ch[8]; int nch = 0;
        if (root.left < 0) { ch[nch++] = task.bnode; }
        else { ch[nch++] = root.left; ch[nch++] = root.right; }
        while (nch < 8) {
            int pick = -1; float best = -1.f;
            for (int i = 0; i < nch; ++i) {
                const BinaryNode& c = b.nodes[ch[i]];
                if (c.left < 0) continue;
                const float a = box_half_area(c.box);
                if (a > best) { best = a; pick = i; }
            }
            if (pick < 0) break;
            const BinaryNode& c = b.nodes[ch[pick]];
            ch[pick] = c.left; ch[nch++] = c.right;
        }
        // ---- slot assignment for octant-ordered traversal ---------------------------------------
        // Slot s stands for the corner direction D_s = (s&4 ? + : -, s&2 ? + : -, s&1 ? + : -); a child is
        // placed in the slot whose direction its centroid offset points along most, greedily by score.
        const Aabb& nb = root.box;
        float ctr[3]; for (int k = 0; k < 3; ++k) ctr[k] = 0.5f * (nb.lo[k] + nb.hi[k]);
        float score[8][8];
        for (int i = 0; i < nch; ++i) {
            const Aabb& cb = b.nodes[ch[i]].box;
            float d[3]; for (int k = 0; k < 3; ++k) d[k] = 0.5f * (cb.lo[k] + cb.hi[k]) - ctr[k];
            for (int s = 0; s < 8; ++s)
                score[i][s] = ((s & 4) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 1) ? d[2] : -d[2]);
        }
        int slot_of[8]; bool slot_used[8] = {false}, child_done[8] = {false};
        for (int round = 0; round < nch; ++round) {
            int bi = -1, bs = -1; float bv = -kInf;
            for (int i = 0; i < nch; ++i) {
                if (child_done[i]) continue;
                for (int s = 0; s < 8; ++s) {
                    if (slot_used[s]) continue;
                    if (score[i][s] > bv) { bv = score[i][s]; bi = i; bs = s; }
                }
            }
            if (bi < 0) {   // non-finite scores (degenerate boxes): first free child into first free slot
                for (int i = 0; i < nch && bi < 0; ++i) if (!child_done[i]) bi = i;
                for (int s = 0; s < 8 && bs < 0; ++s) if (!slot_used[s]) bs = s;
            }
            slot_of[bi] = bs; slot_used[bs] = true; child_done[bi] = true;
        }
        int child_in_slot[8]; for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
        for (int i = 0; i < nch; ++i) child_in_slot[slot_of[i]] = ch[i];

        // ---- encode ------------------------------------------------------------------------------
        Cwbvh8Node fn; std::memset(&fn, 0, sizeof fn);
        float cell[3];
        for (int k = 0; k < 3; ++k) {
            fn.p[k] = nb.lo[k];
            const float ext = nb.hi[k] - nb.lo[k];
            int e;
            if (!(ext > 0.f) || !std::isfinite(ext)) e = 1;
            else {
                // smallest power of two with 255 * 2^x >= ext (checked in binary32, as the device evaluates it)
                int x = (int)std::ceil(std::log2((double)ext / 255.0));
                while ((double)fn.p[k] + 255.0 * std::ldexp(1.0, x) < (double)nb.hi[k]) ++x;
                e = x + 127;
                if (e < 1) e = 1;
                if (e > 254) e = 254;
            }
            fn.e[k] = (uint8_t)e;
            cell[k] = std::ldexp(1.0f, e - 127);
        }
        uint32_t n_internal = 0;
        for (int s = 0; s < 8; ++s) if (child_in_slot[s] >= 0 && b.nodes[child_in_slot[s]].left >= 0) n_internal++;
        fn.child_base = (uint32_t)out.nodes8.size();
        fn.tri_base = (uint32_t)out.order.size();
        if (n_internal) out.nodes8.resize(out.nodes8.size() + n_internal);
        uint32_t tri_off = 0;
        for (int s = 0; s < 8; ++s) {
            const int32_t c = child_in_slot[s];
            if (c < 0) {
                fn.meta[s] = 0;
                fn.qlox[s] = fn.qloy[s] = fn.qloz[s] = 255;   // inverted box: never hit
                fn.qhix[s] = fn.qhiy[s] = fn.qhiz[s] = 0;
                continue;
            }
            const BinaryNode& cn = b.nodes[c];
            uint8_t* qlo[3] = {fn.qlox, fn.qloy, fn.qloz};
            uint8_t* qhi[3] = {fn.qhix, fn.qhiy, fn.qhiz};
            for (int k = 0; k < 3; ++k) {
                double lo = std::floor(((double)cn.box.lo[k] - (double)fn.p[k]) / (double)cell[k]);
                double hi = std::ceil(((double)cn.box.hi[k] - (double)fn.p[k]) / (double)cell[k]);
                lo = std::min(std::max(lo, 0.0), 255.0); hi = std::min(std::max(hi, 0.0), 255.0);
                // re-check against the binary32 reconstruction p + q*cell and widen if rounding bit us
                while (lo > 0.0 && fn.p[k] + (float)lo * cell[k] > cn.box.lo[k]) lo -= 1.0;
                while (hi < 255.0 && fn.p[k] + (float)hi * cell[k] < cn.box.hi[k]) hi += 1.0;
                qlo[k][s] = (uint8_t)lo; qhi[k][s] = (uint8_t)hi;
            }
            if (cn.left >= 0) {
                fn.imask |= (uint8_t)(1u << s);
                fn.meta[s] = (uint8_t)(0x20 | (24 + s));
            } else {
                const uint32_t cnt = cn.count;  // <= 3 by construction
                fn.meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | tri_off);
                for (uint32_t t = 0; t < cnt; ++t) out.order.push_back(b.order[cn.first + t]);
                tri_off += cnt;
            }
        }
        // children in slot order get consecutive indices; register their tasks in that order
        uint32_t rank = 0;
        for (int s = 0; s < 8; ++s) {
            const int32_t c = child_in_slot[s];
            if (c >= 0 && b.nodes[c].left >= 0) {
                const uint32_t idx = fn.child_base + rank++;
                if (tasks.size() <= idx) tasks.resize(idx + 1);
                tasks[idx] = {c, task.depth + 1};
            }
        }
        out.nodes8[ti] = fn;
    }
}

void make_tri_records(const float* v, const std::vector<uint32_t>& order, std::vector<TriRecord>& out)
{
    out.resize(order.size());
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)order.size(); ++i) {
        const uint32_t p = order[i];
        const float* a = v + 9 * (size_t)p;
        TriRecord r;
        r.ax = a[0]; r.ay = a[1]; r.az = a[2]; r.prim_id = p;
        // B - A and C - A in binary32, exactly the BmA / CmA of Triangle.cpp:150
        r.e1x = a[3] - a[0]; r.e1y = a[4] - a[1]; r.e1z = a[5] - a[2];
        r.e2x = a[6] - a[0]; r.e2y = a[7] - a[1]; r.e2z = a[8] - a[2];
        // normal = cross(BmA, CmA), Triangle.cpp:151 (Vector3.h cross: y*v.z - z*v.y, ...), each product and the
        // difference rounded to binary32 separately (this file is compiled with -ffp-contract=off)
        const float p0 = r.e1y * r.e2z, p1 = r.e1z * r.e2y, p2 = r.e1z * r.e2x, p3 = r.e1x * r.e2z, p4 = r.e1x * r.e2y, p5 = r.e1y * r.e2x;
        r.nx = p0 - p1; r.ny = p2 - p3; r.nz = p4 - p5;
        r.pad[0] = r.pad[1] = r.pad[2] = 0.f;
        out[i] = r;
    }
}

}  // namespace mirogpu
