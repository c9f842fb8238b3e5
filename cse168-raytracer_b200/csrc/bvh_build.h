// bvh_build.h -- host-side acceleration-structure construction for the B200 engine.
//
// Replaces BVH::build (reference BVH.cpp:60-339) on the product path.  The reference grows a pointer
// tree by a 32-step binary search on the split plane per axis; here a binned-SAH binary tree is built
// over conservative triangle bounds and then flattened into one of two GPU layouts:
//   * BVH2  -- 64-byte nodes holding both children's boxes (two children per fetch), and
//   * CWBVH8 -- 80-byte 8-wide nodes with child boxes quantised to 8 bits on a per-node grid
//     (compressed wide BVH), children slotted for octant-ordered traversal.
// Triangles are re-ordered into leaf (spatial) order so a leaf's triangles are one contiguous 48-byte-
// strided run in HBM.
//
// Box contract (what makes the flat tree a drop-in for the reference's): every box contains, with
// margin, every point at which Triangle::intersect (Triangle.cpp:136-169) can accept a hit on the
// triangles below it -- i.e. the triangle grown by the reference's epsilon slop in barycentric space
// (beta, gamma >= -eps, beta+gamma <= 1+eps) plus an absolute pad >= the reference's own epsilon pad
// (BVH.cpp:75-79).  tests/test_flat_bvh.py checks this for every node of both layouts.
#ifndef MIROGPU_BVH_BUILD_H
#define MIROGPU_BVH_BUILD_H

#include <cstdint>
#include <cmath>
#include <vector>
#include "qbvh4_config.h"

// four-wide layouts: the first MIRO_TOP_NODES nodes are the top levels in breadth-first order (1 + 4 + 16 + 64)
#define MIRO_TOP_NODES 85

namespace mirogpu {

struct Aabb {
    float lo[3], hi[3];
};

struct BinaryNode {
    Aabb box;
    int32_t left, right;    // children (internal) or -1 (leaf)
    uint32_t first, count;  // leaf: range in BinaryBvh::order
};

struct BinaryBvh {
    std::vector<BinaryNode> nodes;  // nodes[0] is the root
    std::vector<uint32_t> order;    // primitive ids in leaf order
    uint32_t num_leaves = 0;
    uint32_t max_depth = 0;
};

// Conservative bounds of one triangle (9 floats A,B,C): slop-grown triangle + absolute pad.
void triangle_bounds(const float* v9, Aabb& out);

BinaryBvh build_binary_sah(const float* tri_vertices, uint32_t ntris, int max_leaf, int bins);

// ---- BVH2 flat layout (64 B / node) -----------------------------------------------------------------
// f[0..3]  = child0 lo.x, hi.x, lo.y, hi.y      f[4..7] = child1 lo.x, hi.x, lo.y, hi.y
// f[8..11] = child0 lo.z, hi.z, child1 lo.z, hi.z
// link[0], link[1] = child references: >= 0 internal node index; < 0 leaf: ~((first << 3) | (count-1))
struct Bvh2Node {
    float f[12];
    int32_t link[4];
};
static_assert(sizeof(Bvh2Node) == 64, "Bvh2Node must be 64 bytes");

// ---- BVH4 flat layout (128 B / node = one cache line = four 32-byte sectors) --------------------------------------
// The binary SAH tree collapsed to four children per node (always opening the child with the largest surface area),
// boxes in full binary32, structure-of-arrays so one 256-bit load brings both planes of one axis for all four
// children.  Half the dependent fetches per ray of BVH2 for the same number of box tests.
//   lox[4] hix[4] | loy[4] hiy[4] | loz[4] hiz[4] | link[4] pad[4]
// link: >= 0 internal node index; < 0 leaf: ~((first << 3) | (count-1)); empty slot: boxes at +inf (never hit).
struct Bvh4Node {
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
    int32_t link[4];
    int32_t pad[4];
};
static_assert(sizeof(Bvh4Node) == 128, "Bvh4Node must be 128 bytes");

// ---- QBVH4 flat layout (64 B / node = two 32-byte sectors) ---------------------------------------------------------
// The same four-wide collapse as BVH4 with the child boxes quantised to 8 bits per plane on a per-node grid: grid
// origin = the node's min corner (binary32; the MIRO_QDIRECT variant stores it 2^15 cells low, qbvh4_stored_origin), one
// power-of-two cell size per axis (biased exponent byte), planes = origin + q * cell with q rounded outward (plus a safety margin that covers the decode's rounding).  Half the
// sectors per visit of BVH4 / the same as BVH2 at half the visits -- the traversal is bound by L1 wavefronts per ray.
//   origin[3] | ex,ey,ez,0 | qlo.x[4] qhi.x[4] qlo.y[4] qhi.y[4] || qlo.z[4] qhi.z[4] | link[4] | cell[2]
// Empty slot: qlo = 255, qhi = 0 on every axis (an inverted interval never passes the slab test).
struct Qbvh4Node {
    float origin[3];
    uint8_t e[3], pad0;
    uint8_t qlox[4], qhix[4], qloy[4], qhiy[4];
    uint8_t qloz[4], qhiz[4];
    int32_t link[4];
    uint32_t cell[2];   // the three cell sizes times 2^24 as the traversal multiplies them (qbvh4_cell_words), derived from e[]
};
static_assert(sizeof(Qbvh4Node) == 64, "Qbvh4Node must be 64 bytes");

// The traversal needs 2^24 * cell per axis as a binary32 (qbvh4_node_step: the planes arrive as q * 2^-24).  A power of two is
// its exponent field alone, so the upper 16 bits of the binary32 are enough: x and y share cell[0] (low / high half), z is
// cell[1] whole -- one shift and one mask per node instead of three shift / mask / add chains on the exponent bytes.
// e[] <= 227 (the builders clamp the exponent), so e + 24 is a finite exponent.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void qbvh4_cell_words(Qbvh4Node& q)
{
    q.cell[0] = (((uint32_t)q.e[0] + (uint32_t)MIRO_QSHIFT) << 7) | (((uint32_t)q.e[1] + (uint32_t)MIRO_QSHIFT) << 23);
    q.cell[1] = ((uint32_t)q.e[2] + (uint32_t)MIRO_QSHIFT) << 23;
}

// MIRO_QDIRECT (qbvh4_config.h; a measured, slower variant, off by default): the traversal reads a plane byte as the binary32 1 + q 2^-15 and multiplies by 2^15 cell, i.e. it evaluates
// (2^15 + q) cell above the stored origin -- so the node stores origin' = the largest binary32 with origin' + 2^15 cell <= mn, and
// the builders quantise against the grid that origin' really gives, qbvh4_grid_origin(origin', e) = origin' + 2^15 cell (exact in
// binary64), never above mn.  Where 2^15 cell exceeds |mn| the grid moves down by up to 2^-9 cell, which the outward rounding of q
// absorbs like any other offset; the traversal's own rounding of (origin' - o) / d is then at most 2^-9 cell too, inside the
// builders' 0.02-cell margin.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline float qbvh4_stored_origin(float mn, int e)
{
#if MIRO_QDIRECT
    const double off = ldexp(1.0, e + 15);
    float o = (float)((double)mn - off);
    while ((double)o + off > (double)mn) o = nextafterf(o, -INFINITY);
    return o;
#else
    (void)e;
    return mn;
#endif
}
#if defined(__CUDACC__)
__host__ __device__
#endif
inline double qbvh4_grid_origin(float stored, int e)
{
#if MIRO_QDIRECT
    return (double)stored + ldexp(1.0, e + 15);
#else
    (void)e;
    return (double)stored;
#endif
}

// ---- CWBVH8 flat layout (80 B / node) ----------------------------------------------------------------
struct Cwbvh8Node {
    float p[3];          // quantisation grid origin (node box min)
    uint8_t e[3];        // per-axis biased exponent: cell size = 2^(e-127)
    uint8_t imask;       // bit s set: slot s holds an internal child
    uint32_t child_base; // index of the first internal child; child of slot s = base + popc(imask & ((1<<s)-1))
    uint32_t tri_base;   // first triangle of this node's leaf children
    uint8_t meta[8];     // 0 empty | internal: 0x20 | (24+s) | leaf: (unary count)<<5 | triangle offset
    uint8_t qlox[8], qloy[8], qloz[8];
    uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Cwbvh8Node) == 80, "Cwbvh8Node must be 80 bytes");

// 64-byte triangle record in leaf order (two 32-byte sectors, fetched with two 256-bit loads): A.xyz + prim id
// bits, B-A, C-A (the two edge vectors exactly as Triangle.cpp:150 forms them in binary32) and their cross product
// (the plane normal Triangle.cpp:151 recomputes on every call; same operations, so the same bits).
struct TriRecord {
    float ax, ay, az;
    uint32_t prim_id;
    float e1x, e1y, e1z, nx;
    float e2x, e2y, e2z, ny;
    float nz, pad[3];
};
static_assert(sizeof(TriRecord) == 64, "TriRecord must be 64 bytes");

struct FlatBvh {
    int layout = 0;
    std::vector<Bvh2Node> nodes2;
    std::vector<Cwbvh8Node> nodes8;
    std::vector<Bvh4Node> nodes4;
    std::vector<Qbvh4Node> nodesq;
    std::vector<uint32_t> order;  // flat triangle slot -> prim id
    Aabb root;
    uint32_t max_depth = 0;
    uint32_t max_stack = 0;   // BVH4: most entries a walk can have on its stack (sum over a root-to-leaf path of children - 1)
};

void flatten_bvh2(const BinaryBvh& b, FlatBvh& out);
void flatten_cwbvh8(const BinaryBvh& b, FlatBvh& out);
void flatten_bvh4(const BinaryBvh& b, FlatBvh& out);
void flatten_qbvh4(const BinaryBvh& b, FlatBvh& out);   // flatten_bvh4, then quantise each node
void make_tri_records(const float* tri_vertices, const std::vector<uint32_t>& order, std::vector<TriRecord>& out);

}  // namespace mirogpu
#endif
