// tests/cpu_emu/emu.cu -- TEST-ONLY host emulation of the per-ray traversal cores.
//
// The product's traversal functions (cse168-raytracer_b200/csrc/traverse.cuh) are __host__ __device__;
// this file calls them on the CPU, one ray at a time, over the same flat BVH the product's builder
// produces, so the traversal / flattening logic can be checked against the oracle in the CPU test
// tier (no GPU in the build container).  It is compiled into tests/cpu_emu/libmiro_emu.so and is
// never linked into libmirogpu.so: the product has no CPU path.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../cse168-raytracer_b200/csrc/bvh_build.h"
#include "../../cse168-raytracer_b200/csrc/traverse.cuh"

using namespace mirogpu;

extern "C" int emu_trace(const float* tri_vertices, uint32_t ntris, int layout, int max_leaf, const mirogpu_ray* rays,
                         long n, mirogpu_hit* hits, int any_hit, unsigned long long* counters3, uint32_t* info4)
{
    // layout 2: the BVH2 tree walked through the single-step functions of the hybrid kernel; layout 3: BVH4
    // layout 5: QBVH4 walked through the default hybrid kernel's steps -- qbvh4_node_step + one triangle per leaf step
    const int walk = layout == 2 ? 2 : 0;
    const bool wide4 = layout == 3, quant4 = layout == 4 || layout == 5, steps4 = layout == 5;
    if (layout >= 2) layout = MIROGPU_LAYOUT_BVH2;
    if (max_leaf <= 0) max_leaf = layout == MIROGPU_LAYOUT_CWBVH8 ? 3 : 4;
    if (layout == MIROGPU_LAYOUT_CWBVH8 && max_leaf > 3) max_leaf = 3;
    BinaryBvh bin = build_binary_sah(tri_vertices, ntris, max_leaf, 32);
    FlatBvh flat;
    if (quant4) flatten_qbvh4(bin, flat); else if (wide4) flatten_bvh4(bin, flat); else if (layout == MIROGPU_LAYOUT_BVH2) flatten_bvh2(bin, flat); else flatten_cwbvh8(bin, flat);
    std::vector<TriRecord> tris;
    make_tri_records(tri_vertices, flat.order, tris);
    if ((wide4 || quant4) && flat.max_stack > MIRO_STACK4) return 1;
    if (info4) {
        info4[0] = (uint32_t)((wide4 || quant4) ? flat.nodes4.size() : layout == MIROGPU_LAYOUT_BVH2 ? flat.nodes2.size() : flat.nodes8.size());
        info4[1] = (uint32_t)bin.nodes.size(); info4[2] = bin.num_leaves; info4[3] = flat.max_depth;
    }
    const float4* nodes = quant4 ? reinterpret_cast<const float4*>(flat.nodesq.data())
                          : wide4 ? reinterpret_cast<const float4*>(flat.nodes4.data())
                          : layout == MIROGPU_LAYOUT_BVH2 ? reinterpret_cast<const float4*>(flat.nodes2.data())
                                                          : reinterpret_cast<const float4*>(flat.nodes8.data());
    const float4* tr = reinterpret_cast<const float4*>(tris.data());
    unsigned long long cn = 0, cb = 0, ct = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : cn, cb, ct)
    for (long i = 0; i < n; ++i) {
        BestHit best;
        TraceCounters c = {0, 0, 0};
        if (steps4) {
            Bvh2Walk st;
            LocalStack<MIRO_STACK4 + 1> stack;
            bvh2_begin(rays[i], st, best);
            while (st.node != MIRO_BVH2_DONE) {
                if (st.node >= 0) { qbvh4_node_step<0>(nodes, tr, rays[i], st, stack, best); ++c.nodes; c.boxes += 4; }
                else {
                    ++c.tris;
                    if (any_hit) bvh2_leaf_step_one<true, false>(tr, rays[i], st, stack, best); else bvh2_leaf_step_one<false, false>(tr, rays[i], st, stack, best);
                }
            }
        } else if (quant4) {
            if (any_hit) trace_qbvh4<true, true>(nodes, tr, rays[i], best, &c); else trace_qbvh4<false, true>(nodes, tr, rays[i], best, &c);
        } else if (wide4) {
            if (any_hit) trace_bvh4<true, true>(nodes, tr, rays[i], best, &c); else trace_bvh4<false, true>(nodes, tr, rays[i], best, &c);
        } else if (walk) {
            Bvh2Walk st;
            LocalStack<MIRO_STACK + 1> stack;
            bvh2_begin(rays[i], st, best);
            while (st.node != MIRO_BVH2_DONE) {
                if (st.node >= 0) { bvh2_node_step<0>(nodes, tr, rays[i], st, stack, best); ++c.nodes; }
                else {
                    ++c.tris;
                    if (any_hit) bvh2_leaf_step<true, false>(tr, rays[i], st, stack, best); else bvh2_leaf_step<false, false>(tr, rays[i], st, stack, best);
                }
            }
        } else if (layout == MIROGPU_LAYOUT_BVH2) {
            if (any_hit) trace_bvh2<true, true>(nodes, tr, rays[i], best, &c); else trace_bvh2<false, true>(nodes, tr, rays[i], best, &c);
        } else {
            const uint4* n8 = reinterpret_cast<const uint4*>(nodes);
            if (any_hit) trace_cwbvh8<true, true>(n8, tr, rays[i], best, &c); else trace_cwbvh8<false, true>(n8, tr, rays[i], best, &c);
        }
        hits[i].t = best.t; hits[i].prim_id = best.prim; hits[i].beta = best.beta; hits[i].gamma = best.gamma;
        cn += c.nodes; cb += c.boxes; ct += c.tris;
    }
    if (counters3) { counters3[0] = cn; counters3[1] = cb; counters3[2] = ct; }
    return 0;
}

// Flat layout produced by the product's builder, for structural checks on the CPU tier.
// Call with out_nodes = NULL to get sizes: sizes[0] = node bytes, sizes[1] = triangles, sizes[2] = binary nodes,
// sizes[3] = binary leaves, sizes[4] = depth.
extern "C" int emu_build(const float* tri_vertices, uint32_t ntris, int layout, int max_leaf, void* out_nodes, uint32_t* out_order,
                         void* out_tris, uint64_t* sizes)
{
    if (max_leaf <= 0) max_leaf = layout == MIROGPU_LAYOUT_CWBVH8 ? 3 : 4;
    if (layout == MIROGPU_LAYOUT_CWBVH8 && max_leaf > 3) max_leaf = 3;
    const auto t0 = std::chrono::steady_clock::now();
    BinaryBvh bin = build_binary_sah(tri_vertices, ntris, max_leaf, 32);
    const auto t1 = std::chrono::steady_clock::now();
    FlatBvh flat;
    if (layout == 5) layout = 4;
    if (layout == 4) flatten_qbvh4(bin, flat); else if (layout == 3) flatten_bvh4(bin, flat); else if (layout == MIROGPU_LAYOUT_BVH2) flatten_bvh2(bin, flat); else flatten_cwbvh8(bin, flat);
    if (sizes) {   // seconds of the two phases, as IEEE doubles
        const double bs = std::chrono::duration<double>(t1 - t0).count(), fs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
        memcpy(&sizes[5], &bs, 8); memcpy(&sizes[6], &fs, 8);
    }
    const size_t nb = layout == 4 ? flat.nodesq.size() * sizeof(Qbvh4Node) : layout == 3 ? flat.nodes4.size() * sizeof(Bvh4Node)
                                  : layout == MIROGPU_LAYOUT_BVH2 ? flat.nodes2.size() * sizeof(Bvh2Node) : flat.nodes8.size() * sizeof(Cwbvh8Node);
    sizes[0] = nb; sizes[1] = flat.order.size(); sizes[2] = bin.nodes.size(); sizes[3] = bin.num_leaves; sizes[4] = layout >= 3 ? flat.max_stack : flat.max_depth;
    if (out_nodes) memcpy(out_nodes, layout == 4 ? (const void*)flat.nodesq.data() : layout == 3 ? (const void*)flat.nodes4.data() : layout == MIROGPU_LAYOUT_BVH2 ? (const void*)flat.nodes2.data() : (const void*)flat.nodes8.data(), nb);
    if (out_order) memcpy(out_order, flat.order.data(), flat.order.size() * sizeof(uint32_t));
    if (out_tris) {
        std::vector<TriRecord> tris;
        make_tri_records(tri_vertices, flat.order, tris);
        memcpy(out_tris, tris.data(), tris.size() * sizeof(TriRecord));
    }
    return 0;
}
