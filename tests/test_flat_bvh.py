"""Structure of the flat GPU layouts the product's builder emits (CPU tier, via the emulation library).

The box contract of bvh_build.h -- "its node tests checked against the reference BVH.cpp": the reference pads
every node by epsilon around the plain bounds of its triangles (BVH.cpp:14-38,75-79).  Here every flat box,
at every level, must contain those reference bounds (triangle bounds + epsilon) of every triangle below it,
so the flat tree can never cull a hit the reference's tree would have reached.
"""
import numpy as np
import pytest

import objio
from emu_helpers import emu_build

EPS = np.float32(1e-4)


def _verts(oracle, scenes, name):
    scenes.realise(oracle, name, objio.obj_path)
    return np.ascontiguousarray(oracle.dump_triangles()[:, :9])


def _tri_bounds(V):
    P = V.reshape(-1, 3, 3)
    return P.min(1) - EPS, P.max(1) + EPS


@pytest.mark.parametrize("name", ["testobj", "cornell", "teapot", "bunny_teapot"])
def test_bvh2_boxes_contain_reference_bounds(emu, oracle, scenes, name):
    V = _verts(oracle, scenes, name)
    nodes, order, tris, info = emu_build(emu, V, 0)
    assert sorted(order.tolist()) == list(range(V.shape[0]))            # a permutation: every triangle exactly once
    assert np.array_equal(tris["prim_id"], order)
    assert np.array_equal(tris["a"], V[order, 0:3]) and np.array_equal(tris["e1"], V[order, 3:6] - V[order, 0:3])
    # the stored plane normal is cross(B-A, C-A) in separately rounded binary32 operations (Triangle.cpp:151)
    e1, e2 = tris["e1"], tris["e2"]
    assert np.array_equal(tris["nx"], e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1])
    assert np.array_equal(tris["ny"], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2])
    assert np.array_equal(tris["nz"], e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
    tlo, thi = _tri_bounds(V)
    tlo, thi = tlo[order], thi[order]

    def check(ref, lo, hi):
        """returns the bounds of the subtree, asserting containment on the way up"""
        if ref < 0:
            r = ~ref
            first, count = r >> 3, (r & 7) + 1
            assert count <= 4
            slo, shi = tlo[first:first + count].min(0), thi[first:first + count].max(0)
        else:
            f, link = nodes["f"][ref], nodes["link"][ref]
            parts = []
            for c in range(2):
                clo = np.array([f[4 * c + 0], f[4 * c + 2], f[8 + 2 * c + 0]])
                chi = np.array([f[4 * c + 1], f[4 * c + 3], f[8 + 2 * c + 1]])
                if np.isinf(clo).all():
                    continue  # empty slot (lo = hi = +inf)
                parts.append(check(int(link[c]), clo, chi))
            slo, shi = np.min([p[0] for p in parts], 0), np.max([p[1] for p in parts], 0)
        if lo is not None:
            assert np.all(lo <= slo) and np.all(hi >= shi)
        return slo, shi

    import sys
    sys.setrecursionlimit(10000)
    check(0, None, None)


@pytest.mark.parametrize("name", ["testobj", "cornell", "teapot", "bunny_teapot"])
def test_cwbvh8_boxes_contain_reference_bounds(emu, oracle, scenes, name):
    V = _verts(oracle, scenes, name)
    nodes, order, tris, info = emu_build(emu, V, 1)
    assert sorted(order.tolist()) == list(range(V.shape[0]))
    tlo, thi = _tri_bounds(V)
    tlo, thi = tlo[order], thi[order]
    seen_nodes = np.zeros(len(nodes), bool)
    seen_tris = np.zeros(len(order), bool)

    def visit(ni):
        """bounds (reference-padded) of all triangles under flat node ni; asserts every child box contains its subtree"""
        assert not seen_nodes[ni]
        seen_nodes[ni] = True
        n = nodes[ni]
        cell = np.ldexp(np.float32(1), n["e"].astype(np.int32) - 127).astype(np.float32)
        rank = 0
        los, his = [], []
        for s in range(8):
            meta = int(n["meta"][s])
            if meta == 0:
                assert not (n["imask"] >> s) & 1
                continue
            # the box exactly as the device reconstructs it: p + q * cell in binary32
            clo = n["p"] + np.array([n["qlox"][s], n["qloy"][s], n["qloz"][s]], np.float32) * cell
            chi = n["p"] + np.array([n["qhix"][s], n["qhiy"][s], n["qhiz"][s]], np.float32) * cell
            if (n["imask"] >> s) & 1:
                assert meta == (0x20 | (24 + s))
                slo, shi = visit(int(n["child_base"]) + rank)
                rank += 1
            else:
                cnt = {1: 1, 3: 2, 7: 3}[meta >> 5]
                first = int(n["tri_base"]) + (meta & 31)
                assert not seen_tris[first:first + cnt].any()
                seen_tris[first:first + cnt] = True
                slo, shi = tlo[first:first + cnt].min(0), thi[first:first + cnt].max(0)
            assert np.all(clo <= slo) and np.all(chi >= shi)
            los.append(slo); his.append(shi)
        return np.min(los, 0), np.max(his, 0)

    visit(0)
    assert seen_nodes.all() and seen_tris.all()


def test_cwbvh8_is_smaller_than_bvh2(emu, oracle, scenes):
    V = _verts(oracle, scenes, "bunny_teapot")
    n2, *_ = emu_build(emu, V, 0)
    n8, *_ = emu_build(emu, V, 1)
    assert n8.nbytes < 0.5 * n2.nbytes


@pytest.mark.parametrize("name", ["testobj", "cornell", "teapot", "bunny_teapot"])
def test_bvh4_boxes_contain_reference_bounds(emu, oracle, scenes, name):
    """The 4-wide collapse: every triangle exactly once, every child box contains the reference bounds (triangle bounds +
    epsilon) of everything below it, and the stack need the builder reports fits the kernels' stack."""
    V = _verts(oracle, scenes, name)
    nodes, order, tris, info = emu_build(emu, V, 3)
    assert sorted(order.tolist()) == list(range(V.shape[0]))
    assert info["depth"] <= 96                                            # FlatBvh::max_stack vs MIRO_STACK4
    tlo, thi = _tri_bounds(V)
    tlo, thi = tlo[order], thi[order]
    seen = np.zeros(len(order), bool)

    def check(ref, lo, hi):
        if ref < 0:
            r = ~ref
            first, count = r >> 3, (r & 7) + 1
            assert count <= 4 and not seen[first:first + count].any()
            seen[first:first + count] = True
            slo, shi = tlo[first:first + count].min(0), thi[first:first + count].max(0)
        else:
            nd = nodes[ref]
            parts = []
            for c in range(4):
                clo = np.array([nd["lox"][c], nd["loy"][c], nd["loz"][c]]); chi = np.array([nd["hix"][c], nd["hiy"][c], nd["hiz"][c]])
                if np.isinf(clo).all():
                    continue
                parts.append(check(int(nd["link"][c]), clo, chi))
            assert parts
            slo, shi = np.min([p[0] for p in parts], 0), np.max([p[1] for p in parts], 0)
        if lo is not None:
            assert np.all(lo <= slo) and np.all(hi >= shi)
        return slo, shi

    import sys
    sys.setrecursionlimit(10000)
    check(0, None, None)
    assert seen.all()


@pytest.mark.parametrize("name", ["testobj", "cornell", "teapot", "bunny_teapot"])
def test_qbvh4_decoded_boxes_contain_the_full_precision_boxes(emu, oracle, scenes, name):
    """QBVH4 = the BVH4 tree with child boxes quantised to 8 bits per plane: same topology and links, and every decoded
    box (grid origin + q * 2^(e-127)) contains the full-precision BVH4 box it encodes, with the builder's margin to spare."""
    V = _verts(oracle, scenes, name)
    wide, order, _, _ = emu_build(emu, V, 3)
    quant, order_q, _, _ = emu_build(emu, V, 4)
    assert np.array_equal(order, order_q) and len(wide) == len(quant)
    assert np.array_equal(wide["link"], quant["link"])
    cell = np.ldexp(1.0, quant["e"].astype(np.int64) - 127)                 # (n, 3)
    # the last two words hold 2^24 cell per axis ready-made for the traversal (bvh_build.h, qbvh4_cell_words: x and y as the upper
    # halves of their binary32, z whole)
    grid = quant["origin"].astype(np.float64)
    cw = quant["cell"].astype(np.uint64)
    assert np.array_equal((cw[:, 0] & 0xffff) << 16, (quant["e"][:, 0].astype(np.uint64) + 24) << 23)
    assert np.array_equal(cw[:, 0] & 0xffff0000, (quant["e"][:, 1].astype(np.uint64) + 24) << 23)
    assert np.array_equal(cw[:, 1], (quant["e"][:, 2].astype(np.uint64) + 24) << 23)
    for a, (lo, hi, qlo, qhi) in enumerate((("lox", "hix", "qlox", "qhix"), ("loy", "hiy", "qloy", "qhiy"), ("loz", "hiz", "qloz", "qhiz"))):
        used = np.isfinite(wide[lo])
        dlo = grid[:, a:a + 1] + quant[qlo].astype(np.float64) * cell[:, a:a + 1]
        dhi = grid[:, a:a + 1] + quant[qhi].astype(np.float64) * cell[:, a:a + 1]
        assert (dlo[used] <= wide[lo][used]).all() and (dhi[used] >= wide[hi][used]).all()
        assert (quant[qlo][~used] == 255).all() and (quant[qhi][~used] == 0).all()      # empty slots: inverted interval
        # not needlessly loose: within two cells of the box it encodes
        assert ((wide[lo][used] - dlo[used]) <= 2.05 * np.broadcast_to(cell[:, a:a + 1], used.shape)[used]).all()
