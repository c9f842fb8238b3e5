"""CPU tier: the product's traversal cores and builder, executed on the host through the test-only emulation
library, against the oracle.  The device kernels call the same __host__ __device__ functions, so this pins the
flattening and traversal LOGIC before a GPU is involved; the GPU tier (test_gpu_parity.py) repeats the
comparison through the C ABI on the real kernels.

Contract: for every ray the result equals the oracle's exhaustive search with the reference's triangle test
(ties in t to the smaller primitive id) -- bit for bit in t, beta, gamma, identical ids.
"""
import numpy as np
import pytest

import objio
from conftest import bits, random_rays, subsample_rays
from emu_helpers import emu_trace, hit_ids

LAYOUTS = [0, 1, 2, 3, 4, 5]   # BVH2, CWBVH8, BVH2 via the single-step functions of the hybrid kernel, BVH4, QBVH4, QBVH4 via the default kernel's steps (one triangle per leaf step)


def _scene(oracle, scenes, name):
    scenes.realise(oracle, name, objio.obj_path)
    oracle.precalc()
    tri = oracle.dump_triangles()
    return np.ascontiguousarray(tri[:, :9])


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("name", ["testobj", "cornell", "teapot"])
def test_small_scenes_equal_exhaustive_search(emu, oracle, scenes, name, layout):
    V = _scene(oracle, scenes, name)
    w, h = 96, 96
    rays = np.concatenate([oracle.eye_rays(w, h), random_rays(6000, V.reshape(-1, 3).min(0), V.reshape(-1, 3).max(0), 11)])
    bt, bid, bP, bN = oracle.trace_brute(rays)
    hits, cnt, info = emu_trace(emu, V, rays, layout)
    assert np.array_equal(hit_ids(hits), bid)
    assert np.array_equal(bits(hits["t"]), bits(bt))
    assert cnt["tris"] > 0


@pytest.mark.parametrize("layout", LAYOUTS)
def test_bunny_teapot_equal_exhaustive_search_and_reference_semantics(emu, oracle, scenes, layout):
    V = _scene(oracle, scenes, "bunny_teapot")
    rays = np.concatenate([subsample_rays(oracle.eye_rays(512, 512), 512, 512, 8), random_rays(2000, [-6, 0, -3], [3, 4, 4], 5)])
    bt, bid, _, _ = oracle.trace_brute(rays)
    hits, cnt, info = emu_trace(emu, V, rays, layout)
    assert np.array_equal(hit_ids(hits), bid)
    assert np.array_equal(bits(hits["t"]), bits(bt))
    # against the reference's own BVH traversal order (Scene::trace): identical except documented tie / cull cases
    rt, rid, _, _ = oracle.trace(rays)
    mism = int((hit_ids(hits) != rid).sum())
    assert mism <= max(1, int(1e-4 * rays.shape[0])), mism


@pytest.mark.parametrize("layout", LAYOUTS)
def test_any_hit_agrees_with_closest_hit(emu, oracle, scenes, layout):
    V = _scene(oracle, scenes, "teapot")
    rays = random_rays(5000, [-4, 0, -3], [4, 3, 3], 23)
    rays[:, 7] = np.random.default_rng(1).random(5000, dtype=np.float32) * 6  # bounded tmax like shadow rays
    closest, _, _ = emu_trace(emu, V, rays, layout)
    anyh, _, _ = emu_trace(emu, V, rays, layout, any_hit=True)
    assert np.array_equal(hit_ids(closest) >= 0, hit_ids(anyh) >= 0)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_edge_cases(emu, layout):
    # empty scene
    rays = random_rays(64, [-1, -1, -1], [1, 1, 1], 3)
    hits, _, _ = emu_trace(emu, np.zeros((0, 9), np.float32), rays, layout)
    assert (hit_ids(hits) == -1).all() and np.array_equal(hits["t"], rays[:, 7])
    # one triangle, axis-parallel rays (zero direction components), tmin/tmax windows, degenerate triangle
    V = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0], [5, 5, 5, 5, 5, 5, 5, 5, 5]], np.float32)
    r = np.zeros((5, 8), np.float32)
    r[:, 0:3] = [0.25, 0.25, 1.0]; r[:, 4:7] = [0, 0, -1]; r[:, 7] = 1e12
    r[1, 7] = 0.5            # tmax in front of the triangle -> miss
    r[2, 3] = 1.5            # tmin behind it -> miss
    r[3, 4:7] = [0, 0, 1]    # pointing away -> miss
    r[4, 7] = 1.0            # t == tmax exactly -> hit (the reference's range test is inclusive, Triangle.cpp:158)
    hits, _, _ = emu_trace(emu, V, r, layout)
    assert list(hit_ids(hits)) == [0, -1, -1, -1, 0]
    assert hits["t"][0] == 1.0 and hits["t"][1] == 0.5 and hits["t"][4] == 1.0
    # the epsilon slop: a ray just outside an edge (beta = -5e-5) is accepted, one further out (-2e-4) is not
    r2 = np.zeros((2, 8), np.float32)
    r2[:, 0:3] = [[-5e-5, 0.3, 1.0], [-2e-4, 0.3, 1.0]]; r2[:, 4:7] = [0, 0, -1]; r2[:, 7] = 1e12
    hits, _, _ = emu_trace(emu, V, r2, layout)
    assert list(hit_ids(hits)) == [0, -1]


@pytest.mark.parametrize("layout", LAYOUTS)
def test_duplicate_geometry_ties_go_to_smaller_id(emu, layout):
    tri = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)
    V = np.stack([tri] * 7)
    r = np.zeros((1, 8), np.float32); r[0, 0:3] = [0.2, 0.2, 2]; r[0, 4:7] = [0, 0, -1]; r[0, 7] = 1e12
    hits, _, _ = emu_trace(emu, V, r, layout)
    assert hit_ids(hits)[0] == 0 and hits["t"][0] == 2.0
