"""Pins the oracle (oracle/miro_oracle.cpp) against the reference itself.

(1) The published known-answer test: the reference's BVH over bunny.obj has 42 881 nodes / 21 441 leaves
    (writeup/A2/Readme.tex:96).  (2) Where oracle/_ref exists (the unmodified reference compiled in place),
    every output of the restatement -- loader, eye rays, Scene::trace hits, -DSTATS counters, shading,
    photon-map balance and gather -- must be bit-identical on identical inputs.
"""
import ctypes

import numpy as np
import pytest

import miro_driver as md
import objio
from conftest import bits, random_rays, subsample_rays


def _both(reference, oracle, scenes, name):
    for d in (reference, oracle):
        scenes.realise(d, name, objio.obj_path)
        d.precalc()


def test_kat_bunny_node_counts(oracle, scenes):
    scenes.realise(oracle, "bunny1", objio.obj_path)
    oracle.precalc()
    st = oracle.stats()
    assert (st["nodes"], st["leaves"]) == (42881, 21441)


def test_kat_bunny_node_counts_reference(reference_stats, scenes):
    scenes.realise(reference_stats, "bunny1", objio.obj_path)
    reference_stats.precalc()
    st = reference_stats.stats()
    assert (st["nodes"], st["leaves"]) == (42881, 21441)


@pytest.mark.parametrize("name,step", [("testobj", 1), ("cornell", 4), ("teapot", 4), ("bunny_teapot", 8)])
def test_trace_bit_exact(reference_stats, oracle, scenes, name, step):
    R, O = reference_stats, oracle
    _both(R, O, scenes, name)
    assert np.array_equal(bits(R.dump_triangles()), bits(O.dump_triangles()))
    sr, so = R.stats(), O.stats()
    assert (sr["nodes"], sr["leaves"]) == (so["nodes"], so["leaves"])
    w, h = scenes.SCENES[name]["size"]
    w, h = min(w, 512), min(h, 512)
    rays_o = O.eye_rays(w, h)
    rays = subsample_rays(rays_o, w, h, step)
    R.stats_reset_rays(); O.stats_reset_rays()
    a, b = R.trace(rays, 1), O.trace(rays, 1)
    assert np.array_equal(a[1], b[1])
    for x, y in zip((a[0], a[2], a[3]), (b[0], b[2], b[3])):
        assert np.array_equal(bits(x), bits(y))
    sr, so = R.stats(), O.stats()
    assert (sr["ray_box"], sr["ray_tri"]) == (so["ray_box"], so["ray_tri"])
    # incoherent second-generation rays from the hits
    u = np.random.default_rng(7).random((rays.shape[0], 2), dtype=np.float32)
    br = np.zeros_like(rays)
    O.lib.orc_diffuse_rays(md._fp(b[2]), md._fp(b[3]), md._fp(b[1]), md._fp(u), ctypes.c_long(rays.shape[0]), md._fp(br))
    a2, b2 = R.trace(br, 1), O.trace(br, 1)
    assert np.array_equal(a2[1], b2[1]) and np.array_equal(bits(a2[0]), bits(b2[0]))


def test_shading_bit_exact(reference, oracle, scenes):
    """Scene::traceScene (Phong::shade + shadow query) on the cornell scene, pixel-centre rays."""
    _both(reference, oracle, scenes, "cornell")
    rays = subsample_rays(oracle.eye_rays(128, 128), 128, 128, 2)
    a = reference.trace_scene(rays, depth=10, nthreads=1)
    b = oracle.trace_scene(rays, depth=10, nthreads=1)
    assert np.array_equal(bits(a), bits(b))
    assert a.max() > 0


def test_photon_map_bit_exact(reference, oracle):
    rng = np.random.default_rng(3)
    n = 5000
    pos = rng.random((n, 3), dtype=np.float32) * 4
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pw = rng.random((n, 3), dtype=np.float32)
    reference.new_scene(); oracle.new_scene()
    wr, wo = reference.pm_new(n), oracle.pm_new(n)
    for D, w in ((reference, wr), (oracle, wo)):
        D.pm_store(w, pw, pos, d); D.pm_scale(w, 1.0 / n); D.pm_balance(w)
    pr, po = reference.pm_dump(wr), oracle.pm_dump(wo)
    half = n // 2 - 1
    for f in ("pos", "power", "theta", "phi"):
        assert np.array_equal(pr[f][1:], po[f][1:]), f
    assert np.array_equal(pr["plane"][1:half], po["plane"][1:half])
    q = rng.random((300, 3), dtype=np.float32) * 4
    qn = rng.normal(size=(300, 3)).astype(np.float32); qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    for k in (1, 50, 500):
        a = reference.pm_irradiance(wr, q, qn, 1e10, k, 1)
        b = oracle.pm_irradiance(wo, q, qn, 1e10, k, 1)
        assert np.array_equal(bits(a), bits(b))
    a = reference.pm_irradiance(wr, q, qn, 0.5, 50, 1)
    b = oracle.pm_irradiance(wo, q, qn, 0.5, 50, 1)
    assert np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("name", ["spiral", "spheres_teapot"])
def test_spheres_and_planes_bit_exact(reference_stats, oracle, scenes, name):
    """Sphere::intersect (Sphere.cpp:28-69) inside the tree, Plane::intersect (Plane.cpp:33-48) in Scene::trace's unbounded
    loop (Scene.cpp:219-230): restatement against the reference's own classes -- hits, P, N, counters, shaded radiance."""
    R, O = reference_stats, oracle
    _both(R, O, scenes, name)
    sr, so = R.stats(), O.stats()
    assert (sr["nodes"], sr["leaves"]) == (so["nodes"], so["leaves"])
    w, h = scenes.SCENES[name]["size"]
    rays = subsample_rays(O.eye_rays(w, h), w, h, 2)
    R.stats_reset_rays(); O.stats_reset_rays()
    a, b = R.trace(rays, 1), O.trace(rays, 1)
    assert np.array_equal(a[1], b[1])
    nbounded = O.num_objects()
    kinds = set(np.unique(np.where(b[1] >= nbounded, 2, np.where(b[1] < 0, -1, 0))))
    assert 2 in kinds and 0 in kinds          # planes and tree primitives are both hit
    for x, y in zip((a[0], a[2], a[3]), (b[0], b[2], b[3])):
        assert np.array_equal(bits(x), bits(y))
    sr, so = R.stats(), O.stats()
    assert (sr["ray_box"], sr["ray_tri"]) == (so["ray_box"], so["ray_tri"])
    # incoherent rays from inside the scene (spheres entered from within: the far root)
    rr = random_rays(20000, (-3, -2, -3), (3, 4, 3), 11)
    a, b = R.trace(rr, 1), O.trace(rr, 1)
    assert np.array_equal(a[1], b[1])
    for x, y in zip((a[0], a[2], a[3]), (b[0], b[2], b[3])):
        assert np.array_equal(bits(x), bits(y))
    # bounded ranges exercise the strict (tMin, tMax) of the sphere and the inclusive one of the plane
    rr2 = rr.copy(); rr2[:, 3] = 0.5; rr2[:, 7] = 6.0
    a, b = R.trace(rr2, 1), O.trace(rr2, 1)
    assert np.array_equal(a[1], b[1]) and np.array_equal(bits(a[0]), bits(b[0]))
    # shaded radiance through Scene::traceScene (reflection / refraction on the spheres of spheres_teapot)
    sub = subsample_rays(O.eye_rays(w, h), w, h, 8)
    ra = R.trace_scene(sub, depth=10, nthreads=1)
    rb = O.trace_scene(sub, depth=10, nthreads=1)
    assert np.array_equal(bits(ra), bits(rb))
