"""GPU tier, SURVEY 8f-4: spheres as leaf primitives of the device tree and planes as its post-walk list, against the oracle
(whose Sphere / Plane restatement is pinned bit for bit to the reference's classes, tests/test_oracle_vs_reference.py).

  * closest hits: ids equal the oracle's exhaustive search, t bit-identical, for every layout and kernel variant;
  * P and N (Sphere.cpp:62-64, Plane.cpp:42-44, then Scene::trace's normalisation) bit-identical to the oracle's Scene::trace;
  * the host layer's Scene::trace / traceBatch (Sphere and Plane objects, BVH::build -> mirogpu_scene_create_ex) agree;
  * Whitted frames with mirror / glass spheres over a plane within the image tolerance of the oracle's Scene::traceScene;
  * a handle replicated on several devices renders the same frame bit for bit (skipped with one GPU).
"""
import ctypes

import numpy as np
import pytest

import miro_driver as md
import objio
from conftest import bits, random_rays, subsample_rays

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def ids_of(hits):
    i = hits["prim_id"].astype(np.int64)
    i[i == 0xFFFFFFFF] = -1
    return i


def build(pkg, scenes, oracle, name, layout=3):
    H = pkg.HostScene(layout)
    scenes.realise(H, name, objio.obj_path)
    H.precalc()
    scenes.realise(oracle, name, objio.obj_path)
    oracle.precalc()
    return H, H.scene()


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
@pytest.mark.parametrize("name", ["spiral", "spheres_teapot"])
def test_sphere_plane_hits_equal_the_oracle(pkg, scenes, oracle, name, layout):
    H, S = build(pkg, scenes, oracle, name, layout)
    w, h = scenes.SCENES[name]["size"]
    rays = np.concatenate([subsample_rays(oracle.eye_rays(w, h), w, h, 2), random_rays(30000, (-3, -2, -3), (3, 4, 3), 5)])
    bounded = rays.copy(); bounded[:, 3] = 0.5; bounded[:, 7] = 6.0
    for rr in (rays, bounded):
        t, ids, P, N = H.trace(rr)                        # host layer: Scene::traceBatch -> device -> Sphere / Plane::fillHit
        bt, bid, bP, bN = oracle.trace_brute(rr)
        assert np.array_equal(ids, bid)
        assert np.array_equal(bits(t), bits(bt))
        assert np.array_equal(bits(P), bits(bP)) and np.array_equal(bits(N), bits(bN))
        rt, rid, _, _ = oracle.trace(rr)                  # the reference's own traversal: same answers here (no ties in these scenes)
        assert (ids != rid).mean() <= 1e-4 and np.array_equal(bits(t[ids == rid]), bits(rt[ids == rid]))
        kinds = np.where(ids >= oracle.num_objects(), 2, 0)
        assert (kinds == 2).any()                        # planes answer some rays
    # every kernel variant, and the device-side resolve (needs the rays: P = o + t d)
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    base = None
    for variant in (-1, 0, 1, 2):
        if variant == 2 and layout == 1:
            continue
        S.set_kernel_variant(variant)
        for mode in (pkg.CLOSEST_HIT, pkg.CLOSEST_HIT | pkg.HINT_COHERENT):
            S.intersect_device(d_rays, d_hits, mode=mode)
            torch.cuda.synchronize()
            got = d_hits.cpu().numpy().copy()
            if base is None:
                base = got
            assert np.array_equal(base.view(np.uint32), got.view(np.uint32))
    S.set_kernel_variant(-1)
    anyh = S.intersect(rays, mode=pkg.ANY_HIT)
    assert np.array_equal(ids_of(anyh) >= 0, ids_of(base.view(pkg.HIT_DTYPE).reshape(-1)) >= 0)
    d_P = torch.empty((rays.shape[0], 3), dtype=torch.float32, device="cuda"); d_N = torch.empty_like(d_P)
    with pytest.raises(pkg.MiroGpuError):
        S.resolve_hits_device(d_hits, d_P, d_N)           # without the rays a sphere's hit point cannot be formed
    S.intersect_device(d_rays, d_hits)
    S.resolve_hits_device(d_hits, d_P, d_N, d_rays=d_rays)
    torch.cuda.synchronize()
    t, ids, P, N = oracle.trace_brute(rays)
    hit = ids >= 0
    assert np.array_equal(bits(d_P.cpu().numpy()[hit]), bits(P[hit])) and np.array_equal(bits(d_N.cpu().numpy()[hit]), bits(N[hit]))


@pytest.mark.parametrize("name", ["spiral", "spheres_teapot"])
def test_frames_with_spheres_and_planes(pkg, scenes, oracle, name):
    H, S = build(pkg, scenes, oracle, name)
    w, h = 192, 160
    sc = scenes.SCENES[name]
    p = S.render_params(w, h, mode=pkg.RENDER_WHITTED, tonemap=0, bg=sc.get("bg", (0, 0, 0)))
    img = S.render(H.camera(), p)
    ref = oracle.trace_scene(oracle.eye_rays(w, h), depth=10).reshape(h, w, 3)
    close = np.isclose(img, ref, rtol=2e-3, atol=2e-4)
    assert close.all(axis=2).mean() > 0.99, close.all(axis=2).mean()
    # the diffuse-bounce estimator (fused two-wave path) runs on these primitives too and equals the general wavefront
    import os
    q = S.render_params(w, h, mode=pkg.RENDER_DIFFUSE_BOUNCE, jitter=1, spp=2, seed=3, shadows=0, bg=sc.get("bg", (0, 0, 0)))
    fused = S.render(H.camera(), q)
    os.environ["MIROGPU_GENERAL_WAVEFRONT"] = "1"
    try:
        general = S.render(H.camera(), q)
    finally:
        del os.environ["MIROGPU_GENERAL_WAVEFRONT"]
    # equal up to subnormal terms: a channel with kd = 0 holds only the highlight, pow(c, 500) ~ 1e-40, which the general path's
    # float atomics (RED.ADD.F32.FTZ) flush to zero and the fused path's plain additions keep, like the reference's x86 arithmetic
    assert np.allclose(fused, general, rtol=0, atol=1e-30)
    big = np.abs(general) > 1e-30
    assert np.array_equal(bits(fused)[big], bits(general)[big])


def test_path_queue_overflow_is_retried_not_dropped(pkg, scenes, oracle):
    """Refractive hits spawn up to three children per level (Scene.cpp:301-335); the wavefront's queues start at a multiple of the
    primary items and must grow when a wave overflows them -- never drop children.  Started at 1x (far too small for a glass
    sphere at depth 10) the frame must equal the one rendered with roomy queues (same rays, same sums up to atomic order)."""
    import os
    name, (w, h) = "spheres_teapot", (224, 160)
    frames = []
    for mult in ("1", "64"):
        os.environ["MIROGPU_QUEUE_MULT"] = mult
        try:
            H, S = build(pkg, scenes, oracle, name)
        finally:
            del os.environ["MIROGPU_QUEUE_MULT"]
        p = S.render_params(w, h, mode=pkg.RENDER_WHITTED, tonemap=0, bg=(0.2, 0.3, 0.5), max_depth=10)
        frames.append((S.render(H.camera(), p), S.last_call_stats()[0]))
    # several path items land on one pixel here and are summed by float atomics in arrival order: equal up to that rounding
    assert np.allclose(frames[0][0], frames[1][0], rtol=2e-5, atol=1e-7)
    assert frames[0][1] == frames[1][1] > 2 * w * h        # exactly the same rays traced, secondary ones included


def test_multi_device_handle_renders_the_same_frame(pkg, scenes, oracle):
    n = pkg.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    name = "spheres_teapot"
    w, h = 320, 203
    frames = {}
    for ndev in (1, min(n, 2), min(n, 8)):
        H = pkg.HostScene()
        scenes.realise(H, name, objio.obj_path)
        H.set_device_count(ndev)
        H.precalc()
        S = H.scene()
        assert S.devices() == list(range(ndev))
        for mode, spp in ((pkg.RENDER_WHITTED, 1), (pkg.RENDER_DIFFUSE_BOUNCE, 3)):
            p = S.render_params(w, h, mode=mode, jitter=1 if spp > 1 else 0, spp=spp, seed=9, shadows=1 if mode == pkg.RENDER_WHITTED else 0, bg=(0.2, 0.3, 0.5))
            f32 = S.render(H.camera(), p)
            u8 = S.render_rgb8(H.camera(), p)
            frames.setdefault((mode, "f32"), []).append(f32)
            frames.setdefault((mode, "u8"), []).append(u8)
    for (mode, kind), fs in frames.items():
        for f in fs[1:]:
            if mode == pkg.RENDER_DIFFUSE_BOUNCE:
                assert np.array_equal(fs[0], f), (mode, kind)              # one owner per (pixel, sample) slot: deterministic sums
            elif kind == "f32":
                assert np.allclose(fs[0], f, rtol=2e-5, atol=1e-7)           # mirror / glass: several items per pixel, atomics in arrival order
            else:
                assert np.abs(fs[0].astype(int) - f.astype(int)).max() <= 1
