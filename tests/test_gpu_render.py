"""GPU tier: whole-frame rendering (Scene::raytraceImage on the device) and the photon-map gather, through the
C ABI and the host API layer, against the oracle's restatement of Scene::traceScene / Phong::shade /
Photon_map::irradiance_estimate (pinned bit-exactly to the real reference in test_oracle_vs_reference.py).

Tolerances: images max |diff| <= 2/255 per channel after the tone map for the deterministic
configs; PSNR >= 40 dB where secondary rays make isolated pixels flip across geometric edges; photon gather
bit-identical (the device walk keeps the reference's visiting order and heap, so even the summation order is
the same).
"""
import ctypes

import numpy as np
import pytest

import miro_driver as md
import objio
from conftest import bits

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def tonemap_u8(O, rgb):
    out = np.zeros(rgb.shape, np.uint8)
    O.lib.orc_tonemap(md._fp(np.ascontiguousarray(rgb, np.float32)), ctypes.c_long(rgb.shape[0] * rgb.shape[1] if rgb.ndim == 3 else rgb.shape[0]), md._fp(out))
    return out


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def build_pair(pkg, scenes, oracle, name, layout=1):
    H = pkg.HostScene(layout)
    for d in (oracle, H):
        scenes.realise(d, name, objio.obj_path)
        d.precalc()
    return H, H.scene()


@pytest.mark.parametrize("name,size", [("cornell", (160, 120)), ("teapot", (128, 128)), ("bunny_teapot", (256, 256))])
def test_whitted_render_matches_reference_semantics(pkg, scenes, oracle, name, size):
    H, S = build_pair(pkg, scenes, oracle, name)
    w, h = size
    p = S.render_params(w, h, mode=pkg.RENDER_WHITTED, tonemap=0)
    img = S.render(H.camera(), p)
    ref = oracle.trace_scene(oracle.eye_rays(w, h), depth=10).reshape(h, w, 3)
    close = np.isclose(img, ref, rtol=2e-4, atol=2e-5)
    assert close.mean() > 0.9995, close.mean()
    a, b = tonemap_u8(oracle, img), tonemap_u8(oracle, ref)
    diff = np.abs(a.astype(int) - b.astype(int))
    assert (diff <= 2).mean() > 0.9995 and psnr(a, b) >= 40
    # the host layer's Scene::raytraceImage (device tone map + Image::setPixel) gives the same 8-bit image
    H.set_render(spp=1, jitter=0, mode=pkg.RENDER_WHITTED, shadows=1)
    u8 = H.render(w, h)
    assert (np.abs(u8.astype(int) - b.astype(int)) <= 2).mean() > 0.9995
    # the same 8-bit frame straight from the C ABI, and the device tone map of a float frame
    u8b = S.render_rgb8(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, tonemap=1))
    assert np.array_equal(u8, u8b)
    d_rgb = torch.from_numpy(img).cuda(); d_u8 = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    S.tonemap_rgb8_device(d_rgb, d_u8)
    torch.cuda.synchronize()
    assert np.array_equal(d_u8.cpu().numpy(), u8b)
    # the tone map of a frame sharded by rows: per-shard maxima, combined (what the ranks all-reduce), then per-shard mapping
    d_rgb[h // 3, w // 2, 1] = float("nan")                    # a NaN pixel takes the frame-wide maximum (Scene.cpp:157-164)
    S.tonemap_rgb8_device(d_rgb, d_u8)
    whole = d_u8.cpu().numpy().copy()
    d_m = torch.empty(3, dtype=torch.float32, device="cuda"); d_u8.zero_()
    shards = [(0, h, 3, r) for r in range(3)]
    for r, rows in enumerate(shards):
        S.frame_max_device(d_rgb, rows, d_m[r:r + 1])
    d_all = d_m.max().reshape(1)
    assert float(d_all) == float(np.nanmax(d_rgb.cpu().numpy()))
    for rows in shards:
        S.tonemap_rows_rgb8_device(d_rgb, rows, d_all, d_u8)
    torch.cuda.synchronize()
    assert np.array_equal(d_u8.cpu().numpy(), whole)
    # no shadows == the reference's -DDISABLE_SHADOWS build: brighter or equal everywhere
    p2 = S.render_params(w, h, mode=pkg.RENDER_WHITTED, tonemap=0, shadows=0)
    img2 = S.render(H.camera(), p2)
    assert (img2 >= img - 1e-6).all() and (img2 > img + 1e-4).any()


def test_render_row_shards_are_bit_identical_to_the_full_frame(pkg, scenes, oracle):
    H, S = build_pair(pkg, scenes, oracle, "teapot")
    w, h = 200, 150
    for mode in (pkg.RENDER_WHITTED, pkg.RENDER_DIFFUSE_BOUNCE):
        full = S.render(H.camera(), S.render_params(w, h, mode=mode, jitter=1, spp=2, seed=5))
        for nshard in (2, 3, 8):
            acc = np.full((h, w, 3), np.nan, np.float32)
            for r in range(nshard):
                S.render(H.camera(), S.render_params(w, h, mode=mode, jitter=1, spp=2, seed=5, rows=(0, h, nshard, r)), out=acc)
            assert np.array_equal(bits(acc), bits(full))
        # contiguous blocks of rows as well
        acc = np.full((h, w, 3), np.nan, np.float32)
        for (a, b) in ((0, 40), (40, 41), (41, 150)):
            S.render(H.camera(), S.render_params(w, h, mode=mode, jitter=1, spp=2, seed=5, rows=(a, b, 1, 0)), out=acc)
        assert np.array_equal(bits(acc), bits(full))


def test_diffuse_bounce_render(pkg, scenes, oracle):
    """BASELINE config 3's estimator on small geometry: jittered primary + one Ray::diffuse bounce, radiance =
    direct(P0) + kd * direct(P1), rebuilt from oracle pieces with the same uniforms."""
    H, S = build_pair(pkg, scenes, oracle, "teapot")
    w, h = 160, 120
    n = w * h
    p = S.render_params(w, h, mode=pkg.RENDER_DIFFUSE_BOUNCE, jitter=1, spp=1, seed=168, shadows=1)
    img = S.render(H.camera(), p).reshape(n, 3)
    assert S.last_call_stats()[0] >= 2 * n                                   # primaries + shadow + bounce rays were traced
    uj = pkg.rng_uniforms(168, 0, 0, 0, n)
    rays = np.zeros((n, 8), np.float32)
    oracle.lib.orc_eye_rays_jitter(w, h, md._fp(uj), md._fp(rays))
    t, ids, P, N = oracle.trace(rays)
    d0 = oracle.trace_scene(rays, depth=0)
    ub = pkg.rng_uniforms(168, 0, 1, 0, n)
    brays = np.zeros((n, 8), np.float32)
    oracle.lib.orc_diffuse_rays(md._fp(P), md._fp(N), md._fp(ids), md._fp(ub), ctypes.c_long(n), md._fp(brays))
    d1 = oracle.trace_scene(brays, depth=0)
    d1[ids < 0] = 0
    ref = d0 + d1          # kd = 1
    close = np.isclose(img, ref, rtol=1e-3, atol=1e-4)
    assert close.all(axis=1).mean() > 0.998, close.all(axis=1).mean()


def test_fused_bounce_waves_equal_the_general_wavefront(pkg, scenes, oracle):
    """MIROGPU_RENDER_DIFFUSE_BOUNCE without shadow rays runs as two fused waves (k_bounce_wave0 / k_bounce_wave1: implicit
    pixel / weight, camera ray recomputed, plain stores into the per-sample planes).  Frames must equal the general
    wavefront's bit for bit -- full frame, several sample batches, row shards -- and the reported ray count is primary +
    LIVE bounce rays (dead slots of missed pixels are not rays)."""
    import os
    for name, (w, h) in (("teapot", (200, 150)), ("cornell", (96, 64))):
        H, S = build_pair(pkg, scenes, oracle, name)
        cam = H.camera()
        for spp in (1, 3, 20):
            p = S.render_params(w, h, mode=pkg.RENDER_DIFFUSE_BOUNCE, jitter=1, spp=spp, seed=7, shadows=0, bg=(0.1, 0.2, 0.3))
            fused = S.render(cam, p)
            rays_fused = S.last_call_stats()[0]
            os.environ["MIROGPU_GENERAL_WAVEFRONT"] = "1"
            try:
                general = S.render(cam, p)
                rays_general = S.last_call_stats()[0]
            finally:
                del os.environ["MIROGPU_GENERAL_WAVEFRONT"]
            assert np.array_equal(bits(fused), bits(general))
            assert rays_fused == rays_general
            # primary + live bounce rays, counted independently: one bounce ray per camera ray that hit something (kd > 0 everywhere)
            n = w * h * spp
            d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
            S.generate_primary(cam, w, h, d_rays, jitter=1, seed=7, sample=0, samples=spp)
            S.intersect_device(d_rays, d_hits)
            torch.cuda.synchronize()
            live = int((d_hits.view(torch.int32)[:, 1] != -1).sum().item())
            assert rays_fused == n + live
        p = S.render_params(w, h, mode=pkg.RENDER_DIFFUSE_BOUNCE, jitter=1, spp=3, seed=7, shadows=0, bg=(0.1, 0.2, 0.3))
        full = S.render(cam, p)
        acc = np.full((h, w, 3), np.nan, np.float32)
        for r in range(3):
            S.render(cam, S.render_params(w, h, mode=pkg.RENDER_DIFFUSE_BOUNCE, jitter=1, spp=3, seed=7, shadows=0, bg=(0.1, 0.2, 0.3), rows=(0, h, 3, r)), out=acc)
        assert np.array_equal(bits(acc), bits(full))
        u8 = S.render_rgb8(cam, p)
        assert np.abs(u8.astype(int) - tonemap_u8(oracle, full).astype(int)).max() <= 1   # expf on the device vs libm


def test_specular_and_refractive_materials(pkg, scenes, oracle):
    """Reflection / Fresnel / refraction recursion (Scene.cpp:302-336) and the refractive-occluder shadow rule
    (Phong.cpp:99-113): cornell box + a glass sphere + a mirror teapot."""
    H = pkg.HostScene()
    T = scenes.translate
    for d in (oracle, H):
        d.new_scene()
        d.new_material((1, 1, 1), (0, 0, 0), (0, 0, 0), 1.0, 1.0)
        d.new_material((0.1, 0.1, 0.1), (0, 0, 0), (1, 1, 1), 5.0, 1.5)        # glass: Phong(kd, 0, 1, 5, 1.5)
        d.new_material((0.2, 0.2, 0.2), (0.8, 0.8, 0.8), (0, 0, 0), -1.0, 1.0)  # mirror, infinite shininess
        d.add_obj(objio.obj_path("cornell_box"), None, 0)
        d.add_obj(objio.obj_path("sphere"), (T(1.5, 1.2, -1.0) @ scenes.scale(0.8, 0.8, 0.8)).astype(np.float32), 1)
        d.add_obj(objio.obj_path("teapot"), (T(3.6, 0.0, -2.0) @ scenes.scale(0.6, 0.6, 0.6)).astype(np.float32), 2)
        d.add_point_light((2.5, 4.9, -1), (1, 1, 1), 160)
        d.set_bg_color((0.0, 0.0, 0.2))
        d.set_camera((2.5, 3, 3), (2.5, 2.5, 0), (0, 1, 0), 90)
        d.precalc()
    S = H.scene()
    w, h = 192, 144
    img = S.render(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, max_depth=10, bg=(0, 0, 0.2)))
    ref = oracle.trace_scene(oracle.eye_rays(w, h), depth=10).reshape(h, w, 3)
    assert np.isfinite(img).mean() > 0.999
    a, b = tonemap_u8(oracle, np.nan_to_num(img)), tonemap_u8(oracle, np.nan_to_num(ref))
    diff = np.abs(a.astype(int) - b.astype(int)).max(axis=2)
    assert (diff <= 2).mean() > 0.99, (diff <= 2).mean()
    assert psnr(a, b) >= 35, psnr(a, b)
    assert S.last_call_stats()[0] > 2 * w * h + 1000     # primaries + shadow slots + the secondary generations


def _photon_cloud(n, seed):
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, 3), dtype=np.float32) * np.float32(5)).astype(np.float32)
    pos[:, 1] *= 0.02                                      # mostly on a floor, like stored photons
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pw = rng.random((n, 3), dtype=np.float32)
    return pw, pos, d


def knn_estimate(ph, q, qn, max_dist, k):
    """Irradiance by exhaustive search in numpy: the k nearest photons that pass the direction filter (PhotonMap.cpp:183-186)
    among those the reference's walk can reach, summed and divided by pi r_k^2 -- or, when at most k exist inside max_dist, all
    of them over pi max_dist^2 (the reference's radius only shrinks once its k-set overflows, PhotonMap.cpp:211-239)."""
    stored = len(ph) - 1
    half = stored // 2 - 1
    reach = max(1, 2 * half - 1) if half >= 1 else 1          # children are only followed below half_stored (PhotonMap.cpp:161)
    P = ph["pos"][1:reach + 1].astype(np.float32); W = ph["power"][1:reach + 1].astype(np.float32)
    ang = np.arange(256, dtype=np.float64) * (1.0 / 256.0) * np.pi
    ct, st_, cp, sp_ = np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32), np.cos(2 * ang).astype(np.float32), np.sin(2 * ang).astype(np.float32)
    th, phi = ph["theta"][1:reach + 1], ph["phi"][1:reach + 1]
    D = np.stack([st_[th] * cp[phi], st_[th] * sp_[phi], ct[th]], axis=1).astype(np.float32)
    out = np.zeros((len(q), 3), np.float32)
    r2max = np.float32(max_dist) * np.float32(max_dist)
    for i in range(len(q)):
        dx = P[:, 0] - q[i, 0]; dy = P[:, 1] - q[i, 1]; dz = P[:, 2] - q[i, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        dots = (D[:, 0] * qn[i, 0] + D[:, 1] * qn[i, 1]) + D[:, 2] * qn[i, 2]
        ok = (d2 < r2max) & (dots < 0)
        cand = np.nonzero(ok)[0]
        if len(cand) > k:
            order = cand[np.argsort(d2[cand], kind="stable")[:k]]
            r2 = d2[order[-1]]
        else:
            order, r2 = cand, r2max
        tmp = np.float32((1.0 / np.pi) / np.float64(r2))
        out[i] = W[order].astype(np.float64).sum(0).astype(np.float32) * tmp
    return out


@pytest.mark.parametrize("nphot", [1, 7, 1000, 60000])
def test_photon_gather(pkg, scenes, oracle, nphot):
    """exact mode (one query per thread, the reference's search verbatim): bit-identical to the oracle.
    default mode (one query per warp): the k nearest photons, checked against an exhaustive search; it equals the reference's
    result except where the reference's own quirk bites -- its first overflow evicts the farthest of the first k photons it
    visited even when the newcomer is farther (PhotonMap.cpp:211-239), which loses a true neighbour exactly when those first k
    were the k nearest: common for k = 1, never seen for k = 500 on 200 000 photons."""
    H, S = build_pair(pkg, scenes, oracle, "testobj")
    pw, pos, d = _photon_cloud(nphot, 4)
    w = oracle.pm_new(nphot)
    oracle.pm_store(w, pw, pos, d); oracle.pm_scale(w, 1.0 / nphot); oracle.pm_balance(w)
    ph = oracle.pm_dump(w)
    S.photon_upload(0, ph)
    rng = np.random.default_rng(5)
    nq = 3000 if nphot <= 1000 else 600
    q = (rng.random((nq, 3), dtype=np.float32) * np.float32(5)).astype(np.float32); q[:, 1] *= 0.02
    qn = np.tile(np.array([[0, 1, 0]], np.float32), (nq, 1))
    qn[::3] = rng.normal(size=(len(qn[::3]), 3)).astype(np.float32)
    for k, md_ in ((500, 1e10), (50, 1e10), (1, 1e10), (50, 0.3), (512, 1e10)):
        a = oracle.pm_irradiance(w, q, qn, md_, k)
        S.photon_set_exact(0, True)
        b = S.photon_gather(0, q, qn, md_, k)
        assert np.array_equal(bits(a), bits(b)), (nphot, k, md_)
        S.photon_set_exact(0, False)
        c = S.photon_gather(0, q, qn, md_, k)
        e = knn_estimate(ph, q, qn, md_, k)
        assert np.allclose(c, e, rtol=2e-5, atol=0), (nphot, k, md_, float(np.abs(c - e).max()))
        agree = np.isclose(c, a, rtol=2e-5, atol=0).all(axis=1).mean()
        assert agree > (0.999 if k >= 500 else 0.5), (nphot, k, md_, agree)
    assert oracle.pm_irradiance(w, q, qn, 1e10, 50).max() > 0 or nphot < 50
    # through the host layer's Photon_map (store / scale / balance on the host, gather on the device)
    H.pm_store(1, pw, pos, d); H.pm_scale(1, 1.0 / nphot); H.pm_balance(1); H.pm_attach(1)
    S.photon_set_exact(1, True)
    c = H.pm_irradiance(1, q, qn, 1e10, 100)
    assert np.array_equal(bits(c), bits(oracle.pm_irradiance(w, q, qn, 1e10, 100)))


def test_photon_gather_neighbouring_queries(pkg, scenes, oracle):
    """Neighbouring queries (pixel order) seed each other's search radius inside a chunk: same k nearest photons as the
    exhaustive search, for smooth runs, for runs whose normal flips (the seed must be dropped) and for k-sets that tie the
    photon count (a seeded walk that ends with exactly k photons is repeated unseeded)."""
    H, S = build_pair(pkg, scenes, oracle, "testobj")
    nphot = 30000
    pw, pos, d = _photon_cloud(nphot, 11)
    w = oracle.pm_new(nphot)
    oracle.pm_store(w, pw, pos, d); oracle.pm_scale(w, 1.0 / nphot); oracle.pm_balance(w)
    ph = oracle.pm_dump(w)
    S.photon_upload(0, ph)
    S.photon_set_exact(0, False)
    rng = np.random.default_rng(12)
    nq = 640
    t = np.arange(nq, dtype=np.float32)
    q = np.stack([0.5 + 0.006 * t, 0.01 + 0.0 * t, 1.0 + 0.004 * t], axis=1).astype(np.float32)      # a scanline across the floor
    qn = np.tile(np.array([[0, 1, 0]], np.float32), (nq, 1)) + rng.normal(size=(nq, 3)).astype(np.float32) * np.float32(1e-3)
    qn[200:230] *= -1                 # a stretch seen from below: other photons face it
    qn[400] = [1, 0, 0]
    for k, md_ in ((500, 1e10), (60, 1e10), (60, 0.05), (1, 1e10)):
        c = S.photon_gather(0, q, qn, md_, k)
        e = knn_estimate(ph, q, qn, md_, k)
        assert np.allclose(c, e, rtol=2e-5, atol=0), (k, md_, float(np.abs(c - e).max()))
        assert c.max() > 0
    # exactly k facing photons in the whole map: never an overflow, the estimate divides by max_dist^2
    few = 40
    pw2, pos2, d2 = _photon_cloud(few, 13)
    d2[:, 1] = -np.abs(d2[:, 1]) - 0.1
    w2 = oracle.pm_new(few)
    oracle.pm_store(w2, pw2, pos2, d2); oracle.pm_scale(w2, 1.0 / few); oracle.pm_balance(w2)
    ph2 = oracle.pm_dump(w2)
    S.photon_upload(1, ph2)
    qn2 = np.tile(np.array([[0, 1, 0]], np.float32), (nq, 1))
    e = knn_estimate(ph2, q, qn2, 1e3, few)
    reach = int((e != 0).any(axis=1).sum())
    for k in (few, few - 3, few - 9):
        c = S.photon_gather(1, q, qn2, 1e3, k)
        assert np.allclose(c, knn_estimate(ph2, q, qn2, 1e3, k), rtol=2e-5, atol=0), k
    assert reach > 0


def test_render_with_photon_maps(pkg, scenes, oracle):
    """Config 5's gather inside the frame: irradiance of both maps added at diffuse hits (Scene.cpp:286-299)."""
    H, S = build_pair(pkg, scenes, oracle, "cornell")
    n = 20000
    pw, pos, d = _photon_cloud(n, 8)
    pw2, pos2, d2 = _photon_cloud(n // 4, 9)
    d[:, 1] = -np.abs(d[:, 1]); d2[:, 1] = -np.abs(d2[:, 1])      # arriving from above, so floor normals accept them
    for which, (a, b, c) in enumerate(((pw, pos, d), (pw2, pos2, d2))):
        oracle.lib.orc_pm_reset(which, ctypes.c_int(len(b)))
        oracle.pm_store(which, a, b, c); oracle.pm_scale(which, 1.0 / len(b)); oracle.pm_balance(which)
        S.photon_upload(which, oracle.pm_dump(which))
    w, h = 96, 72
    img = S.render(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, use_photon_maps=1))
    ref = oracle.trace_scene(oracle.eye_rays(w, h), depth=10).reshape(h, w, 3)
    base = S.render(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, use_photon_maps=0))
    assert (img - base).max() > 1e-3                                   # the maps contribute
    assert np.isclose(img, ref, rtol=2e-4, atol=2e-5).mean() > 0.999
    for which in (0, 1):
        oracle.lib.orc_pm_reset(which, ctypes.c_int(1))


def test_config4_flower_refractive_scene(pkg, scenes, oracle):
    """BASELINE config 4 (Petals2 + Stem + Leaf + WaterDrops, water = Phong(1, 0, 1, 250, 1.33), DirectionalAreaLight):
    the reflect / Fresnel / refract tree to depth 10 with closest-hit shadow rays (refractive occluders, Phong.cpp:99-113),
    against the oracle's Scene::traceScene at reduced resolution, plus 4 jittered samples at a larger size for sanity."""
    H, S = build_pair(pkg, scenes, oracle, "flower", layout=0)
    assert S.info.num_triangles == 14784 + 1664 + 6144 + 20160
    w, h = 384, 256
    img = S.render(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, max_depth=10, bg=(1, 1, 1)))
    ref = oracle.trace_scene(oracle.eye_rays(w, h), depth=10).reshape(h, w, 3)
    assert np.isfinite(img).mean() > 0.999
    a, b = tonemap_u8(oracle, np.nan_to_num(img)), tonemap_u8(oracle, np.nan_to_num(ref))
    diff = np.abs(a.astype(int) - b.astype(int)).max(axis=2)
    assert (diff <= 2).mean() > 0.99, (diff <= 2).mean()
    assert psnr(a, b) >= 35, psnr(a, b)
    rays_1spp = S.last_call_stats()[0]
    assert rays_1spp > 2 * w * h                                       # primaries + shadow slots + secondary generations
    # 4 spp, jittered (the config's sampling): every sample is an independent frame, so the ray count scales with it
    p4 = S.render_params(768, 512, spp=4, jitter=1, mode=pkg.RENDER_WHITTED, max_depth=10, bg=(1, 1, 1))
    img4 = S.render(H.camera(), p4)
    assert np.isfinite(img4).mean() > 0.999 and S.last_call_stats()[0] > 3.5 * 4 * rays_1spp * 0.9
    small = img4.reshape(256, 2, 384, 2, 3).mean(axis=(1, 3))           # box-filtered back to 384 x 256
    assert np.abs(tonemap_u8(oracle, np.nan_to_num(small)).astype(int) - a.astype(int)).mean() < 6
    # the config's own size, 2048 x 1365 at 4 spp: same picture (block means agree with the small frame), ray count scales
    W, Hh = scenes.SCENES["flower"]["size"]
    full = S.render(H.camera(), S.render_params(W, Hh, spp=4, jitter=1, mode=pkg.RENDER_WHITTED, max_depth=10, bg=(1, 1, 1)))
    assert full.shape == (Hh, W, 3) and np.isfinite(full).mean() > 0.999
    assert S.last_call_stats()[0] > 0.9 * 4 * rays_1spp * (W * Hh) / (w * h)
    fb = np.nan_to_num(full)
    assert abs(float(np.clip(fb, 0, 4).mean()) / float(np.clip(np.nan_to_num(img), 0, 4).mean()) - 1) < 0.05


def test_config5_photon_map_render(pkg, scenes, oracle):
    """BASELINE config 5 end to end on the device: Scene::preCalc traces both photon maps (device walks, host balance),
    then the frame adds the kNN irradiance of both maps at every diffuse hit (Scene.cpp:286-299).  Checked against the
    oracle's traceScene over the SAME maps (the host layer's balanced arrays loaded into the oracle)."""
    H = pkg.HostScene(0)
    for d in (oracle, H):
        scenes.realise(d, "cornell_drops", objio.obj_path)
    H.set_photon_counts(200000, 200000)
    oracle.precalc(); H.precalc()
    S = H.scene()
    for which in (0, 1):
        ph = H.pm_dump(which)
        assert 200000 <= len(ph) - 1 <= 200005
        oracle.lib.orc_pm_reset(which, ctypes.c_int(len(ph)))
        oracle.lib.orc_pm_load(which, md._fp(ph.view(np.uint8)), ctypes.c_int(len(ph) - 1))
    w, h = 128, 128
    img = S.render(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, max_depth=10, use_photon_maps=1))
    base = S.render(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, max_depth=10, use_photon_maps=0))
    ref = oracle.trace_scene(oracle.eye_rays(w, h), depth=10).reshape(h, w, 3)
    assert (np.nan_to_num(img) - np.nan_to_num(base)).max() > 1e-3     # the maps contribute
    a, b = tonemap_u8(oracle, np.nan_to_num(img)), tonemap_u8(oracle, np.nan_to_num(ref))
    diff = np.abs(a.astype(int) - b.astype(int)).max(axis=2)
    assert (diff <= 2).mean() > 0.99, (diff <= 2).mean()
    assert psnr(a, b) >= 35, psnr(a, b)
    # the config's own size, 512 x 512: finite, and the same picture as the small frame on average
    Wc, Hc = scenes.SCENES["cornell_drops"]["size"]
    full = S.render(H.camera(), S.render_params(Wc, Hc, mode=pkg.RENDER_WHITTED, max_depth=10, use_photon_maps=1))
    assert np.isfinite(full).mean() > 0.999
    assert abs(float(np.clip(np.nan_to_num(full), 0, 4).mean()) / float(np.clip(np.nan_to_num(img), 0, 4).mean()) - 1) < 0.05
    for which in (0, 1):
        oracle.lib.orc_pm_reset(which, ctypes.c_int(1))


def test_page_locked_host_buffers_and_image_reuse(pkg, scenes, oracle):
    """mirogpu_host_alloc / mirogpu_host_free: a page-locked frame buffer takes the same bytes as a pageable one, and
    Camera::click on the same Image (Image::resize to the same size) keeps its pixels across frames."""
    H, S = build_pair(pkg, scenes, oracle, "cornell")
    w, h = 96, 64
    p = S.render_params(w, h, spp=1, jitter=0, mode=pkg.RENDER_WHITTED, tonemap=1, shadows=1)
    want = S.render_rgb8(H.camera(), p)
    ptr = pkg.lib.mirogpu_host_alloc(w * h * 3)
    assert ptr
    try:
        buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_ubyte)), shape=(h, w, 3))
        buf[:] = 7
        got = S.render_rgb8(H.camera(), p, out=buf)
        assert np.array_equal(got, want) and want.max() > 0
    finally:
        pkg.lib.mirogpu_host_free(ptr)
    pkg.lib.mirogpu_host_free(None)
    H.set_render(spp=1, jitter=0, mode=pkg.RENDER_WHITTED, shadows=1, seed=168, use_photon_maps=0)
    a = H.render(w, h)
    out = np.zeros((h, w, 3), np.uint8)
    b = H.render(w, h, out=out)
    assert b is out and np.array_equal(a, b) and np.array_equal(a, want)
    assert H.last_render_seconds > 0
    c = H.render(w // 2, h // 2)      # a different size reallocates
    assert c.shape == (h // 2, w // 2, 3) and c.max() > 0


def test_frame_pipeline_with_several_handles_delivers_the_single_call_frames(pkg, scenes, oracle):
    """sharding.FramePipeline (bench.py's multi-rank e2e path) with three handles of the scene -- three consecutive frames
    rendering concurrently on three streams, exchanges in frame order on the side stream: every delivered host frame equals
    the frame of one synchronous mirogpu_render_rgb8 call with the same seed, and the ray counts agree.  One-rank NCCL group."""
    import importlib
    import os
    import torch.distributed as dist
    sharding = importlib.import_module(pkg.__name__ + ".sharding")
    H, S = build_pair(pkg, scenes, oracle, "teapot", layout=pkg.LAYOUT_QBVH4)
    extra = [scenes.handle_replica(pkg, H, "teapot", pkg.LAYOUT_QBVH4) for _ in range(2)]   # the host layer holds ONE scene
    w, h = 192, 128
    dev = torch.device("cuda", 0)
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        created = True

    def params(it):
        return S.render_params(w, h, spp=4, jitter=1, max_depth=10, mode=pkg.RENDER_DIFFUSE_BOUNCE, seed=40 + it, tonemap=0, shadows=0)
    try:
        pipe = sharding.FramePipeline(S, h, w, 1, 0, dev, replicas=extra)
        assert len(pipe.slots) == 4 and len(pipe.scenes) == 3
        cam = H.camera()
        nframes = 7
        slots = [pipe.submit(cam, params(it)) for it in range(nframes)]      # queued without waiting on any frame
        pipe.drain()
        # the four slots now hold the last four frames
        for it in range(nframes - 4, nframes):
            ref = np.zeros((h, w, 3), np.uint8)
            S.render_rgb8(cam, params(it), out=ref)
            assert np.array_equal(pipe.host_frame(slots[it]).numpy(), ref), it
        # ray counts: read per frame (waits for that frame's render)
        for it in range(3):
            got = pipe.rays_traced(pipe.submit(cam, params(it)))
            S.render_rgb8(cam, params(it), out=np.zeros((h, w, 3), np.uint8))
            assert got == S.last_call_stats()[0] and got > w * h * 4
        pipe.drain()
    finally:
        if created:
            dist.destroy_process_group()
