"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): all layouts x all kernel variants,
ray generation, a Whitted frame with shadows and glass, a diffuse-bounce frame, photon tracing, gather."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import objio
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
ref = None
for layout in (pkg.LAYOUT_BVH2, pkg.LAYOUT_CWBVH8, pkg.LAYOUT_BVH4):
    H = pkg.HostScene(layout)
    scenes.realise(H, "cornell_drops", objio.obj_path)
    H.set_photon_counts(3000, 500)
    H.precalc()
    S = H.scene(); cam = H.camera()
    w = h = 64
    n = w * h * 2
    d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_b = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_h2 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    S.generate_primary(cam, w, h, d_rays, jitter=1, samples=2)
    outs = []
    for variant in (-1, 0, 1, 2):
        S.set_kernel_variant(variant)
        for mode in (pkg.CLOSEST_HIT, pkg.ANY_HIT, pkg.CLOSEST_HIT | pkg.HINT_COHERENT):
            S.intersect_device(d_rays, d_hits, mode=mode)
            S.generate_bounce(d_rays, d_hits, d_b)
            S.intersect_device(d_b, d_h2, mode=mode)
            torch.cuda.synchronize()
            if mode != pkg.ANY_HIT:
                outs.append((d_hits.cpu().numpy().copy(), d_h2.cpu().numpy().copy()))
    for a, b in outs[1:]:
        assert np.array_equal(a.view(np.uint32), outs[0][0].view(np.uint32)) and np.array_equal(b.view(np.uint32), outs[0][1].view(np.uint32))
    if ref is None:
        ref = outs[0]
    else:
        assert np.array_equal(ref[0].view(np.uint32), outs[0][0].view(np.uint32))
    S.set_kernel_variant(-1)
    img = S.render(cam, S.render_params(48, 40, mode=pkg.RENDER_WHITTED, max_depth=6, use_photon_maps=1))
    img2 = S.render(cam, S.render_params(48, 40, spp=2, jitter=1, mode=pkg.RENDER_DIFFUSE_BOUNCE, shadows=0))
    c, r = S.photon_trace(0, 1, 7, 0, 2000)
    q = np.array([[2.5, 0.0, -2.5]], np.float32)
    irr = S.photon_gather(0, q, np.array([[0, 1, 0]], np.float32), 1e10, 50)
    counted, cnt = S.intersect_counted(d_rays.cpu().numpy()[:500])
    print("layout", layout, "ok: hits", int((outs[0][0].view(np.uint32)[:, 1] != 0xFFFFFFFF).sum()), "photons", H.pm_stored(0), H.pm_stored(1), "irr", irr, "finite", np.isfinite(img).mean(), np.isfinite(img2).mean())
print("SANITIZE RUN COMPLETE")
