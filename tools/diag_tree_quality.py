"""Node visits and triangle tests per ray (the instrumented one-thread-per-ray kernel behind mirogpu_intersect_batch_counted) through
the trees of the three builders on the bench scene: what separates the device-built trees from the SAH tree."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import objio
pkg = importlib.import_module("cse168-raytracer_b200"); scenes = importlib.import_module("cse168-raytracer_b200.scenes")
saved = os.dup(1); os.dup2(2, 1)
H = pkg.HostScene(); scenes.realise(H, "bunny20", objio.obj_path); H.precalc()
V = np.ascontiguousarray(H.dump_triangles()[:, :9])
S0 = H.scene(); cam = H.camera()
w, h = 960, 540
n = w * h
d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda"); d_b = torch.empty((n, 8), dtype=torch.float32, device="cuda")
S0.generate_primary(cam, w, h, d_rays, jitter=1, seed=168, sample=0, samples=1)
S0.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
S0.generate_bounce(d_rays, d_hits, d_b, seed=168, sample=0)
torch.cuda.synchronize()
rp = d_rays.cpu().numpy(); rb = d_b.cpu().numpy(); rb = np.ascontiguousarray(rb[rb[:, 7] >= rb[:, 3]])
out = {}
for name, b, ml in (("sah_host", pkg.BUILDER_SAH_HOST, 0), ("ploc_leaf1", pkg.BUILDER_PLOC_DEVICE, 0), ("ploc_leaf4", pkg.BUILDER_PLOC_DEVICE, 4), ("lbvh_leaf1", pkg.BUILDER_LBVH_DEVICE, 0)):
    S = pkg.MiroScene(V, layout=pkg.LAYOUT_QBVH4, builder=b, max_leaf=ml)
    res = {"nodes": int(S.info.num_nodes)}
    for kind, rays in (("primary", rp), ("bounce", rb)):
        hits, c = S.intersect_counted(rays)
        res[kind] = {"node_visits_per_ray": c.node_visits / c.rays, "triangle_tests_per_ray": c.triangle_tests / c.rays, "box_tests_per_ray": c.box_tests / c.rays}
    out[name] = res
    del S
os.dup2(saved, 1)
print(json.dumps(out, indent=1))
