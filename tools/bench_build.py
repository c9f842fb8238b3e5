"""BVH::build on the bench scene (bunny x 20 + floor, 1 389 021 triangles): seconds per builder, two builds each in one process
(the second shows the steady figure: CUDA module load and the device builders' scratch allocation are paid once), plus the
closest-hit rate of 2 M camera rays through each tree.  One JSON object."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import objio
import torch
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
saved = os.dup(1); os.dup2(2, 1)
H = pkg.HostScene()
scenes.realise(H, "bunny20", objio.obj_path)
H.precalc_host_only() if hasattr(H, "precalc_host_only") else H.precalc()
V = np.ascontiguousarray(H.dump_triangles()[:, :9])
rays = H.eye_rays(1920, 1080)
d_rays = torch.from_numpy(rays).cuda(); d_hits = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
out = {"triangles": int(V.shape[0]), "builders": {}}
ONLY = os.environ.get("MIRO_BUILD_ONLY")
for name, b in (("sah_host", pkg.BUILDER_SAH_HOST), ("lbvh_device", pkg.BUILDER_LBVH_DEVICE), ("ploc_device", pkg.BUILDER_PLOC_DEVICE)):
    if ONLY and name != ONLY:
        continue
    res = {"build_s": [], "wall_s": []}
    for rep in range(3):
        t0 = time.perf_counter()
        S = pkg.MiroScene(V, layout=pkg.LAYOUT_QBVH4, builder=b, max_leaf=int(os.environ.get("MIRO_BUILD_MAX_LEAF", "0")))
        res["wall_s"].append(time.perf_counter() - t0)            # mirogpu_scene_create: build + flatten / collapse + uploads
        res["build_s"].append(S.info.build_seconds)
    res["nodes"] = int(S.info.num_nodes)
    for _ in range(2):
        S.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        S.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
    e1.record(); torch.cuda.synchronize()
    res["camera_mrays_s"] = 5 * rays.shape[0] / (e0.elapsed_time(e1) * 1e-3) / 1e6
    out["builders"][name] = res
    del S
os.dup2(saved, 1)
print(json.dumps(out))
