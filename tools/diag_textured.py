import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miro_driver as md, objio
pkg = importlib.import_module("cse168-raytracer_b200"); scenes = importlib.import_module("cse168-raytracer_b200.scenes")
name = sys.argv[1] if len(sys.argv) > 1 else "flower_textured"
saved = os.dup(1); os.dup2(2, 1)
R = md.reference("scalar"); H = pkg.HostScene(pkg.LAYOUT_QBVH4)
for d in (R, H):
    scenes.realise(d, name, objio.obj_path); d.precalc()
S = H.scene(); w, h = 384, 256
sc = scenes.SCENES[name]
p = S.render_params(w, h, mode=pkg.RENDER_WHITTED, tonemap=0, bg=sc.get("bg", (0, 0, 0)), shadows=int(os.environ.get("SH", "1")), max_depth=int(os.environ.get("DEPTH", "10")))
img = S.render(H.camera(), p)
rays = R.eye_rays(w, h)
ref = R.trace_scene(rays, depth=int(os.environ.get("DEPTH", "10"))).reshape(h, w, 3)
t, ids, P, N = R.trace(rays, 0)
dev = torch.device("cuda", 0)
d_rays = torch.from_numpy(rays).to(dev); d_hits = torch.empty((rays.shape[0], 4), dtype=torch.float32, device=dev)
S.intersect_device(d_rays, d_hits)
d_P = torch.empty((rays.shape[0], 3), dtype=torch.float32, device=dev); d_N = torch.empty_like(d_P); d_m = torch.empty(rays.shape[0], dtype=torch.int32, device=dev)
S.resolve_hits_device(d_hits, d_P, d_N, d_mat=d_m, d_rays=d_rays)
mat = d_m.cpu().numpy(); gN = d_N.cpu().numpy()
close = np.isclose(img, ref, rtol=2e-3, atol=2e-4).all(axis=2).reshape(-1)
os.dup2(saved, 1)
print("overall", close.mean())
for m in np.unique(mat):
    sel = mat == m
    nclose = np.isclose(gN[sel], N[sel], rtol=1e-4, atol=1e-4).all(axis=1).mean() if m >= 0 else float("nan")
    print("material", m, "pixels", sel.sum(), "frame close", close[sel].mean(), "normal close", nclose, "|N| mean", np.linalg.norm(N[sel], axis=1).mean() if m >= 0 else 0)
    bad = np.flatnonzero(sel & ~close)[:3]
    for i in bad:
        print("   px", i, "gpu", img.reshape(-1, 3)[i], "ref", ref.reshape(-1, 3)[i], "P", P[i], "N", N[i], "gN", gN[i])
