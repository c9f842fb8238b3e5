"""The e2e call of bench.py on its own (Scene::raytraceImage -> mirogpu_render_rgb8, diffuse-bounce mode, pinned host framebuffer),
for tuning the render path: ms per frame and Grays/s over 10 frames after 3 warm-ups.  One JSON line; environment knobs apply."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, objio
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
saved = os.dup(1); os.dup2(2, 1)
H = pkg.HostScene(); scenes.realise(H, "bunny20", objio.obj_path); H.precalc()
S = H.scene(); cam = H.camera()
W, Hh, SPP = 1920, 1080, 16
p = S.render_params(W, Hh, spp=SPP, jitter=1, max_depth=10, mode=pkg.RENDER_DIFFUSE_BOUNCE, seed=168, tonemap=0, shadows=0)
fb = torch.empty((Hh, W, 3), dtype=torch.uint8).pin_memory()
for it in range(3):
    p.seed = it; S.render_rgb8(cam, p, out=fb.numpy())
t0 = time.perf_counter(); rays = 0
for it in range(10):
    p.seed = 100 + it; S.render_rgb8(cam, p, out=fb.numpy()); rays += S.last_call_stats()[0]
dt = time.perf_counter() - t0
os.dup2(saved, 1)
print(json.dumps({"ms_per_frame": dt / 10 * 1e3, "grays_s": rays / dt / 1e9, "rays_per_frame": rays / 10,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("MIROGPU_")}}))
