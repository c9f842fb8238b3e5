"""Scene::raytraceImage on ONE handle replicated over N devices (mirogpu_scene_create_ex with a device list, one host process):
the bench frame (bunny20, 1920x1080, 16 spp, diffuse-bounce mode, pinned 8-bit host framebuffer) for N = 1, 2, 4, 8 as far as
the box has GPUs.  Frames are checked bit-identical across N.  One JSON line."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, objio
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
saved = os.dup(1); os.dup2(2, 1)
W, Hh, SPP = 1920, 1080, 16
out = {"frame": f"bunny20 {W}x{Hh}, {SPP} spp, diffuse-bounce, mirogpu_render_rgb8 into pinned host memory", "runs": []}
first = None
ngpu = pkg.device_count()
for n in (1, 2, 4, 8):
    if n > ngpu:
        break
    H = pkg.HostScene(); scenes.realise(H, "bunny20", objio.obj_path); H.set_device_count(n)
    t0 = time.perf_counter(); H.precalc(); pre = time.perf_counter() - t0
    S = H.scene(); cam = H.camera()
    p = S.render_params(W, Hh, spp=SPP, jitter=1, max_depth=10, mode=pkg.RENDER_DIFFUSE_BOUNCE, seed=168, tonemap=0, shadows=0)
    fb = torch.empty((Hh, W, 3), dtype=torch.uint8).pin_memory()
    for it in range(3):
        p.seed = it; S.render_rgb8(cam, p, out=fb.numpy())
    p.seed = 7; S.render_rgb8(cam, p, out=fb.numpy())
    if first is None:
        first = fb.numpy().copy()
    same = bool(np.array_equal(first, fb.numpy()))
    t0 = time.perf_counter(); rays = 0
    for it in range(10):
        p.seed = 100 + it; S.render_rgb8(cam, p, out=fb.numpy()); rays += S.last_call_stats()[0]
    dt = time.perf_counter() - t0
    out["runs"].append({"devices": n, "ms_per_frame": dt / 10 * 1e3, "grays_s": rays / dt / 1e9, "precalc_s": pre, "frame_equals_one_device_frame": same})
os.dup2(saved, 1)
print(json.dumps(out))
