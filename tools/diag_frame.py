"""Where a small frame's time goes: device-only (CUDA events around mirogpu_render_device), the C-ABI call into a pinned
host framebuffer, Scene::raytraceImage's own timer, and the Python wall clock around HostScene.render.  One JSON object."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import objio
import torch
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
out = []
saved = os.dup(1); os.dup2(2, 1)
for name, (w, h) in [("cornell", (512, 512)), ("bunny_teapot", (1024, 1024))]:
    H = pkg.HostScene()
    scenes.realise(H, name, objio.obj_path)
    H.precalc()
    H.set_render(spp=1, jitter=0, mode=pkg.RENDER_WHITTED, shadows=1, seed=168, use_photon_maps=0)
    S, cam = H.scene(), H.camera()
    p = S.render_params(w, h, spp=1, jitter=0, max_depth=10, mode=pkg.RENDER_WHITTED, seed=168, tonemap=1, shadows=1)
    d_rgb = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
    pinned = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()
    pageable = np.zeros((h, w, 3), np.uint8)
    res = {"scene": name, "w": w, "h": h}
    def med(f, n=9):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
        return 1e3 * float(np.median(ts[2:]))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dts = []
    for _ in range(9):
        torch.cuda.synchronize(); ev[0].record(); S.render_device(cam, p, d_rgb); ev[1].record(); torch.cuda.synchronize()
        dts.append(ev[0].elapsed_time(ev[1]))
    res["device_events_ms"] = float(np.median(dts[2:]))
    res["launches"] = int(S.last_call_stats()[1])
    def dev_wall():
        S.render_device(cam, p, d_rgb); torch.cuda.synchronize()
    res["device_wall_ms"] = med(dev_wall)
    res["abi_rgb8_pinned_ms"] = med(lambda: S.render_rgb8(cam, p, out=pinned.numpy()))
    res["abi_rgb8_pageable_ms"] = med(lambda: S.render_rgb8(cam, p, out=pageable))
    res["host_render_wall_ms"] = med(lambda: H.render(w, h, out=pageable))
    res["host_render_own_timer_ms"] = 1e3 * H.last_render_seconds
    res["host_render_fresh_out_ms"] = med(lambda: H.render(w, h))
    out.append(res)
os.dup2(saved, 1)
print(json.dumps(out, indent=1))
