"""GPU experiment: does a tiled pixel order of the camera rays (instead of scanlines) speed up the primary trace?
Rays are permuted with torch (untimed); only the trace launch is timed."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
H = bench.build_host_scene(pkg, scenes, pkg.LAYOUT_QBVH4)
S = H.scene(); cam = H.camera()
W, Hh, SPP = 1920, 1080, 4
n = W * Hh * SPP
dev = torch.device("cuda", 0)
d_rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
hb = torch.empty((n, 4), dtype=torch.float32, device=dev)
S.generate_primary(cam, W, Hh, d_rays, rows=(0, Hh, 1, 0), jitter=1, seed=168, sample=0, samples=SPP)
idx = torch.arange(n, device=dev, dtype=torch.int64)
s = idx // (W * Hh); pix = idx % (W * Hh); px = pix % W; py = pix // W
def tile(tw, th, samples_inner=False):
    t = (py // th) * ((W + tw - 1) // tw) + (px // tw)
    inner = (py % th) * tw + (px % tw)
    if samples_inner:
        return (t * (tw * th) + inner) * SPP + s
    return (s * (10 ** 7) + t) * (tw * th) + inner
keys = {"scanline (today)": idx, "tiles 8x4": tile(8, 4), "tiles 8x8": tile(8, 8), "tiles 16x4": tile(16, 4), "tiles 32x2": tile(32, 2),
        "tiles 8x8, 4 samples of a pixel adjacent": tile(8, 8, True), "tiles 4x4, samples adjacent": tile(4, 4, True)}
for name, k in keys.items():
    perm = torch.sort(k, stable=True).indices
    r = d_rays[perm].contiguous()
    line = f"{name:44s}"
    for mode, label in ((pkg.CLOSEST_HIT, "hybrid"), (pkg.CLOSEST_HIT | pkg.HINT_COHERENT, "hint")):
        for _ in range(2):
            S.intersect_device(r, hb, mode=mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            S.intersect_device(r, hb, mode=mode)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        line += f"  {label}: {ms:6.3f} ms {n / ms / 1e3:7.0f} Mrays/s"
    print(line, flush=True)
