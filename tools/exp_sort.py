"""GPU experiment: how much does ray ordering change closest-hit throughput on the bench's bounce rays?
Rays are permuted with torch (untimed); only the trace launch is timed."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
H = bench.build_host_scene(pkg, scenes, pkg.LAYOUT_BVH2)
S = H.scene(); cam = H.camera()
W, Hh, SPP = 1920, 1080, 4
n = W * Hh * SPP
dev = torch.device("cuda", 0)
d_rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
d_b = torch.empty((n, 8), dtype=torch.float32, device=dev)
h0 = torch.empty((n, 4), dtype=torch.float32, device=dev)
S.set_kernel_variant(0)
S.generate_primary(cam, W, Hh, d_rays, rows=(0, Hh, 1, 0), jitter=1, seed=168, sample=0, samples=SPP)
S.intersect_device(d_rays, h0)
S.generate_bounce(d_rays, h0, d_b, seed=168, sample=0, index_base=0)
torch.cuda.synchronize()
live = d_b[:, 7] > 0
nlive = int(live.sum())
print("rays", n, "live", nlive)

def morton3(q, bits):
    out = torch.zeros(q.shape[0], dtype=torch.int64, device=dev)
    for b in range(bits):
        for a in range(3):
            out |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return out

o = d_b[:, 0:3]; d = d_b[:, 4:7]
lo = o[live].min(0).values; hi = o[live].max(0).values
def oq(bits):
    return ((o - lo) / (hi - lo).clamp_min(1e-9) * (2 ** bits - 1)).clamp(0, 2 ** bits - 1).to(torch.int64)
octant = ((d[:, 0] < 0).to(torch.int64) | ((d[:, 1] < 0).to(torch.int64) << 1) | ((d[:, 2] < 0).to(torch.int64) << 2))
ad = d.abs(); face_axis = ad.argmax(1)
m = ad.gather(1, face_axis[:, None]).squeeze(1).clamp_min(1e-20)
sgn = (d.gather(1, face_axis[:, None]).squeeze(1) < 0).to(torch.int64)
ua = d.gather(1, ((face_axis + 1) % 3)[:, None]).squeeze(1) / m
ub = d.gather(1, ((face_axis + 2) % 3)[:, None]).squeeze(1) / m
def cube(res):
    ia = ((ua * 0.5 + 0.5) * res).clamp(0, res - 1).to(torch.int64); ib = ((ub * 0.5 + 0.5) * res).clamp(0, res - 1).to(torch.int64)
    return ((face_axis * 2 + sgn) * res + ia) * res + ib
idx = torch.arange(n, device=dev, dtype=torch.int64)
dead = (~live).to(torch.int64)
pix = idx % (W * Hh); px = pix % W; py = pix // W
def tile2d(ts):
    return (py // ts) * ((W + ts - 1) // ts) + (px // ts)
keys = {
    "identity": idx,
    "tile2d 16 | cube2x2": (tile2d(16) * 2 + dead) * 32 + cube(2),
    "tile2d 16 | cube4x4": (tile2d(16) * 2 + dead) * 128 + cube(4),
    "tile2d 32 | cube2x2": (tile2d(32) * 2 + dead) * 32 + cube(2),
    "tile2d 32 | cube4x4": (tile2d(32) * 2 + dead) * 128 + cube(4),
    "tile2d 64 | cube4x4": (tile2d(64) * 2 + dead) * 128 + cube(4),
    "tile2d 64 | cube8x8": (tile2d(64) * 2 + dead) * 512 + cube(8),
    "tile2d 128 | cube8x8": (tile2d(128) * 2 + dead) * 512 + cube(8),
    "tile2d 32 only (pixel-tile order)": tile2d(32) * 2 + dead,
    "cube4x4 | tile2d 16": (dead * 128 + cube(4)) * (1 << 20) + tile2d(16),
    "cube8x8 | origin morton 7b": (dead * 512 + cube(8)) * (1 << 21) + morton3(oq(7), 7),
    "cube8x8 | origin morton 10b": (dead * 512 + cube(8)) * (1 << 30) + morton3(oq(10), 10),
    "cube16x16 | origin morton 10b": (dead * 2048 + cube(16)) * (1 << 30) + morton3(oq(10), 10),
    "dead-last (compaction only)": dead * n + idx,
    "octant, then index": (dead * 8 + octant) * n + idx,
    "origin morton 10b": dead * (1 << 40) + morton3(oq(10), 10),
    "cube4x4 | origin morton 7b": (dead * 128 + cube(4)) * (1 << 21) + morton3(oq(7), 7),
    "origin morton 4b | cube4x4 | morton 10b": ((dead * (1 << 12) + morton3(oq(4), 4)) * 128 + cube(4)) * (1 << 30) + morton3(oq(10), 10),
}
variants = [int(v) for v in (sys.argv[1:] or ["0", "3"])]
hb = torch.empty((n, 4), dtype=torch.float32, device=dev)
for name, k in keys.items():
    perm = torch.sort(k, stable=True).indices
    rb = d_b[perm].contiguous()
    line = f"{name:42s}"
    for v in variants:
        S.set_kernel_variant(v)
        for _ in range(2):
            S.intersect_device(rb, hb)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            S.intersect_device(rb, hb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        line += f"  v{v}: {ms:6.3f} ms {nlive / ms / 1e3:7.0f} Mrays/s"
    print(line, flush=True)
