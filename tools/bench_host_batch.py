"""The host-buffer batch API of the boundary (mirogpu_intersect_batch: BVH::intersect over n rays, HOST rays in, HOST hits out, copies
inside the call) on the bench scene: Mrays/s for a large incoherent batch from pageable and from page-locked buffers, and the
latency of the reference's own call pattern -- one ray per BVH::intersect call.  One JSON object."""
import ctypes, importlib, json, os, sys, time
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
H.precalc()
S = H.scene(); cam = H.camera()
W, Hh = 1920, 1080
n = W * Hh * 4
d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda"); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
d_b = torch.empty((n, 8), dtype=torch.float32, device="cuda")
S.generate_primary(cam, W, Hh, d_rays, jitter=1, samples=4)
S.intersect_device(d_rays, d_hits, mode=pkg.CLOSEST_HIT | pkg.HINT_COHERENT)
S.generate_bounce(d_rays, d_hits, d_b)
torch.cuda.synchronize()
rb = d_b.cpu().numpy(); rb = np.ascontiguousarray(rb[rb[:, 7] >= rb[:, 3]])
out = {"scene": "bunny20", "rays": int(rb.shape[0]), "kind": "incoherent Ray::diffuse bounce rays"}
hits = np.zeros(rb.shape[0], pkg.HIT_DTYPE)
for name, rays, h in (("pageable", rb, hits),):
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); S.intersect(rays, out=h); ts.append(time.perf_counter() - t0)
    out[name + "_mrays_s"] = rays.shape[0] / min(ts[1:]) / 1e6
pr = torch.from_numpy(rb).pin_memory(); ph = torch.empty((rb.shape[0], 4), dtype=torch.float32).pin_memory()
ts = []
for _ in range(4):
    t0 = time.perf_counter(); S.intersect(pr, out=ph); ts.append(time.perf_counter() - t0)
out["pinned_mrays_s"] = rb.shape[0] / min(ts[1:]) / 1e6
assert np.array_equal(ph.numpy().view(pkg.HIT_DTYPE).reshape(-1), hits)
# single-ray calls: the reference's own call pattern (BVH::intersect per ray)
one = np.ascontiguousarray(rb[:1]); h1 = np.zeros(1, pkg.HIT_DTYPE)
lib = pkg.lib
fn = lib.mirogpu_intersect_batch
args = (S._h, ctypes.c_void_p(one.ctypes.data), ctypes.c_size_t(1), ctypes.c_void_p(h1.ctypes.data), 0)
for _ in range(200):
    fn(*args)
t0 = time.perf_counter()
for _ in range(2000):
    fn(*args)
out["single_ray_call_us"] = (time.perf_counter() - t0) / 2000 * 1e6
os.dup2(saved, 1)
print(json.dumps(out))
