"""GPU check: every kernel variant returns bit-identical hits on the bench's own primary + bounce rays."""
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
W, Hh = 1920, 1080
n = W * Hh
dev = torch.device("cuda", 0)
d_rays = torch.empty((n, 8), dtype=torch.float32, device=dev)
d_b = torch.empty((n, 8), dtype=torch.float32, device=dev)
ref = None
variants = [int(v) for v in (sys.argv[1:] or ["0", "1", "2"])]
S.set_kernel_variant(0)
S.generate_primary(cam, W, Hh, d_rays, rows=(0, Hh, 1, 0), jitter=1, seed=168, sample=0, samples=1)
h0 = torch.empty((n, 4), dtype=torch.float32, device=dev)
S.intersect_device(d_rays, h0)
S.generate_bounce(d_rays, h0, d_b, seed=168, sample=0, index_base=0)
out = {}
for v in variants:
    S.set_kernel_variant(v)
    for any_hit in (0, 1):
        hp = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        hb = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        S.intersect_device(d_rays, hp, mode=any_hit)
        S.intersect_device(d_b, hb, mode=any_hit)
        torch.cuda.synchronize()
        out[(v, any_hit)] = (hp.cpu().numpy().view(np.uint32), hb.cpu().numpy().view(np.uint32))
ok = True
for v in variants[1:]:
    for k in (0, 1):
        a, b = out[(variants[0], 0)][k], out[(v, 0)][k]
        same = np.array_equal(a, b)
        ok &= same
        print(f"variant {v} vs {variants[0]} closest {'primary' if k == 0 else 'bounce'}: {'identical' if same else 'DIFFERENT %d' % int((a != b).any(1).sum())}")
        # any-hit: same hit/miss classification as closest hit
        ah = out[(v, 1)][k][:, 1] != 0xFFFFFFFF
        ch = out[(variants[0], 0)][k][:, 1] != 0xFFFFFFFF
        same = np.array_equal(ah, ch)
        ok &= same
        print(f"variant {v} any-hit {'primary' if k == 0 else 'bounce'} hit/miss: {'identical' if same else 'DIFFERENT %d' % int((ah != ch).sum())}")
print("VARIANTS", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
