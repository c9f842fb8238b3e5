"""Quick device-side timing probe (not the bench): one frame of gen-primary -> closest-hit -> gen-bounce ->
closest-hit per layout / kernel variant, CUDA events on torch's current stream."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
import objio  # noqa: E402


def main(scene="bunny20", w=1920, h=1080, iters=10):
    for layout in (1, 0):
        H = pkg.HostScene(layout)
        t0 = time.time()
        scenes.realise(H, scene, objio.obj_path)
        t1 = time.time()
        H.precalc()
        t2 = time.time()
        S = H.scene()
        i = S.info
        print(f"layout {layout}: load {t1-t0:.2f}s precalc {t2-t1:.2f}s (build {i.build_seconds:.2f} flatten {i.flatten_seconds:.2f} upload {i.upload_seconds:.2f}) "
              f"tris {i.num_triangles} nodes {i.num_nodes} node MB {i.node_bytes/1e6:.1f} tri MB {i.triangle_bytes/1e6:.1f} depth {i.max_depth}", flush=True)
        n = w * h
        cam = H.camera()
        d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        d_b = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        d_h2 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        for variant in (0, 2, 1):
            S.set_kernel_variant(variant)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            acc = np.zeros(4)
            for it in range(iters + 3):
                ev[0].record(); S.generate_primary(cam, w, h, d_rays, jitter=1, sample=it)
                ev[1].record(); S.intersect_device(d_rays, d_hits)
                ev[2].record(); S.generate_bounce(d_rays, d_hits, d_b, sample=it)
                ev[3].record(); S.intersect_device(d_b, d_h2)
                ev[4].record(); torch.cuda.synchronize()
                if it >= 3:
                    acc += [ev[k].elapsed_time(ev[k + 1]) for k in range(4)]
            acc /= iters
            hits1 = int((d_hits[:, 1].view(torch.int32) != -1).sum()); hits2 = int((d_h2[:, 1].view(torch.int32) != -1).sum())
            print(f"  variant {variant}: gen {acc[0]:.3f} ms | primary trace {acc[1]:.3f} ms = {n/acc[1]/1e3:.1f} Mrays/s | bounce gen {acc[2]:.3f} ms | "
                  f"bounce trace {acc[3]:.3f} ms = {hits1/acc[3]/1e3:.1f} Mrays/s live ({n/acc[3]/1e3:.1f} incl. dead) | hits {hits1} {hits2} | "
                  f"frame {acc.sum():.3f} ms = {(n+hits1)/acc.sum()/1e3:.1f} Mrays/s", flush=True)
        # instrumented pass on a subsample
        sub = d_rays.cpu().numpy()[::16]
        _, c = S.intersect_counted(sub)
        subb = d_b.cpu().numpy()[::16]
        live = subb[:, 7] >= subb[:, 3]
        _, cb = S.intersect_counted(subb[live])
        print(f"  primary: nodes/ray {c.node_visits/c.rays:.2f} tris/ray {c.triangle_tests/c.rays:.2f} bytes/ray {c.bytes_fetched/c.rays:.0f} | "
              f"bounce(live): nodes/ray {cb.node_visits/cb.rays:.2f} tris/ray {cb.triangle_tests/cb.rays:.2f} bytes/ray {cb.bytes_fetched/cb.rays:.0f}", flush=True)


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["bunny20"]))
