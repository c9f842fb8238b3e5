"""Generates tests/golden/ref_hits.npz with the REAL reference (oracle/_ref, compiled in place from
/root/reference) -- run in the build container:   python tests/golden/make_golden_hits.py

Per scene: a primary-ray image (pixel centres, Camera::eyeRay) at reduced resolution, one Ray::diffuse-style
bounce ray per hit (directions from fixed uniforms through the oracle's restatement of
alignHemisphereToVector; the reference itself only offers rand()-driven generation), and for both ray
sets the reference's Scene::trace outputs (t, prim id, P, N), plus the -DSTATS counters and BVH node counts.
One process per scene: Camera::eyeRay caches its basis in function statics (Camera.cpp:106-125).
"""
import importlib
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {"cornell": (64, 64), "teapot": (96, 96), "bunny_teapot": (128, 128), "testobj": (32, 32)}


def run(name, q):
    import miro_driver as md
    import objio
    scenes = importlib.import_module("cse168-raytracer_b200.scenes")
    w, h = CASES[name]
    out = {}
    R = md.reference("stats")
    scenes.realise(R, name, objio.obj_path)
    R.precalc()
    st = R.stats()
    out["nodes"] = np.array([st["nodes"], st["leaves"]], np.int64)
    rays = R.eye_rays(w, h)
    R.stats_reset_rays()
    t, ids, P, N = R.trace(rays, 1)
    st = R.stats()
    out.update(primary_rays=rays, primary_t=t, primary_id=ids, primary_P=P, primary_N=N,
               primary_counters=np.array([st["ray_box"], st["ray_tri"]], np.int64))
    O = md.oracle()
    scenes.realise(O, name, objio.obj_path)
    u = np.random.default_rng(168).random((rays.shape[0], 2), dtype=np.float32)
    brays = np.zeros_like(rays)
    import ctypes
    O.lib.orc_diffuse_rays(md._fp(P), md._fp(N), md._fp(ids), md._fp(u), ctypes.c_long(rays.shape[0]), md._fp(brays))
    R.stats_reset_rays()
    t2, ids2, P2, N2 = R.trace(brays, 1)
    st = R.stats()
    out.update(bounce_u=u, bounce_rays=brays, bounce_t=t2, bounce_id=ids2, bounce_P=P2, bounce_N=N2,
               bounce_counters=np.array([st["ray_box"], st["ray_tri"]], np.int64))
    q.put((name, out))


if __name__ == "__main__":
    allout = {}
    for name in CASES:
        q = mp.Queue()
        p = mp.Process(target=run, args=(name, q))
        p.start()
        n, out = q.get()
        p.join()
        for k, v in out.items():
            allout[f"{n}__{k}"] = v
        print(n, {k: (v.shape if v.ndim else v) for k, v in out.items() if "counters" in k or k == "nodes"}, out["nodes"], out["primary_counters"], out["bounce_counters"],
              "hit fraction", (out["primary_id"] >= 0).mean(), (out["bounce_id"] >= 0).mean())
    path = os.path.join(HERE, "ref_hits.npz")
    np.savez_compressed(path, **allout)
    print("wrote", path, os.path.getsize(path))
