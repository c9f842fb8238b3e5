import importlib, os, sys, time, ctypes
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import objio
pkg = importlib.import_module("cse168-raytracer_b200"); scenes = importlib.import_module("cse168-raytracer_b200.scenes")
saved = os.dup(1); os.dup2(2, 1)
H = pkg.HostScene(); scenes.realise(H, "cornell_drops", objio.obj_path); H.set_photon_counts(200000, 0); H.precalc(); S = H.scene()
q = np.array([[2.5, 0.0, -2.5]], np.float32); qn = np.array([[0, 1, 0]], np.float32)
a = S.photon_gather(0, q, qn, 1e10, 500)
qs = np.repeat(q, 2000, 0); ns = np.repeat(qn, 2000, 0)
b = S.photon_gather(0, qs, ns, 1e10, 500)
assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)), (a[0], b[0])
for _ in range(50): S.photon_gather(0, q, qn, 1e10, 500)
t0 = time.perf_counter()
for _ in range(500): S.photon_gather(0, q, qn, 1e10, 500)
dt = (time.perf_counter() - t0) / 500
os.dup2(saved, 1)
print("single irradiance_estimate call us", dt * 1e6, a[0])
