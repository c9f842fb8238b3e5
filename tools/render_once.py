"""Renders one BASELINE config frame a few times (for ncu launch lists): python tools/render_once.py <scene> <w> <h> <spp> [photon maps 0/1]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: F401
import objio
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
pm = int(sys.argv[5]) if len(sys.argv) > 5 else 0
H = pkg.HostScene()
scenes.realise(H, name, objio.obj_path)
if pm:
    H.set_photon_counts(200000, 200000)
H.precalc()
H.set_render(spp=spp, jitter=1 if spp > 1 else 0, mode=pkg.RENDER_WHITTED, shadows=1, seed=168, use_photon_maps=pm)
for _ in range(3):
    img = H.render(w, h)
print("ok", img.shape, H.scene().last_call_stats())
