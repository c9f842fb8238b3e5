"""Diagnostic: config 2 at its own size, final 8-bit frames of every layout against the oracle's."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import miro_driver as md, objio
pkg = importlib.import_module("cse168-raytracer_b200")
scenes = importlib.import_module("cse168-raytracer_b200.scenes")
name = sys.argv[1] if len(sys.argv) > 1 else "bunny_teapot"
w = h = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
O = md.oracle()
scenes.realise(O, name, objio.obj_path); O.precalc()
rgb = O.trace_scene(O.eye_rays(w, h), depth=10).reshape(h, w, 3)
ref = np.zeros((h, w, 3), np.uint8)
O.lib.orc_tonemap(md._fp(np.ascontiguousarray(rgb, np.float32)), ctypes.c_long(w * h), md._fp(ref))
for layout in (0, 1, 2, 3):
    H = pkg.HostScene(layout)
    scenes.realise(H, name, objio.obj_path); H.precalc()
    S = H.scene()
    for shadows in (1, 0):
        H.set_render(spp=1, jitter=0, mode=pkg.RENDER_WHITTED, shadows=shadows)
        u8 = H.render(w, h)
        f = S.render(H.camera(), S.render_params(w, h, mode=pkg.RENDER_WHITTED, tonemap=0, shadows=shadows))
        d = np.abs(u8.astype(int) - ref.astype(int)).max(axis=2)
        close = np.isclose(f, rgb, rtol=2e-4, atol=2e-5).all(axis=2)
        ys, xs = np.nonzero(~close)
        print(f"layout {layout} shadows {shadows}: u8 within 2: {(d <= 2).mean():.5f}  float close: {close.mean():.5f}  first bad {list(zip(ys[:4].tolist(), xs[:4].tolist()))}",
              [(f[y, x].tolist(), rgb[y, x].tolist()) for y, x in zip(ys[:2], xs[:2])] if shadows else "")
