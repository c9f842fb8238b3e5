"""Generates tests/golden/ref_textures.npz: colours and bump heights of the REAL reference's texture classes (Texture.h,
Texture.cpp over lib/src/Perlin.cpp and lib/src/Worley.cpp, compiled in place into oracle/_ref) at seeded coordinates.
Run where oracle/_ref is built:   python tests/golden/make_textures.py
kind numbers and parameter order: include/mirogpu.h (MIROGPU_TEX_*)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import miro_driver as md  # noqa: E402

CASES = {   # name: (kind, constructor arguments, coordinate half-range)
    "checker": (1, [1, 0.5, 0.2, 0, 0.1, 0.9, 2.5], 6.0),
    "stone": (2, [3.0], 6.0),
    "stone20": (2, [20.0], 1.0),
    "stem": (3, [30.0], 1.0),
    "petal": (4, [0, 0, 0, 7.0], 6.0),
    "leaf": (5, [1.0], 6.0),
    "flower_center": (6, [-0.1, -0.35, 0, 1.1], 1.5),
}

if __name__ == "__main__":
    R = md.reference("scalar")
    rng = np.random.default_rng(168)
    out = {}
    for name, (kind, tp, half) in CASES.items():
        c = ((rng.random((1500, 3), dtype=np.float32) * 2 - 1) * np.float32(half)).astype(np.float32)
        rgb, bump = R.texture_lookup(kind, tp, c, bump=True)
        out[name + "__kind"] = np.int32(kind); out[name + "__params"] = np.asarray(tp, np.float32)
        out[name + "__coords"] = c; out[name + "__rgb"] = rgb; out[name + "__bump"] = bump
        print(name, rgb.mean(axis=0), float(bump.mean()))
    path = os.path.join(HERE, "ref_textures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
