"""Generates tests/golden/texcoords.npz: the texture coordinates (vt records and per-face texture indices) of the models that the
2-D procedural textures are applied to (StemTexture on Stem.obj, assignment3.cpp:96).  Run where /root/reference exists:
    python tests/golden/make_texcoords.py
Parsed records only (see tests/objio.py); no reference file is copied."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import objio  # noqa: E402

MODELS = ["Stem"]

if __name__ == "__main__":
    out = {}
    for m in MODELS:
        rec = objio.parse_texcoords(f"/root/reference/models/{m}.obj")
        for k, v in rec.items():
            out[f"{m}__{k}"] = v
        print(m, {k: v.shape for k, v in rec.items()})
    np.savez_compressed(objio.TEXCOORDS_NPZ, **out)
    print("wrote", objio.TEXCOORDS_NPZ, os.path.getsize(objio.TEXCOORDS_NPZ), "bytes")
