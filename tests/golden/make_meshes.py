"""Generates tests/golden/meshes.npz from the reference's models (run in the build container, where
/root/reference exists):   python tests/golden/make_meshes.py
Only parsed geometry records are stored (see tests/objio.py); no reference file is copied."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import objio  # noqa: E402

MODELS = ["bunny", "teapot", "cornell_box", "cornell_box_1", "cornell_box_2", "cornell_box_3", "cornell_box_4", "testobj", "sphere",
          "Petals2", "Stem", "Leaf", "WaterDrops"]   # BASELINE configs 4 and 5

if __name__ == "__main__":
    out = {}
    for m in MODELS:
        rec = objio.parse_obj(f"/root/reference/models/{m}.obj")
        for k, v in rec.items():
            out[f"{m}__{k}"] = v
        print(m, {k: v.shape for k, v in rec.items()})
    np.savez_compressed(objio.MESHES_NPZ, **out)
    print("wrote", objio.MESHES_NPZ, os.path.getsize(objio.MESHES_NPZ), "bytes")
