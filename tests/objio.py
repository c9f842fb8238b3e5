"""Lossless binary form of the reference's small OBJ models, and the writer that turns it back into text.

/root/reference does not exist on the GPU box, and reference files are never copied into this repo.  The
geometry fixtures in tests/golden/meshes.npz hold, per model, the parsed records of the OBJ file in file
order (vertex / normal / face index triples); `write_obj` regenerates an equivalent .obj whose numbers
parse (sscanf %f) to exactly the same binary32 values, so the reference's own loader (in oracle/_ref), the
oracle's restatement and the product's host loader can all ingest it anywhere.
"""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MESHES_NPZ = os.path.join(GOLDEN_DIR, "meshes.npz")


def parse_obj(path):
    """Records in file order.  kinds: 0 = v, 1 = vn, 2 = f (first three tokens, as the reference reads them)."""
    kinds, v, vn, fv, fn = [], [], [], [], []
    with open(path, "rb") as fh:
        for raw in fh:
            line = raw.decode("latin-1").strip()
            if line.startswith("vn"):
                x = line[2:].split()
                vn.append([float(x[0]), float(x[1]), float(x[2])]); kinds.append(1)
            elif line.startswith("vt"):
                continue
            elif line.startswith("v"):
                x = line[1:].split()
                v.append([float(x[0]), float(x[1]), float(x[2])]); kinds.append(0)
            elif line.startswith("f"):
                toks = line[1:].split()[:3]
                a, b = [], []
                for t in toks:
                    parts = t.split("/")
                    a.append(int(parts[0]))
                    b.append(int(parts[2]) if len(parts) > 2 and parts[2] else 0)
                fv.append(a); fn.append(b); kinds.append(2)
    return dict(kinds=np.asarray(kinds, np.uint8), v=np.asarray(v, np.float32).reshape(-1, 3),
                vn=np.asarray(vn, np.float32).reshape(-1, 3), fv=np.asarray(fv, np.int32).reshape(-1, 3),
                fn=np.asarray(fn, np.int32).reshape(-1, 3))


TEXCOORDS_NPZ = os.path.join(GOLDEN_DIR, "texcoords.npz")


def parse_texcoords(path):
    """The vt records and the texture index of the first three face tokens (0 = none), in file order."""
    vt, ft = [], []
    with open(path, "rb") as fh:
        for raw in fh:
            line = raw.decode("latin-1").strip()
            if line.startswith("vt"):
                x = line[2:].split()
                vt.append([float(x[0]), float(x[1])])
            elif line.startswith("f"):
                toks = line[1:].split()[:3]
                ft.append([int(t.split("/")[1]) if len(t.split("/")) > 1 and t.split("/")[1] else 0 for t in toks])
    return dict(vt=np.asarray(vt, np.float32).reshape(-1, 2), ft=np.asarray(ft, np.int32).reshape(-1, 3))


def write_obj(path, mesh):
    """Regenerates OBJ text (lines < 80 chars, the reference's fgets limit) in the original record order; texture coordinates
    (models listed in texcoords.npz) come first -- the loader only appends them to a table, their place among the other
    records does not matter."""
    iv = ivn = iff = 0
    out = []
    v, vn, fv, fn = mesh["v"], mesh["vn"], mesh["fv"], mesh["fn"]
    ft = mesh.get("ft")
    if ft is not None:
        for t in mesh["vt"]:
            out.append("vt %.9g %.9g\n" % (float(t[0]), float(t[1])))
    for k in mesh["kinds"]:
        if k == 0:
            out.append("v %.9g %.9g %.9g\n" % tuple(float(x) for x in v[iv])); iv += 1
        elif k == 1:
            out.append("vn %.9g %.9g %.9g\n" % tuple(float(x) for x in vn[ivn])); ivn += 1
        else:
            a, b = fv[iff], fn[iff]
            if ft is not None:
                c = ft[iff]
                out.append("f " + " ".join("%d/%s/%s" % (a[k], c[k] if c[k] else "", b[k] if b[k] else "") for k in range(3)) + "\n")
            elif b[2] != 0 or b[0] != 0 or b[1] != 0:
                out.append("f %d//%d %d//%d %d//%d\n" % (a[0], b[0], a[1], b[1], a[2], b[2]))
            else:
                out.append("f %d %d %d\n" % (a[0], a[1], a[2]))
            iff += 1
    with open(path, "w") as fh:
        fh.writelines(out)


def load_meshes():
    z = np.load(MESHES_NPZ)
    names = sorted({k.split("__")[0] for k in z.files})
    out = {n: {f: z[f"{n}__{f}"] for f in ("kinds", "v", "vn", "fv", "fn")} for n in names}
    if os.path.exists(TEXCOORDS_NPZ):
        t = np.load(TEXCOORDS_NPZ)
        for n in sorted({k.split("__")[0] for k in t.files}):
            out[n]["vt"] = t[f"{n}__vt"]; out[n]["ft"] = t[f"{n}__ft"]
    return out


_cache = {}


def obj_path(name, tmpdir=None):
    """Path of a regenerated .obj for fixture `name` (written once per process into a temp dir)."""
    import tempfile
    if name in _cache and os.path.exists(_cache[name]):
        return _cache[name]
    d = tmpdir or tempfile.mkdtemp(prefix="miro_obj_")
    p = os.path.join(d, name + ".obj")
    write_obj(p, load_meshes()[name])
    _cache[name] = p
    return p
