"""CPU tier: the procedural textures (csrc/texture.cuh -- Perlin noise, Worley cells, the reference's colour and bump formulas),
evaluated on the host through mirogpu_texture_lookup / mirogpu_texture_bump -- the same __host__ __device__ code the shading
kernels run -- against the REAL reference's Texture classes:

  * golden vectors generated from the reference compiled in place (tests/golden/ref_textures.npz, make_textures.py), so the check
    also runs where /root/reference does not exist;
  * live against oracle/_ref where it is built, on more points.

The arithmetic follows the reference operation by operation (binary32 / double as its expressions promote), so on the same libm
the colours are bit-identical; the assertion allows last-ulp libm differences (powf, expf, acosf) on a few points."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref_textures.npz")
NAMES = ["checker", "stone", "stone20", "stem", "petal", "leaf", "flower_center"]


def _check(ref_rgb, ref_bump, got_rgb, got_bump):
    assert np.isfinite(got_rgb).all()
    close = np.isclose(got_rgb, ref_rgb, rtol=1e-5, atol=1e-6).all(axis=1)
    assert close.mean() >= 0.999, close.mean()
    assert np.isclose(got_bump, ref_bump, rtol=1e-5, atol=1e-6).mean() >= 0.999


@pytest.mark.parametrize("name", NAMES)
def test_textures_equal_the_reference_golden_vectors(pkg, name):
    z = np.load(GOLD)
    kind, tp, c = int(z[name + "__kind"]), z[name + "__params"], z[name + "__coords"]
    got = pkg.texture_lookup(kind, tp, c)
    bump = pkg.texture_bump(kind, tp, c[:, :2])
    _check(z[name + "__rgb"], z[name + "__bump"], got, bump)
    # on this libm the agreement is exact
    assert (got.view(np.uint32) == z[name + "__rgb"].view(np.uint32)).all(axis=1).mean() > 0.99


@pytest.mark.parametrize("name", NAMES)
def test_textures_equal_the_reference_live(pkg, reference, name):
    z = np.load(GOLD)
    kind, tp = int(z[name + "__kind"]), z[name + "__params"]
    rng = np.random.default_rng(7)
    c = ((rng.random((6000, 3), dtype=np.float32) * 2 - 1) * np.float32(8.0)).astype(np.float32)
    ref, rb = reference.texture_lookup(kind, tp, c, bump=True)
    _check(ref, rb, pkg.texture_lookup(kind, tp, c), pkg.texture_bump(kind, tp, c[:, :2]))


def test_texture_queries_reject_bad_kinds(pkg):
    with pytest.raises(pkg.MiroGpuError):
        pkg.texture_lookup(0, [1.0], np.zeros((1, 3), np.float32))
    with pytest.raises(pkg.MiroGpuError):
        pkg.texture_lookup(99, [1.0], np.zeros((1, 3), np.float32))
